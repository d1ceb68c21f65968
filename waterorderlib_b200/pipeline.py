"""Host-fed trajectory pipeline: frames live in (pinned) host memory, as they do when a trajectory reader
hands them over (structureLibs/orderParam_lib.py:1312-1316 reads one pytraj frame per iteration); the GPU
sees them in batches.  Separate streams overlap the host->device copies of the next batches (n_slots - 1 of them in
flight), the kernels of batch i and the device->host copy of batch i-1's per-water results; consecutive batches run
on alternating kernel streams (each with its own workspace) so that the short, latency-bound tail of one batch -- the
widened search of the few centres the 27-cell sweep could not finish -- overlaps the next batch's sweep.

This is the call bench.py's end-to-end leg and host applications that stream a trajectory make.
"""
import numpy as np
import torch

from . import engine
from ._capi import WOL_NSTATS


class FramePipeline:
    """Fused q + three-body analysis of host-resident frames.

    n_atoms, frames_per_batch fix the device buffers; dtype is the storage type of the host frames
    (float64 as pytraj gives them, or float32 as trajectory files store them).
    """

    def __init__(self, n_atoms, frames_per_batch, dtype=np.float64, device=None, *, do_q=True, do_3body=True,
                 low3=0.0, high3=3.413, lowq=0.0, highq=10.0, nbins=500, bin_range=(0.0, 180.0), q_nbins=500,
                 precision="fp64", hist_per_frame=False, r_cell=None, want_q=True, want_n3=True, want_nn=False, n_slots=4,
                 n_run_streams=2):
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.n_atoms, self.fpb = int(n_atoms), int(frames_per_batch)
        self.tdtype = torch.float64 if np.dtype(dtype) == np.float64 else torch.float32
        self.kw = dict(do_q=do_q, do_3body=do_3body, low3=low3, high3=high3, lowq=lowq, highq=highq, nbins=nbins,
                       bin_range=bin_range, q_nbins=q_nbins, precision=precision, hist_per_frame=False, r_cell=r_cell)
        self.hist_per_frame = hist_per_frame
        self.do_q, self.do_3body = do_q, do_3body
        self.nbins, self.q_nbins = nbins, q_nbins
        self.qdtype = torch.float64 if precision == "fp64" else torch.float32
        self.want_q, self.want_n3, self.want_nn = want_q and do_q, want_n3 and do_3body, want_nn and do_q
        with torch.cuda.device(self.device):
            self.s_in, self.s_out = torch.cuda.Stream(), torch.cuda.Stream()
            self.s_runs = [torch.cuda.Stream() for _ in range(max(1, int(n_run_streams)))]
            B, N = self.fpb, self.n_atoms
            self.slots = []
            self.n_slots = max(2, int(n_slots))
            for _ in range(self.n_slots):
                slot = dict(pos=torch.empty((B, N, 3), dtype=self.tdtype, device=self.device),
                            q=torch.empty((B, N), dtype=self.qdtype, device=self.device) if self.want_q else None,
                            n3=torch.empty((B, N), dtype=torch.int32, device=self.device) if self.want_n3 else None,
                            nn=torch.empty((B, N, 4), dtype=torch.int32, device=self.device) if self.want_nn else None,
                            loaded=torch.cuda.Event(), computed=torch.cuda.Event(), drained=torch.cuda.Event())
                self.slots.append(slot)
            self.wss = [engine.Workspace(self.device) for _ in self.s_runs]
        self.launches = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0
        self.trace = False  # True: time every batch's copy-in / kernels / copy-out with CUDA events (see timeline())
        self._trace = []

    def timeline(self):
        """After a run with ``trace = True``: one row per batch, (copy-in start, end, kernels start, end, copy-out start,
        end) in ms since the first batch's copy-in started.  Synchronises the device."""
        torch.cuda.synchronize(self.device)
        if not self._trace:
            return np.zeros((0, 6))
        t0 = self._trace[0][0]
        return np.array([[t0.elapsed_time(e) for e in row] for row in self._trace])

    def _mark(self, row, k, stream):
        if self.trace:
            e = torch.cuda.Event(enable_timing=True)
            e.record(stream)
            row[k] = e

    def run(self, pos_host, box, out_q=None, out_n3=None, out_nn=None):
        """pos_host (F,N,3) torch CPU tensor (pinned for overlap) or numpy array; box (3,) / (F,3).
        Returns dict(ang_hist, q_hist, frame_stats [, q, n3, nn_idx]) of HOST tensors; per-water arrays
        are written into out_q / out_n3 / out_nn (pinned CPU tensors) when given.  The call returns when every
        result has landed in host memory (it waits for the last device->host copy); a list that overflowed even
        the large-capacity path raises WolError, as the per-call API does.  n_widened / n_overflow of the run are
        in the result."""
        pos_host = pos_host if isinstance(pos_host, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(pos_host))
        if pos_host.dtype != self.tdtype:
            raise ValueError("pipeline was built for %s frames, got %s" % (self.tdtype, pos_host.dtype))
        F, N = int(pos_host.shape[0]), int(pos_host.shape[1])
        if N != self.n_atoms:
            raise ValueError("pipeline was built for %d atoms per frame, got %d" % (self.n_atoms, N))
        box_h = engine.as_host_boxes(box, F)
        dev = self.device
        pin = pos_host.is_pinned()
        if self.want_q and out_q is None:
            out_q = torch.empty((F, N), dtype=self.qdtype, pin_memory=True)
        if self.want_n3 and out_n3 is None:
            out_n3 = torch.empty((F, N), dtype=torch.int32, pin_memory=True)
        if self.want_nn and out_nn is None:
            out_nn = torch.empty((F, N, 4), dtype=torch.int32, pin_memory=True)
        H = F if self.hist_per_frame else 1
        self.launches = self.h2d_bytes = self.d2h_bytes = 0
        self._trace = [[None] * 6 for _ in range((F + self.fpb - 1) // self.fpb)] if self.trace else []
        with torch.cuda.device(dev):
            main = torch.cuda.current_stream()
            acc = {"frame_stats": torch.zeros((F, WOL_NSTATS), dtype=torch.float64, device=dev)}
            if self.do_3body:
                acc["ang_hist"] = torch.zeros((H, self.nbins), dtype=torch.int64, device=dev)
            if self.do_q:
                acc["q_hist"] = torch.zeros((H, self.q_nbins), dtype=torch.int64, device=dev)
            box_d = torch.from_numpy(box_h.copy()).to(dev)  # all boxes once: nothing in the loop blocks the host
            for s in [self.s_in, self.s_out] + self.s_runs:
                s.wait_stream(main)
            starts = list(range(0, F, self.fpb))

            def load(i):
                f0 = starts[i]
                nb = min(self.fpb, F - f0)
                slot = self.slots[i % self.n_slots]
                with torch.cuda.stream(self.s_in):
                    self.s_in.wait_event(slot["computed"])  # the kernels that last read this buffer
                    if self.trace:
                        self._mark(self._trace[i], 0, self.s_in)
                    slot["pos"][:nb].copy_(pos_host[f0:f0 + nb], non_blocking=pin)
                    slot["loaded"].record(self.s_in)
                    if self.trace:
                        self._mark(self._trace[i], 1, self.s_in)
                self.h2d_bytes += nb * N * 3 * pos_host.element_size()

            ahead = self.n_slots - 1  # copies in flight ahead of the kernels
            for i in range(min(ahead, len(starts))):
                load(i)
            for i, f0 in enumerate(starts):
                nb = min(self.fpb, F - f0)
                slot = self.slots[i % self.n_slots]
                if i + ahead < len(starts):
                    load(i + ahead)
                s_run, ws = self.s_runs[i % len(self.s_runs)], self.wss[i % len(self.s_runs)]
                with torch.cuda.stream(s_run):
                    s_run.wait_event(slot["loaded"])
                    s_run.wait_event(slot["drained"])  # the D2H of this slot's previous results
                    out = {"frame_stats": acc["frame_stats"][f0:f0 + nb]}
                    if self.do_3body:
                        out["ang_hist"] = acc["ang_hist"][f0:f0 + nb] if self.hist_per_frame else acc["ang_hist"]
                        if self.want_n3:
                            out["n3"] = slot["n3"][:nb]
                    if self.do_q:
                        out["q_hist"] = acc["q_hist"][f0:f0 + nb] if self.hist_per_frame else acc["q_hist"]
                        if self.want_q:
                            out["q"] = slot["q"][:nb]
                        if self.want_nn:
                            out["nn_idx"] = slot["nn"][:nb]
                    if self.trace:
                        self._mark(self._trace[i], 2, s_run)
                    want = tuple(out.keys())
                    kw = dict(self.kw)
                    kw["hist_per_frame"] = self.hist_per_frame
                    r = engine.q3b_frames(slot["pos"][:nb], box_h[f0:f0 + nb], out=out, want=want, workspace=ws,
                                          device=dev, check_status=False, box_device=box_d[f0:f0 + nb], **kw)
                    self.launches += r["launches"]
                    slot["computed"].record(s_run)
                    if self.trace:
                        self._mark(self._trace[i], 3, s_run)
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(slot["computed"])
                    if self.trace:
                        self._mark(self._trace[i], 4, self.s_out)
                    if self.want_q:
                        out_q[f0:f0 + nb].copy_(slot["q"][:nb], non_blocking=True)
                        self.d2h_bytes += nb * N * out_q.element_size()
                    if self.want_n3:
                        out_n3[f0:f0 + nb].copy_(slot["n3"][:nb], non_blocking=True)
                        self.d2h_bytes += nb * N * 4
                    if self.want_nn:
                        out_nn[f0:f0 + nb].copy_(slot["nn"][:nb], non_blocking=True)
                        self.d2h_bytes += nb * N * 16
                    slot["drained"].record(self.s_out)
                    if self.trace:
                        self._mark(self._trace[i], 5, self.s_out)
            res = {}
            with torch.cuda.stream(self.s_out):
                for s_run in self.s_runs:
                    self.s_out.wait_stream(s_run)
                for k, t in acc.items():
                    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
                    h.copy_(t, non_blocking=True)
                    self.d2h_bytes += t.numel() * t.element_size()
                    res[k] = h
            done = torch.cuda.Event()
            done.record(self.s_out)
            for s in [self.s_out, self.s_in] + self.s_runs:
                main.wait_stream(s)
            # the returned tensors are host memory filled by asynchronous copies: wait for the last of them
            done.synchronize()
            # device-side list capacity: the batches ran with check_status=False (no host sync inside the loop)
            widened = overflow = 0
            for k, ws in enumerate(self.wss):
                nb = [min(self.fpb, F - f0) for i, f0 in enumerate(starts) if i % len(self.s_runs) == k]
                if not nb:
                    continue
                w, o = engine.workspace_status(ws, nb[-1], N, N, self.kw["r_cell"] if self.kw["r_cell"] is not None else
                                               engine.default_r_cell(self.do_q, self.do_3body, self.kw["high3"], self.kw["highq"]),
                                               box_h[:nb[-1]], stream=self.s_runs[k])
                widened += w
                overflow += o
        res["q"], res["n3"], res["nn_idx"] = out_q, out_n3, out_nn
        res["n_widened_last_batches"], res["n_overflow_last_batches"] = widened, overflow
        res["_device"] = acc
        return res
