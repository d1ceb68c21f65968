"""Host side of the fused q / three-body path: owns the device buffers (PyTorch tensors), plans the cell
grid, and calls the C ABI (include/wol_capi.h) on the current CUDA stream.

PyTorch is plumbing only here: device memory, streams, host<->device copies.  Every number comes out of
libwol.so's sm_100a kernels; nothing in this module computes on the CPU.
"""
import ctypes

import numpy as np
import torch

from . import _capi
from ._capi import WOL_F32, WOL_F64, WOL_NSTATS, WOL_PREC_FP32, WOL_PREC_FP64, WOL_TABLE_EXTRA, Q3bArgs, check, lib

_I3 = ctypes.c_int32 * 3


def _dtype_code(t):
    if t.dtype == torch.float64:
        return WOL_F64
    if t.dtype == torch.float32:
        return WOL_F32
    raise ValueError("positions must be float64 or float32, got %s" % t.dtype)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def as_device_positions(a, device):
    """(F,N,3) or (N,3) array/tensor -> contiguous (F,N,3) CUDA tensor, dtype kept if float32/float64."""
    if isinstance(a, torch.Tensor):
        t = a
    else:
        arr = np.asarray(a)
        if arr.dtype not in (np.float32, np.float64):
            arr = arr.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(arr))
    if t.dtype not in (torch.float32, torch.float64):
        t = t.to(torch.float64)
    if t.dim() == 2:
        t = t.unsqueeze(0)
    if t.dim() != 3 or t.shape[-1] != 3:
        raise ValueError("positions must have shape (n,3) or (frames,n,3), got %s" % (tuple(t.shape),))
    return t.to(device, non_blocking=True).contiguous()


def as_host_boxes(box, n_frames):
    """BoxDims as the reference accepts it ((3,), (1,3), or per frame (F,3)) -> float64 (F,3) numpy."""
    if isinstance(box, torch.Tensor):
        box = box.detach().cpu().numpy()
    b = np.asarray(box, dtype=np.float64)
    if b.size == 3:
        b = np.broadcast_to(b.reshape(1, 3), (n_frames, 3))
    b = np.ascontiguousarray(b.reshape(-1, 3))
    if b.shape[0] != n_frames:
        raise ValueError("need one box per frame: %d boxes for %d frames" % (b.shape[0], n_frames))
    return b


_TABLE_CACHE = {}


def angle_table(hist_lo, hist_hi, nbins, device, tet_lo=100.0, tet_hi=120.0):
    """Device copy of wol_angle_table for this histogram spec (cached per device)."""
    key = (float(hist_lo), float(hist_hi), int(nbins), float(tet_lo), float(tet_hi), str(device))
    t = _TABLE_CACHE.get(key)
    if t is None:
        host = np.zeros(nbins + 1 + WOL_TABLE_EXTRA, dtype=np.float64)
        check(lib().wol_angle_table(hist_lo, hist_hi, nbins, tet_lo, tet_hi, host.ctypes.data_as(ctypes.c_void_p)),
              "wol_angle_table")
        t = torch.from_numpy(host).to(device)
        _TABLE_CACHE[key] = t
    return t


def effective_boxes(box_host, pos_d, cen_d, reach, device):
    """The reference's "negative edge = axis not periodic" (fortran/waterlib.f90:41) for the cell-list paths: every
    negative edge becomes an equivalent period (wol_effective_box: bit-identical arithmetic, no image within `reach`).
    Boxes without a negative edge come back unchanged, without touching the device."""
    if not (box_host < 0.0).any():
        return box_host
    F = int(box_host.shape[0])
    out = np.empty_like(box_host)
    with torch.cuda.device(device):
        scratch = torch.empty(F * 6, dtype=torch.int64, device=device)
        check(lib().wol_effective_box(_ptr(pos_d), _dtype_code(pos_d), F, int(pos_d.shape[1]),
                                      _ptr(cen_d) if cen_d is not None else None, _dtype_code(cen_d) if cen_d is not None else 0,
                                      int(cen_d.shape[1]) if cen_d is not None else 0,
                                      box_host.ctypes.data_as(ctypes.c_void_p), float(reach), _ptr(scratch),
                                      out.ctypes.data_as(ctypes.c_void_p), _stream_ptr(device)), "wol_effective_box")
    return out


def plan_grid(box_host, r_cell):
    nc = _I3()
    edge = ctypes.c_double(0.0)
    bmax = ctypes.c_double(0.0)
    check(lib().wol_plan_grid(box_host.ctypes.data_as(ctypes.c_void_p), box_host.shape[0], float(r_cell),
                              ctypes.byref(nc), ctypes.byref(edge), ctypes.byref(bmax)), "wol_plan_grid")
    return nc, edge.value, bmax.value


class Q3bResult(dict):
    """Outputs of one fused call (torch tensors on the device) plus launch bookkeeping."""
    __getattr__ = dict.__getitem__


class Workspace:
    """Caller-owned scratch, grown on demand and reused between calls."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes + 256:
            old = self.buf
            self.buf = torch.empty(int(nbytes) + 512, dtype=torch.uint8, device=self.device)
            self.buf[:512].zero_()  # only the counters (first 256 bytes after alignment) must start at zero
            if old is not None:
                # the device counters (first 256 bytes, incl. the sticky overflow flag) move with the workspace
                o0, o1 = (-old.data_ptr()) % 256, (-self.buf.data_ptr()) % 256
                self.buf[o1:o1 + 256].copy_(old[o0:o0 + 256])
        off = (-self.buf.data_ptr()) % 256
        return self.buf.data_ptr() + off, self.buf.numel() - off


_GRAPHS = {}          # signature -> captured call (see q3b_frames_graphed)
_GRAPH_CACHE_MAX = 16
GRAPH_MAX_ATOMS = 1 << 18  # below this many atoms per call the ~10 launches of a call cost more than its kernels


def q3b_frames_graphed(pos, box, centres=None, *, want=("q", "nn_idx", "n3", "ang_hist", "q_hist", "frame_stats"), device=None,
                       check_status=True, **opts):
    """q3b_frames for small, repeated calls (a driver looping over frames of a few thousand waters): the whole launch
    sequence -- cell build, sweep, queued passes -- is captured ONCE into a CUDA graph per call signature (shapes, dtypes,
    options and the exact box: cutoffs and rounding margins derived from the box are baked into the kernels' arguments)
    and replayed afterwards, which replaces ~10 launches by one.  Inputs are copied into the graph's own buffers; the
    returned tensors ARE the graph's output buffers and are overwritten by the next call with the same signature --
    copy what must survive (the numpy-returning wrappers do).  Histograms and frame_stats start from zero each call."""
    if device is None:
        device = pos.device if isinstance(pos, torch.Tensor) and pos.is_cuda else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    pos_t = pos if isinstance(pos, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(pos)))
    if pos_t.dtype not in (torch.float32, torch.float64):
        pos_t = pos_t.to(torch.float64)
    if pos_t.dim() == 2:
        pos_t = pos_t.unsqueeze(0)
    cen_t = None
    if centres is not None:
        cen_t = centres if isinstance(centres, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(centres)))
        if cen_t.dtype not in (torch.float32, torch.float64):
            cen_t = cen_t.to(torch.float64)
        if cen_t.dim() == 2:
            cen_t = cen_t.unsqueeze(0)
    F = int(pos_t.shape[0])
    box_h = as_host_boxes(box, F)
    key = (str(device), tuple(pos_t.shape), pos_t.dtype, None if cen_t is None else (tuple(cen_t.shape), cen_t.dtype),
           tuple(want), tuple(sorted((k, repr(v)) for k, v in opts.items())), box_h.tobytes())
    g = _GRAPHS.get(key)
    if g is None:
        if len(_GRAPHS) >= _GRAPH_CACHE_MAX:
            _GRAPHS.pop(next(iter(_GRAPHS)))
        g = {"pos": torch.empty(pos_t.shape, dtype=pos_t.dtype, device=device),
             "cen": None if cen_t is None else torch.empty(cen_t.shape, dtype=cen_t.dtype, device=device),
             "box": torch.from_numpy(box_h.copy()).to(device), "ws": Workspace(device)}
        g["pos"].copy_(pos_t)
        if cen_t is not None:
            g["cen"].copy_(cen_t)
        # eager once: loads the kernels, sizes the workspace, builds the bin table, creates the output buffers
        r0 = q3b_frames(g["pos"], box_h, g["cen"], want=want, workspace=g["ws"], device=device, check_status=False,
                        box_device=g["box"], **opts)
        g["out"] = {k: r0[k] for k in want if k in r0}
        torch.cuda.synchronize(device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for k in ("ang_hist", "q_hist", "frame_stats"):
                if k in g["out"]:
                    g["out"][k].zero_()
            rg = q3b_frames(g["pos"], box_h, g["cen"], want=tuple(g["out"].keys()), out=g["out"], workspace=g["ws"], device=device,
                            check_status=False, box_device=g["box"], **opts)
        g["graph"], g["meta"] = graph, {k: rg[k] for k in ("launches", "nc", "edge_min")}
        _GRAPHS[key] = g
    with torch.cuda.device(device):
        g["pos"].copy_(pos_t, non_blocking=True)
        if cen_t is not None:
            g["cen"].copy_(cen_t, non_blocking=True)
        g["graph"].replay()
        res = Q3bResult(g["out"])
        res.update(g["meta"])
        res["graph"] = True
        if check_status:
            M = int(cen_t.shape[1]) if cen_t is not None else int(pos_t.shape[1])
            ws_ptr, _ = g["ws"].get(0)
            st = (ctypes.c_int32 * 4)()
            nc = _I3(*g["meta"]["nc"])
            check(lib().wol_status(ctypes.c_void_p(ws_ptr), F, int(pos_t.shape[1]), M, ctypes.byref(nc), _stream_ptr(device),
                                   ctypes.byref(st)), "wol_status")
            res["n_widened"], res["n_overflow"], res["n_slow_pairs"] = int(st[0]), int(st[1]), int(st[3])
    return res


def q3b_frames(pos, box, centres=None, *, do_q=True, do_3body=True, low3=0.0, high3=3.413, lowq=0.0, highq=10.0,
               nbins=500, bin_range=(0.0, 180.0), q_nbins=500, precision="fp64", hist_per_frame=False, r_cell=None,
               want=("q", "nn_idx", "n3", "ang_hist", "q_hist", "frame_stats"), out=None, workspace=None,
               device=None, check_status=True, timing_events=None, box_device=None, n_valid=None, reuse_cells=False):
    """Fused tetrahedral q + three-body angle histogram for a batch of frames.

    pos      (F,N,3) positions of all atoms that can be neighbours (reference: `Pos`)
    box      (3,), (1,3) or (F,3) orthorhombic box edges (reference: `BoxDims`)
    centres  None = every atom of pos is a centre (reference: subPos is Pos); else (F,M,3) (`subPos`)
    out      optional dict of preallocated output tensors to accumulate into / overwrite
    box_device  optional (F,3) float64 CUDA tensor holding the same boxes (skips the upload)
    n_valid  optional (F,) counts: frame f evaluates only its first n_valid[f] centres (ragged sub-populations padded to M)
    reuse_cells  the workspace already holds the cell list of exactly these pos / box / r_cell (skip the build)
    Returns a Q3bResult of device tensors; histograms and frame_stats ACCUMULATE into `out` if given.
    """
    if device is None:
        device = pos.device if isinstance(pos, torch.Tensor) and pos.is_cuda else torch.device("cuda", torch.cuda.current_device())
    L = lib()
    pos_d = as_device_positions(pos, device)
    F, N = int(pos_d.shape[0]), int(pos_d.shape[1])
    box_h = as_host_boxes(box, F)
    cen_d = None
    M = N
    if centres is not None:
        cen_d = as_device_positions(centres, device)
        if cen_d.shape[0] != F:
            raise ValueError("centres and pos must hold the same number of frames")
        M = int(cen_d.shape[1])
    prec = {"fp64": WOL_PREC_FP64, "fp32": WOL_PREC_FP32}[precision]
    if r_cell is None:
        r_cell = default_r_cell(do_q, do_3body, high3, highq)
    if (box_h < 0.0).any():  # open axes (the reference's negative edges): equivalent periods, see effective_boxes
        if box_device is not None:
            raise ValueError("box_device cannot be combined with non-periodic (negative) box edges")
        box_h = effective_boxes(box_h, pos_d, cen_d, max(high3 if do_3body else 0.0, highq if do_q else 0.0, r_cell), device)
    # a pageable host->device copy blocks the host behind everything queued on the stream: callers that pipeline
    # batches upload all boxes once and pass the slice (box_device)
    box_d = box_device if box_device is not None else torch.from_numpy(box_h.copy()).to(device)
    nc, edge_min, box_max = plan_grid(box_h, r_cell)
    ws = workspace if workspace is not None else Workspace(device)
    need = L.wol_workspace_bytes(F, N, M, ctypes.byref(nc))
    if reuse_cells and (ws.buf is None or ws.buf.numel() < need + 256):
        raise ValueError("reuse_cells: this workspace does not hold a cell list for a batch of this shape")
    ws_ptr, ws_bytes = ws.get(need)
    stream = _stream_ptr(device)
    launches = 0
    with torch.cuda.device(device):
        if not reuse_cells:
            check(L.wol_cell_build(_ptr(pos_d), _dtype_code(pos_d), _ptr(box_d), F, N, ctypes.byref(nc), prec,
                                   ctypes.c_void_p(ws_ptr), ws_bytes, stream), "wol_cell_build")
            launches += L.wol_last_launch_count()
        res = Q3bResult()
        out = out or {}
        qdtype = torch.float64 if prec == WOL_PREC_FP64 else torch.float32
        H = F if hist_per_frame else 1

        def buf(name, shape, dtype, zero):
            if name not in want:
                return None
            t = out.get(name)
            if t is None:
                t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=device)
            res[name] = t
            return t

        q = buf("q", (F, M), qdtype, True) if do_q else None
        nn = buf("nn_idx", (F, M, 4), torch.int32, False) if do_q else None
        n3 = buf("n3", (F, M), torch.int32, True) if do_3body else None
        ah = buf("ang_hist", (H, nbins), torch.int64, True) if do_3body else None
        qh = buf("q_hist", (H, q_nbins), torch.int64, True) if do_q else None
        fs = buf("frame_stats", (F, WOL_NSTATS), torch.float64, True)
        table = angle_table(bin_range[0], bin_range[1], nbins, device) if do_3body else None

        a = Q3bArgs()
        a.struct_size = ctypes.sizeof(Q3bArgs)
        a.precision = prec
        a.n_frames, a.n_pos, a.n_centres = F, N, M
        a.centre_dtype = _dtype_code(cen_d) if cen_d is not None else WOL_F64
        a.centres = cen_d.data_ptr() if cen_d is not None else None
        a.box = box_d.data_ptr()
        a.workspace, a.workspace_bytes = ws_ptr, ws_bytes
        a.nc = nc
        a.hist_per_frame = 1 if hist_per_frame else 0
        a.edge_min = edge_min
        a.box_max = box_max
        a.low3, a.high3, a.lowq, a.highq = float(low3), float(high3), float(lowq), float(highq)
        a.do_q, a.do_3body = int(bool(do_q)), int(bool(do_3body))
        a.nbins, a.q_nbins = int(nbins), int(q_nbins)
        a.hist_lo, a.hist_hi = float(bin_range[0]), float(bin_range[1])
        a.angle_table = table.data_ptr() if table is not None else None
        for name, t in (("q", q), ("nn_idx", nn), ("n3", n3), ("ang_hist", ah), ("q_hist", qh), ("frame_stats", fs)):
            setattr(a, name, t.data_ptr() if t is not None else None)
        nv_d = None
        if n_valid is not None:
            if cen_d is None:
                raise ValueError("n_valid needs explicit centres")
            nv_d = n_valid if isinstance(n_valid, torch.Tensor) else torch.as_tensor(np.asarray(n_valid, dtype=np.int32))
            nv_d = nv_d.to(device=device, dtype=torch.int32).contiguous()
            if nv_d.numel() != F:
                raise ValueError("n_valid must hold one count per frame")
            a.n_valid = nv_d.data_ptr()
        if timing_events is not None:  # (begin, end) torch.cuda.Event pair, already recorded once
            a.timing_event_begin, a.timing_event_end = timing_events[0].cuda_event, timing_events[1].cuda_event
        check(L.wol_q3b_frames(ctypes.byref(a), stream), "wol_q3b_frames")
        launches += L.wol_last_launch_count()
        res["launches"] = launches
        res["nc"] = tuple(nc)
        res["edge_min"] = edge_min
        if check_status:
            st = (ctypes.c_int32 * 4)()
            check(L.wol_status(ctypes.c_void_p(ws_ptr), F, N, M, ctypes.byref(nc), stream, ctypes.byref(st)), "wol_status")
            res["n_widened"], res["n_overflow"] = int(st[0]), int(st[1])
            res["n_slow_pairs"] = int(st[3])  # brick path: pairs decided by the exact re-evaluation
    # keep inputs alive until the stream has consumed them
    res["_keep"] = (pos_d, box_d, cen_d, ws, table, nv_d)
    return res


def default_r_cell(do_q, do_3body, high3, highq):
    """Cell edge request: at least the three-body cutoff (the 27-cell sweep must contain it).  For q the 4th
    nearest oxygen of liquid water sits inside ~3.8 A for all but a few per cent of the molecules; measured on
    1M-water boxes (B200): 3.5 A sends 4 % (jittered ice) / 20 % (liquid-like) of the centres to the widened
    search, 3.8 A 0.2 % / 1.5 %, for ~25 % more candidates per sweep -- a net gain of 2 % / 30 %."""
    r = 0.0
    if do_3body:
        r = max(r, float(high3))
    if do_q:
        r = max(r, min(float(highq), 3.8))
    return max(r, 1e-3)


def workspace_status(ws, n_frames, n_pos, n_centres, r_cell, box, stream=None):
    """(widened, overflow) of the last evaluation on `ws`; raises WolError if a list overflowed the
    large-capacity path at any point since the previous call (the flag is sticky).  Synchronises `stream`
    (default: the current stream of the workspace's device)."""
    box_h = as_host_boxes(box, n_frames)
    nc, _edge, _bmax = plan_grid(box_h, r_cell)
    need = lib().wol_workspace_bytes(n_frames, n_pos, n_centres, ctypes.byref(nc))
    ws_ptr, _ = ws.get(need)
    st = (ctypes.c_int32 * 4)()
    with torch.cuda.device(ws.device):
        sp = ctypes.c_void_p(stream.cuda_stream) if stream is not None else _stream_ptr(ws.device)
        check(lib().wol_status(ctypes.c_void_p(ws_ptr), n_frames, n_pos, n_centres, ctypes.byref(nc), sp,
                               ctypes.byref(st)), "wol_status")
    return int(st[0]), int(st[1])
