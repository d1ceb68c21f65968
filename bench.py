#!/usr/bin/env python
"""Benchmark of the per-frame water-structure hot path: tetrahedral q + three-body angle histogram.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[4], the one the metric is quoted on): synthetic jittered-ice boxes of
1,000,000 waters at liquid density (L = 310.3 A), fp64 arithmetic.  A STEP is one pass of the hot path
(cell-list build + fused sweep) over one batch of `--frames-per-step` frames per GPU; frames are sharded
by frame across ranks (weak scaling: per-GPU work fixed), with ONE NCCL all-reduce of the packed int64
histograms after the last step, inside the timed region.

  value        water-frames/s with the batch already resident in HBM when the timed region starts
  e2e          the same through the host-fed per-water API (waterorderlib_b200.pipeline.FramePipeline): float64
               frames start in pinned HOST memory, per-water q / neighbour counts and the histograms end in host
               memory; copies are inside the timed region
  e2e_driver   what the frame drivers (tetOrderCalc / threeBodyCalc) need: float32 host frames as a trajectory
               file stores them in, histograms and per-frame sums out -- nothing per water crosses PCIe
  roofline     HBM roofline of the dominant kernel, timed alone with CUDA events recorded by the library
               around its launch; roofline_fp the same against an FMA peak measured in this run
  cpu_baseline the reference's own compiled Fortran + its per-water Python loops (oracle/) on host cores;
               cpu_like_for_like repeats the GPU on exactly those frames

`--impl reference` times the reference CPU path alone (rank 0 only).
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "water-frames/sec for tetrahedral q + 3-body angle histogram"
UNIT = "water-frames/s"
N_CELLS_1M = 50          # 8 * 50^3 = 1,000,000 waters
SIGMA = 0.25             # jittered ice
SIGMA_LIQUID = 0.6      # liquid-like
BYTES_PER_WF_FP64 = 36   # read xyz 24 B, write q 8 B + neighbour count 4 B
R3, RQ = 3.413, 10.0     # three-body cutoff, q highCut (reference defaults)
RHO = 0.033456           # liquid-water number density the synthetic boxes are built at (water_properties.py:55)


# ---------------------------------------------------------------------------------------------------
# CPU reference leg

def ref_seeds(frames_per_proc, procs):
    return [[1000 + p * frames_per_proc + k for k in range(frames_per_proc)] for p in range(procs)]


def _ref_worker(args):
    """One process: frames of an m^3-cell box through the reference's CPU path.  Returns (waters, seconds spent in the
    reference calls, summed angle histogram, summed q) -- the last two let the GPU leg on the same frames be checked."""
    m, seeds = args
    from oracle import ref_driver, ref_fortran
    from waterorderlib_b200 import synth
    wl = ref_fortran.RefWaterlib()
    t_used = 0.0
    waters = 0
    hist = np.zeros(500, dtype=np.int64)
    q_sum = 0.0
    for seed in seeds:
        pos, box = synth.water_box(m, sigma=SIGMA, seed=seed)
        t0 = time.perf_counter()
        q = ref_driver.get_order_param_q(wl, pos, pos, box)
        ang, _num = ref_driver.get_cos_angs(wl, pos, pos, box)
        h, _ = np.histogram(ang, bins=500, range=[0.0, 180.0])
        t_used += time.perf_counter() - t0
        waters += pos.shape[0]
        hist += h
        q_sum += float(q.sum())
        assert q.shape[0] == pos.shape[0]
    return waters, t_used, hist, q_sum


def reference_sample(m, frames_per_proc, procs):
    """Frame-parallel run of the reference path (it is single-threaded; frames are independent).
    Returns dict(value, seconds, kind, sample, hist, q_sum, seeds)."""
    import multiprocessing as mp
    from oracle import build_oracle, ref_fortran
    build_oracle.build(verbose=False)
    kind = "reference" if ref_fortran.reference_available() else "port"
    seeds = ref_seeds(frames_per_proc, procs)
    hist, q_sum = None, None
    t0 = time.perf_counter()
    if kind == "reference":
        jobs = [(m, s) for s in seeds]
        if procs > 1:
            with mp.get_context("fork").Pool(procs) as pool:
                out = pool.map(_ref_worker, jobs)
        else:
            out = [_ref_worker(jobs[0])]
        waters = sum(o[0] for o in out)
        hist = sum(o[2] for o in out)
        q_sum = sum(o[3] for o in out)
    else:
        # the reference's binary is not staged: time the C restatement (cell list, OpenMP) instead
        from oracle import port
        from waterorderlib_b200 import synth
        waters = 0
        for sl in seeds:
            for seed in sl:
                pos, box = synth.water_box(m, sigma=SIGMA, seed=seed)
                port.order_param_q(pos, pos, box)
                port.three_body(pos, pos, box, materialize=False)
                waters += pos.shape[0]
    dt = time.perf_counter() - t0
    n = 8 * m ** 3
    sample = ("%d frames of a %d-water jittered-ice box, one process per host core (%d), q (highCut 10) + 3-body (3.413) + "
              "histogram per frame; the 1M-water frame of the GPU arm needs a 3.6 TiB dense neighbour matrix in the reference"
              % (frames_per_proc * procs, n, procs))
    return {"value": waters / dt, "seconds": dt, "kind": kind, "sample": sample, "hist": hist, "q_sum": q_sum,
            "seeds": [s for sl in seeds for s in sl], "n_waters": n, "frames": frames_per_proc * procs}


# ---------------------------------------------------------------------------------------------------

class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.sm, self.reasons, self.max_mhz = index, False, [], set(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.05)

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.sm)}


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            return local_rank
    return local_rank


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_traffic(n_waters, frames):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if it matches."""
    path = os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")
    try:
        d = json.load(open(path))
        if d.get("n_waters") == n_waters and "brick" in d.get("kernel", ""):
            return float(d["dram_bytes_per_water_frame"]) * n_waters * frames
    except Exception:  # noqa: BLE001
        pass
    return None


def fma_peaks(dev):
    """FP64 / FP32 FMA throughput of this GPU right now (TFLOP/s), from the library's probe kernel timed with events."""
    import torch
    from waterorderlib_b200._capi import WOL_F32, WOL_F64, check, lib
    sink = torch.zeros(1, dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    out = {}
    for name, code, iters in (("fp64", WOL_F64, 4096), ("fp32", WOL_F32, 8192)):
        blocks = 148 * 16
        best = 0.0
        for rep in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib().wol_fma_probe(code, blocks, iters, ctypes.c_void_p(sink.data_ptr()), stream), "wol_fma_probe")
            e1.record()
            torch.cuda.synchronize()
            if rep:
                best = max(best, 2.0 * 8 * iters * 256 * blocks / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        out[name] = best
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=N_CELLS_1M, help="diamond-cubic cells per edge (50 -> 1M waters)")
    ap.add_argument("--frames-per-step", type=int, default=16, help="frames per GPU per step")
    ap.add_argument("--e2e-batch", type=int, default=0, help="frames per pipeline batch of the end-to-end legs (0: 1 for the per-water legs, 2 for the driver leg)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-cells", type=int, default=8, help="reference sample box: 8 -> 4096 waters")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_waters = 8 * args.cells ** 3
    workload = "%d-water jittered-ice box (sigma %.2f A, L %.2f A), %d frames per GPU per step, sharded by frame" % (
        n_waters, SIGMA, args.cells * 6.2069, args.frames_per_step)
    config = {"workload": workload, "n_waters": n_waters, "frames_per_gpu_per_step": args.frames_per_step,
              "cutoffs": {"three_body": R3, "q_high": RQ}, "bins": 500,
              "cache_policy": "inputs_larger_than_l2 (%.0f MB per step per GPU vs 126 MB L2)" % (
                  args.frames_per_step * n_waters * 24 / 1e6),
              "parallelism": "frames sharded over %d GPU(s), one NCCL all-reduce of histograms at the end" % world}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        vals = []
        t_all = time.perf_counter()
        s = None
        for _ in range(args.warmup + args.steps):
            s = reference_sample(args.ref_cells, 1, cores)
            vals.append((s["value"], s["seconds"]))
        vals = vals[args.warmup:]
        value = float(np.mean([v for v, _ in vals]))
        ms = float(np.mean([dt for _, dt in vals]) * 1e3)
        # what this arm really ran: the reference's dense N x N search cannot take the GPU arm's 1M-water frames
        ref_config = dict(config)
        ref_config["workload"] = ("reference CPU path (compiled Fortran + its per-water Python loops): %d-water jittered-ice frames, "
                                  "one frame per host core per step (%d frames per step); bounded stand-in for the GPU arm's "
                                  "%d-water frames, which need a 3.6 TiB dense neighbour matrix in the reference"
                                  % (s["n_waters"], cores, n_waters))
        ref_config["n_waters"] = s["n_waters"]
        ref_config["frames_per_step"] = cores
        ref_config["gpu_arm_workload"] = workload
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": ref_config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": s["kind"], "sample": s["sample"]},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}
        print(json.dumps(line), flush=True)
        return 0

    ref = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        # before CUDA is initialised: the sample forks one worker per host core
        cores = os.cpu_count() or 1
        ref = reference_sample(args.ref_cells, 2, cores)
        ref["cores"] = cores

    import torch
    import torch.distributed as dist
    from waterorderlib_b200 import distributed as wdist
    from waterorderlib_b200 import engine, synth
    from waterorderlib_b200.pipeline import FramePipeline

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.frames_per_step
    # this rank's frames: generated on the device from absolute frame indices, so results do not depend on the rank count
    pos_d, box = synth.device_frames(args.cells, rank * B, (rank + 1) * B, sigma=SIGMA, device=dev)
    pos_h = torch.empty((B, n_waters, 3), dtype=torch.float64, pin_memory=True)
    pos_h.copy_(pos_d)
    box_h = np.broadcast_to(np.asarray(box, dtype=np.float64), (B, 3)).copy()
    box_d = torch.from_numpy(box_h).to(dev)
    r_cell = engine.default_r_cell(True, True, R3, RQ)
    ws = engine.Workspace(dev)
    out = {"q": torch.zeros((B, n_waters), dtype=torch.float64, device=dev),
           "n3": torch.zeros((B, n_waters), dtype=torch.int32, device=dev),
           "ang_hist": torch.zeros((1, 500), dtype=torch.int64, device=dev),
           "q_hist": torch.zeros((1, 500), dtype=torch.int64, device=dev),
           "frame_stats": torch.zeros((B, 8), dtype=torch.float64, device=dev)}
    want = tuple(out.keys())
    ev_k0, ev_k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev_k0.record()
    ev_k1.record()
    torch.cuda.synchronize()

    def step(timing=None):
        return engine.q3b_frames(pos_d, box_h, out=out, want=want, workspace=ws, device=dev, check_status=False,
                                 timing_events=timing, box_device=box_d)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    launches_per_step = 0
    for _ in range(max(args.warmup, 3)):
        launches_per_step = step()["launches"]
    # dominant-kernel duration, timed alone by the library's events (a few launches, inputs > L2)
    barrier()
    k_ms = []
    for _ in range(5):
        step(timing=(ev_k0, ev_k1))
        torch.cuda.synchronize()
        k_ms.append(ev_k0.elapsed_time(ev_k1))
    kernel_ms = float(np.mean(k_ms))
    fma = fma_peaks(dev)

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    # ---- device-resident timed region -----------------------------------------------------------
    for t in out.values():
        t.zero_()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    wdist.reduce_histograms(out["ang_hist"], out["q_hist"])  # one packed all-reduce (identity on one rank)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    st = engine.workspace_status(ws, B, n_waters, n_waters, r_cell, box_h)
    n_angles = int(out["ang_hist"].sum().item())

    # ---- multi-GPU: the all-reduced histograms must be what ONE rank gets from all the frames ---------------------
    hist_ok = None
    if world > 1:
        ref_ang = torch.zeros((1, 500), dtype=torch.int64, device=dev)
        ref_q = torch.zeros((1, 500), dtype=torch.int64, device=dev)
        if rank == 0:
            ws_chk = engine.Workspace(dev)
            for r in range(world):
                p_r = pos_d if r == 0 else synth.device_frames(args.cells, r * B, (r + 1) * B, sigma=SIGMA, device=dev)[0]
                engine.q3b_frames(p_r, box_h, out={"ang_hist": ref_ang, "q_hist": ref_q}, want=("ang_hist", "q_hist"),
                                  workspace=ws_chk, device=dev, check_status=False, box_device=box_d)
                del p_r
            hist_ok = bool(torch.equal(ref_ang * args.steps, out["ang_hist"]) and torch.equal(ref_q * args.steps, out["q_hist"]))
            del ws_chk
        barrier()

    # ---- FP32 arithmetic mode, same frames, device-resident (north_star: both modes reported) ---------
    out32 = {"q": torch.zeros((B, n_waters), dtype=torch.float32, device=dev),
             "n3": torch.zeros((B, n_waters), dtype=torch.int32, device=dev),
             "ang_hist": torch.zeros((1, 500), dtype=torch.int64, device=dev),
             "q_hist": torch.zeros((1, 500), dtype=torch.int64, device=dev),
             "frame_stats": torch.zeros((B, 8), dtype=torch.float64, device=dev)}
    ws32 = engine.Workspace(dev)

    def step32():
        return engine.q3b_frames(pos_d, box_h, out=out32, want=want, workspace=ws32, device=dev, check_status=False,
                                 precision="fp32", box_device=box_d)

    for _ in range(3):
        step32()
    out32["ang_hist"].zero_()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        step32()
    g1.record()
    barrier()
    ms32 = g0.elapsed_time(g1)
    # accuracy of the mode on one frame, untimed: where both modes pick the same four neighbours q must agree to
    # 1e-4; waters whose 4th/5th neighbour distances agree to ~1e-7 relative may pick differently
    v64 = engine.q3b_frames(pos_d[:1], box, want=("q", "nn_idx"), device=dev, do_3body=False)
    v32 = engine.q3b_frames(pos_d[:1], box, want=("q", "nn_idx"), device=dev, do_3body=False, precision="fp32")
    same_nn = (v64["nn_idx"] == v32["nn_idx"]).all(dim=-1)
    dq = (v32["q"].double() - v64["q"]).abs()
    q_err32 = float(dq[same_nn].max().item())
    flips32 = float((~same_nn).double().mean().item())
    del v64, v32
    hist_mine = out["ang_hist"] if world == 1 else None
    hist_l1 = None
    if hist_mine is not None:
        hist_l1 = float((out32["ang_hist"] - hist_mine).abs().sum().item()) / max(1.0, float(hist_mine.sum().item()))
    del out32, ws32

    # ---- liquid-like box (sigma 0.6 A: broader neighbour-count distribution, ~2 % of the centres need the widened
    # search), same size and density, fp64, device-resident (north_star: jittered-ice / liquid-density boxes) ----------
    pos_l = synth.device_frames(args.cells, rank * B, (rank + 1) * B, sigma=SIGMA_LIQUID, device=dev)[0]

    def step_liquid():
        return engine.q3b_frames(pos_l, box_h, out=out_l, want=want, workspace=ws, device=dev, check_status=False, box_device=box_d)

    out_l = {k: torch.zeros_like(t) for k, t in out.items()}  # (the jittered-ice results in `out` are checked further down)
    for _ in range(3):
        step_liquid()
    barrier()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(args.steps):
        step_liquid()
    l1.record()
    barrier()
    ms_liquid = l0.elapsed_time(l1)
    st_liquid = engine.workspace_status(ws, B, n_waters, n_waters, r_cell, box_h)
    del pos_l, out_l

    # ---- end-to-end legs (host buffers in, host results out; copies inside the timed region) ----------------------
    def e2e_leg(dtype, per_water):
        # the per-water legs are bound by the host link, where a one-frame first batch starts the kernels soonest; the
        # driver leg is bound by the kernels, which like two frames per batch better (scripts/e2e_batch_sweep.py)
        eb = max(1, min(B, args.e2e_batch if args.e2e_batch > 0 else (1 if per_water else 2)))
        pipe = FramePipeline(n_waters, eb, dtype=dtype, device=dev, want_q=per_water, want_n3=per_water)
        src = pos_h if dtype == np.float64 else pos_h32
        qh = torch.empty((B, n_waters), dtype=torch.float64, pin_memory=True) if per_water else None
        nh = torch.empty((B, n_waters), dtype=torch.int32, pin_memory=True) if per_water else None
        for _ in range(2):
            pipe.run(src, box_h, out_q=qh, out_n3=nh)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host = time.perf_counter()
        f0.record()
        launches = 0
        r = None
        for _ in range(args.steps):
            r = pipe.run(src, box_h, out_q=qh, out_n3=nh)
            launches += pipe.launches
        wdist.reduce_histograms(r["_device"]["ang_hist"], r["_device"]["q_hist"])
        f1.record()
        barrier()
        wall_ms = (time.perf_counter() - t_host) * 1e3
        ms = wdist.max_over_ranks(f0.elapsed_time(f1), dev)
        ok = bool(torch.equal(qh.to(dev), out["q"])) if per_water else None
        hist_same = bool(torch.equal(r["ang_hist"].to(dev) * args.steps, out["ang_hist"])) if world == 1 else None
        leg = {"value": float(world) * B * n_waters * args.steps / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / args.steps,
               "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes, "frames_per_batch": eb,
               "host_wall_ms_per_step": wall_ms / args.steps}
        if ok is not None:
            leg["matches_device_run"] = ok
        if hist_same is not None:
            leg["histogram_matches_device_run"] = hist_same
        del pipe
        return leg, launches

    pos_h32 = torch.empty((B, n_waters, 3), dtype=torch.float32, pin_memory=True)
    pos_h32.copy_(pos_h)
    e2e, l1 = e2e_leg(np.float64, True)
    e2e["api"] = "waterorderlib_b200.pipeline.FramePipeline.run (float64 host frames in; q, neighbour counts, histograms out)"
    e2e32, l2 = e2e_leg(np.float32, True)
    e2e32["note"] = "float32 host frames as a NetCDF trajectory stores them; fp64 arithmetic, identical results"
    e2e_drv, l3 = e2e_leg(np.float32, False)
    e2e_drv["api"] = ("FramePipeline.run(want_q=False, want_n3=False): float32 host frames in, histograms + per-frame sums out -- "
                      "what tetOrderCalc / threeBodyCalc consume (reference orderParam_lib.py:1484-1501, :1355-1382)")
    e2e_launches = l1 + l2 + l3
    sampler.stop_flag = True
    sampler.join(timeout=2.0)

    # ---- like for like with the CPU reference: the GPU on exactly the frames the reference sample ran -------------
    like = None
    if ref is not None and ref["hist"] is not None:
        frames = [synth.water_box(args.ref_cells, sigma=SIGMA, seed=s) for s in ref["seeds"]]
        xyz = np.stack([f[0] for f in frames])
        bxs = np.stack([f[1] for f in frames])
        host = torch.from_numpy(xyz).pin_memory()
        small = FramePipeline(xyz.shape[1], xyz.shape[0], dtype=np.float64, device=dev)
        for _ in range(3):
            small.run(host, bxs)
        torch.cuda.synchronize()
        reps = 20
        t0 = time.perf_counter()
        for _ in range(reps):
            rs = small.run(host, bxs)
        dt = (time.perf_counter() - t0) / reps
        gpu_wf_s = xyz.shape[0] * xyz.shape[1] / dt
        like = {"frames": int(xyz.shape[0]), "n_waters": int(xyz.shape[1]), "gpu_wf_s": gpu_wf_s, "cpu_wf_s": ref["value"],
                "cores": ref["cores"], "ratio": gpu_wf_s / ref["value"],
                "gpu_path": "FramePipeline.run on host float64 frames, wall clock incl. copies and the host sync, one batch",
                "histogram_equals_reference": bool(np.array_equal(rs["ang_hist"][0].numpy(), ref["hist"])),
                "q_sum_rel_diff": abs(float(rs["frame_stats"][:, 0].sum()) - ref["q_sum"]) / abs(ref["q_sum"])}

    ms_total = wdist.max_over_ranks(ms_total, dev)
    ms32 = wdist.max_over_ranks(ms32, dev)
    ms_liquid = wdist.max_over_ranks(ms_liquid, dev)
    kernel_ms = wdist.max_over_ranks(kernel_ms, dev)
    wf_per_step = float(world) * B * n_waters
    value = wf_per_step * args.steps / (ms_total * 1e-3)
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_WF_FP64 * B * n_waters / (kernel_ms * 1e-3) / 1e9

    # FP roofline of the dominant kernel, SURVEY 8(d): 17 flop per candidate pair of a 27-cell stencil of edge r_c
    # (C = 27 r_c^3 rho) + 37 per angle (three-body angles, measured, + 6 for q).  The sweep really visits cells of the
    # planned edge (>= 3.8 A for the q search): those extra candidates are reported as overhead, not as useful work.
    nc_used = step()["nc"]
    cand_alg = 27.0 * R3 ** 3 * RHO
    cand_swept = 27.0 * n_waters / float(nc_used[0] * nc_used[1] * nc_used[2])
    n_ang_local = n_angles / float(world) if world > 1 else n_angles
    ang_per_wf = n_ang_local / (float(B) * n_waters * args.steps) + 6.0
    flop_per_wf = 17.0 * cand_alg + 37.0 * ang_per_wf
    fp_achieved = flop_per_wf * B * n_waters / (kernel_ms * 1e-3) / 1e12

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "frames_per_s": value / n_waters,
            "e2e": e2e, "e2e_f32_host_frames": e2e32, "e2e_driver": e2e_drv,
            "gpu_launches": int(launches_per_step * args.steps + e2e_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": kernel_traffic(n_waters, B), "kernel": "wol::q3b_brick_kernel",
                         "kernel_ms": kernel_ms, "bytes_per_water_frame": BYTES_PER_WF_FP64, "peak_source": peak_src,
                         "note": "the sweep is issue/FP-bound by construction (about 1.1 kFLOP per water-frame, SURVEY 8d); see roofline_fp"},
            "roofline_fp": {"bound": "fp64 FMA pipe", "flop_per_water_frame": flop_per_wf,
                            "candidates_per_water_algorithmic": cand_alg, "candidates_per_water_swept": cand_swept,
                            "angles_per_water": ang_per_wf, "achieved": fp_achieved, "peak": fma["fp64"], "unit": "TFLOP/s",
                            "frac": fp_achieved / fma["fp64"] if fma["fp64"] > 0 else None,
                            "peak_source": "measured in this run (wol_fma_probe, DFMA, 8 independent chains per thread)",
                            "fp32_fma_peak_measured": fma["fp32"],
                            "note": "the float prefilter runs on the FP32 pipe, the exact re-evaluation and the angles on the FP64 pipe"},
            "clocks": sampler.summary(),
            "fp32_mode": {"value": float(world) * B * n_waters * args.steps / (ms32 * 1e-3), "unit": UNIT,
                          "ms_per_step": ms32 / args.steps, "max_abs_q_error_vs_fp64_same_neighbours": q_err32,
                          "fraction_with_different_4nn": flips32,
                          "angle_hist_L1_distance_vs_fp64": hist_l1, "tolerance": 1e-4},
            "liquid_like": {"value": float(world) * B * n_waters * args.steps / (ms_liquid * 1e-3), "unit": UNIT,
                            "ms_per_step": ms_liquid / args.steps, "sigma": SIGMA_LIQUID,
                            "widened_per_step": st_liquid[0], "overflow": st_liquid[1],
                            "note": "same box, density and arithmetic (fp64) with a jitter of 0.6 A, inputs resident in HBM"},
            "checks": {"angles_binned": n_angles, "widened": st[0], "overflow": st[1]}}
    if hist_ok is not None:
        line["checks"]["hist_equals_single_rank"] = hist_ok
    if ref is not None:
        line["cpu_baseline"] = {"value": ref["value"], "unit": UNIT, "cores": ref["cores"], "kind": ref["kind"], "sample": ref["sample"],
                                "seconds": ref["seconds"]}
    if like is not None:
        line["cpu_like_for_like"] = like
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
