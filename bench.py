#!/usr/bin/env python
"""Benchmark of the per-frame water-structure hot path: tetrahedral q + three-body angle histogram.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[4], the one the metric is quoted on): synthetic jittered-ice boxes of
1,000,000 waters at liquid density (L = 310.3 A), fp64 arithmetic.  A STEP is one pass of the hot path
(cell-list build + fused sweep) over one batch of `--frames-per-step` frames per GPU; frames are sharded
by frame across ranks (weak scaling: per-GPU work fixed), with one NCCL all-reduce of the int64
histograms after the last step, inside the timed region.

  value      water-frames/s with the batch already resident in HBM when the timed region starts
  e2e        the same through the public host-fed API (waterorderlib_b200.pipeline.FramePipeline): the
             batch starts in pinned HOST memory, the per-water q / neighbour counts and the histograms end
             in host memory; copies are inside the timed region
  roofline   HBM roofline of the dominant kernel (the fused sweep), timed alone with CUDA events recorded
             by the library around its launch
  cpu_baseline  the reference's own compiled Fortran + its per-water Python loops (oracle/) on host cores

`--impl reference` times the reference CPU path alone (rank 0 only).
"""
import argparse
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
BYTES_PER_WF_FP64 = 36   # read xyz 24 B, write q 8 B + neighbour count 4 B


# ---------------------------------------------------------------------------------------------------
# CPU reference leg

def _ref_worker(args):
    """One process: `n` frames of an m^3-cell box through the reference's CPU path."""
    m, seeds = args
    from oracle import ref_driver, ref_fortran
    from waterorderlib_b200 import synth
    wl = ref_fortran.RefWaterlib()
    t_used = 0.0
    waters = 0
    for seed in seeds:
        pos, box = synth.water_box(m, sigma=SIGMA, seed=seed)
        t0 = time.perf_counter()
        q = ref_driver.get_order_param_q(wl, pos, pos, box)
        ang, _num = ref_driver.get_cos_angs(wl, pos, pos, box)
        np.histogram(ang, bins=500, range=[0.0, 180.0])
        t_used += time.perf_counter() - t0
        waters += pos.shape[0]
        assert q.shape[0] == pos.shape[0]
    return waters, t_used


def reference_sample(m, frames_per_proc, procs):
    """Frame-parallel run of the reference path (it is single-threaded; frames are independent).
    Returns (water-frames/s, seconds, description)."""
    import multiprocessing as mp
    from oracle import build_oracle, ref_fortran
    build_oracle.build(verbose=False)
    kind = "reference" if ref_fortran.reference_available() else "port"
    t0 = time.perf_counter()
    if kind == "reference":
        jobs = [(m, [1000 + p * frames_per_proc + k for k in range(frames_per_proc)]) for p in range(procs)]
        if procs > 1:
            with mp.get_context("fork").Pool(procs) as pool:
                out = pool.map(_ref_worker, jobs)
        else:
            out = [_ref_worker(jobs[0])]
        waters = sum(o[0] for o in out)
    else:
        # the reference's binary is not staged: time the C restatement (cell list, OpenMP) instead
        from oracle import port
        from waterorderlib_b200 import synth
        waters = 0
        for k in range(frames_per_proc * procs):
            pos, box = synth.water_box(m, sigma=SIGMA, seed=1000 + k)
            port.order_param_q(pos, pos, box)
            port.three_body(pos, pos, box, materialize=False)
            waters += pos.shape[0]
    dt = time.perf_counter() - t0
    n = 8 * m ** 3
    sample = ("%d frames of a %d-water box (the 1M-water frame needs a 3.6 TiB dense neighbour matrix in the reference), "
              "%d processes, q (highCut 10) + 3-body (3.413) + histogram per frame" % (frames_per_proc * procs, n, procs))
    return waters / dt, dt, kind, sample


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
        if d.get("n_waters") == n_waters:
            return float(d["dram_bytes_per_water_frame"]) * n_waters * frames
    except Exception:  # noqa: BLE001
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=N_CELLS_1M, help="diamond-cubic cells per edge (50 -> 1M waters)")
    ap.add_argument("--frames-per-step", type=int, default=16, help="frames per GPU per step")
    ap.add_argument("--e2e-batch", type=int, default=1, help="frames per pipeline batch of the end-to-end leg")
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
              "cutoffs": {"three_body": 3.413, "q_high": 10.0}, "bins": 500,
              "cache_policy": "inputs_larger_than_l2 (%.0f MB per step per GPU vs 126 MB L2)" % (
                  args.frames_per_step * n_waters * 24 / 1e6),
              "parallelism": "frames sharded over %d GPU(s), one NCCL all-reduce of histograms at the end" % world}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cores = os.cpu_count() or 1
        vals = []
        t_all = time.perf_counter()
        for _ in range(args.warmup + args.steps):
            v, dt, kind, sample = reference_sample(args.ref_cells, 1, cores)
            vals.append((v, dt))
        vals = vals[args.warmup:]
        value = float(np.mean([v for v, _ in vals]))
        ms = float(np.mean([dt for _, dt in vals]) * 1e3)
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0, "wall_s": time.perf_counter() - t_all}
        print(json.dumps(line), flush=True)
        return 0

    cpu_baseline = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        # before CUDA is initialised: the sample forks one worker per host core
        cores = os.cpu_count() or 1
        v, dt, kind, sample = reference_sample(args.ref_cells, 2, cores)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "seconds": dt}

    import torch
    import torch.distributed as dist
    from waterorderlib_b200 import engine, synth
    from waterorderlib_b200.pipeline import FramePipeline

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.frames_per_step
    # this rank's frames: absolute seeds, so results do not depend on the number of ranks
    pos_h = torch.empty((B, n_waters, 3), dtype=torch.float64, pin_memory=True)
    box = None
    for b in range(B):
        p, box = synth.water_box(args.cells, sigma=SIGMA, seed=rank * B + b)
        pos_h[b].copy_(torch.from_numpy(p))
    pos_d = pos_h.to(dev)
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
        return engine.q3b_frames(pos_d, box, out=out, want=want, workspace=ws, device=dev, check_status=False,
                                 timing_events=timing)

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
    if world > 1:
        dist.all_reduce(out["ang_hist"])
        dist.all_reduce(out["q_hist"])
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    st = engine.workspace_status(ws, B, n_waters, n_waters, engine.default_r_cell(True, True, 3.413, 10.0), box)
    n_angles = int(out["ang_hist"].sum().item())

    # ---- FP32 arithmetic mode, same frames, device-resident (north_star: both modes reported) ---------
    out32 = {"q": torch.zeros((B, n_waters), dtype=torch.float32, device=dev),
             "n3": torch.zeros((B, n_waters), dtype=torch.int32, device=dev),
             "ang_hist": torch.zeros((1, 500), dtype=torch.int64, device=dev),
             "q_hist": torch.zeros((1, 500), dtype=torch.int64, device=dev),
             "frame_stats": torch.zeros((B, 8), dtype=torch.float64, device=dev)}
    ws32 = engine.Workspace(dev)

    def step32():
        return engine.q3b_frames(pos_d, box, out=out32, want=want, workspace=ws32, device=dev, check_status=False,
                                 precision="fp32")

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
    hist_l1 = float((out32["ang_hist"] - out["ang_hist"]).abs().sum().item()) / max(1.0, float(out["ang_hist"].sum().item()))
    del out32, ws32

    # ---- end-to-end timed region (host buffers in, host results out) -------------------------------
    pipe = FramePipeline(n_waters, max(1, min(B, args.e2e_batch)), dtype=np.float64, device=dev)
    q_h = torch.empty((B, n_waters), dtype=torch.float64, pin_memory=True)
    n3_h = torch.empty((B, n_waters), dtype=torch.int32, pin_memory=True)
    for _ in range(2):
        pipe.run(pos_h, box, out_q=q_h, out_n3=n3_h)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_launches = 0
    for _ in range(args.steps):
        r = pipe.run(pos_h, box, out_q=q_h, out_n3=n3_h)
        e2e_launches += pipe.launches
    if world > 1:
        dist.all_reduce(r["_device"]["ang_hist"])
    f1.record()
    barrier()
    e2e_ms_total = f0.elapsed_time(f1)

    # the same leg fed float32 host frames -- the values an AMBER NetCDF trajectory stores (amber_io.NetCDFTrajectory
    # hands them on without upcasting); arithmetic stays fp64, results are identical; half the host->device bytes
    pipe32 = FramePipeline(n_waters, max(1, min(B, args.e2e_batch)), dtype=np.float32, device=dev)
    pos_h32 = torch.empty((B, n_waters, 3), dtype=torch.float32, pin_memory=True)
    pos_h32.copy_(pos_h)
    q_h32 = torch.empty((B, n_waters), dtype=torch.float64, pin_memory=True)
    for _ in range(2):
        pipe32.run(pos_h32, box, out_q=q_h32, out_n3=n3_h)
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        r32 = pipe32.run(pos_h32, box, out_q=q_h32, out_n3=n3_h)
        e2e_launches += pipe32.launches
    if world > 1:
        dist.all_reduce(r32["_device"]["ang_hist"])
    g1.record()
    barrier()
    e2e32_ms_total = g0.elapsed_time(g1)
    sampler.stop_flag = True
    sampler.join(timeout=2.0)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_total = max_over_ranks(ms_total)
    ms32 = max_over_ranks(ms32)
    e2e_ms_total = max_over_ranks(e2e_ms_total)
    e2e32_ms_total = max_over_ranks(e2e32_ms_total)
    kernel_ms = max_over_ranks(kernel_ms)
    wf_per_step = float(world) * B * n_waters
    value = wf_per_step * args.steps / (ms_total * 1e-3)
    e2e_value = wf_per_step * args.steps / (e2e_ms_total * 1e-3)
    peak, peak_src = measured_peak()
    achieved = BYTES_PER_WF_FP64 * B * n_waters / (kernel_ms * 1e-3) / 1e9
    q_ok = bool(torch.equal(q_h.to(dev), out["q"]))
    q32_ok = bool(torch.equal(q_h32.to(dev), out["q"]))

    # FP roofline of the dominant kernel: 17 flop per candidate pair evaluation + 37 per angle, candidates
    # = 27 cells * cell volume * number density, angles = three-body angles (measured) + 6 for q
    nc_used = step()["nc"]
    cand = 27.0 * n_waters / float(nc_used[0] * nc_used[1] * nc_used[2])
    ang_per_wf = n_angles / (float(B) * n_waters * args.steps) + 6.0
    flop_per_wf = 17.0 * cand + 37.0 * ang_per_wf
    fp64_peak = 148 * 64 * 2 * 1.965e9 / 1e12  # nominal FP64 FMA peak of a B200 at 1965 MHz, TFLOP/s
    fp_achieved = flop_per_wf * B * n_waters / (kernel_ms * 1e-3) / 1e12

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config,
            "frames_per_s": value / n_waters,
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms_total / args.steps,
                    "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
                    "api": "waterorderlib_b200.pipeline.FramePipeline.run", "matches_device_run": q_ok},
            "e2e_f32_host_frames": {"value": wf_per_step * args.steps / (e2e32_ms_total * 1e-3), "unit": UNIT,
                                    "ms_per_step": e2e32_ms_total / args.steps, "h2d_bytes_per_step": pipe32.h2d_bytes,
                                    "d2h_bytes_per_step": pipe32.d2h_bytes, "matches_device_run": q32_ok,
                                    "note": "float32 host frames as a NetCDF trajectory stores them; fp64 arithmetic, identical results"},
            "gpu_launches": int(launches_per_step * args.steps + e2e_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": kernel_traffic(n_waters, B), "kernel": "wol::q3b_tpc_kernel",
                         "kernel_ms": kernel_ms, "bytes_per_water_frame": BYTES_PER_WF_FP64, "peak_source": peak_src,
                         "note": "the sweep is issue/FP64-bound by construction (about 1.3 kFLOP per water-frame, SURVEY 8d); see roofline_fp"},
            "roofline_fp": {"bound": "fp64 pipe (nominal, 148 SM x 64 FMA lanes x 2 x 1.965 GHz)", "flop_per_water_frame": flop_per_wf,
                            "candidates_per_water": cand, "angles_per_water": ang_per_wf, "achieved": fp_achieved,
                            "peak": fp64_peak, "unit": "TFLOP/s", "frac": fp_achieved / fp64_peak},
            "clocks": sampler.summary(),
            "fp32_mode": {"value": float(world) * B * n_waters * args.steps / (ms32 * 1e-3), "unit": UNIT,
                          "ms_per_step": ms32 / args.steps, "max_abs_q_error_vs_fp64_same_neighbours": q_err32,
                          "fraction_with_different_4nn": flips32,
                          "angle_hist_L1_distance_vs_fp64": hist_l1, "tolerance": 1e-4},
            "checks": {"angles_binned": n_angles, "widened": st[0], "overflow": st[1]}}
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
