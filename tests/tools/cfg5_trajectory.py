"""BASELINE config 5 end to end: a 1,000,000-water, 1,000-frame synthetic trajectory sharded by frame over the GPUs of
one box, q + three-body histograms accumulated on each device, ONE NCCL all-reduce of the int64 histograms and one
all-gather of the per-frame rows at the end.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
        tests/tools/cfg5_trajectory.py [--frames 1000] [--cells 50]

Frames are jittered-ice boxes generated ON THE DEVICE from per-frame seeds (torch.Generator, Box-Muller in torch), so
the result does not depend on the number of ranks; generation is outside the timed region (a batch is generated, then
analysed).  Frame 0 is checked against the CPU oracle (tests/tools because it runs the checker)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from waterorderlib_b200 import distributed as wdist  # noqa: E402
from waterorderlib_b200 import engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=1000)
ap.add_argument("--cells", type=int, default=50)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--sigma", type=float, default=0.25)
ap.add_argument("--check", type=int, default=1)
args = ap.parse_args()

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

lattice, box = synth.diamond_lattice(args.cells)
box = box.astype(np.float32).astype(np.float64)
lat_d = torch.from_numpy(lattice).to(dev)
box_d = torch.from_numpy(box).to(dev)
N = lattice.shape[0]


def make_frames(f0, f1):
    """(f1 - f0, N, 3) float32-representable float64 positions; frame f depends only on its absolute index."""
    out = torch.empty((f1 - f0, N, 3), dtype=torch.float64, device=dev)
    for k, f in enumerate(range(f0, f1)):
        g = torch.Generator(device=dev)
        g.manual_seed(1_000_003 * f + 17)
        p = lat_d + args.sigma * torch.randn((N, 3), generator=g, device=dev, dtype=torch.float64)
        p = p - box_d * torch.floor(p / box_d)
        out[k] = p.to(torch.float32).to(torch.float64)
    return out


begin, end = wdist.shard_frames(args.frames)
ang_hist = torch.zeros((1, 500), dtype=torch.int64, device=dev)
q_hist = torch.zeros((1, 500), dtype=torch.int64, device=dev)
stats = torch.zeros((end - begin, 8), dtype=torch.float64, device=dev)
ws = engine.Workspace(dev)
q = torch.zeros((args.batch, N), dtype=torch.float64, device=dev)
n3 = torch.zeros((args.batch, N), dtype=torch.int32, device=dev)
t_gpu = 0.0
first = None
if end > begin:  # untimed warm-up: workspace allocation, module load, bin table
    wpos = make_frames(begin, min(end, begin + args.batch))
    for _ in range(3):
        engine.q3b_frames(wpos, box, workspace=ws, device=dev, check_status=False, want=("q", "n3"))
    del wpos
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t_wall = time.perf_counter()
for b0 in range(begin, end, args.batch):
    b1 = min(end, b0 + args.batch)
    pos = make_frames(b0, b1)
    if first is None and rank == 0:
        first = pos[0].cpu().numpy()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = {"q": q[: b1 - b0], "n3": n3[: b1 - b0], "ang_hist": ang_hist, "q_hist": q_hist, "frame_stats": stats[b0 - begin:b1 - begin]}
    engine.q3b_frames(pos, box, out=out, want=tuple(out.keys()), workspace=ws, device=dev, check_status=False)
    e1.record()
    torch.cuda.synchronize()
    t_gpu += e0.elapsed_time(e1) * 1e-3
st = engine.workspace_status(ws, min(args.batch, end - begin), N, N, engine.default_r_cell(True, True, 3.413, 10.0), box) if end > begin else (0, 0)
c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
c0.record()
wdist.reduce_histograms(ang_hist, q_hist)
rows = wdist.gather_frame_rows(stats, args.frames)
c1.record()
torch.cuda.synchronize()
t_coll = c0.elapsed_time(c1) * 1e-3
t_wall = time.perf_counter() - t_wall
t_gpu_max = wdist.max_over_ranks(t_gpu + t_coll, dev)

if rank == 0:
    res = {"config": 5, "n_gpus": world, "frames": args.frames, "waters_per_frame": N, "analysis_seconds_max_over_ranks": t_gpu_max,
           "collective_seconds": t_coll, "wall_seconds_incl_generation": t_wall,
           "water_frames_per_s": args.frames * N / t_gpu_max, "angles_binned": int(ang_hist.sum().item()),
           "q_binned": int(q_hist.sum().item()), "mean_q": float((rows[:, 0].sum() / rows[:, 2].sum()).item()),
           "frames_in_rows": int(rows.shape[0]), "overflow": st[1]}
    if args.check:
        from oracle import port  # the checker
        tb = port.three_body(first, first, box, materialize=False)
        qr, _, _ = port.order_param_q(first, first, box)
        chk = engine.q3b_frames(first, box, device=dev)
        res["frame0_parity"] = bool(np.array_equal(chk.ang_hist.cpu().numpy()[0], tb["hist"]) and np.array_equal(chk.n3.cpu().numpy()[0], tb["numAngs"])
                                    and np.allclose(chk.q.cpu().numpy()[0], qr, rtol=1e-6, atol=1e-9))
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
