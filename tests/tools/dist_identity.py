"""Multi-GPU identity check, run under torchrun (one rank per GPU, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 \
        tests/tools/dist_identity.py [--cells 16] [--frames 12]

A trajectory of `frames` jittered-ice frames (generated on the device from absolute frame indices) is sharded by frame;
each rank accumulates int64 angle / q histograms and per-frame rows on its device; ONE all-reduce and ONE all-gather
combine them.  Rank 0 then analyses the whole trajectory alone and checks that the combined histograms are
bit-identical and the gathered rows equal, and checks frame 0 against the CPU oracle (tests/tools: it runs the
checker).  Prints one JSON line; exit code 1 on any mismatch.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from waterorderlib_b200 import distributed as wdist  # noqa: E402
from waterorderlib_b200 import engine, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=16)
ap.add_argument("--frames", type=int, default=12)
ap.add_argument("--sigma", type=float, default=0.3)
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)


def analyse(f0, f1):
    ang = torch.zeros((1, 500), dtype=torch.int64, device=dev)
    qh = torch.zeros((1, 500), dtype=torch.int64, device=dev)
    rows = torch.zeros((f1 - f0, 8), dtype=torch.float64, device=dev)
    if f1 > f0:
        pos, box = synth.device_frames(args.cells, f0, f1, sigma=args.sigma, device=dev)
        engine.q3b_frames(pos, box, out={"ang_hist": ang, "q_hist": qh, "frame_stats": rows}, want=("ang_hist", "q_hist", "frame_stats"),
                          device=dev)
    return ang, qh, rows


begin, end = wdist.shard_frames(args.frames)
ang, qh, rows = analyse(begin, end)
wdist.reduce_histograms(ang, qh)
all_rows = wdist.gather_frame_rows(rows, args.frames)
ok = True
res = {"n_gpus": world, "frames": args.frames, "waters": 8 * args.cells ** 3}
if rank == 0:
    ang1, qh1, rows1 = analyse(0, args.frames)
    res["hist_equals_single_rank"] = bool(torch.equal(ang, ang1) and torch.equal(qh, qh1))
    res["rows_equal_single_rank"] = bool(torch.allclose(all_rows, rows1, rtol=1e-13, atol=0.0)) and all_rows.shape == rows1.shape
    from oracle import port  # the checker
    pos0, box = synth.device_frames(args.cells, 0, 1, sigma=args.sigma, device=dev)
    p = pos0[0].cpu().numpy()
    tb = port.three_body(p, p, box, materialize=False)
    chk = engine.q3b_frames(pos0, box, device=dev)
    q_ref, nn4, _ = port.order_param_q(p, p, box)
    res["frame0_parity"] = bool(np.array_equal(chk.ang_hist.cpu().numpy()[0], tb["hist"]) and np.array_equal(chk.n3.cpu().numpy()[0], tb["numAngs"])
                                and np.array_equal(chk.nn_idx.cpu().numpy()[0], nn4) and np.allclose(chk.q.cpu().numpy()[0], q_ref, rtol=1e-6, atol=1e-9))
    res["angles_binned"] = int(ang.sum().item())
    ok = res["hist_equals_single_rank"] and res["rows_equal_single_rank"] and res["frame0_parity"]
    print(json.dumps(res), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
