"""Host time per engine.q3b_frames call (what a per-frame batch of the pipeline pays before any kernel runs).

    python scripts/host_overhead.py
"""
import cProfile
import pstats
import time

import numpy as np
import torch

from waterorderlib_b200 import engine, synth

p, box = synth.water_box(4, sigma=0.25, seed=0)
dev = torch.device("cuda", 0)
pos = torch.from_numpy(p[None]).to(dev)
ws = engine.Workspace(dev)
out = {"q": torch.zeros((1, p.shape[0]), dtype=torch.float64, device=dev), "n3": torch.zeros((1, p.shape[0]), dtype=torch.int32, device=dev),
       "ang_hist": torch.zeros((1, 500), dtype=torch.int64, device=dev), "q_hist": torch.zeros((1, 500), dtype=torch.int64, device=dev),
       "frame_stats": torch.zeros((1, 8), dtype=torch.float64, device=dev)}
box_d = torch.from_numpy(np.asarray(box)[None].copy()).to(dev)


def call():
    return engine.q3b_frames(pos, box, out=out, want=tuple(out), workspace=ws, device=dev, check_status=False, box_device=box_d)


for _ in range(20):
    call()
torch.cuda.synchronize()
n = 500
t0 = time.perf_counter()
for _ in range(n):
    call()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print("host time per call: %.1f us (enqueue only), %.1f us incl. final sync; launches per call %d" % ((t1 - t0) / n * 1e6, (t2 - t0) / n * 1e6, call()["launches"]))
pr = cProfile.Profile()
pr.enable()
for _ in range(200):
    call()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
