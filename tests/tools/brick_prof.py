"""A few launches of the dominant fp64 kernel on 1M-water frames, for ncu:  python tests/tools/brick_prof.py [frames] [sigma]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from waterorderlib_b200 import engine, synth  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 4
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 0.25
dev = torch.device("cuda", 0)
frames = []
for f in range(F):
    p, box = synth.water_box(50, sigma=sigma, seed=f)
    frames.append(p)
pos = torch.from_numpy(np.stack(frames)).to(dev)
ws = engine.Workspace(dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
for e in ev:
    e.record()
for it in range(4):
    r = engine.q3b_frames(pos, box, workspace=ws, timing_events=ev, check_status=(it == 3))
    torch.cuda.synchronize()
    print("kernel ms", ev[0].elapsed_time(ev[1]), flush=True)
print("widened", r["n_widened"], "overflow", r["n_overflow"], "slow pairs", r["n_slow_pairs"], "angles", int(r["ang_hist"].sum()))
