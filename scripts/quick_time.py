"""Rough device-side timing of the fused path (development aid, not the bench)."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from waterorderlib_b200 import engine, synth

m = int(sys.argv[1]) if len(sys.argv) > 1 else 50
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sigma = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
prec = sys.argv[4] if len(sys.argv) > 4 else "fp64"
pos, box = synth.water_box(m, sigma=sigma, seed=0)
pos = np.stack([pos] * frames)
pos_d = torch.from_numpy(pos).cuda()
ws = engine.Workspace(torch.device("cuda"))
for dt in (torch.float64, torch.float32):
    p = pos_d.to(dt)
    for it in range(3):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = engine.q3b_frames(p, box, workspace=ws, precision=prec, check_status=False)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    r = engine.q3b_frames(p, box, workspace=ws, precision=prec)
    n = pos.shape[1] * frames
    print("N=%d frames=%d store=%s prec=%s: %.3f ms/call  %.3e water-frames/s  widened=%d overflow=%d <q>=%.5f angles=%d nc=%s"
          % (pos.shape[1], frames, dt, prec, ms, n / ms * 1e3, r.n_widened, r.n_overflow, float(r.q.double().mean()),
             int(r.ang_hist.sum()), r.nc))
