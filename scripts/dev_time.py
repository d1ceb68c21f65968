"""Development timing of the fused path: whole call and the dominant kernel alone (library timing events).
usage: dev_time.py [cells=50] [frames=8] [sigma=0.25] [reps=5]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from waterorderlib_b200 import engine, synth

m = int(sys.argv[1]) if len(sys.argv) > 1 else 50
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sigma = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
want = tuple(sys.argv[5].split(",")) if len(sys.argv) > 5 else ("q", "nn_idx", "n3", "ang_hist", "q_hist", "frame_stats")
import os
r_cell = float(os.environ["WOL_RCELL"]) if "WOL_RCELL" in os.environ else None
pos = np.stack([synth.water_box(m, sigma=sigma, seed=s)[0] for s in range(frames)])
box = synth.water_box(m, sigma=0.0, seed=0)[1]
pos_d = torch.from_numpy(pos).cuda()
ws = engine.Workspace(torch.device("cuda"))
k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
k0.record(); k1.record()
tot, ker = [], []
for it in range(reps + 2):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = engine.q3b_frames(pos_d, box, workspace=ws, check_status=False, timing_events=(k0, k1), want=want, r_cell=r_cell)
    e1.record()
    torch.cuda.synchronize()
    if it >= 2:
        tot.append(e0.elapsed_time(e1)); ker.append(k0.elapsed_time(k1))
r = engine.q3b_frames(pos_d, box, workspace=ws, r_cell=r_cell)
n = pos.shape[1] * frames
print(("N=%d F=%d sigma=%.2f want=" + ",".join(want) + ": call %.3f ms (min %.3f)  main kernel %.3f ms  rest %.3f ms  %.3e wf/s  widened=%d overflow=%d <q>=%.6f angles=%d nc=%s")
      % (pos.shape[1], frames, sigma, np.mean(tot), np.min(tot), np.mean(ker), np.mean(tot) - np.mean(ker), n / np.mean(tot) * 1e3,
         r.n_widened, r.n_overflow, float(r.q.mean()), int(r.ang_hist.sum()), r.nc))
