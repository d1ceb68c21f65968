"""Marginal cost of the phases of the fused sweep: both observables, three-body only, q only (library timing events).
usage: phase_times.py [cells=50] [frames=8] [sigma=0.25]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from waterorderlib_b200 import engine, synth

m = int(sys.argv[1]) if len(sys.argv) > 1 else 50
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sigma = float(sys.argv[3]) if len(sys.argv) > 3 else 0.25
pos = np.stack([synth.water_box(m, sigma=sigma, seed=s)[0] for s in range(frames)])
box = synth.water_box(m, sigma=0.0, seed=0)[1]
pos_d = torch.from_numpy(pos).cuda()
ws = engine.Workspace(torch.device("cuda"))
k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
k0.record(); k1.record()
for name, kw, want in (("q + three-body", {}, ("q", "n3", "ang_hist", "q_hist", "frame_stats")),
                       ("three-body only", {"do_q": False}, ("n3", "ang_hist", "frame_stats")),
                       ("q only", {"do_3body": False}, ("q", "q_hist", "frame_stats")),
                       ("q only, r_cell 3.8 explicit", {"do_3body": False, "r_cell": 3.8}, ("q", "q_hist", "frame_stats"))):
    ker = []
    for it in range(6):
        torch.cuda.synchronize()
        engine.q3b_frames(pos_d, box, workspace=ws, check_status=False, timing_events=(k0, k1), want=want, **kw)
        torch.cuda.synchronize()
        ker.append(k0.elapsed_time(k1))
    print("%-30s main kernel %.3f ms per %d frames" % (name, min(ker[1:]), frames))
