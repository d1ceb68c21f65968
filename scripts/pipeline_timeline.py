"""Per-batch timeline of the host-fed pipeline (copy-in / kernels / copy-out), from CUDA events.

    python scripts/pipeline_timeline.py [--cells 50] [--frames 16] [--batch 1] [--dtype f32|f64]
"""
import argparse
import time

import numpy as np
import torch

from waterorderlib_b200 import synth
from waterorderlib_b200.pipeline import FramePipeline

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=50)
ap.add_argument("--frames", type=int, default=16)
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--dtype", default="f64")
ap.add_argument("--slots", type=int, default=4)
ap.add_argument("--run-streams", type=int, default=2)
a = ap.parse_args()
dt = np.float64 if a.dtype == "f64" else np.float32
p, box = synth.water_box(a.cells, sigma=0.25, seed=0)
n = p.shape[0]
pos_h = torch.empty((a.frames, n, 3), dtype=torch.float64 if dt == np.float64 else torch.float32, pin_memory=True)
for f in range(a.frames):
    pos_h[f].copy_(torch.from_numpy(p))
pipe = FramePipeline(n, a.batch, dtype=dt, n_slots=a.slots, n_run_streams=a.run_streams)
q_h = torch.empty((a.frames, n), dtype=torch.float64, pin_memory=True)
n3_h = torch.empty((a.frames, n), dtype=torch.int32, pin_memory=True)
for _ in range(3):
    pipe.run(pos_h, box, out_q=q_h, out_n3=n3_h)
torch.cuda.synchronize()
pipe.trace = True
t0 = time.perf_counter()
pipe.run(pos_h, box, out_q=q_h, out_n3=n3_h)
host_ms = (time.perf_counter() - t0) * 1e3
tl = pipe.timeline()
print("host time to enqueue the run: %.2f ms; last copy-out ends at %.2f ms" % (host_ms, tl[:, 5].max()))
print("batch   in0    in1 |  run0   run1 |  out0   out1   (ms)   in  run  out")
for i, r in enumerate(tl):
    print("%4d %6.2f %6.2f | %6.2f %6.2f | %6.2f %6.2f        %4.2f %4.2f %4.2f" % ((i,) + tuple(r) + (r[1] - r[0], r[3] - r[2], r[5] - r[4])))
