"""Wall-clock throughput of the reference-facing frame drivers (tetOrderCalc, threeBodyCalc) on a large in-memory
trajectory with the reference's atom layout (O, H, H per water): what a user of orderParam_lib sees end to end.
usage: driver_time.py [cells=50] [frames=8] [dtype=f32|f64]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from waterorderlib_b200 import synth
from waterorderlib_b200.structureLibs import orderParam_lib as opl
from waterorderlib_b200.structureLibs.TrajObject import ArrayTrajectory, Topology

m = int(sys.argv[1]) if len(sys.argv) > 1 else 50
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dt = np.float64 if (len(sys.argv) > 3 and sys.argv[3] == "f64") else np.float32
where = sys.argv[4] if len(sys.argv) > 4 else "numpy"   # numpy | pinned | cuda
n_w = 8 * m ** 3
o, box = synth.water_box(m, sigma=0.25, seed=0)
h = synth.add_hydrogens(o, seed=0)
one = np.empty((3 * n_w, 3), dtype=dt)
one[0::3], one[1::3], one[2::3] = o, h[0::2], h[1::2]
xyz = np.broadcast_to(one, (frames,) + one.shape).copy()
top = Topology.water_box(n_w)
if where == "pinned":
    xyz = torch.from_numpy(xyz).pin_memory()
elif where == "cuda":
    xyz = torch.from_numpy(xyz).cuda()
traj = ArrayTrajectory(xyz, np.tile(box, (frames, 1)), top=top)
os.chdir(tempfile.mkdtemp())
for name, fn in (("tetOrderCalc", opl.tetOrderCalc), ("threeBodyCalc", opl.threeBodyCalc)):
    fn(top, traj)  # warm-up (allocations, first-use initialisation)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    r = fn(top, traj)
    torch.cuda.synchronize()
    dtm = time.perf_counter() - t0
    print("%-14s %d waters x %d frames (%s, %s): %.1f ms per frame, %.3g water-frames/s; first value %.6f"
          % (name, n_w, frames, np.dtype(dt).name, where, dtm / frames * 1e3, n_w * frames / dtm, r[0][0][0]))
if len(sys.argv) > 5 and sys.argv[5] == "hb":
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = opl.hbCalc(top, traj)
        torch.cuda.synchronize()
        dtm = time.perf_counter() - t0
    print("hbCalc         %d waters x %d frames (%s, %s): %.1f ms per frame, %.3g water-frames/s; H-bonds per water %.4f"
          % (n_w, frames, np.dtype(dt).name, where, dtm / frames * 1e3, n_w * frames / dtm, r[0]))
