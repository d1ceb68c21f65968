"""In-process A/B of the frame drivers' host staging: tetOrderCalc on an in-memory numpy trajectory with 1 and 4 staging
threads, alternating, best of several repetitions each (box-to-box and run-to-run noise of the host is larger than the
effect).  usage: driver_ab.py [cells=50] [frames=16]"""
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
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 16
n_w = 8 * m ** 3
o, box = synth.water_box(m, sigma=0.25, seed=0)
h = synth.add_hydrogens(o, seed=0)
one = np.empty((3 * n_w, 3), dtype=np.float32)
one[0::3], one[1::3], one[2::3] = o, h[0::2], h[1::2]
xyz = np.broadcast_to(one, (frames,) + one.shape).copy()
top = Topology.water_box(n_w)
os.chdir(tempfile.mkdtemp())
res = {}
for where in ("numpy", "pinned"):
    src = xyz if where == "numpy" else torch.from_numpy(xyz).pin_memory()
    traj = ArrayTrajectory(src, np.tile(box, (frames, 1)), top=top)
    for rep in range(4):
        for threads in ((1, 4) if where == "numpy" else (4,)):
            opl._COPY_THREADS, opl._COPY_POOL = threads, None
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            opl.tetOrderCalc(top, traj)
            torch.cuda.synchronize()
            ms = (time.perf_counter() - t0) / frames * 1e3
            res.setdefault((where, threads), []).append(ms)
for k, v in res.items():
    print("%s, %d staging threads: best %.2f ms per frame (all: %s)" % (k[0], k[1], min(v[1:]), " ".join("%.2f" % x for x in v)))
