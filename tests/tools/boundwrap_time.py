"""Wall-clock cost of the per-frame population split (getBoundWrap via boundWrapPopulations) on a solvated cosolvent box.
usage: boundwrap_time.py [cells=16] [frames=6] [n_sol=40]"""
import os
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from waterorderlib_b200.structureLibs import orderParam_lib as opl  # noqa: E402

import test_gpu_drivers as helper  # noqa: E402  (make_system only; needs the oracle importable, not used)

m = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n_sol = int(sys.argv[3]) if len(sys.argv) > 3 else 40
top, traj = helper.make_system(m, T, n_sol=n_sol)
os.chdir(tempfile.mkdtemp())
opl.boundWrapPopulations(top, traj, cacheFile=None)
torch.cuda.synchronize()
t0 = time.perf_counter()
sub = opl.boundWrapPopulations(top, traj, cacheFile=None)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print("boundWrapPopulations: %d waters + %d cosolvent molecules, %d frames: %.1f ms per frame; shell %d bound %d"
      % (8 * m ** 3, n_sol, T, dt / T * 1e3, len(sub[0][2]), len(sub[0][0])))
import cProfile, pstats
cProfile.run("opl.boundWrapPopulations(top, traj, cacheFile=None)", "/tmp/bw.prof")
pstats.Stats("/tmp/bw.prof").sort_stats("cumtime").print_stats(14)
