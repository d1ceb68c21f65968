"""Typical use: the reference's analysis script with the backend swapped in.

    python examples/analyse_trajectory.py [system.parm7 trajectory.nc]

Without arguments a small synthetic water box with a cosolvent is generated in memory.  Needs a CUDA device.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from waterorderlib_b200 import synth  # noqa: E402
from waterorderlib_b200.structureLibs import orderParam_lib as opl  # noqa: E402  (was: import orderParam_lib as opl)
from waterorderlib_b200.structureLibs import water_properties as wp  # noqa: E402  (was: import water_properties as wp)
from waterorderlib_b200.structureLibs.TrajObject import ArrayTrajectory, Topology, TrajObject  # noqa: E402

if len(sys.argv) == 3:
    topFile, trajFile = sys.argv[1], sys.argv[2]
else:
    m, T = 6, 40
    n_w = 8 * m ** 3
    topFile = Topology.water_box(n_w)
    xyz = np.zeros((T, 3 * n_w, 3))
    boxes = np.zeros((T, 3))
    for t in range(T):
        o, box = synth.water_box(m, sigma=0.4, seed=t)
        h = synth.add_hydrogens(o, seed=t)
        xyz[t, 0::3], xyz[t, 1::3], xyz[t, 2::3], boxes[t] = o, h[0::2], h[1::2], box
    trajFile = ArrayTrajectory(xyz, boxes, top=topFile)

# per-frame API, exactly the reference's call sites (orderParam_lib.py:1325, :1471)
obj = TrajObject(topFile, trajFile)
watInds, watHInds, lenWat = obj.getWatInds()
frame = obj.traj[0]
watPos, thisbox = np.array(frame.xyz)[watInds], np.array(frame.box.values[:3])
q = wp.getOrderParamq(watPos, watPos, thisbox)
angles, numbers = wp.getCosAngs(watPos, watPos, thisbox)
angDist, bins, pTet, avgCos, varCos, entropy = wp.tetrahedralMetrics(angles)
print("frame 0: %d waters, <q> = %.4f, %d three-body angles, tetrahedral fraction %.3f" % (len(q), q.mean(), len(angles), pTet))

# trajectory drivers (write qDistribution_0.txt, 3bDistribution_0.txt, hbDistribution_water.txt like the reference)
avgQ, varQ = opl.tetOrderCalc(topFile, trajFile)
pTet, avgCos, varCos, entropy, nWats = opl.threeBodyCalc(topFile, trajFile)
print("trajectory: <q> = %.4f +- %.4f, tetrahedral fraction = %.4f +- %.4f" % (avgQ[0][0], avgQ[1][0], pTet[0][0], pTet[1][0]))
