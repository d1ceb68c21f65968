"""Timings of the auxiliary kernels at production sizes (development aid; prints one line per kernel)."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from waterorderlib_b200 import routines, synth


def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


dev = torch.device("cuda")
O, box = synth.water_box(50, sigma=0.25, seed=1)
H = synth.add_hydrogens(O, seed=1)
O_d, H_d = torch.from_numpy(O).to(dev)[None], torch.from_numpy(H).to(dev)[None]
D_d = O_d.repeat_interleave(2, dim=1).contiguous()
ms, r = timeit(lambda: routines.hbond_counts(O_d, D_d, H_d, box, 3.5, 120.0))
print("hbond_counts 1M waters (1M acceptors x 2M donors, 3.5 A / 120 deg): %.3f ms, %.3f bonds per water" % (ms, r["acc_count"].sum().item() / 1e6))
pos, sbox, z_lo, z_hi = synth.slab_box(32, 32, 8, sigma=0.3, seed=11)
g = [np.linspace(0, sbox[k], 80, endpoint=False) for k in range(3)]
pos_d = torch.from_numpy(pos).to(dev)
ms, r = timeit(lambda: routines.willard_density(pos_d, sbox, 2.4, grid=g))
print("willard_density 65536 waters, 80^3 grid, sigma 2.4: %.3f ms (max density %.4f)" % (ms, r[0].max().item()))
ms, r = timeit(lambda: routines.lsi(None, O_d, box))
print("lsi 1M waters: %.3f ms" % ms)
ms, r = timeit(lambda: routines.pair_hist(1, O_d[0], None, box, 0.1, 150))
print("radialdistsame 1M waters, 150 bins of 0.1 A: %.3f ms (%d pairs)" % (ms, int(r.sum().item())))
O2, box2 = synth.water_box(16, sigma=0.25, seed=1)
ms, r = timeit(lambda: routines.psi(None, torch.from_numpy(O2).to(dev), box2, 0.0, 7.0))
print("psi 32768 atoms, cutoff 7 A: %.3f ms" % ms)
