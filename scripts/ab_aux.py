"""A/B timing of the auxiliary kernels between two builds of libwol.so (development aid).

    python scripts/ab_aux.py [path/to/other/libwol.so]     # default: the in-tree library

Prints the routine time (CUDA events around the whole call: cell build + kernel) of H-bond counts, LSI, the CSR
neighbour list and the shell selection on a 1M-water frame, in lattice order and with the atoms shuffled.
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from waterorderlib_b200 import _capi

if len(sys.argv) > 1:
    _capi.LIB_PATH = sys.argv[1]
from waterorderlib_b200 import routines, synth  # noqa: E402


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


dev = torch.device("cuda")
O, box = synth.water_box(50, sigma=0.25, seed=1)
H = synth.add_hydrogens(O, seed=1)
for order in ("lattice", "shuffled"):
    if order == "shuffled":
        perm = np.random.default_rng(5).permutation(O.shape[0])
        O, H = O[perm], H.reshape(-1, 2, 3)[perm].reshape(-1, 3)
    O_d, H_d = torch.from_numpy(O).to(dev)[None], torch.from_numpy(H).to(dev)[None]
    D_d = O_d.repeat_interleave(2, dim=1).contiguous()
    ms, r = timeit(lambda: routines.hbond_counts(O_d, D_d, H_d, box, 3.5, 120.0))
    print("%-9s hbond_counts 1M waters: %.3f ms (%d bonds)" % (order, ms, int(r["acc_count"].sum().item())))
    ms, r = timeit(lambda: routines.lsi(None, O_d, box))
    print("%-9s lsi 1M waters: %.3f ms" % (order, ms))
    ms, r = timeit(lambda: routines.neighbors_csr(None, O_d[0], box, 0.0, 3.5))
    print("%-9s neighbors_csr 1M waters, 3.5 A: %.3f ms (%d pairs)" % (order, ms, int(r[1].numel())))
    ms, r = timeit(lambda: routines.three_body_angles(None, O_d, box))
    print("%-9s getCosAngs values 1M waters: %.3f ms (%d angles)" % (order, ms, int(r[0].numel())))
    ms, r = timeit(lambda: routines.pair_hist(1, O_d[0], None, box, 0.1, 150), reps=2)
    print("%-9s radialdistsame 1M waters, 150 bins of 0.1 A: %.3f ms (%d pairs)" % (order, ms, int(r.sum().item())))
    sol = O_d[0, :4096].contiguous()
    ms, r = timeit(lambda: routines.shell_mask(sol, O_d[0], box, 4.0))
    print("%-9s shell of 4096 solute atoms in 1M waters: %.3f ms" % (order, ms))

sp, sbox, z_lo, z_hi = synth.slab_box(32, 32, 8, sigma=0.3, seed=11)
gp, gn = synth.plane_interface(sbox, z_lo, z_hi, spacing=2.0)
sp_d, gp_d, gn_d = torch.from_numpy(sp).to(dev), torch.from_numpy(gp).to(dev), torch.from_numpy(gn).to(dev)
for surf in (False, True):
    ms, r = timeit(lambda: routines.interface_water(sp_d, gp_d, gn_d, 0.0, sbox, want_surfclose=surf))
    print("interface_water %d waters x %d surface points, surfclose %s: %.3f ms" % (sp.shape[0], gp.shape[0], surf, ms))
