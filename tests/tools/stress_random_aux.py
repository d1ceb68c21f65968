"""More seeds of tests/test_gpu_random_aux.py (auxiliary routines against the oracle), plus the CSR neighbour list,
water orientation and iso-surface vertices on the same random systems.   usage: stress_random_aux.py [lo=12] [hi=212]"""
import sys

import numpy as np

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from test_gpu_random_aux import case, test_random_aux  # noqa: E402

from oracle import port  # noqa: E402
from waterorderlib_b200 import routines, synth  # noqa: E402

lo = int(sys.argv[1]) if len(sys.argv) > 1 else 12
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 212
bad = 0
for seed in range(lo, hi):
    try:
        test_random_aux(seed)
        rng, pos, box = case(seed)
        cut = float(rng.uniform(3.0, 0.45 * box.min()))
        off, idx = routines.neighbors_csr(None, pos, box, 0.0, cut)
        mat = port.neighbor_matrix(pos, pos, box, 0.0, cut)
        assert np.array_equal(off.cpu().numpy(), np.concatenate([[0], np.cumsum(mat.sum(axis=1))]))
        assert np.array_equal(idx.cpu().numpy(), np.nonzero(mat)[1])
        h = synth.add_hydrogens(pos, seed=seed)
        ref = rng.normal(size=3)
        d, p = routines.water_orient(pos, h, box, ref)
        rd, rp = port.watorient(pos, h, ref, box)
        assert np.allclose(d[0].cpu().numpy(), rd, rtol=0, atol=1e-11) and np.allclose(p[0].cpu().numpy(), rp, rtol=0, atol=1e-11)
        grid = [(np.arange(n) + 0.5) * (box[k] / n) for k, n in enumerate((9, 11, 8))]
        dens, _ = routines.willard_density(pos, box, 2.4, grid=grid, want_normals=False)
        lvl = float(np.quantile(dens.cpu().numpy(), 0.5))
        assert np.array_equal(routines.iso_points(dens, grid, lvl).cpu().numpy(), port.iso_points(dens.cpu().numpy(), *grid, lvl))
    except AssertionError as e:
        bad += 1
        print("MISMATCH seed", seed, repr(e)[:200])
print("aux stress: seeds %d..%d, %d mismatches" % (lo, hi - 1, bad))
