"""Randomised parity sweep of the auxiliary routines (H-bond counts, shell mask, LSI, psi, pair-distance histograms,
materialised angles, dense matrices) against the CPU oracle: random orthorhombic boxes, densities and cutoffs,
unwrapped coordinates."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import port  # noqa: E402  (the checker)
from waterorderlib_b200 import routines, synth  # noqa: E402
from waterorderlib_b200.structureLibs import water_properties as wp  # noqa: E402
from waterorderlib_b200.structureLibs import waterlib as wl  # noqa: E402


def case(seed):
    rng = np.random.default_rng(1000 + seed)
    box = rng.uniform(12.0, 40.0, size=3)
    n = int(np.clip(10 ** rng.uniform(-2.2, -1.35) * box.prod(), 30, 2500))
    pos = rng.random((n, 3)) * box
    if seed % 2:
        pos = pos + box * rng.integers(-3, 4, size=(n, 3))
    return rng, pos.astype(np.float32).astype(np.float64), box


@pytest.mark.parametrize("seed", range(12))
def test_random_aux(seed):
    rng, pos, box = case(seed)
    n = pos.shape[0]
    # hydrogen bonds with random hydrogens around each heavy atom
    hpos = synth.add_hydrogens(pos, seed=seed)
    don = np.repeat(pos, 2, axis=0)
    dc, ac = float(rng.uniform(2.8, 3.6)), float(rng.uniform(100.0, 160.0))
    r = routines.hbond_counts(pos, don, hpos, box, dc, ac)
    a_ref, d_ref = port.hbonds(pos, don, hpos, box, dc, ac)
    assert np.array_equal(r["acc_count"].cpu().numpy()[0], a_ref) and np.array_equal(r["don_count"].cpu().numpy()[0], d_ref)
    # shell
    sol = (rng.random((20, 3)) * box).astype(np.float32).astype(np.float64)
    cut = float(rng.uniform(3.0, 5.5))
    assert np.array_equal(routines.shell_mask(sol, pos, box, cut).cpu().numpy()[0], port.shell_mask(sol, pos, box, cut))
    # LSI and psi on a subset of centres
    sub = pos[rng.choice(n, min(n, 60), replace=False)]
    v, num = wp.getLSI(sub, pos, box, 0.0, 3.7)
    v_ref, num_ref = port.getLSI(sub, pos, box, 0.0, 3.7)
    assert np.array_equal(num, num_ref) and np.allclose(v, v_ref, rtol=1e-10, atol=1e-16)
    hi = float(rng.uniform(3.5, min(6.0, 0.45 * box.min())))
    assert np.allclose(wp.getOrderParamPsi(sub, pos, box, 0.0, hi), port.getOrderParamPsi(sub, pos, box, 0.0, hi), rtol=1e-9, atol=1e-13)
    # pair-distance histograms
    bw, nb = float(rng.uniform(0.05, 0.3)), int(rng.integers(20, 80))
    assert np.array_equal(wl.radialdistsame(pos, bw, nb, 1.0, box), port.radialdistsame(pos, bw, nb, 1.0, box))
    assert np.array_equal(wl.radialdist(sol, pos, bw, nb, 0.5, box), port.radialdist(sol, pos, bw, nb, 0.5, box))
    # materialised angles (order and values) and dense matrices
    ang, cnt = wp.getCosAngs(sub, pos, box, 0.0, 3.4)
    ang_ref, cnt_ref = port.getCosAngs(sub, pos, box, 0.0, 3.4)
    assert np.array_equal(cnt, cnt_ref) and ang.shape == ang_ref.shape and np.allclose(ang, ang_ref, rtol=1e-12, atol=1e-10)
    assert np.array_equal(wl.nearneighbors(sub, pos, box, 0.5, cut), port.neighbor_matrix(sub, pos, box, 0.5, cut))
