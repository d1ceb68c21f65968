"""Pins the C oracle on the LIVE reference (compiled Fortran via ctypes + AST-extracted Python bodies).
Only runs where /root/reference exists (the build container); the GPU box relies on tests/golden."""
import numpy as np
import pytest

from oracle import port, ref_fortran
from waterorderlib_b200 import synth

pytestmark = pytest.mark.live_reference


@pytest.fixture(scope="module")
def ref():
    from oracle import build_oracle
    build_oracle.build(verbose=False)
    wl = ref_fortran.RefWaterlib()
    return wl, ref_fortran.load_reference_functions(wl=wl)


@pytest.mark.parametrize("m,sigma,seed", [(2, 0.3, 1), (3, 0.6, 2), (4, 0.25, 1234), (5, 0.45, 9)])
def test_q_bit_exact(ref, m, sigma, seed):
    wl, fn = ref
    pos, box = synth.water_box(m, sigma=sigma, seed=seed)
    highq = min(10.0, 0.49 * box.min())
    assert np.array_equal(port.getOrderParamq(pos, pos, box, 0.0, highq), fn["getOrderParamq"](pos, pos, box, 0.0, highq))


@pytest.mark.parametrize("m,sigma,seed,high", [(2, 0.3, 1, 3.413), (4, 0.6, 2, 3.413), (4, 0.25, 3, 3.7), (6, 0.5, 4, 3.0)])
def test_three_body_bit_exact(ref, m, sigma, seed, high):
    wl, fn = ref
    pos, box = synth.water_box(m, sigma=sigma, seed=seed)
    a_ref, n_ref = fn["getCosAngs"](pos, pos, box, 0.0, high)
    a, n = port.getCosAngs(pos, pos, box, 0.0, high)
    assert np.array_equal(a, a_ref) and np.array_equal(n, n_ref)
    r = port.three_body(pos, pos, box, 0.0, high)
    assert np.array_equal(r["hist"], fn["tetrahedralMetrics"](a_ref)[0])


def test_neighbor_matrix_and_subpop(ref):
    wl, fn = ref
    rng = np.random.default_rng(5)
    box = np.array([13.0, 15.0, 11.5])
    pos = rng.random((300, 3)) * box * 3.0 - box  # unwrapped
    sub = rng.random((37, 3)) * box
    assert np.array_equal(port.neighbor_matrix(pos, pos, box, 0.0, 3.5), wl.allnearneighbors(pos, box, 0.0, 3.5))
    assert np.array_equal(port.neighbor_matrix(sub, pos, box, 1.0, 4.5), wl.nearneighbors(sub, pos, box, 1.0, 4.5))
    assert np.array_equal(port.getOrderParamq(sub, pos, box, 0.0, 5.0), fn["getOrderParamq"](sub, pos, box, 0.0, 5.0))
    a_ref, n_ref = fn["getCosAngs"](sub, pos, box, 0.5, 4.0)
    a, n = port.getCosAngs(sub, pos, box, 0.5, 4.0)
    assert np.array_equal(a, a_ref) and np.array_equal(n, n_ref)


def test_hbonds_and_shell(ref):
    wl, fn = ref
    opos, box = synth.water_box(3, sigma=0.35, seed=11)
    hpos = synth.add_hydrogens(opos, seed=11)
    don = np.repeat(opos, 2, axis=0)
    mat = wl.generalhbonds(opos, don, hpos, box, 3.5, 120.0)
    ac, dc, m2 = port.hbonds(opos, don, hpos, box, 3.5, 120.0, dense=True)
    assert np.array_equal(m2, mat)
    assert np.array_equal(ac, mat.sum(axis=1)) and np.array_equal(dc, mat.sum(axis=0))
    sol = synth.solute_grid(box, n_side=2)
    ref_mask = wl.nearneighbors(sol, opos, box, 0.0, 4.0).any(axis=0)
    assert np.array_equal(port.shell_mask(sol, opos, box, 4.0).astype(bool), ref_mask)


def test_angle_quirks(ref):
    wl, _ = ref
    assert wl.cosangle3([1, 0, 0], [0, 0, 0], [-1, 0, 0]) == -180.0
    assert wl.cosangle3([1, 0, 0], [0, 0, 0], [2, 0, 0]) == 0.0
    # collinear simple-cubic neighbours: the -180 angles drop out of np.histogram, in the oracle too
    g = np.arange(4) * 3.0
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    box = np.array([12.0, 12.0, 12.0])
    r = port.three_body(pos, pos, box, 0.0, 3.2)
    assert np.all(r["numAngs"] == 6) and r["n_angles"] == 64 * 15
    assert (r["angVals"] == -180.0).sum() == 64 * 3
    assert r["hist"].sum() == 64 * 12


def test_ref_driver_matches_live(ref):
    """oracle/ref_driver.py (used for the CPU baseline on the GPU box) == the live reference bodies."""
    from oracle import ref_driver
    wl, fn = ref
    pos, box = synth.water_box(4, sigma=0.5, seed=21)
    assert np.array_equal(ref_driver.get_order_param_q(wl, pos, pos, box), fn["getOrderParamq"](pos, pos, box))
    a, n = ref_driver.get_cos_angs(wl, pos, pos, box)
    a_ref, n_ref = fn["getCosAngs"](pos, pos, box)
    assert np.array_equal(a, a_ref) and np.array_equal(n, n_ref)
    sub = pos[5:60:3] + 0.125
    assert np.array_equal(ref_driver.get_order_param_q(wl, sub, pos, box, 0.0, 7.0),
                          fn["getOrderParamq"](sub, pos, box, 0.0, 7.0))


def slab_case(seed=3):
    pos, box, z_lo, z_hi = synth.slab_box(3, 3, 2, sigma=0.3, seed=seed)
    return pos, box, z_lo, z_hi


def test_willard_density_matches_live(ref):
    wl, _ = ref
    pos, box, z_lo, z_hi = slab_case()
    gx = np.linspace(0.0, box[0], 7, endpoint=False)
    gy = np.linspace(0.0, box[1], 6, endpoint=False)
    gz = np.linspace(0.0, box[2], 19, endpoint=False)
    d_ref, n_ref = wl.willarddensityfield(pos, gx, gy, gz, box, 2.4)
    d, n = port.willard_density_field(pos, gx, gy, gz, box, 2.4)
    assert np.allclose(d, d_ref, rtol=1e-13, atol=1e-18)
    both = np.isfinite(n_ref)
    assert np.array_equal(np.isfinite(n), both)  # 0/0 far from every water, in both
    assert np.allclose(n[both], n_ref[both], rtol=0, atol=1e-12)
    pts = np.random.default_rng(1).random((50, 3)) * box
    d_ref, n_ref = wl.willarddensitypoints(pos, pts, box, 2.4)
    d, n = port.willard_density_points(pos, pts, box, 2.4)
    ok = np.isfinite(n_ref).all(axis=1)
    assert np.allclose(d, d_ref, rtol=1e-13, atol=1e-18) and np.allclose(n[ok], n_ref[ok], rtol=0, atol=1e-12)


def test_interface_water_matches_live(ref):
    wl, _ = ref
    pos, box, z_lo, z_hi = slab_case(5)
    gp, gn = synth.plane_interface(box, z_lo, z_hi, spacing=2.0)
    wc_ref, sc_ref, nw_ref, dist_ref = wl.interfacewater(pos, gp, gn, 3.0, box)
    wc, sc, nw, dist = port.interface_water(pos, gp, gn, 3.0, box)
    assert np.array_equal(wc + 1, wc_ref) and np.array_equal(sc + 1, sc_ref)  # the Fortran returns 1-based indices
    assert nw == nw_ref and np.array_equal(dist, dist_ref)


def test_histrr3b_matches_live(ref):
    wl, _ = ref
    pos, box = synth.water_box(3, sigma=0.35, seed=8)
    h_ref = wl.histrr3b(pos, box, 0.5, 8, 5.0, 36)
    h = port.histrr3b(pos, box, 0.5, 8, 5.0, 36)
    assert h.sum() > 1000 and np.array_equal(h.astype(np.float64), np.ascontiguousarray(h_ref))


def test_lsi_matches_live(ref):
    """getLSI incl. its non-periodic choice of the next neighbour (water_properties.py:289)."""
    wl, fn = ref
    for m, sigma, seed in ((3, 0.4, 5), (4, 0.6, 6)):
        pos, box = synth.water_box(m, sigma=sigma, seed=seed)
        v_ref, n_ref = fn["getLSI"](pos, pos, box)
        v, n = port.getLSI(pos, pos, box)
        assert np.array_equal(n, n_ref) and v.shape == v_ref.shape and np.allclose(v, v_ref, rtol=1e-12, atol=1e-18)
    sub = pos[::7] + 0.3
    v_ref, n_ref = fn["getLSI"](sub, pos, box, 0.5, 3.5)
    v, n = port.getLSI(sub, pos, box, 0.5, 3.5)
    assert np.array_equal(n, n_ref) and np.allclose(v, v_ref, rtol=1e-12, atol=1e-18)


def test_pair_histograms_match_live(ref):
    """RadialDist / RadialDistSame / PairDistanceHistogram incl. the single-precision (4./3.) and truncated pi."""
    wl, _ = ref
    pos, box = synth.water_box(3, sigma=0.4, seed=2)
    rng = np.random.default_rng(2)
    sol = rng.random((23, 3)) * box
    assert np.array_equal(port.radialdistsame(pos, 0.1, 90, 1.0, box), wl.radialdistsame(pos, 0.1, 90, 1.0, box))
    assert np.array_equal(port.radialdist(sol, pos, 0.1, 90, 0.0334, box), wl.radialdist(sol, pos, 0.1, 90, 0.0334, box))
    assert np.array_equal(port.radialdistsame(pos, 0.25, 400, 1.0, box.reshape(1, 3)), wl.radialdistsame(pos, 0.25, 400, 1.0, box))
    assert np.array_equal(port.pairdistancehistogram(sol, pos, 0.2, 60, box), wl.pairdistancehistogram(sol, pos, 0.2, 60, box))
    assert np.array_equal(port.pairdistancehistogram(pos, pos, 0.2, 60, box), wl.pairdistancehistogram(pos, pos, 0.2, 60, box))


def test_psi_matches_live(ref):
    wl, fn = ref
    pos, box = synth.water_box(3, sigma=0.3, seed=4)
    for sub, lo, hi in ((pos, 0.0, 4.5), (pos[::5] + 0.2, 1.0, 6.0)):
        want = fn["getOrderParamPsi"](sub, pos, box, lo, hi)
        got = port.getOrderParamPsi(sub, pos, box, lo, hi)
        assert np.allclose(got, want, rtol=1e-10, atol=1e-14) and want.max() > 0.05


def test_density_field_matches_live(ref):
    wl, _ = ref
    pos, box, z_lo, z_hi = slab_case(7)
    gx = np.linspace(0.0, box[0], 9, endpoint=False) + 0.3
    gy = np.linspace(0.0, box[1], 7, endpoint=False)
    gz = np.linspace(0.0, box[2], 15, endpoint=False)
    want = wl.densityfield(pos, gx, gy, gz, box)
    assert want.max() > 0 and np.array_equal(port.density_field(pos, gx, gy, gz, box), want)


def test_watorient_and_binongrid_match_live(ref):
    """watOrient (through the stub's internal_pack: the Fortran passes strided sections) and binOnGrid, bit for bit."""
    wl, _ = ref
    o, box = synth.water_box(4, sigma=0.4, seed=3)
    h = synth.add_hydrogens(o, seed=3)
    h[::5] += box          # some hydrogens stored in the neighbouring image: the minimum image must bring them back
    for refvec in ([0.0, 0.0, 1.0], [1.0, 2.0, -0.5], [0.0, -3.0, 0.0]):
        a, b = wl.watorient(o, h, refvec, box)
        pa, pb = port.watorient(o, h, refvec, box)
        assert np.array_equal(a, pa) and np.array_equal(b, pb) and a.min() >= 0.0 and a.max() <= 180.0
    edges = np.arange(0.0, 24.0, 3.0)
    for shift in (0.0, 0.7):
        r = wl.binongrid(o, edges + shift, edges, edges - 1.0)
        assert np.array_equal(np.ascontiguousarray(r), port.binongrid(o, edges + shift, edges, edges - 1.0)) and r.sum() > 50


def test_clusters_restatement_matches_the_live_getClusters():
    """oracle/ref_driver.get_clusters against the reference's unmodified getClusters body (orderParam_lib.py:123-156),
    both over the reference's compiled depthfirstsort (fortran/sortlib.f90:26-72)."""
    from oracle import ref_driver
    sl = ref_fortran.RefSortlib()
    live = ref_fortran.load_reference_driver_functions(["getClusters"], sl)["getClusters"]
    rng = np.random.default_rng(11)
    for n, p in ((25, 0.04), (40, 0.03), (16, 0.3), (9, 0.0), (33, 0.06)):
        m = np.triu((rng.random((n, n)) < p).astype(int), 1)
        m = m + m.T
        a, b = live(m), ref_driver.get_clusters(sl, m)
        assert len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))
        assert sorted(np.concatenate(b).tolist()) == list(range(n)) or len(b[-1]) == n


def test_radialdistplane_restatement_matches_live(ref):
    """RadialDistPlane (fortran/waterlib.f90:237-314; its matmul runs in the stub runtime with libgfortran 5's order)."""
    wl, _ = ref
    rng = np.random.default_rng(4)
    L = np.array([38.0, 41.0, 35.0])
    p1 = np.array([[2.0, 1.5, 1.0], [2.4, 4.4, 1.2], [5.1, 1.2, 1.5]])
    p2 = rng.random((800, 3)) * np.array([13.0, 13.0, 11.0]) + np.array([3.5, 3.5, 0.5])
    for bw, nb in ((0.5, 36), (0.8, 20)):
        mine, bad = port.radialdistplane(p1, p2, bw, nb, 0.03, L)
        assert bad == 0 and mine.sum() > 50
        assert np.array_equal(wl.radialdistplane(p1, p2, bw, nb, 0.03, L), mine)
