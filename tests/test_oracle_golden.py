"""The CPU oracle (oracle/wol_oracle.c via oracle/port.py) against the committed golden fixtures, which
hold outputs of the reference's own compiled Fortran + unmodified Python (tests/golden/make_golden.py).
Bit-exact everywhere: the oracle restates the same fp64 operations in the same order."""
import glob
import os

import numpy as np
import pytest

from oracle import port

Q3B_CASES = ["cfg1_n512_ice", "cfg1_n512_liq", "lattice_n216", "random_n160_noncubic", "subpop_n512_m97",
             "cfg2_n4096_frame0"]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", Q3B_CASES)
def test_q_and_selection(golden_dir, name):
    g = load(golden_dir, name)
    q, nn4, nq = port.order_param_q(g["sub"], g["pos"], g["box"], float(g["lowq"]), float(g["highq"]))
    assert np.array_equal(nq, g["nq"])
    assert np.array_equal(nn4, g["nn4"])
    assert np.array_equal(q, g["q"])  # bit-exact


@pytest.mark.parametrize("name", Q3B_CASES)
def test_three_body(golden_dir, name):
    g = load(golden_dir, name)
    r = port.three_body(g["sub"], g["pos"], g["box"], float(g["low3"]), float(g["high3"]))
    assert np.array_equal(r["numAngs"], g["n3"])
    assert r["n_angles"] == int(g["n_angles"])
    assert np.array_equal(r["hist"], g["hist"])
    if "angVals" in g.files:
        assert np.array_equal(r["angVals"], g["angVals"])
    angDist, _bins, frac, avg, var, ent = port.tetrahedralMetrics(r["angVals"])
    assert np.array_equal(angDist, g["hist"])
    assert frac == float(g["fracTet"])
    assert avg == float(g["avgCos"]) and var == float(g["varCos"]) and ent == float(g["entropy"])
    # the streaming sums the CUDA path also produces
    assert r["tet"][0] == round(frac * r["n_angles"])
    assert abs(r["tet"][1] / r["tet"][0] - avg) < 1e-13


def test_lattice_is_tetrahedral(golden_dir):
    g = load(golden_dir, "lattice_n216")
    assert np.all(g["n3"] == 4) and np.all(g["nq"] >= 4)
    assert np.allclose(g["q"], 1.0, atol=1e-6)
    assert np.allclose(g["angVals"], 109.4712206, atol=1e-4)


def test_histogram_matches_numpy():
    rng = np.random.default_rng(3)
    x = np.concatenate([rng.random(20000) * 200.0 - 10.0, np.linspace(0.0, 180.0, 501), [-180.0, 180.0, 0.0]])
    assert np.array_equal(port.histogram(x, 500, 0.0, 180.0), np.histogram(x, bins=500, range=[0.0, 180.0])[0])
    q = rng.random(5000)
    assert np.array_equal(port.histogram(q, 500, 0.0, 1.0), np.histogram(q, bins=500, range=[0.0, 1.0])[0])


@pytest.mark.parametrize("tag", ["35_120", "30_150"])
def test_hbonds(golden_dir, tag):
    g = load(golden_dir, "hbonds_n512_" + tag)
    ac, dc = port.hbonds(g["acc"], g["don"], g["donh"], g["box"], float(g["distcut"]), float(g["angcut"]))
    assert np.array_equal(ac, g["acc_count"]) and np.array_equal(dc, g["don_count"])
    assert int(ac.sum()) == int(g["n_bonds"])


def test_shell(golden_dir):
    g = load(golden_dir, "shell_n4096")
    mask = port.shell_mask(g["sol"], g["pos"], g["box"], float(g["cutoff"]))
    assert np.array_equal(np.nonzero(mask)[0].astype(np.int32), g["shell"])


def test_fixtures_match_generator(golden_dir):
    """The synthetic generator must keep producing the inputs the fixtures were made from."""
    from waterorderlib_b200 import synth
    g = load(golden_dir, "cfg1_n512_ice")
    pos, box = synth.water_box(4, sigma=0.25, seed=1234)
    assert np.array_equal(pos, g["pos"]) and np.array_equal(box, g["box"])
    assert len(glob.glob(os.path.join(golden_dir, "*.npz"))) >= 10


def test_slab_routines_against_golden(golden_dir):
    """Willard-Chandler density, InterfaceWater (fortran/waterlib.f90:1286-1469)."""
    g = np.load(os.path.join(golden_dir, "slab_n256.npz"))
    d, n = port.willard_density_field(g["pos"], g["gx"], g["gy"], g["gz"], g["box"], float(g["smoothlen"]))
    fin = np.isfinite(g["norms"])
    assert np.allclose(d, g["dens"], rtol=1e-13, atol=1e-18) and np.array_equal(np.isfinite(n), fin)
    assert np.allclose(n[fin], g["norms"][fin], rtol=0, atol=1e-12)
    d, n = port.willard_density_points(g["pos"], g["pts"], g["box"], float(g["smoothlen"]))
    assert np.allclose(d, g["pdens"], rtol=1e-13, atol=1e-18)
    wc, sc, nw, dist = port.interface_water(g["pos"], g["gridpos"], g["gridnorm"], float(g["cutoff"]), g["box"])
    assert np.array_equal(wc + 1, g["watclose"]) and np.array_equal(sc + 1, g["surfclose"])
    assert nw == int(g["numwater"]) and np.array_equal(dist, g["allwatdists"])


def test_histrr3b_against_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "histrr3b_n216.npz"))
    h = port.histrr3b(g["pos"], g["box"], float(g["dwidth"]), int(g["dnum"]), float(g["awidth"]), int(g["anum"]))
    assert np.array_equal(h, g["hist"])


def test_lsi_against_golden(golden_dir):
    """getLSI (structureLibs/water_properties.py:252-311), values from the reference's own function body."""
    g = np.load(os.path.join(golden_dir, "lsi_n512.npz"))
    v, n = port.getLSI(g["pos"], g["pos"], g["box"])
    assert np.array_equal(n, g["num"]) and np.allclose(v, g["lsi"], rtol=1e-12, atol=1e-18)
    v, n = port.getLSI(g["sub"], g["pos"], g["box"], float(g["low_sub"]), float(g["high_sub"]))
    assert np.array_equal(n, g["num_sub"]) and np.allclose(v, g["lsi_sub"], rtol=1e-12, atol=1e-18)


def test_pairs_against_golden(golden_dir):
    """RadialDistSame / RadialDist / PairDistanceHistogram (waterlib.f90:193-389), getOrderParamPsi (:393-433)."""
    g = np.load(os.path.join(golden_dir, "pairs_n512.npz"))
    assert np.array_equal(port.radialdistsame(g["pos"], 0.1, 120, 1.0, g["box"]), g["rdf_same"])
    assert np.array_equal(port.radialdist(g["sol"], g["pos"], 0.1, 120, 0.0334, g["box"]), g["rdf_cross"])
    assert np.array_equal(port.pairdistancehistogram(g["sol"], g["pos"], 0.25, 40, g["box"]), g["pdh"])
    assert np.allclose(port.getOrderParamPsi(g["pos"], g["pos"], g["box"], 0.0, 4.5), g["psi_all"], rtol=1e-10, atol=1e-14)
    assert np.allclose(port.getOrderParamPsi(g["sol"], g["pos"], g["box"], 1.0, 6.0), g["psi_sub"], rtol=1e-10, atol=1e-14)


def test_density_field_against_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "slab_n256.npz"))
    assert np.array_equal(port.density_field(g["pos"], g["gx"] + 0.3, g["gy"], g["gz"], g["box"]), g["voxel"])


def test_iso_points_properties():
    """The iso-surface restatement (parity unpinned: skimage is not installed) pinned by its own properties: every
    vertex lies on a grid edge, between two nodes whose values straddle the level, at the linear interpolation; a
    linear field is interpolated exactly; one vertex per straddling edge."""
    from oracle import port
    g = np.linspace(0.0, 1.0, 9)
    field = np.add.outer(np.add.outer(g, 2.0 * g), 3.0 * g)  # x + 2y + 3z
    pts = port.iso_points(field, g, g, g, 2.2)
    assert np.allclose(pts[:, 0] + 2.0 * pts[:, 1] + 3.0 * pts[:, 2], 2.2, atol=1e-12)
    on_node = np.isclose(pts[:, :, None], g[None, None, :], atol=1e-12).any(axis=2)
    assert np.all(on_node.sum(axis=1) >= 2)  # at most one coordinate is off the grid lines
    above = field > 2.2
    n_edges = sum(int(np.sum(np.diff(above, axis=a) != 0)) for a in range(3))
    assert len(pts) == n_edges > 0
    X, Y, Z = np.meshgrid(g - 0.5, g - 0.5, g - 0.5, indexing="ij")
    r = np.sqrt(X * X + Y * Y + Z * Z)
    sph = port.iso_points(r, g, g, g, 0.3)
    rr = np.linalg.norm(sph - 0.5, axis=1)
    assert len(sph) > 50 and np.all(rr <= 0.3 + 1e-12) and np.all(rr > 0.27)  # chords of a convex field lie inside
    assert port.iso_points(r, g, g, g, 5.0).shape == (0, 3)


def test_watorient_and_binongrid_golden(golden_dir):
    from oracle import port
    g = np.load(os.path.join(golden_dir, "orient_n512.npz"))
    dip, plane = port.watorient(g["opos"], g["hpos"], g["refvec"], g["box"])
    assert np.array_equal(dip, g["angdip"]) and np.array_equal(plane, g["angplane"])
    assert np.array_equal(port.binongrid(g["opos"], g["xbins"], g["ybins"], g["zbins"]), g["occupancy"])
    with pytest.raises(ValueError):
        port.binongrid(g["opos"], g["xbins"], g["ybins"] * 1.5, g["zbins"])
    with pytest.raises(ValueError):
        port.watorient(g["opos"], g["hpos"][:-1], g["refvec"], g["box"])


def test_clusters_and_rdfplane_against_golden(golden_dir):
    """getClusters through the staged reference DFS (oracle/_ref) and the RadialDistPlane restatement, against fixtures made
    by the reference's own getClusters body / compiled RadialDistPlane (tests/golden/make_golden_extras.py)."""
    from oracle import port, ref_driver, ref_fortran
    g = np.load(os.path.join(golden_dir, "clusters.npz"))
    if ref_fortran.sortlib_available():
        sl = ref_fortran.RefSortlib()
        for k in range(int(g["n_cases"])):
            cl = ref_driver.get_clusters(sl, g["mat%d" % k])
            assert [len(c) for c in cl] == list(g["sizes%d" % k]) and np.array_equal(np.concatenate(cl), g["members%d" % k])
    r = np.load(os.path.join(golden_dir, "rdfplane.npz"))
    for tag in ("a", "b"):
        rdf, bad = port.radialdistplane(r["pos1"], r["pos2"], float(r["binwidth_" + tag]), int(r["totbins_" + tag]), float(r["bulkdens"]), r["box"])
        assert bad == 0 and np.array_equal(rdf, r["rdf_" + tag])
    _, bad = port.radialdistplane(r["pos1"], r["pos2"] - 6.0, 0.5, 40, float(r["bulkdens"]), r["box"])
    assert bad > 0
