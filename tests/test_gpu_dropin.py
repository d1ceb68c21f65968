"""Parity of the drop-in Python API (waterorderlib_b200.structureLibs: water_properties + the f2py-compatible
waterlib shim) against the golden fixtures generated from the reference's compiled Fortran + its own Python,
and against the CPU oracle on seeded inputs.  The calls read like the reference's own call sites
(structureLibs/orderParam_lib.py:1325, :1471, :805-807, :495-498).

Integer outputs (neighbour matrices, counts, histogram bins, bond lists) are bit-exact.  q within 1e-6
relative (north_star, fp64 mode).  Returned angle VALUES go through the device acos instead of glibc's:
they agree to a few ulps (tolerance 1e-12 relative, 1e-10 degrees absolute); bin membership of the fused
histogram path does not depend on that (tests/test_gpu_q3b.py).
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import port  # noqa: E402  (the checker)
from waterorderlib_b200 import routines, synth  # noqa: E402
from waterorderlib_b200.structureLibs import water_properties as wp  # noqa: E402
from waterorderlib_b200.structureLibs import waterlib as wl  # noqa: E402

Q_RTOL = 1e-6
ANG_RTOL, ANG_ATOL = 1e-12, 1e-10
Q3B_CASES = ["cfg1_n512_ice", "cfg1_n512_liq", "lattice_n216", "random_n160_noncubic", "subpop_n512_m97"]


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", Q3B_CASES + ["cfg2_n4096_frame0"])
def test_getOrderParamq(golden_dir, name):
    g = load(golden_dir, name)
    q = wp.getOrderParamq(g["sub"], g["pos"], g["box"], float(g["lowq"]), float(g["highq"]))
    assert isinstance(q, np.ndarray) and q.dtype == np.float64 and q.shape == g["q"].shape
    assert np.allclose(q, g["q"], rtol=Q_RTOL, atol=1e-9)


def test_getOrderParamq_defaults_and_box_shapes(golden_dir):
    g = load(golden_dir, "cfg1_n512_ice")
    q1 = wp.getOrderParamq(g["pos"], g["pos"], g["box"])               # (3,) box, default cutoffs 0 / 10
    q2 = wp.getOrderParamq(g["pos"], g["pos"], g["box"].reshape(1, 3))  # (1,3) box (orderParam_lib.py:183)
    assert np.array_equal(q1, q2) and np.allclose(q1, g["q"], rtol=Q_RTOL, atol=1e-9)


@pytest.mark.parametrize("name", Q3B_CASES)
def test_getCosAngs(golden_dir, name):
    g = load(golden_dir, name)
    ang, num = wp.getCosAngs(g["sub"], g["pos"], g["box"], float(g["low3"]), float(g["high3"]))
    assert num.dtype == np.float64 and np.array_equal(num, g["n3"].astype(np.float64))
    assert ang.shape == g["angVals"].shape
    assert np.allclose(ang, g["angVals"], rtol=ANG_RTOL, atol=ANG_ATOL)  # same order, same values


@pytest.mark.parametrize("name", Q3B_CASES)
def test_tetrahedralMetrics(golden_dir, name):
    g = load(golden_dir, name)
    angDist, bins, fracTet, avgCos, varCos, entropy = wp.tetrahedralMetrics(g["angVals"])
    assert np.array_equal(angDist, g["hist"])
    assert np.array_equal(bins, np.linspace(0.0, 180.0, 501))
    assert fracTet == float(g["fracTet"])
    assert abs(avgCos - float(g["avgCos"])) < 1e-12 and abs(varCos - float(g["varCos"])) < 1e-12
    assert abs(entropy - float(g["entropy"])) < 1e-12
    # other bin specs
    d2, b2, *_ = wp.tetrahedralMetrics(g["angVals"], nBins=90, binRange=[30.0, 150.0])
    assert np.array_equal(d2, np.histogram(g["angVals"], bins=90, range=[30.0, 150.0])[0]) and b2.size == 91


def test_tetrahedralMetrics_empty_raises():
    with pytest.raises(ZeroDivisionError):
        wp.tetrahedralMetrics(np.array([]))


def test_collinear_angles_are_minus_180():
    """CosAngle3 returns -180 for an exactly antiparallel pair."""
    g = np.arange(4) * 3.0
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    box = np.array([12.0, 12.0, 12.0])
    ang, num = wp.getCosAngs(pos, pos, box, 0.0, 3.2)
    ref, refnum = port.getCosAngs(pos, pos, box, 0.0, 3.2)
    assert np.array_equal(num, refnum) and np.array_equal(ang == -180.0, ref == -180.0)
    assert (ang == -180.0).sum() == 64 * 3 and np.allclose(ang, ref, rtol=ANG_RTOL, atol=ANG_ATOL)


def test_torch_in_torch_out(golden_dir):
    g = load(golden_dir, "cfg1_n512_liq")
    pos = torch.from_numpy(g["pos"]).cuda()
    q = wp.getOrderParamq(pos, pos, g["box"])
    assert isinstance(q, torch.Tensor) and q.is_cuda
    assert np.allclose(q.cpu().numpy(), g["q"], rtol=Q_RTOL, atol=1e-9)
    ang, num = wp.getCosAngs(pos, pos, g["box"])
    assert ang.is_cuda and np.array_equal(num.cpu().numpy(), g["n3"].astype(np.float64))


# ---- f2py-compatible waterlib shim -----------------------------------------------------------------

def test_neighbor_matrices_vs_oracle():
    rng = np.random.default_rng(5)
    box = np.array([15.0, 11.0, 13.0])
    pos = rng.random((300, 3)) * box * 1.5 - 2.0
    sub = rng.random((40, 3)) * box
    pos[7] = pos[3]  # coincident pair: excluded by lowCut = 0
    pos[11] = pos[10] + np.array([7.5, 0.0, 0.0])  # exactly L/2 apart along x: anint rounds half away
    a = wl.allnearneighbors(pos, box, 0.0, 3.5)
    assert a.dtype == np.int32 and a.shape == (300, 300) and a.flags.f_contiguous
    assert np.array_equal(a, port.neighbor_matrix(pos, pos, box, 0.0, 3.5))
    assert np.array_equal(a, a.T) and a[3, 7] == 0
    for lo, hi in ((0.0, 4.0), (2.5, 7.4), (0.0, 7.5)):
        b = wl.nearneighbors(sub, pos, box, lo, hi)
        assert np.array_equal(b, port.neighbor_matrix(sub, pos, box, lo, hi))
    # negative edge = that axis is not periodic (waterlib.f90:41)
    nb = np.array([15.0, -1.0, 13.0])
    assert np.array_equal(wl.nearneighbors(sub, pos, nb, 0.0, 5.0), port.neighbor_matrix(sub, pos, nb, 0.0, 5.0))


def test_reimage_tetracosang_lsidists_golden(golden_dir):
    g = load(golden_dir, "routines")
    assert np.array_equal(wl.reimage(g["neigh"], g["ref"], g["box"]), g["reimaged"])
    assert np.array_equal(wl.lsidists(g["ref"], g["neigh"], g["box"]), g["lsid"])
    a = wl.tetracosang(g["ref"], g["neigh"], g["box"])
    off = ~np.eye(7, dtype=bool)
    assert np.allclose(a[off], g["angs"][off], rtol=ANG_RTOL, atol=ANG_ATOL) and np.all(np.diag(a) == 0.0)
    assert np.array_equal(a[off] == -180.0, g["angs"][off] == -180.0)


@pytest.mark.parametrize("tag", ["35_120", "30_150"])
def test_generalhbonds_golden(golden_dir, tag):
    g = load(golden_dir, "hbonds_n512_" + tag)
    m = wl.generalhbonds(g["acc"], g["don"], g["donh"], g["box"], float(g["distcut"]), float(g["angcut"]))
    assert m.dtype == np.int32 and m.shape == (512, 1024)
    assert np.array_equal(m.sum(axis=1), g["acc_count"]) and np.array_equal(m.sum(axis=0), g["don_count"])
    assert int(m.sum()) == int(g["n_bonds"])
    _, _, ref = port.hbonds(g["acc"], g["don"], g["donh"], g["box"], float(g["distcut"]), float(g["angcut"]), dense=True)
    assert np.array_equal(m, ref)


def test_generalhbonds_mismatched_donors():
    with pytest.raises(ValueError):
        wl.generalhbonds(np.zeros((2, 3)), np.zeros((3, 3)), np.zeros((2, 3)), np.ones(3) * 10, 3.5, 120.0)


def test_HBondsGeneral(golden_dir):
    g = load(golden_dir, "hbonds_n512_35_120")
    acc, don, donh, box = g["acc"], g["don"], g["donh"], g["box"]
    accInds = np.arange(0, 3 * 512, 3)
    donInds = np.repeat(accInds, 2)
    donHInds = np.arange(3 * 512)[np.arange(3 * 512) % 3 != 0]
    n, lst, loc = wp.HBondsGeneral(acc, don, donh, box, accInds, donInds, donHInds, 3.5, 120.0)
    assert n == int(g["n_bonds"]) and lst.shape == (n, 2) and loc.shape == (n, 3)
    _, _, mat = port.hbonds(acc, don, donh, box, 3.5, 120.0, dense=True)
    ii, jj = np.nonzero(mat)  # row-major: the order the reference walks its matrix in
    assert np.array_equal(lst[:, 0], accInds[ii]) and np.array_equal(lst[:, 1], donInds[jj])
    d = donh[jj] - acc[ii]
    d = d - box * np.round(d / box)
    assert np.allclose(loc, 0.5 * ((acc[ii] + d) + acc[ii]), rtol=0, atol=1e-12)


def test_batched_hbond_counts_vs_oracle():
    """cfg2 shape: several frames in one call; per-water sums as hbCalc takes them (orderParam_lib.py:867-870)."""
    F = 3
    res_acc, res_don = [], []
    O, H = [], []
    for f in range(F):
        o, box = synth.water_box(6, sigma=0.3, seed=50 + f)
        h = synth.add_hydrogens(o, seed=50 + f)
        O.append(o); H.append(h)
        a, d = port.hbonds(o, np.repeat(o, 2, axis=0), h, box, 3.5, 120.0)
        res_acc.append(a); res_don.append(d)
    O, H = np.stack(O), np.stack(H)
    r = routines.hbond_counts(O, np.repeat(O, 2, axis=1), H, box, 3.5, 120.0)
    assert np.array_equal(r["acc_count"].cpu().numpy(), np.stack(res_acc))
    assert np.array_equal(r["don_count"].cpu().numpy(), np.stack(res_don))


def test_shell_mask_golden(golden_dir):
    g = load(golden_dir, "shell_n4096")
    mask = routines.shell_mask(g["sol"], g["pos"], g["box"], float(g["cutoff"]))
    assert np.array_equal(np.nonzero(mask.cpu().numpy()[0])[0].astype(np.int32), g["shell"])
    nb = wl.nearneighbors(g["sol"], g["pos"], g["box"], 0.0, float(g["cutoff"]))
    assert np.array_equal(np.unique(np.where(nb == 1)[1]).astype(np.int32), g["shell"])  # orderParam_lib.py:495-496


def test_getLSI_golden(golden_dir):
    g = load(golden_dir, "lsi_n512")
    v, n = wp.getLSI(g["pos"], g["pos"], g["box"])
    assert np.array_equal(n, g["num"]) and v.shape == g["lsi"].shape
    assert np.allclose(v, g["lsi"], rtol=1e-10, atol=1e-16)
    v, n = wp.getLSI(g["sub"], g["pos"], g["box"], float(g["low_sub"]), float(g["high_sub"]))
    assert np.array_equal(n, g["num_sub"]) and np.allclose(v, g["lsi_sub"], rtol=1e-10, atol=1e-16)


def test_getLSI_vs_oracle_unwrapped_coordinates():
    """The next-shell neighbour is chosen by NON-periodic distance (water_properties.py:289): unwrapped inputs."""
    pos, box = synth.water_box(5, sigma=0.5, seed=12)
    rng = np.random.default_rng(12)
    pos = pos + box * rng.integers(-1, 2, size=pos.shape)
    v, n = wp.getLSI(pos, pos, box)
    v_ref, n_ref = port.getLSI(pos, pos, box)
    assert np.array_equal(n, n_ref) and np.allclose(v, v_ref, rtol=1e-10, atol=1e-16)
    # dilute gas: centres with fewer than two neighbours (or an empty next shell) have no value -- skipped in
    # lsiVals, 0 in numLSI
    gas = rng.random((400, 3)) * np.array([40.0, 42.0, 38.0])
    v, n = wp.getLSI(gas, gas, np.array([40.0, 42.0, 38.0]))
    v_ref, n_ref = port.getLSI(gas, gas, np.array([40.0, 42.0, 38.0]))
    assert (n_ref == 0).any() and (n_ref > 0).any() and v.shape == v_ref.shape
    assert np.array_equal(n, n_ref) and np.allclose(v, v_ref, rtol=1e-10, atol=1e-16)


def test_pair_histograms_golden(golden_dir):
    """radialdistsame / radialdist / pairdistancehistogram: bit-exact g(r) incl. the Fortran's normalisation."""
    g = load(golden_dir, "pairs_n512")
    assert np.array_equal(wl.radialdistsame(g["pos"], 0.1, 120, 1.0, g["box"]), g["rdf_same"])
    assert np.array_equal(wl.radialdist(g["sol"], g["pos"], 0.1, 120, 0.0334, g["box"]), g["rdf_cross"])
    assert np.array_equal(wl.pairdistancehistogram(g["sol"], g["pos"], 0.25, 40, g["box"]), g["pdh"])
    # a range the cell grid cannot hold (> L / 3): every cell is enumerated, still each pair once
    big = wl.radialdistsame(g["pos"], 0.5, 40, 1.0, g["box"].reshape(1, 3))
    assert np.array_equal(big, port.radialdistsame(g["pos"], 0.5, 40, 1.0, g["box"]))
    # coincident atoms (distance 0) are not counted; 4096 waters through the real cell list
    pos, box = synth.water_box(8, sigma=0.4, seed=6)
    pos[5] = pos[4]
    assert np.array_equal(wl.radialdistsame(pos, 0.1, 100, 1.0, box), port.radialdistsame(pos, 0.1, 100, 1.0, box))
    assert np.array_equal(wl.pairdistancehistogram(pos[:300], pos, 0.2, 50, box), port.pairdistancehistogram(pos[:300], pos, 0.2, 50, box))


def test_getOrderParamPsi_golden(golden_dir):
    g = load(golden_dir, "pairs_n512")
    assert np.allclose(wp.getOrderParamPsi(g["pos"], g["pos"], g["box"], 0.0, 4.5), g["psi_all"], rtol=1e-9, atol=1e-13)
    assert np.allclose(wp.getOrderParamPsi(g["sol"], g["pos"], g["box"], 1.0, 6.0), g["psi_sub"], rtol=1e-9, atol=1e-13)
    # default cutoff 10 A (~140 neighbours, ~10^4 pairs per centre) and the fewer-than-two-neighbours rule
    pos, box = synth.water_box(4, sigma=0.3, seed=3)
    assert np.allclose(wp.getOrderParamPsi(pos[:40], pos, box), port.getOrderParamPsi(pos[:40], pos, box), rtol=1e-9, atol=1e-13)
    lone = np.array([[1.0, 1.0, 1.0], [3.0, 1.0, 1.0], [20.0, 20.0, 20.0]])
    assert np.array_equal(wp.getOrderParamPsi(lone, lone, np.array([40.0, 40.0, 40.0]), 0.0, 5.0), np.zeros(3))


def test_integration_md_ctypes_stub_runs(golden_dir):
    """The ctypes binding printed in INTEGRATION.md section 3 is real code: extract it, point it at the built library and
    check its output against the golden fixture."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = next(b for b in blocks if "def q_and_three_body" in b)
    stub = stub.replace("/path/to/waterorderlib_b200/libwol.so", os.path.join(root, "waterorderlib_b200", "libwol.so"))
    ns = {}
    exec(compile(stub, "INTEGRATION.md", "exec"), ns)
    g = load(golden_dir, "cfg1_n512_liq")
    q, n3, hist = ns["q_and_three_body"](g["pos"], g["box"])
    assert np.allclose(q, g["q"], rtol=Q_RTOL, atol=1e-9) and np.array_equal(n3, g["n3"]) and np.array_equal(hist, g["hist"])


def test_hist_allreduce_over_a_caller_owned_nccl_communicator():
    """wol_hist_allreduce with a raw ncclComm_t (single rank here; scripts/dist_check.py covers two GPUs): the sum over
    one rank is the input, in place, for int64 bins and for doubles; bad arguments are refused."""
    import ctypes
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "tools"))
    import raw_nccl
    from waterorderlib_b200._capi import lib
    nccl = raw_nccl.load()
    torch.cuda.set_device(0)
    comm = raw_nccl.comm_init(nccl, raw_nccl.unique_id(nccl), 1, 0)
    try:
        hist = torch.arange(1011, dtype=torch.int64, device="cuda") * 3_000_000_007
        rows = torch.linspace(-1.0, 1.0, 77, dtype=torch.float64, device="cuda")
        h0, r0 = hist.clone(), rows.clone()
        s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        assert lib().wol_hist_allreduce(comm, ctypes.c_void_p(hist.data_ptr()), hist.numel(), 0, s) == 0
        assert lib().wol_hist_allreduce(comm, ctypes.c_void_p(rows.data_ptr()), rows.numel(), 1, s) == 0
        torch.cuda.synchronize()
        assert torch.equal(hist, h0) and torch.equal(rows, r0)
        assert lib().wol_hist_allreduce(comm, None, 0, 0, s) == 0
        assert lib().wol_hist_allreduce(None, ctypes.c_void_p(hist.data_ptr()), 4, 0, s) != 0
        assert lib().wol_hist_allreduce(comm, ctypes.c_void_p(hist.data_ptr()), 4, 7, s) != 0
    finally:
        nccl.ncclCommDestroy(comm)


def _csr_from_dense(mat):
    off = np.concatenate([[0], np.cumsum(mat.sum(axis=1))])
    return off, np.nonzero(mat)[1].astype(np.int32)


def test_neighbors_csr_matches_dense_reference_matrices():
    """wol_neighbors_csr against the dense matrices of allNearNeighbors / nearNeighbors (oracle restatement, pinned on
    the compiled Fortran): same pairs, ascending atom index within a centre; several frames, sub-populations, float32
    centres, cut-offs from 3.4 to 10 A, boxes below four cells per edge, empty inputs."""
    pos, box = synth.water_box(4, sigma=0.4, seed=12)
    for low, high in ((0.0, 3.413), (0.0, 10.0), (2.8, 5.5)):
        off, idx = routines.neighbors_csr(None, pos, box, low, high)
        ro, ri = _csr_from_dense(port.neighbor_matrix(pos, pos, box, low, high))
        assert np.array_equal(off.cpu().numpy(), ro) and np.array_equal(idx.cpu().numpy(), ri)
    rng = np.random.default_rng(5)
    sub = (rng.random((37, 3)) * box).astype(np.float32)
    off, idx = routines.neighbors_csr(torch.from_numpy(sub).cuda(), pos, box, 0.0, 4.0)
    ro, ri = _csr_from_dense(port.neighbor_matrix(sub.astype(np.float64), pos, box, 0.0, 4.0))
    assert np.array_equal(off.cpu().numpy(), ro) and np.array_equal(idx.cpu().numpy(), ri)
    xyz, boxes = synth.trajectory(3, 4, sigma=0.5, seed0=8)     # 216 waters: three cells per edge at 6 A
    off, idx = routines.neighbors_csr(None, xyz, boxes, 0.0, 6.0)
    off, idx = off.cpu().numpy(), idx.cpu().numpy()
    n = xyz.shape[1]
    for f in range(4):
        ro, ri = _csr_from_dense(port.neighbor_matrix(xyz[f], xyz[f], boxes[f], 0.0, 6.0))
        seg = off[f * n:(f + 1) * n + 1]
        assert np.array_equal(seg - seg[0], ro) and np.array_equal(idx[seg[0]:seg[-1]], ri)
    off, idx = routines.neighbors_csr(np.zeros((0, 3)), pos, box, 0.0, 3.5)
    assert off.tolist() == [0] and idx.numel() == 0


def test_neighbors_csr_large_counts_agree_with_the_fused_sweep():
    pos, box = synth.water_box(30, sigma=0.3, seed=2)            # 216 000 waters
    off, idx = routines.neighbors_csr(None, pos, box, 0.0, 3.413)
    r = engine_q3b(pos, box)
    counts = (off[1:] - off[:-1]).to(torch.int32)
    assert torch.equal(counts, r["n3"][0]) and int(off[-1]) == idx.numel()
    seg = torch.repeat_interleave(torch.arange(counts.numel(), device=idx.device), counts.to(torch.int64))
    i = idx.to(torch.int64)
    assert bool(((i[1:] > i[:-1]) | (seg[1:] != seg[:-1])).all())   # ascending atom index inside every segment
    for c in (0, 777, 215_999):                                  # a few centres against the oracle's dense row
        row = port.neighbor_matrix(pos[c:c + 1], pos, box, 0.0, 3.413)[0]
        assert np.array_equal(idx[off[c]:off[c + 1]].cpu().numpy(), np.nonzero(row)[0])


def engine_q3b(pos, box):
    from waterorderlib_b200 import engine
    return engine.q3b_frames(pos, box, do_q=False, want=("n3",))


def test_water_orientation_and_sphere_occupancy(golden_dir):
    """watorient / binongrid and their callers waterOrientation, waterOrientationBinZ, binnedVolumePofN
    (reference water_properties.py:578-676): golden values from the reference's compiled Fortran, the oracle at a larger
    size and over several frames.  Angles to 1e-12 degrees (device acos vs glibc's), occupancy counts exact."""
    from waterorderlib_b200.structureLibs import water_properties as wp
    from waterorderlib_b200.structureLibs import waterlib as wl
    g = np.load(os.path.join(golden_dir, "orient_n512.npz"))
    dip, plane = wl.watorient(g["opos"], g["hpos"], g["refvec"], g["box"])
    assert np.allclose(dip, g["angdip"], rtol=0, atol=1e-12) and np.allclose(plane, g["angplane"], rtol=0, atol=1e-12)
    occ = wl.binongrid(g["opos"], g["xbins"], g["ybins"], g["zbins"])
    assert occ.dtype == np.int32 and np.array_equal(occ, g["occupancy"])
    xyz, boxes = synth.trajectory(6, 3, sigma=0.5, seed0=60)
    hyd = np.stack([synth.add_hydrogens(xyz[f], seed=60 + f) for f in range(3)])
    d, p = routines.water_orient(xyz, hyd, boxes, (0.3, -1.0, 2.0))
    for f in range(3):
        rd, rp = port.watorient(xyz[f], hyd[f], (0.3, -1.0, 2.0), boxes[f])
        assert np.allclose(d[f].cpu().numpy(), rd, rtol=0, atol=1e-12) and np.allclose(p[f].cpu().numpy(), rp, rtol=0, atol=1e-12)
    da, pa = wp.waterOrientation(xyz[0], hyd[0], boxes[0])
    rd, rp = port.watorient(xyz[0], hyd[0], (0.0, 0.0, 1.0), boxes[0])
    assert isinstance(da, np.ndarray) and np.allclose(da, rd, atol=1e-12) and np.allclose(pa, rp, atol=1e-12)
    plane2d, dip2d = wp.waterOrientationBinZ(xyz[0], hyd[0], boxes[0], angBins=np.arange(0.0, 180.001, 5.0))
    z = xyz[0][:, 2]
    want, _, _ = np.histogram2d(rd, z, bins=[np.arange(0.0, 180.001, 5.0), np.arange(z.min(), z.max(), 0.2)])
    assert np.array_equal(dip2d, want) and plane2d.shape == want.shape and plane2d.sum() == dip2d.sum()
    edges = np.linspace(0.0, boxes[0][0], 9)
    pn = wp.binnedVolumePofN(xyz[0], (edges, edges, edges), np.arange(0, 12))
    ref_occ = port.binongrid(xyz[0], edges, edges, edges)
    assert np.array_equal(pn, np.histogram(ref_occ.flatten(), bins=np.arange(0, 12))[0]) and pn.sum() == 8 ** 3
    with pytest.raises(ValueError):
        wl.binongrid(xyz[0], edges, edges * 1.1, edges)
    with pytest.raises(ValueError):
        wl.watorient(xyz[0], hyd[0][:-1], (0, 0, 1.0), boxes[0])


def test_getClusters_matches_the_reference_depth_first_search(golden_dir):
    """Fixture: the reference's own getClusters body (orderParam_lib.py:123-156) over its compiled
    sortlib.depthfirstsort (fortran/sortlib.f90:26-72): same clusters, same order, members ascending."""
    from waterorderlib_b200.structureLibs import orderParam_lib as opl
    g = np.load(os.path.join(golden_dir, "clusters.npz"))
    for k in range(int(g["n_cases"])):
        got = opl.getClusters(g["mat%d" % k])
        sizes, members = g["sizes%d" % k], g["members%d" % k]
        assert [len(c) for c in got] == list(sizes)
        assert np.array_equal(np.concatenate(got), members)


def test_radialdistplane_matches_the_compiled_fortran(golden_dir):
    g = np.load(os.path.join(golden_dir, "rdfplane.npz"))
    for tag in ("a", "b"):
        rdf = wl.radialdistplane(g["pos1"], g["pos2"], float(g["binwidth_" + tag]), int(g["totbins_" + tag]), float(g["bulkdens"]), g["box"])
        assert rdf.shape == g["rdf_" + tag].shape and np.array_equal(rdf, g["rdf_" + tag]) and rdf.sum() > 100
    # an atom behind the plane's origin indexes bin <= 0 in the Fortran (out-of-bounds write): reported, not performed
    with pytest.raises(ValueError):
        wl.radialdistplane(g["pos1"], g["pos2"] - 6.0, 0.5, 40, float(g["bulkdens"]), g["box"])
    with pytest.raises(ValueError):
        wl.radialdistplane(g["pos1"][:2], g["pos2"], 0.5, 40, 1.0, g["box"])


def test_histogram2d_is_numpy_histogram2d():
    from waterorderlib_b200 import routines
    rng = np.random.default_rng(5)
    x = rng.integers(-2, 14, size=20000).astype(float)
    y = rng.random(20000) * 190.0 - 5.0
    y[:50] = 180.0          # the last edge belongs to the last bin
    y[50:100] = 0.0
    xe, ye = np.arange(-1.5, 13.5, 1), np.linspace(0, 180, 500)
    y[100:600] = ye[rng.integers(0, 500, size=500)]  # values exactly on edges go to the bin on their right
    want = np.histogram2d(x, y, bins=(xe, ye))[0]
    got = routines.histogram2d(x, y, xe, ye)
    assert np.array_equal(got.cpu().numpy(), want.astype(np.int64))
    routines.histogram2d(x, y, xe, ye, out=got)
    assert np.array_equal(got.cpu().numpy(), 2 * want.astype(np.int64))


def test_pair_histograms_on_a_grid_with_many_cells():
    """Pair-distance histograms where every axis has more than 3 cells -- the path with the float pass, the periodic
    image taken from cell adjacency, the warp's queue and the bin edges in distance^2 (csrc/wol_pairs.cu) -- against
    the oracle: shuffled atoms, a non-cubic box, atoms left outside the box by whole periods, all three modes."""
    pos, box = synth.water_box(12, sigma=0.3, seed=21)  # 13 824 waters, L = 74.5 A
    rng = np.random.default_rng(3)
    pos = pos[rng.permutation(pos.shape[0])]
    box = box * np.array([1.0, 1.07, 0.93])
    pos = pos * np.array([1.0, 1.07, 0.93])
    pos[::7] += box * rng.integers(-2, 3, size=(pos[::7].shape[0], 3))
    sub = rng.uniform(-0.5, 1.5, size=(300, 3)) * box
    for bw, nb in ((0.1, 120), (0.25, 60), (0.37, 9)):
        g = routines.pair_hist(1, pos, None, box, bw, nb).cpu().numpy()
        assert np.array_equal(g, port._pair_hist(1, pos, pos, box, bw, nb)), (bw, nb)
        g0 = routines.pair_hist(0, sub, pos, box, bw, nb).cpu().numpy()
        assert np.array_equal(g0, port._pair_hist(0, sub, pos, box, bw, nb)), (bw, nb)
        g2 = routines.pair_hist(2, sub, pos, box, bw, nb).cpu().numpy()
        assert np.array_equal(g2, port._pair_hist(2, sub, pos, box, bw, nb)), (bw, nb)
