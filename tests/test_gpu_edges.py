"""Edge cases of the fused path and the drop-in API, against the CPU oracle: empty and tiny inputs, boxes smaller
than the cell grid the fast path needs (cutoffs beyond L/2, where the exact round-half-away anint matters), atoms
exactly on cell boundaries and box faces, coincident atoms with and without a lower cutoff, per-frame boxes
(NPT), ragged sub-populations, and the error paths."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import port  # noqa: E402  (the checker)
from waterorderlib_b200 import engine, routines, synth  # noqa: E402
from waterorderlib_b200._capi import WolError  # noqa: E402
from waterorderlib_b200.structureLibs import water_properties as wp  # noqa: E402


def check(pos, box, sub=None, **cut):
    r = engine.q3b_frames(pos, box, sub, low3=cut.get("low3", 0.0), high3=cut.get("high3", 3.413), lowq=cut.get("lowq", 0.0),
                          highq=cut.get("highq", 10.0))
    torch.cuda.synchronize()
    c = pos if sub is None else sub
    q, nn4, _ = port.order_param_q(c, pos, box, cut.get("lowq", 0.0), cut.get("highq", 10.0))
    tb = port.three_body(c, pos, box, cut.get("low3", 0.0), cut.get("high3", 3.413), materialize=False)
    assert np.array_equal(r.nn_idx.cpu().numpy()[0], nn4)
    assert np.array_equal(r.n3.cpu().numpy()[0], tb["numAngs"])
    assert np.array_equal(r.ang_hist.cpu().numpy()[0], tb["hist"])
    assert np.allclose(r.q.cpu().numpy()[0], q, rtol=1e-6, atol=1e-9)
    return r


def test_empty_and_tiny_inputs():
    box = np.array([20.0, 20.0, 20.0])
    assert wp.getOrderParamq(np.zeros((0, 3)), np.zeros((0, 3)), box).shape == (0,)
    one = np.array([[1.0, 2.0, 3.0]])
    assert wp.getOrderParamq(one, one, box)[0] == 0.0  # no neighbours: q = 0 (water_properties.py:358)
    ang, num = wp.getCosAngs(one, one, box)
    assert ang.shape == (0,) and num[0] == 0.0
    sub = np.array([[5.0, 5.0, 5.0], [15.0, 15.0, 15.0]])
    assert np.array_equal(wp.getOrderParamq(sub, np.zeros((0, 3)), box), np.zeros(2))
    two = np.array([[1.0, 1.0, 1.0], [3.5, 1.0, 1.0]])
    check(two, box)
    r = engine.q3b_frames(np.zeros((2, 0, 3)), box)  # frames without atoms
    assert r.ang_hist.sum().item() == 0


@pytest.mark.parametrize("L,n,highq", [(9.0, 30, 10.0), (7.5, 12, 10.0), (12.0, 60, 6.0), (10.5, 25, 5.25)])
def test_small_boxes_cutoff_beyond_half_box(L, n, highq):
    """nc < 4 cells per axis: generic path; highCut > L/2 keeps only the nearest image of every atom, like the
    reference's min-image loops; pairs exactly L/2 apart exercise anint's round-half-away."""
    rng = np.random.default_rng(int(L * 10) + n)
    box = np.array([L, L * 1.1, L * 0.9])
    pos = (rng.random((n, 3)) * box).astype(np.float32).astype(np.float64)
    pos[1] = pos[0] + np.array([L / 2, 0.0, 0.0])  # exactly half a box apart along x
    check(pos, box, highq=highq, high3=min(3.413, 0.45 * L))


def test_atoms_on_cell_boundaries_and_faces():
    box = np.array([28.0, 28.0, 28.0])  # r_cell 3.5 -> exactly 8 cells of 3.5 A
    g = np.arange(8) * 3.5
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)  # every atom on a cell corner
    rng = np.random.default_rng(2)
    extra = rng.random((300, 3)) * box
    extra[:50, 0] = 28.0   # exactly on the upper face
    extra[50:100, 1] = 0.0
    extra[100:120] = -extra[100:120]  # outside the box, negative
    pos = np.concatenate([pos, extra]).astype(np.float32).astype(np.float64)
    check(pos, box, high3=3.5, highq=7.0)


def test_coincident_atoms_and_lower_cutoffs():
    rng = np.random.default_rng(8)
    box = np.array([22.0, 25.0, 21.0])
    pos = (rng.random((700, 3)) * box).astype(np.float32).astype(np.float64)
    pos[10] = pos[3]; pos[11] = pos[3]  # triple coincidence: excluded at lowCut = 0, never a neighbour of itself
    pos[20] = pos[4] + box               # the same point one box over: distance 0 under the minimum image
    check(pos, box)
    check(pos, box, low3=2.0, high3=3.6, lowq=2.5, highq=8.0)


def test_per_frame_boxes_npt():
    frames, boxes = [], []
    for f, scale in enumerate((0.98, 1.0, 1.03)):
        p, b = synth.water_box(6, sigma=0.4, seed=60 + f)
        frames.append((p * scale).astype(np.float32).astype(np.float64))
        boxes.append(b * scale)
    pos, box = np.stack(frames), np.stack(boxes)
    r = engine.q3b_frames(pos, box, hist_per_frame=True)
    torch.cuda.synchronize()
    for f in range(3):
        q, nn4, _ = port.order_param_q(pos[f], pos[f], box[f])
        tb = port.three_body(pos[f], pos[f], box[f], materialize=False)
        assert np.array_equal(r.nn_idx.cpu().numpy()[f], nn4) and np.array_equal(r.n3.cpu().numpy()[f], tb["numAngs"])
        assert np.array_equal(r.ang_hist.cpu().numpy()[f], tb["hist"])
        assert np.allclose(r.q.cpu().numpy()[f], q, rtol=1e-6, atol=1e-9)
        assert np.array_equal(r.q_hist.cpu().numpy()[f], np.histogram(q, bins=500, range=[0.0, 1.0])[0])


def test_ragged_subpopulations_same_workspace():
    pos, box = synth.water_box(5, sigma=0.5, seed=3)
    rng = np.random.default_rng(3)
    ws = engine.Workspace(torch.device("cuda"))
    for m in (1, 7, 130, 999, 2):
        sub = np.concatenate([pos[rng.choice(len(pos), m // 2, replace=False)], rng.random((m - m // 2, 3)) * box])
        r = engine.q3b_frames(pos, box, sub, workspace=ws, highq=7.0)
        q, nn4, _ = port.order_param_q(sub, pos, box, 0.0, 7.0)
        assert np.array_equal(r.nn_idx.cpu().numpy()[0], nn4) and np.allclose(r.q.cpu().numpy()[0], q, rtol=1e-6, atol=1e-9)


def test_error_paths():
    pos, box = synth.water_box(3, sigma=0.3, seed=1)
    with pytest.raises(ValueError):
        engine.q3b_frames(pos, np.array([18.0, 0.0, 18.0]))      # an empty axis (negative = non-periodic is supported, see below)
    with pytest.raises(ValueError):
        engine.q3b_frames(pos, box, r_cell=2.0)                  # three-body cutoff beyond the planned cell edge
    with pytest.raises(ValueError):
        engine.q3b_frames(pos[:, :2], box)
    with pytest.raises(ValueError):
        wp.getOrderParamq(pos, pos, np.array([1.0, 2.0]))
    with pytest.raises(ValueError):
        routines.hbond_counts(pos, pos, pos[:-1], box)           # donors and hydrogens differ in number (waterlib.f90:1171)
    # more neighbours than the large-capacity path holds -> WOL_ERR_CAPACITY, not a silent truncation
    blob = np.random.default_rng(0).random((1500, 3)) * 2.0 + 9.0
    with pytest.raises(WolError):
        engine.q3b_frames(np.concatenate([blob, pos]), box, high3=3.4)


def test_one_workspace_across_batch_shapes():
    """A driver reuses its workspace for a shorter last batch and for other system sizes: the status counters sit at a
    fixed offset, so no stale bytes are mistaken for an overflow flag, and a real overflow is still reported later."""
    ws = engine.Workspace(torch.device("cuda"))
    pos, box = synth.trajectory(5, 5, sigma=0.4, seed0=70)
    for sl in (slice(0, 5), slice(0, 2), slice(3, 4)):
        r = engine.q3b_frames(pos[sl], box[sl], workspace=ws, hist_per_frame=True)
        for k, f in enumerate(range(*sl.indices(5))):
            tb = port.three_body(pos[f], pos[f], box[f], materialize=False)
            assert np.array_equal(r.ang_hist.cpu().numpy()[k], tb["hist"])
    small, sbox = synth.water_box(3, sigma=0.3, seed=1)
    engine.q3b_frames(small, sbox, workspace=ws)          # other N, other grid: same workspace
    big, bbox = synth.trajectory(7, 3, sigma=0.3, seed0=5)
    engine.q3b_frames(big, bbox, workspace=ws)            # grows the buffer: counters are carried over
    blob = np.concatenate([np.random.default_rng(0).random((1500, 3)) * 2.0 + 9.0, small])
    engine.q3b_frames(blob, sbox, workspace=ws, check_status=False, high3=3.4)   # overflows, not checked yet
    with pytest.raises(WolError):
        engine.q3b_frames(small, sbox, workspace=ws)      # the sticky flag surfaces at the next checked call
    engine.q3b_frames(small, sbox, workspace=ws)          # and is cleared by having been read


def test_workspace_needs_only_zero_counters():
    """Only the first 256 bytes of a workspace (the counters) must start at zero: everything behind them is rebuilt by
    every call, so a buffer full of stale bytes gives the same answers on every path (fast sweep, widened search,
    overflow pass, H-bonds)."""
    dev = torch.device("cuda")
    pos, box = synth.trajectory(5, 3, sigma=0.6, seed0=21)   # liquid-like: widened centres and some list overflows
    ref = engine.q3b_frames(pos, box)
    ws = engine.Workspace(dev)
    engine.q3b_frames(pos, box, workspace=ws)
    ws.buf[512:].fill_(0xAB)
    r = engine.q3b_frames(pos, box, workspace=ws)
    assert torch.equal(r.q, ref.q) and torch.equal(r.nn_idx, ref.nn_idx) and torch.equal(r.ang_hist, ref.ang_hist)
    assert r["n_widened"] == ref["n_widened"] > 0
    ws.buf[512:].fill_(0xFF)
    hyd = synth.add_hydrogens(pos[0], seed=4)
    don = np.repeat(pos[0], 2, axis=0)
    cells = routines.CellList(don, box[0], 3.5, workspace=ws)
    hb_poisoned = routines.hbond_counts(pos[0], don, hyd, box[0], 3.5, 120.0, cells=cells)
    hb = routines.hbond_counts(pos[0], don, hyd, box[0], 3.5, 120.0)
    assert torch.equal(hb_poisoned["acc_count"], hb["acc_count"]) and torch.equal(hb_poisoned["don_count"], hb["don_count"])


def test_ragged_centres_batched_with_reused_cells():
    """n_valid: frame f evaluates only its first n_valid[f] centres; the cell list of a previous call on the same batch
    is reused.  Per-frame outputs must equal the oracle's on exactly the valid centres; padded slots stay untouched."""
    pos, box = synth.trajectory(5, 4, sigma=0.5, seed0=90)
    rng = np.random.default_rng(1)
    counts = np.array([17, 0, 63, 5], dtype=np.int32)
    cen = np.zeros((4, 63, 3))
    for f, c in enumerate(counts):
        cen[f, :c] = pos[f][rng.choice(pos.shape[1], c, replace=False)] + rng.normal(0, 0.2, size=(c, 3))
    cen = cen.astype(np.float32).astype(np.float64)
    ws = engine.Workspace(torch.device("cuda"))
    engine.q3b_frames(pos, box, workspace=ws)  # builds the cell list for this batch
    q = torch.full((4, 63), -7.0, dtype=torch.float64, device="cuda")
    n3 = torch.full((4, 63), -7, dtype=torch.int32, device="cuda")
    r = engine.q3b_frames(pos, box, cen, n_valid=counts, reuse_cells=True, workspace=ws, hist_per_frame=True,
                          out={"q": q, "n3": n3}, highq=8.0)
    for f, c in enumerate(counts):
        qr, nn4, _ = port.order_param_q(cen[f, :c], pos[f], box[f], 0.0, 8.0)
        tb = port.three_body(cen[f, :c], pos[f], box[f], materialize=False)
        assert np.allclose(r.q.cpu().numpy()[f, :c], qr, rtol=1e-6, atol=1e-9) and np.all(r.q.cpu().numpy()[f, c:] == -7.0)
        assert np.array_equal(r.n3.cpu().numpy()[f, :c], tb["numAngs"]) and np.all(r.n3.cpu().numpy()[f, c:] == -7)
        assert np.array_equal(r.nn_idx.cpu().numpy()[f, :c], nn4)
        assert np.array_equal(r.ang_hist.cpu().numpy()[f], tb["hist"])
        assert r.frame_stats[f, 2].item() == c
    with pytest.raises(ValueError):
        engine.q3b_frames(pos, box, cen, reuse_cells=True, workspace=engine.Workspace(torch.device("cuda")))


def test_reentrant_from_several_host_threads():
    """The library keeps no global state beyond thread-local error / launch bookkeeping: host threads working on their
    own streams and workspaces at the same time get the answers of a serial run (ctypes releases the GIL in the calls)."""
    import threading
    cases = [synth.water_box(m, sigma=s, seed=k) for k, (m, s) in enumerate(((5, 0.3), (6, 0.5), (4, 0.6), (7, 0.25)))]
    serial = [engine.q3b_frames(p, b) for p, b in cases]
    got, errors = [None] * len(cases), []

    def work(k):
        try:
            stream = torch.cuda.Stream()
            ws = engine.Workspace(torch.device("cuda"))
            with torch.cuda.stream(stream):
                for _ in range(5):
                    r = engine.q3b_frames(cases[k][0], cases[k][1], workspace=ws)
                stream.synchronize()
            got[k] = r
        except Exception as e:  # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=work, args=(k,)) for k in range(len(cases))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors
    for r, s in zip(got, serial):
        assert torch.equal(r.q, s.q) and torch.equal(r.nn_idx, s.nn_idx) and torch.equal(r.ang_hist, s.ang_hist)
        assert torch.equal(r.n3, s.n3)


def test_generic_kernels_agree_with_the_fast_path_at_size(monkeypatch):
    """The group-per-centre kernels normally serve only boxes below four cells per edge; forced on a larger liquid-like
    box (WOL_NO_TPC, read per call) they must reproduce the thread-per-centre path: same neighbours, counts and angle bins, q to rounding (fp64); within the fp32 bar in float mode."""
    pos, box = synth.trajectory(9, 2, sigma=0.6, seed0=31)     # 5832 waters per frame, widened and overflowing centres
    sub = pos[:, ::7]
    for prec in ("fp64", "fp32"):
        fast = engine.q3b_frames(pos, box, precision=prec)
        fast_sub = engine.q3b_frames(pos, box, sub, precision=prec, highq=7.0)
        monkeypatch.setenv("WOL_NO_TPC", "1")
        slow = engine.q3b_frames(pos, box, precision=prec)
        slow_sub = engine.q3b_frames(pos, box, sub, precision=prec, highq=7.0)
        monkeypatch.delenv("WOL_NO_TPC")
        for a, b in ((fast, slow), (fast_sub, slow_sub)):
            if prec == "fp64":
                assert torch.equal(a.nn_idx, b.nn_idx) and torch.equal(a.n3, b.n3) and torch.equal(a.ang_hist, b.ang_hist)
                d = (a.q - b.q).abs().max().item()
                assert d < 1e-12, d
                assert (a.q_hist - b.q_hist).abs().sum().item() <= 2
            else:
                # float arithmetic in a different order (wrapped vs raw coordinates): decisions agree except on the boundary
                same = (a.nn_idx == b.nn_idx).all(dim=-1)
                assert same.double().mean().item() > 0.999 and (a.n3 == b.n3).double().mean().item() > 0.999
                assert torch.allclose(a.q[same], b.q[same], rtol=0, atol=1e-4)
                assert (a.ang_hist - b.ang_hist).abs().sum().item() < 2e-3 * a.ang_hist.sum().item()


# ---- non-periodic axes: the reference's negative box edges (fortran/waterlib.f90:41, :840) -------------------------

def _slab(seed=3):
    pos, box, _, _ = synth.slab_box(6, 6, 3, sigma=0.4, seed=seed)  # 864 waters, vacuum above and below in z
    return pos, box


@pytest.mark.parametrize("open_axes", [(2,), (0, 2), (0, 1, 2)])
def test_open_axes_fused_path_vs_oracle(open_axes):
    """q, 4-NN, neighbour counts and angle bins with negative edges = the oracle's (and the reference Fortran's, which
    the oracle is pinned on) with the same negative edges; shifted far from the origin and in float32 storage too."""
    pos, box = _slab()
    b = box.copy()
    for k in open_axes:
        b[k] = -1.0
    check(pos, b)
    check(pos + np.array([1000.0, -250.0, 3000.0]), b)
    r32 = engine.q3b_frames(pos.astype(np.float32), b)
    q, nn4, _ = port.order_param_q(pos.astype(np.float32).astype(np.float64), pos.astype(np.float32).astype(np.float64), b)
    assert np.array_equal(r32.nn_idx.cpu().numpy()[0], nn4)
    # an open axis is not a periodic one: the atoms at the two faces of the slab must not see each other
    per = engine.q3b_frames(pos, np.array([box[0], box[1], pos[:, 2].max() - pos[:, 2].min() + 1.5]))
    opn = engine.q3b_frames(pos, np.array([box[0], box[1], -1.0]))
    assert int(per.n3.sum()) > int(opn.n3.sum())


def test_open_axes_sub_population_outside_the_atoms_extent():
    pos, box = _slab(seed=5)
    b = np.array([box[0], -box[1], -1.0])
    rng = np.random.default_rng(2)
    sub = np.concatenate([pos[rng.choice(pos.shape[0], 40, replace=False)] + rng.normal(0, 0.3, (40, 3)),
                          pos[:5] + np.array([0.0, 0.0, 60.0]),          # far above the slab: no neighbours at all
                          pos[:5] + np.array([0.0, -45.0, 0.0])])        # outside along the other open axis
    check(pos, b, sub=sub, highq=8.0)
    ang, num = wp.getCosAngs(sub, pos, b)
    ang_o, num_o = port.getCosAngs(sub, pos, b)
    assert np.array_equal(num, num_o) and np.allclose(ang, ang_o, rtol=1e-12, atol=1e-9)


def test_open_axes_neighbour_lists_hbonds_and_shell():
    pos, box = _slab(seed=7)
    b = np.array([box[0], box[1], -box[2]])
    off, idx = routines.neighbors_csr(None, pos, b, 0.0, 3.5)
    mat = port.neighbor_matrix(pos, pos, b, 0.0, 3.5)
    off, idx = off.cpu().numpy(), idx.cpu().numpy()
    for i in range(0, pos.shape[0], 37):
        assert np.array_equal(idx[off[i]:off[i + 1]], np.nonzero(mat[i])[0])
    assert off[-1] == mat.sum()
    H = synth.add_hydrogens(pos, seed=7)
    r = routines.hbond_counts(pos, np.repeat(pos, 2, axis=0), H, b, 3.5, 120.0)
    a, d = port.hbonds(pos, np.repeat(pos, 2, axis=0), H, b, 3.5, 120.0)
    assert np.array_equal(r["acc_count"].cpu().numpy()[0], a) and np.array_equal(r["don_count"].cpu().numpy()[0], d)
    sol = np.array([[5.0, 5.0, pos[:, 2].min() - 2.0], [20.0, 12.0, pos[:, 2].max() + 1.0], [9.0, 30.0, pos[:, 2].mean()]])
    m = routines.shell_mask(sol, pos, b, 4.0).cpu().numpy()[0]
    assert np.array_equal(m.astype(bool), port.shell_mask(sol, pos, b, 4.0).astype(bool))


def test_open_axes_lsi_psi_rdf_histrr3b():
    pos, box = _slab(seed=9)
    b = np.array([box[0], box[1], -1.0])
    rng = np.random.default_rng(4)
    sub = pos[rng.choice(pos.shape[0], 60, replace=False)]
    for c in (pos, sub):
        vals, num = wp.getLSI(c, pos, b)
        ref_vals, ref_num = port.getLSI(c, pos, b)
        assert np.array_equal(num, ref_num) and np.allclose(vals, ref_vals, rtol=1e-10, atol=1e-12)
    psi = wp.getOrderParamPsi(sub, pos, b, 0.0, 7.0)
    assert np.allclose(psi, port.getOrderParamPsi(sub, pos, b, 0.0, 7.0), rtol=1e-9, atol=1e-12, equal_nan=True)
    g = routines.pair_hist(1, pos, None, b, 0.1, 80).cpu().numpy()
    assert np.array_equal(g, port._pair_hist(1, pos, pos, b, 0.1, 80))
    g2 = routines.pair_hist(0, sub, pos, b, 0.1, 80).cpu().numpy()
    assert np.array_equal(g2, port._pair_hist(0, sub, pos, b, 0.1, 80))
    h = routines.histrr3b(pos[:300], b, 0.25, 16, 5.0, 36).cpu().numpy()
    assert np.array_equal(h, port.histrr3b(pos[:300], b, 0.25, 16, 5.0, 36))


def test_open_axes_limits_are_reported():
    pos, box = _slab()
    b = np.array([box[0], box[1], -1.0])
    with pytest.raises(ValueError):
        routines.willard_density(pos, b, 2.4, points=pos[:10])   # needs a periodic (or explicitly bounded) grid
    with pytest.raises((ValueError, WolError)):
        engine.q3b_frames(pos, np.array([box[0], 0.0, box[2]]))  # an empty axis is an error, not "open"


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_odd_atom_counts_over_several_frames(dtype):
    """Frames whose first atom is not 16-byte aligned in the batch (odd atom counts: the cell build's scalar staging path)
    and tiles that end in the middle of a 1024-atom block."""
    rng = np.random.default_rng(11)
    box = np.array([31.0, 29.5, 33.0])
    n = 1027
    pos = (rng.random((3, n, 3)) * box).astype(np.float32).astype(np.float64)
    r = engine.q3b_frames(pos.astype(dtype), box, hist_per_frame=True, highq=8.0)
    torch.cuda.synchronize()
    for f in range(3):
        q, nn4, _ = port.order_param_q(pos[f], pos[f], box, 0.0, 8.0)
        tb = port.three_body(pos[f], pos[f], box, materialize=False)
        assert np.array_equal(r.nn_idx.cpu().numpy()[f], nn4)
        assert np.array_equal(r.n3.cpu().numpy()[f], tb["numAngs"])
        assert np.array_equal(r.ang_hist.cpu().numpy()[f], tb["hist"])
        assert np.allclose(r.q.cpu().numpy()[f], q, rtol=1e-6, atol=1e-9)


def test_effective_box_through_the_c_abi():
    """wol_effective_box called as a C host would: positive edges copied, negative edges replaced by
    2 (extent + reach) + 1 with the extent taken over atoms AND centres of each frame; bad arguments refused."""
    import ctypes
    from waterorderlib_b200._capi import lib
    L = lib()
    rng = np.random.default_rng(3)
    F, N, M = 3, 500, 40
    pos = rng.normal(0.0, 7.0, (F, N, 3))
    cen = rng.normal(50.0, 1.0, (F, M, 3))
    box = np.array([[30.0, -1.0, 25.0], [-2.0, -2.0, -2.0], [30.0, 31.0, 32.0]])
    out = np.zeros_like(box)
    pos_d, cen_d = torch.from_numpy(pos).cuda(), torch.from_numpy(cen.astype(np.float32)).cuda()
    scratch = torch.empty(F * 6, dtype=torch.int64, device="cuda")
    vp = ctypes.c_void_p
    rc = L.wol_effective_box(vp(pos_d.data_ptr()), 0, F, N, vp(cen_d.data_ptr()), 1, M, box.ctypes.data_as(vp), 10.0,
                             vp(scratch.data_ptr()), out.ctypes.data_as(vp), None)
    assert rc == 0, L.wol_last_error()
    allp = np.concatenate([pos, cen.astype(np.float32).astype(np.float64)], axis=1)
    ext = allp.max(axis=1) - allp.min(axis=1)
    want = np.where(box > 0, box, 2.0 * (ext + 10.0) + 1.0)
    assert np.array_equal(out, want)
    # nothing open: no device work, no scratch needed
    rc = L.wol_effective_box(vp(pos_d.data_ptr()), 0, 1, N, None, 0, 0, box[2:].ctypes.data_as(vp), 10.0, None,
                             out[2:].ctypes.data_as(vp), None)
    assert rc == 0 and np.array_equal(out[2], box[2])
    bad = np.array([[30.0, 0.0, 25.0]])
    assert L.wol_effective_box(vp(pos_d.data_ptr()), 0, 1, N, None, 0, 0, bad.ctypes.data_as(vp), 10.0, vp(scratch.data_ptr()),
                               out[:1].ctypes.data_as(vp), None) < 0
    assert L.wol_effective_box(vp(pos_d.data_ptr()), 0, 1, N, None, 0, 0, box[:1].ctypes.data_as(vp), 10.0, None,
                               out[:1].ctypes.data_as(vp), None) < 0  # open axis without scratch
