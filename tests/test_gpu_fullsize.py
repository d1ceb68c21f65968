"""BASELINE.json's configurations at their FULL sizes (1M-water box, 4096 x 100 frames, 32768 waters + solute, 65536-water
slab): the oracle is too slow to redo them whole inside a test run, so they are checked through size-independent
properties -- conservation laws between the outputs, symmetry of the neighbour relation, permutation invariance,
additivity over frames, independence from batching -- and against the oracle on sampled molecules of the same inputs.
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import port  # noqa: E402  (the checker)
from waterorderlib_b200 import engine, routines, synth  # noqa: E402
from waterorderlib_b200.structureLibs import surface_library as sl  # noqa: E402


@pytest.fixture(scope="module")
def million():
    frames = [synth.water_box(50, sigma=s, seed=k) for k, s in ((0, 0.25), (1, 0.6))]
    pos = np.stack([f[0] for f in frames])
    box = np.stack([f[1] for f in frames])
    assert pos.shape == (2, 1_000_000, 3)
    pos_d = torch.from_numpy(pos).cuda()
    r = engine.q3b_frames(pos_d, box, hist_per_frame=True)
    return pos, box, pos_d, r


def test_cfg5_conservation_laws(million):
    pos, box, pos_d, r = million
    N = pos.shape[1]
    n3 = r.n3.to(torch.int64)
    assert torch.equal(r.ang_hist.sum(dim=1), (n3 * (n3 - 1) // 2).sum(dim=1))             # one bin per pair of neighbours
    st = r.frame_stats.cpu().numpy()
    assert np.array_equal(st[:, 2], [N, N]) and np.array_equal(st[:, 7], n3.sum(dim=1).cpu().numpy())
    assert np.array_equal(st[:, 6], r.ang_hist.sum(dim=1).cpu().numpy())
    assert np.allclose(st[:, 0], r.q.sum(dim=1).cpu().numpy(), rtol=1e-12)
    assert bool((n3.sum(dim=1) % 2 == 0).all())                                           # i ~ j  <=>  j ~ i
    nn = r.nn_idx.to(torch.int64)
    me = torch.arange(N, device="cuda")[None, :, None]
    assert bool((nn >= 0).all()) and bool((nn < N).all()) and bool((nn != me).all())
    s = torch.sort(nn, dim=-1).values
    assert bool((s[..., 1:] != s[..., :-1]).all())                                        # four distinct neighbours
    assert float(r.q.max()) <= 1.0 and float(r.q.min()) >= -3.0
    # (the brick path keeps 8 unit vectors per centre: the few liquid-like centres with more go to the large-capacity pass)
    assert r["n_overflow"] < 0.002 * 2 * N and 0 < r["n_widened"] < 0.05 * 2 * N


def test_cfg5_q_histogram_counts_every_value_in_range(million):
    _, _, _, r = million
    inside = ((r.q >= 0.0) & (r.q <= 1.0)).sum(dim=1)
    assert torch.equal(r.q_hist.sum(dim=1), inside)


def test_cfg5_sampled_centres_against_the_oracle(million):
    pos, box, pos_d, r = million
    rng = np.random.default_rng(0)
    for f in range(2):
        pick = np.sort(rng.choice(pos.shape[1], 96, replace=False))
        q, nn4, _ = port.order_param_q(pos[f][pick], pos[f], box[f], 0.0, 10.0)
        tb = port.three_body(pos[f][pick], pos[f], box[f], materialize=False)
        assert np.array_equal(r.nn_idx[f].cpu().numpy()[pick], nn4)
        assert np.allclose(r.q[f].cpu().numpy()[pick], q, rtol=1e-6, atol=1e-9)
        assert np.array_equal(r.n3[f].cpu().numpy()[pick], tb["numAngs"])
        sub = engine.q3b_frames(pos_d[f:f + 1], box[f], pos_d[f:f + 1, pick], do_q=False)
        assert np.array_equal(sub.ang_hist[0].cpu().numpy(), tb["hist"])


def test_cfg5_neighbour_relation_is_symmetric(million):
    pos, box, pos_d, r = million
    N = pos.shape[1]
    off, idx = routines.neighbors_csr(None, pos_d[1:2], box[1], 0.0, 3.413)
    counts = off[1:] - off[:-1]
    assert torch.equal(counts.to(torch.int32), r.n3[1])
    rows = torch.repeat_interleave(torch.arange(N, device="cuda"), counts)
    fwd = rows * N + idx.to(torch.int64)
    bwd = idx.to(torch.int64) * N + rows
    assert torch.equal(fwd, torch.sort(bwd).values)          # CSR is sorted by (row, column): the transpose is the same set


def test_cfg5_permutation_additivity_and_batching(million):
    pos, box, pos_d, r = million
    N = pos.shape[1]
    perm = torch.randperm(N, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    rp = engine.q3b_frames(pos_d[:1, perm], box[0], hist_per_frame=True)
    assert torch.equal(rp.ang_hist[0], r.ang_hist[0]) and torch.equal(rp.q_hist[0], r.q_hist[0])
    assert torch.equal(rp.q[0], r.q[0, perm]) and torch.equal(rp.n3[0], r.n3[0, perm])
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(N, device="cuda")
    assert torch.equal(perm[rp.nn_idx[0].to(torch.int64)].sort(dim=-1).values, r.nn_idx[0, perm].to(torch.int64).sort(dim=-1).values)
    pooled = engine.q3b_frames(pos_d, box)                                               # one histogram for both frames
    assert torch.equal(pooled.ang_hist[0], r.ang_hist.sum(dim=0)) and torch.equal(pooled.q_hist[0], r.q_hist.sum(dim=0))
    single = engine.q3b_frames(pos_d[1:2], box[1])                                        # frame 1 on its own
    assert torch.equal(single.q[0], r.q[1]) and torch.equal(single.nn_idx[0], r.nn_idx[1]) and torch.equal(single.ang_hist[0], r.ang_hist[1])


def test_cfg5_fp32_mode_within_its_bar(million):
    pos, box, pos_d, r = million
    r32 = engine.q3b_frames(pos_d[:1], box[0], precision="fp32")
    same = (r32.nn_idx[0] == r.nn_idx[0]).all(dim=-1)
    assert float(same.double().mean()) > 0.9999
    assert float((r32.q[0].double() - r.q[0]).abs()[same].max()) < 1e-4
    assert float((r32.n3[0] != r.n3[0]).double().mean()) < 1e-4


def test_cfg2_full_size_hbond_bookkeeping():
    F = 100
    O = np.stack([synth.water_box(8, sigma=0.25, seed=1234 + f)[0] for f in range(F)])
    box = synth.water_box(8, sigma=0.0, seed=0)[1]
    H = np.stack([synth.add_hydrogens(O[f], seed=1234 + f) for f in range(F)])
    O_d, H_d = torch.from_numpy(O).cuda(), torch.from_numpy(H).cuda()
    D_d = O_d.repeat_interleave(2, dim=1).contiguous()
    hb = routines.hbond_counts(O_d, D_d, H_d, box, 3.5, 120.0)
    acc, don = hb["acc_count"], hb["don_count"]
    assert acc.shape == (F, 4096) and don.shape == (F, 8192)
    assert torch.equal(acc.sum(dim=1), don.sum(dim=1))                                   # every bond has one acceptor and one hydrogen
    assert int(don.max()) <= 3 and 0.5 < float(acc.sum()) / (F * 4096) < 1.5   # random H orientations: about one accepted bond per water
    for f in (0, 57, 99):
        a_ref, d_ref = port.hbonds(O[f], np.repeat(O[f], 2, axis=0), H[f], box, 3.5, 120.0)
        assert np.array_equal(acc[f].cpu().numpy(), a_ref) and np.array_equal(don[f].cpu().numpy(), d_ref)
    r = engine.q3b_frames(O_d, box, hist_per_frame=True, want=("q", "n3", "ang_hist", "frame_stats"))
    for f in (3, 98):
        tb = port.three_body(O[f], O[f], box, materialize=False)
        assert np.array_equal(r.n3[f].cpu().numpy(), tb["numAngs"]) and np.array_equal(r.ang_hist[f].cpu().numpy(), tb["hist"])


def test_cfg3_full_size_shell_selection():
    pos, box = synth.water_box(16, sigma=0.4, seed=7)
    sol = synth.solute_grid(box)
    assert pos.shape[0] == 32768 and sol.shape[0] == 64
    mask = routines.shell_mask(sol, pos, box, 4.0)[0].cpu().numpy().astype(bool)
    assert np.array_equal(mask, port.shell_mask(sol, pos, box, 4.0).astype(bool)) and 10 < mask.sum() < 500
    idx = np.nonzero(mask)[0]
    shell = engine.q3b_frames(pos, box, pos[idx], do_q=False)
    tb = port.three_body(pos[idx], pos, box, materialize=False)
    assert np.array_equal(shell.ang_hist[0].cpu().numpy(), tb["hist"]) and np.array_equal(shell.n3[0].cpu().numpy(), tb["numAngs"])


def test_cfg4_full_size_slab_profile():
    pos, box, z_lo, z_hi = synth.slab_box(32, 32, 8, sigma=0.3, seed=11)
    assert pos.shape[0] == 65536
    gp, gn = synth.plane_interface(box, z_lo, z_hi, spacing=2.0)
    out = sl.depthBinnedQ(pos, box, gp, gn, binWidth=1.0, depthRange=(-30.0, 6.0))
    depth = out["depth"].cpu().numpy()
    # for flat faces with +-z normals the depth is the signed distance to the nearer face
    z = pos[:, 2]
    want = np.where(z > 0.5 * (z_lo + z_hi), z - gp[-1, 2], gp[0, 2] - z)
    assert np.allclose(depth, want, rtol=0, atol=1e-9)
    inside = (depth >= -30.0) & (depth < 6.0)
    assert out["count"].sum() == inside.sum() and out["numwater"] == int((depth <= 0.0).sum())
    pick = np.sort(np.random.default_rng(1).choice(65536, 1500, replace=False))
    _, _, _, d_ref = port.interface_water(pos[pick], gp, gn, 0.0, box)
    assert np.array_equal(depth[pick], d_ref)
    q_ref, _, _ = port.order_param_q(pos[pick[:200]], pos, box, 0.0, 10.0)
    assert np.allclose(out["q"].cpu().numpy()[pick[:200]], q_ref, rtol=1e-6, atol=1e-9)
    nz = out["count"] > 50
    assert out["q_mean"][nz][-1] < out["q_mean"][nz][: nz.sum() // 2].mean()             # the surface is less tetrahedral than the bulk
