"""Parity of the CUDA fused q / three-body path (through the C ABI) against the golden fixtures (outputs of
the reference's compiled Fortran + Python) and against the CPU oracle on seeded inputs.

fp64 mode: neighbour indices, counts and histogram bins bit-exact; q within 1e-6 relative (in fact ~1e-15).
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import port  # noqa: E402  (the checker)
from waterorderlib_b200 import engine, synth  # noqa: E402

Q_RTOL = 1e-6  # north_star tolerance for fp64 mode
CASES = ["cfg1_n512_ice", "cfg1_n512_liq", "lattice_n216", "random_n160_noncubic", "subpop_n512_m97",
         "cfg2_n4096_frame0"]


def run(sub, pos, box, same, **kw):
    r = engine.q3b_frames(pos, box, None if same else sub, **kw)
    torch.cuda.synchronize()
    return r


def assert_q_close(q, q_ref):
    assert np.allclose(q, q_ref, rtol=Q_RTOL, atol=1e-9), float(np.abs(q - q_ref).max())


@pytest.mark.parametrize("name", CASES)
def test_against_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    same = np.array_equal(g["sub"], g["pos"])
    r = run(g["sub"], g["pos"], g["box"], same, low3=float(g["low3"]), high3=float(g["high3"]),
            lowq=float(g["lowq"]), highq=float(g["highq"]))
    assert np.array_equal(r.n3.cpu().numpy()[0], g["n3"])
    assert np.array_equal(r.ang_hist.cpu().numpy()[0], g["hist"])
    assert np.array_equal(r.nn_idx.cpu().numpy()[0], g["nn4"])
    assert_q_close(r.q.cpu().numpy()[0], g["q"])
    assert np.array_equal(r.q_hist.cpu().numpy()[0], np.histogram(g["q"], bins=500, range=[0.0, 1.0])[0])
    st = r.frame_stats.cpu().numpy()[0]
    assert st[6] == int(g["n_angles"]) and st[7] == g["n3"].sum() and st[2] == g["q"].size
    assert st[3] == round(float(g["fracTet"]) * int(g["n_angles"]))
    assert abs(st[4] / st[3] - float(g["avgCos"])) < 1e-12
    assert abs(st[0] - g["q"].sum()) < 1e-9 * max(1.0, abs(g["q"].sum()))


@pytest.mark.parametrize("m,sigma,seed", [(6, 0.25, 11), (6, 0.6, 12), (10, 0.45, 13)])
def test_against_oracle_batched(m, sigma, seed):
    """Several frames in one call, per-frame histograms, float32 storage of (float32-representable) inputs."""
    pos, box = synth.trajectory(m, 3, sigma=sigma, seed0=seed)
    r = run(None, pos.astype(np.float32), box, True, hist_per_frame=True)
    for f in range(3):
        q, nn4, _ = port.order_param_q(pos[f], pos[f], box[f])
        tb = port.three_body(pos[f], pos[f], box[f], materialize=False)
        assert np.array_equal(r.nn_idx.cpu().numpy()[f], nn4)
        assert np.array_equal(r.n3.cpu().numpy()[f], tb["numAngs"])
        assert np.array_equal(r.ang_hist.cpu().numpy()[f], tb["hist"])
        assert_q_close(r.q.cpu().numpy()[f], q)


def test_unwrapped_and_noncubic_vs_oracle():
    rng = np.random.default_rng(17)
    box = np.array([31.0, 24.5, 40.25])
    n = 1200
    pos = rng.random((n, 3)) * box + box * rng.integers(-3, 4, size=(n, 3))
    r = run(None, pos, box, True, high3=3.7, highq=9.0)
    q, nn4, _ = port.order_param_q(pos, pos, box, 0.0, 9.0)
    tb = port.three_body(pos, pos, box, 0.0, 3.7, materialize=False)
    assert np.array_equal(r.nn_idx.cpu().numpy()[0], nn4)
    assert np.array_equal(r.n3.cpu().numpy()[0], tb["numAngs"])
    assert np.array_equal(r.ang_hist.cpu().numpy()[0], tb["hist"])
    assert_q_close(r.q.cpu().numpy()[0], q)
    assert r.n_widened > 0  # dilute random gas: the widened search is exercised


def test_dense_cluster_overflow_path():
    """More neighbours than the fast list holds -> large-capacity pass; still exact."""
    rng = np.random.default_rng(23)
    box = np.array([30.0, 30.0, 30.0])
    pos = np.concatenate([rng.random((300, 3)) * 4.0 + 10.0, rng.random((500, 3)) * box])
    r = run(None, pos, box, True, high3=3.413, highq=8.0)
    q, nn4, _ = port.order_param_q(pos, pos, box, 0.0, 8.0)
    tb = port.three_body(pos, pos, box, 0.0, 3.413, materialize=False)
    assert r.n_overflow > 0
    assert np.array_equal(r.n3.cpu().numpy()[0], tb["numAngs"])
    assert np.array_equal(r.ang_hist.cpu().numpy()[0], tb["hist"])
    assert np.array_equal(r.nn_idx.cpu().numpy()[0], nn4)
    assert_q_close(r.q.cpu().numpy()[0], q)


def test_collinear_minus_180_quirk():
    g = np.arange(4) * 3.0
    pos = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    box = np.array([12.0, 12.0, 12.0])
    r = run(None, pos, box, True, high3=3.2, highq=5.0)
    tb = port.three_body(pos, pos, box, 0.0, 3.2, materialize=False)
    assert np.array_equal(r.ang_hist.cpu().numpy()[0], tb["hist"])
    assert r.ang_hist.sum().item() == 64 * 12 and r.frame_stats[0, 6].item() == 64 * 15
    q, nn4, _ = port.order_param_q(pos, pos, box, 0.0, 5.0)
    assert np.array_equal(r.nn_idx.cpu().numpy()[0], nn4)  # six equidistant neighbours: ties by index
    assert_q_close(r.q.cpu().numpy()[0], q)


def test_few_neighbours_padding_and_empty():
    box = np.array([40.0, 40.0, 40.0])
    pos = np.array([[1.0, 1.0, 1.0], [3.0, 1.0, 1.0],            # pair: 1 neighbour each
                    [20.0, 20.0, 20.0], [22.0, 20.0, 20.0], [20.0, 22.5, 20.0],  # triple
                    [10.0, 30.0, 5.0]])                            # isolated
    r = run(None, pos, box, True, highq=5.0)
    q, nn4, _ = port.order_param_q(pos, pos, box, 0.0, 5.0)
    assert np.array_equal(r.nn_idx.cpu().numpy()[0], nn4)
    assert_q_close(r.q.cpu().numpy()[0], q)
    assert r.q[0, 5].item() == 0.0 and abs(r.q[0, 0].item()) < 1e-12


def test_graphed_call_replays_and_matches_the_plain_call():
    """engine.q3b_frames_graphed: one captured launch sequence per signature, replayed on new inputs of the same shape;
    every output equals the plain call's, histograms start from zero on every replay."""
    from waterorderlib_b200 import engine, synth
    box = None
    frames = []
    for seed in (1, 2, 3):
        p, box = synth.water_box(4, sigma=0.3 + 0.1 * seed, seed=seed)
        frames.append(p)
    n0 = len(engine._GRAPHS)
    for rep in range(2):
        for p in frames:
            g = engine.q3b_frames_graphed(p, box)
            r = engine.q3b_frames(p, box)
            assert g["graph"] and g["launches"] == r["launches"]
            for k in ("q", "nn_idx", "n3", "ang_hist", "q_hist"):
                assert torch.equal(g[k], r[k]), k
            assert torch.allclose(g["frame_stats"], r["frame_stats"], rtol=1e-13, atol=0)
            assert g["n_widened"] == r["n_widened"] and g["n_overflow"] == r["n_overflow"]
    assert len(engine._GRAPHS) == n0 + 1  # same signature: captured once
    sub = frames[0][:37] + 0.05
    g = engine.q3b_frames_graphed(frames[1], box, sub, do_q=False, want=("n3", "ang_hist"))
    r = engine.q3b_frames(frames[1], box, sub, do_q=False, want=("n3", "ang_hist"))
    assert torch.equal(g["n3"], r["n3"]) and torch.equal(g["ang_hist"], r["ang_hist"])
    # another box = another signature (cutoff margins derived from the box are baked into the captured kernels)
    g2 = engine.q3b_frames_graphed(frames[0] * 1.01, box * 1.01, want=("q",))
    assert torch.equal(g2["q"], engine.q3b_frames(frames[0] * 1.01, box * 1.01, want=("q",))["q"])


@pytest.mark.parametrize("sigma,kw", [(0.25, {}), (0.6, {}), (0.45, {"do_q": False}), (0.45, {"do_3body": False}),
                                      (0.45, {"highq": 3.0, "lowq": 1.0, "low3": 2.5})])
def test_brick_kernels_agree_with_the_thread_per_centre_path(monkeypatch, sigma, kw):
    """The three fp64 kernels of K2 on the same frames: thread-per-centre (WOL_BRICK=0), brick (1) and the opt-in
    warp-specialised brick kernel (3) -- indices, counts and histogram bins identical, q to the last bits; frame 0
    against the oracle."""
    pos, box = synth.trajectory(12, 3, sigma=sigma, seed0=31)  # 3 x 13 824 waters
    out = {}
    for mode in ("0", "1", "3"):
        monkeypatch.setenv("WOL_BRICK", mode)
        out[mode] = run(None, pos, box, True, **kw)
    a = out["0"]
    for mode in ("1", "3"):
        b = out[mode]
        for k in ("nn_idx", "n3", "ang_hist", "q_hist"):
            if k in a and a[k] is not None:
                assert torch.equal(a[k], b[k]), (mode, k)
        if "q" in a and a["q"] is not None:
            assert float((a["q"] - b["q"]).abs().max()) < 1e-12
        assert torch.allclose(a["frame_stats"], b["frame_stats"], rtol=1e-12, atol=1e-9)
    if kw.get("do_3body", True):
        tb = port.three_body(pos[0], pos[0], box[0], kw.get("low3", 0.0), kw.get("high3", 3.413), materialize=False)
        assert np.array_equal(out["3"].n3.cpu().numpy()[0], tb["numAngs"])
    if kw.get("do_q", True):
        q, nn4, _ = port.order_param_q(pos[0], pos[0], box[0], kw.get("lowq", 0.0), kw.get("highq", 10.0))
        assert np.array_equal(out["3"].nn_idx.cpu().numpy()[0], nn4)
        assert_q_close(out["3"].q.cpu().numpy()[0], q)


@pytest.mark.parametrize("sigma", [0.6, 0.9])
def test_widened_search_kernels_agree(monkeypatch, sigma):
    """The widened 4-NN search has a warp-per-centre and a thread-per-centre kernel, and the device-side count of queued
    centres picks one (wol_q3b_tpc.cu: q3b_tpc_widen_launch).  Forced either way (WOL_WIDEN_THREAD=0 / 1) and left to
    the count, the results are the same -- and frame 0 is the oracle's."""
    pos, box = synth.trajectory(12, 2, sigma=sigma, seed0=77)  # 2 x 13 824 waters, hundreds of widened centres each
    out = {}
    for mode in ("0", "1", None):
        if mode is None:
            monkeypatch.delenv("WOL_WIDEN_THREAD", raising=False)
        else:
            monkeypatch.setenv("WOL_WIDEN_THREAD", mode)
        out[mode] = run(None, pos, box, True)
    a = out["0"]
    assert a["n_widened"] > 50
    for mode in ("1", None):
        b = out[mode]
        for k in ("nn_idx", "n3", "ang_hist", "q_hist"):
            assert torch.equal(a[k], b[k]), (mode, k)
        assert torch.equal(a["q"], b["q"])
        assert a["n_widened"] == b["n_widened"]
    q, nn4, _ = port.order_param_q(pos[0], pos[0], box[0])
    assert np.array_equal(out["1"].nn_idx.cpu().numpy()[0], nn4)
    assert_q_close(out["1"].q.cpu().numpy()[0], q)
