"""FP32 arithmetic mode of the fused path (precision="fp32"): the north_star tolerance is 1e-4 on per-molecule q and
cosines; neighbour selection and histogram bins can legitimately differ from the fp64 reference where two
distances (or an angle and a bin edge) agree to ~1e-7 relative, so those are compared statistically."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import port  # noqa: E402  (the checker)
from waterorderlib_b200 import engine, synth  # noqa: E402

Q_ATOL = 1e-4  # north_star tolerance for fp32 mode (q is O(1))


def compare(r, f, pos, box, sub=None, **cut):
    q, nn4, _ = port.order_param_q(pos if sub is None else sub, pos, box, cut.get("lowq", 0.0), cut.get("highq", 10.0))
    tb = port.three_body(pos if sub is None else sub, pos, box, cut.get("low3", 0.0), cut.get("high3", 3.413), materialize=False)
    m = q.shape[0]
    same_nn = np.all(r.nn_idx.cpu().numpy()[f] == nn4, axis=1)
    assert same_nn.mean() > 0.995, same_nn.mean()
    qd = np.abs(r.q.cpu().numpy()[f].astype(np.float64) - q)
    assert qd[same_nn].max() < Q_ATOL, qd[same_nn].max()
    assert (r.n3.cpu().numpy()[f] == tb["numAngs"]).mean() > 0.998
    h = r.ang_hist.cpu().numpy()[f if r.ang_hist.shape[0] > 1 else 0]
    assert abs(int(h.sum()) - int(tb["hist"].sum())) <= 0.002 * tb["hist"].sum() + 5
    assert np.abs(h - tb["hist"]).sum() <= 0.01 * tb["hist"].sum() + 10  # L1 distance: bin-edge flips only
    return m


@pytest.mark.parametrize("name", ["cfg1_n512_ice", "cfg1_n512_liq", "cfg2_n4096_frame0"])
def test_fp32_mode_against_golden_inputs(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    r = engine.q3b_frames(g["pos"].astype(np.float32), g["box"], precision="fp32")
    torch.cuda.synchronize()
    assert r.q.dtype == torch.float32
    compare(r, 0, g["pos"], g["box"])


def test_fp32_mode_batched_large_box():
    """1M-water box edge (310 A): ulp(310 A) in float32 is 3e-5 A -- the mode must stay inside 1e-4 on q."""
    pos, box = synth.water_box(24, sigma=0.3, seed=3)  # 110,592 waters, L = 149 A
    shift = np.floor(np.array([150.0, 0.0, 75.0]) / box) * box  # unwrapped coordinates far from the origin
    r = engine.q3b_frames(np.stack([pos, pos + shift]).astype(np.float32), box, precision="fp32", hist_per_frame=True)
    torch.cuda.synchronize()
    compare(r, 0, pos, box)
    compare(r, 1, pos, box)


def test_fp32_mode_subpopulation_and_widening():
    rng = np.random.default_rng(4)
    box = np.array([40.0, 36.0, 44.0])
    pos = (rng.random((900, 3)) * box).astype(np.float32).astype(np.float64)  # dilute gas (0.014 per A^3): many centres need the widened search
    sub = (rng.random((500, 3)) * box).astype(np.float32).astype(np.float64)
    r = engine.q3b_frames(pos.astype(np.float32), box, sub.astype(np.float32), precision="fp32", highq=9.0)
    torch.cuda.synchronize()
    compare(r, 0, pos, box, sub=sub, highq=9.0)
    assert r.n_widened > 0


def test_fp32_mode_small_box_generic_path():
    """Fewer than 4 cells per axis: the group-per-centre generic kernel runs the float arithmetic."""
    rng = np.random.default_rng(6)
    box = np.array([12.0, 13.0, 11.5])
    pos = (rng.random((70, 3)) * box).astype(np.float32).astype(np.float64)
    r = engine.q3b_frames(pos.astype(np.float32), box, precision="fp32", highq=5.5)
    torch.cuda.synchronize()
    assert r.nc[0] < 4
    compare(r, 0, pos, box, highq=5.5)


def test_fp32_brick_kernel_agrees_with_the_thread_per_centre_kernel(monkeypatch):
    """The two fp32 kernels of K2f on the same frames (WOL_BRICK=0: thread per centre, 1: brick): both are float
    arithmetic on the same wrapped coordinates and may differ only in last-ulp distances (the periodic shift is applied
    to the other operand); each stays inside the mode's bar against the fp64 oracle."""
    pos, box = synth.trajectory(12, 3, sigma=0.45, seed0=41)  # 3 x 13 824 waters
    pos32 = pos.astype(np.float32)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("WOL_BRICK", mode)
        out[mode] = engine.q3b_frames(pos32, box, precision="fp32", hist_per_frame=True)
        torch.cuda.synchronize()
    a, b = out["0"], out["1"]
    same = (a.nn_idx == b.nn_idx).all(dim=-1)
    assert float(same.float().mean()) > 0.9995
    assert float((a.q - b.q).abs()[same].max()) < Q_ATOL
    assert float((a.n3 == b.n3).float().mean()) > 0.9995
    assert float((a.ang_hist - b.ang_hist).abs().sum()) <= 1e-3 * float(a.ang_hist.sum())
    for f in range(3):
        compare(b, f, pos32[f].astype(np.float64), box[f])
