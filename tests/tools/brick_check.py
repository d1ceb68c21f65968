"""Brick path (wol_q3b_brick.cu) against the thread-per-centre path on the same inputs, output by output,
plus kernel timings.  Run on a GPU box:  python tests/tools/brick_check.py [--big]

The WOL_BRICK environment switch is read per call: 0 = never, 1 = the brick kernel whenever the shape allows, 3 = the
warp-specialised brick kernel (wol_q3b_brick_ws.cu) where it applies, else as 1.
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from waterorderlib_b200 import engine, synth  # noqa: E402


def run(pos, box, brick, **kw):
    os.environ["WOL_BRICK"] = str(int(brick))
    r = engine.q3b_frames(pos, box, **kw)
    torch.cuda.synchronize()
    return r


def compare(name, pos, box, **kw):
    a = run(pos, box, 0, **kw)
    ok = True
    msgs = []
    for mode in (1, 3):
        b = run(pos, box, mode, **kw)
        ok &= _same(a, b, msgs, "mode %d: " % mode)
    print("%-34s %s  widened %d/%d overflow %d/%d %s" % (name, "ok" if ok else "MISMATCH", a["n_widened"], b["n_widened"],
                                                        a["n_overflow"], b["n_overflow"], "; ".join(msgs)), flush=True)
    return ok


def _same(a, b, msgs, tag):
    ok = True
    for k in ("nn_idx", "n3", "ang_hist", "q_hist"):
        if k in a and a[k] is not None and k in b:
            same = torch.equal(a[k], b[k])
            if not same:
                ok = False
                msgs.append(tag + "%s differs in %d entries" % (k, int((a[k] != b[k]).sum())))
    if "q" in a:
        d = (a["q"] - b["q"]).abs().max().item()
        if not d < 1e-12:
            ok = False
            msgs.append(tag + "max |dq| = %g" % d)
    fs = torch.allclose(a["frame_stats"], b["frame_stats"], rtol=1e-12, atol=1e-9)
    if not fs:
        ok = False
        msgs.append(tag + "frame_stats differ: %s vs %s" % (a["frame_stats"][0].tolist(), b["frame_stats"][0].tolist()))
    return ok


def timed(pos_d, box, brick, reps=5, **kw):
    os.environ["WOL_BRICK"] = str(int(brick))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for e in ev:
        e.record()
    ws = engine.Workspace(pos_d.device)
    ts = []
    for _ in range(reps + 2):
        engine.q3b_frames(pos_d, box, workspace=ws, timing_events=ev, check_status=False, **kw)
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]))
    return float(np.median(ts[2:]))


def main():
    big = "--big" in sys.argv
    dev = torch.device("cuda", 0)
    ok = True
    rng = np.random.default_rng(7)
    for m, sigma in ((4, 0.25), (4, 0.6), (8, 0.25), (8, 0.6), (16, 0.25), (16, 0.6), (16, 0.0)):
        pos, box = synth.water_box(m, sigma=sigma, seed=100 + m)
        ok &= compare("ice m=%d sigma=%.2f" % (m, sigma), pos, box)
    # several frames, per-frame histograms, q only / three-body only, unwrapped coordinates, non-cubic box
    pos, box = synth.trajectory(8, 5, sigma=0.4, seed0=50)
    ok &= compare("5 frames, per-frame hist", pos, box, hist_per_frame=True)
    ok &= compare("5 frames, q only", pos, box, do_3body=False)
    ok &= compare("5 frames, three-body only", pos, box, do_q=False)
    ok &= compare("5 frames, highq 3.0 lowq 1.0", pos, box, highq=3.0, lowq=1.0, low3=2.5)
    shift = rng.integers(-3, 4, size=pos.shape).astype(np.float64) * box[:, None, :]
    ok &= compare("5 frames, unwrapped +-3 L", pos + shift, box)
    pos, box = synth.water_box(0, sigma=0.5, seed=9, dims=(12, 7, 5))
    ok &= compare("non-cubic 12x7x5", pos, box)
    # uniform random gas (close pairs, empty cells, dense cells) and a slab with vacuum
    for n, L in ((4000, 40.0), (20000, 55.0), (3000, 60.0)):
        p = rng.random((n, 3)) * L
        ok &= compare("random gas n=%d L=%.0f" % (n, L), p, np.array([L, L, L]))
    pos, box, _, _ = synth.slab_box(16, 16, 4, sigma=0.4, seed=3)
    ok &= compare("slab 16x16x4 + vacuum", pos, box)
    # a dense blob: bricks must be split, one-cell-wide columns overflow the stage
    p = np.concatenate([rng.random((6000, 3)) * 12.0 + 20.0, rng.random((4000, 3)) * 60.0])
    ok &= compare("dense blob in a 60 A box", p, np.array([60.0, 60.0, 60.0]), highq=3.4)
    res = {"parity_ok": bool(ok)}
    if big:
        m = 50
        pos, box = synth.water_box(m, sigma=0.25, seed=0)
        ok &= compare("1M waters sigma 0.25", pos, box)
        pos6, _ = synth.water_box(m, sigma=0.6, seed=1)
        ok &= compare("1M waters sigma 0.60", pos6, box)
        res["parity_ok"] = bool(ok)
        F = 8
        pd = torch.from_numpy(np.stack([pos] * F)).to(dev)
        pd6 = torch.from_numpy(np.stack([pos6] * F)).to(dev)
        for label, p in (("ice", pd), ("liquid", pd6)):
            t_old = timed(p, box, 0)
            t_one = timed(p, box, 1)
            t_new = timed(p, box, 3)
            res["ms_%s_tpc" % label] = t_old
            res["ms_%s_brick" % label] = t_one
            res["ms_%s_brick_ws" % label] = t_new
            print("%s: %d x 1M waters  tpc %.3f ms  brick %.3f ms  brick_ws %.3f ms" % (label, F, t_old, t_one, t_new), flush=True)
        r = run(pd, box, 1)
        st = engine.workspace_status  # noqa: F841
        res["slow_pairs_note"] = "see wol_status[3]"
    print(json.dumps(res))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
