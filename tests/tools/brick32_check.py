"""fp32 mode: the brick kernel (wol_q3b_brick32.cu, WOL_BRICK=1) against the thread-per-centre kernel (WOL_BRICK=0) on the
same inputs -- both are float arithmetic on the same wrapped coordinates, so they may differ only where the image shift is
applied to the other operand (last-ulp distances) -- plus kernel timings.  python tests/tools/brick32_check.py [--big]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from waterorderlib_b200 import engine, synth  # noqa: E402


def run(pos, box, brick, **kw):
    os.environ["WOL_BRICK"] = str(int(brick))
    r = engine.q3b_frames(pos, box, precision="fp32", **kw)
    torch.cuda.synchronize()
    return r


def compare(name, pos, box, **kw):
    a, b = run(pos, box, 0, **kw), run(pos, box, 1, **kw)
    msgs = []
    ok = True
    if "nn_idx" in a and a["nn_idx"] is not None:
        same = (a["nn_idx"] == b["nn_idx"]).all(dim=-1)
        frac = float(same.float().mean())
        dq = float((a["q"] - b["q"]).abs()[same].max()) if same.any() else 0.0
        msgs.append("same 4-NN %.6f, max |dq| %.2e" % (frac, dq))
        ok &= frac > 0.9995 and dq < 1e-4  # the mode's own bar on q
    if "n3" in a and a["n3"] is not None:
        f3 = float((a["n3"] == b["n3"]).float().mean())
        ha, hb = a["ang_hist"].double(), b["ang_hist"].double()
        l1 = float((ha - hb).abs().sum() / max(1.0, float(ha.sum())))
        msgs.append("same n3 %.6f, hist L1 %.2e" % (f3, l1))
        ok &= f3 > 0.9995 and l1 < 1e-3
    print("%-34s %s  widened %d/%d overflow %d/%d  %s" % (name, "ok" if ok else "MISMATCH", a["n_widened"], b["n_widened"],
                                                         a["n_overflow"], b["n_overflow"], "; ".join(msgs)), flush=True)
    return ok


def timed(pos_d, box, brick, reps=5):
    os.environ["WOL_BRICK"] = str(int(brick))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for e in ev:
        e.record()
    ws = engine.Workspace(pos_d.device)
    ts = []
    for _ in range(reps + 2):
        engine.q3b_frames(pos_d, box, workspace=ws, timing_events=ev, check_status=False, precision="fp32")
        torch.cuda.synchronize()
        ts.append(ev[0].elapsed_time(ev[1]))
    return float(np.median(ts[2:]))


def main():
    ok = True
    rng = np.random.default_rng(7)
    for m, sigma in ((4, 0.25), (8, 0.6), (16, 0.25), (16, 0.6), (16, 0.0)):
        pos, box = synth.water_box(m, sigma=sigma, seed=100 + m)
        ok &= compare("ice m=%d sigma=%.2f" % (m, sigma), pos.astype(np.float32), box)
    pos, box = synth.trajectory(8, 5, sigma=0.4, seed0=50)
    pos = pos.astype(np.float32)
    ok &= compare("5 frames, per-frame hist", pos, box, hist_per_frame=True)
    ok &= compare("5 frames, q only", pos, box, do_3body=False)
    ok &= compare("5 frames, three-body only", pos, box, do_q=False)
    ok &= compare("5 frames, highq 3.0 lowq 1.0", pos, box, highq=3.0, lowq=1.0, low3=2.5)
    p, b = synth.water_box(0, sigma=0.5, seed=9, dims=(12, 7, 5))
    ok &= compare("non-cubic 12x7x5", p.astype(np.float32), b)
    for n, L in ((4000, 40.0), (20000, 55.0), (3000, 60.0)):
        p = (rng.random((n, 3)) * L).astype(np.float32)
        ok &= compare("random gas n=%d L=%.0f" % (n, L), p, np.array([L, L, L]))
    p, b, _, _ = synth.slab_box(16, 16, 4, sigma=0.4, seed=3)
    ok &= compare("slab 16x16x4 + vacuum", p.astype(np.float32), b)
    p = np.concatenate([rng.random((6000, 3)) * 12.0 + 20.0, rng.random((4000, 3)) * 60.0]).astype(np.float32)
    ok &= compare("dense blob in a 60 A box", p, np.array([60.0, 60.0, 60.0]), highq=3.4)
    res = {"agree": bool(ok)}
    if "--big" in sys.argv:
        dev = torch.device("cuda", 0)
        for label, sigma in (("ice", 0.25), ("liquid", 0.6)):
            pos, box = synth.water_box(50, sigma=sigma, seed=0)
            ok &= compare("1M waters sigma %.2f" % sigma, pos.astype(np.float32), box)
            pd = torch.from_numpy(np.stack([pos.astype(np.float32)] * 8)).to(dev)
            res["ms_%s_tpc32" % label] = timed(pd, box, 0)
            res["ms_%s_brick32" % label] = timed(pd, box, 1)
            print("%s: 8 x 1M waters  tpc32 %.3f ms  brick32 %.3f ms" % (label, res["ms_%s_tpc32" % label], res["ms_%s_brick32" % label]), flush=True)
        res["agree"] = bool(ok)
    print(json.dumps(res))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
