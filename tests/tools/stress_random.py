import sys, numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from test_gpu_random import make_case
from oracle import port
from waterorderlib_b200 import engine
from waterorderlib_b200._capi import WolError
lo = int(sys.argv[1]) if len(sys.argv) > 1 else 36
hi = int(sys.argv[2]) if len(sys.argv) > 2 else 436
bad = skipped = 0
for seed in range(lo, hi):
    pos, box, sub, cut = make_case(seed)
    try:
        r = engine.q3b_frames(pos, box, sub, **cut)
    except WolError as e:
        skipped += 1; continue
    c = pos if sub is None else sub
    q, nn4, _ = port.order_param_q(c, pos, box, cut["lowq"], cut["highq"])
    tb = port.three_body(c, pos, box, cut["low3"], cut["high3"], materialize=False)
    ok = (np.array_equal(r.n3.cpu().numpy()[0], tb["numAngs"]) and np.array_equal(r.ang_hist.cpu().numpy()[0], tb["hist"])
          and np.array_equal(r.nn_idx.cpu().numpy()[0], nn4) and np.allclose(r.q.cpu().numpy()[0], q, rtol=1e-6, atol=1e-9))
    if not ok:
        bad += 1; print("MISMATCH seed", seed, pos.shape, box, cut)
print("stress: seeds %d..%d, %d cases, %d mismatches, %d capacity skips" % (lo, hi - 1, hi - lo, bad, skipped))
