"""A/B timing of the fused q + three-body path between two builds of libwol.so (development aid).

    python scripts/ab_k2.py [path/to/other/libwol.so] [frames=8]

Prints, for a jittered-ice and a liquid-like 1M-water box, the dominant kernel's time (the library's events) and the time
of the whole call (cell build + sweep + queued passes), best and median of 7 runs, plus a checksum of the results.
"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from waterorderlib_b200 import _capi

if len(sys.argv) > 1 and sys.argv[1] != "-":
    _capi.LIB_PATH = sys.argv[1]
from waterorderlib_b200 import engine, synth  # noqa: E402

F = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda", 0)
for sigma in (0.25, 0.6):
    pos, box = synth.device_frames(50, 0, F, sigma=sigma, device=dev)
    ws = engine.Workspace(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for e in ev:
        e.record()
    ks, ts = [], []
    for it in range(10):
        ev[2].record()
        r = engine.q3b_frames(pos, box, workspace=ws, timing_events=(ev[0], ev[1]), check_status=False,
                              want=("q", "n3", "ang_hist", "q_hist", "frame_stats"))
        ev[3].record()
        torch.cuda.synchronize()
        if it >= 3:
            ks.append(ev[0].elapsed_time(ev[1]))
            ts.append(ev[2].elapsed_time(ev[3]))
    print("sigma %.2f  %d frames: kernel best %.4f median %.4f ms | call best %.4f median %.4f ms | angles %d  q sum %.9f  n3 sum %d" % (
        sigma, F, min(ks), float(np.median(ks)), min(ts), float(np.median(ts)), int(r["ang_hist"].sum()), float(r["q"].sum()), int(r["n3"].sum())), flush=True)
    r = engine.q3b_frames(pos, box, workspace=ws, want=("q",))
    print("           widened %d  overflow %d  (of %d centres)" % (r["n_widened"], r["n_overflow"], F * pos.shape[1]), flush=True)
    del pos, ws, r
