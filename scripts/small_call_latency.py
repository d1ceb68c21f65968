"""Per-call cost of the drop-in getOrderParamq on small frames (config 1: 512 waters; config 2: 4096 waters): wall clock per
call with numpy in / numpy out, and device time of the captured launch sequence.

    python scripts/small_call_latency.py
"""
import time

import numpy as np
import torch

from waterorderlib_b200 import engine, synth
from waterorderlib_b200.structureLibs import water_properties as wp

for m in (4, 8):
    p, box = synth.water_box(m, sigma=0.25, seed=1)
    for _ in range(5):
        wp.getOrderParamq(p, p, box)
    torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter()
    for _ in range(n):
        q = wp.getOrderParamq(p, p, box)
    wall = (time.perf_counter() - t0) / n
    pd = torch.from_numpy(p).cuda()
    res = {}
    for name, call in (("graph replay", engine.q3b_frames_graphed), ("plain launches", engine.q3b_frames)):
        for _ in range(5):
            call(pd, box, check_status=False)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            call(pd, box, check_status=False)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / n * 1e3
    print("N=%d: getOrderParamq numpy->numpy %.1f us per call; q + three-body device-resident: %s" % (
        p.shape[0], wall * 1e6, ", ".join("%s %.1f us" % kv for kv in res.items())), flush=True)
