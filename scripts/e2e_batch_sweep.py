"""Driver-level end-to-end leg (float32 host frames in, histograms out) for several pipeline batch sizes.

    python scripts/e2e_batch_sweep.py
"""
import numpy as np
import torch

from waterorderlib_b200 import synth
from waterorderlib_b200.pipeline import FramePipeline

dev = torch.device("cuda", 0)
B = 16
pos_d, box = synth.device_frames(50, 0, B, device=dev)
n = pos_d.shape[1]
pos_h = torch.empty((B, n, 3), dtype=torch.float32, pin_memory=True)
pos_h.copy_(pos_d)
del pos_d
for per_water in (False, True):
    for batch in (1, 2, 4, 8):
        for streams in (1, 2):
            pipe = FramePipeline(n, batch, dtype=np.float32, device=dev, want_q=per_water, want_n3=per_water, n_run_streams=streams)
            qh = torch.empty((B, n), dtype=torch.float64, pin_memory=True) if per_water else None
            nh = torch.empty((B, n), dtype=torch.int32, pin_memory=True) if per_water else None
            for _ in range(2):
                pipe.run(pos_h, box, out_q=qh, out_n3=nh)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                pipe.run(pos_h, box, out_q=qh, out_n3=nh)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print("per_water=%d batch=%d streams=%d: %.2f ms per 16 frames -> %.2fe9 wf/s" % (per_water, batch, streams, ms, B * n / ms / 1e6), flush=True)
            del pipe
