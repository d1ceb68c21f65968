"""K1 alone: the cell-list build of a batch of 1M-water frames, timed with CUDA events on the launching stream.
    python scripts/k1_time.py [frames] [--shuffle] [--f32]
--shuffle: atoms in random order (as in a real MD topology) instead of the lattice order of the synthetic box.
Prints one JSON line: ms per batch, and GB/s at the algorithmic 80 B (fp64 input) / 68 B (fp32 input) per atom."""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from waterorderlib_b200 import engine, synth  # noqa: E402
from waterorderlib_b200._capi import WOL_PREC_FP64, check, lib  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
F = int(args[0]) if args else 16
dev = torch.device("cuda", 0)
pos, box = synth.water_box(50, sigma=0.25, seed=0)
if "--shuffle" in sys.argv:
    pos = pos[np.random.default_rng(1).permutation(pos.shape[0])]
dt = np.float32 if "--f32" in sys.argv else np.float64
pos_d = torch.from_numpy(np.ascontiguousarray(np.stack([pos] * F).astype(dt))).to(dev)
N = pos.shape[0]
box_h = engine.as_host_boxes(box, F)
box_d = torch.from_numpy(box_h.copy()).to(dev)
nc, edge_min, box_max = engine.plan_grid(box_h, engine.default_r_cell(True, True, 3.413, 10.0))
L = lib()
need = L.wol_workspace_bytes(F, N, N, ctypes.byref(nc))
ws = torch.empty(need + 256, dtype=torch.uint8, device=dev)
ws_ptr = (ws.data_ptr() + 255) // 256 * 256
stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for it in range(8):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(L.wol_cell_build(ctypes.c_void_p(pos_d.data_ptr()), engine._dtype_code(pos_d), ctypes.c_void_p(box_d.data_ptr()), F, N,
                           ctypes.byref(nc), WOL_PREC_FP64, ctypes.c_void_p(ws_ptr), need, stream), "wol_cell_build")
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts[2:]))
bpa = 80 if dt == np.float64 else 68
print(json.dumps({"frames": F, "atoms": N, "input": np.dtype(dt).name, "shuffled": "--shuffle" in sys.argv, "ms": ms,
                  "GBps_algorithmic": F * N * bpa / ms / 1e6, "bytes_per_atom": bpa}))
