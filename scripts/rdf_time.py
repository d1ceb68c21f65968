"""RadialDistSame of one 1M-water frame (development aid; run it under ncu for the kernel's own time).

    python scripts/rdf_time.py [path/to/other/libwol.so | -] [shuffle]

Prints the routine time and a checksum of the counts, so that two builds of the library can be compared bin for bin.
"""
import sys

import torch

sys.path.insert(0, ".")
from waterorderlib_b200 import _capi

if len(sys.argv) > 1 and sys.argv[1] != "-":
    _capi.LIB_PATH = sys.argv[1]
from waterorderlib_b200 import routines, synth  # noqa: E402

dev = torch.device("cuda")
O, box = synth.water_box(50, sigma=0.25, seed=1)
if "shuffle" in sys.argv[2:]:  # a real topology: atom order unrelated to position
    import numpy as np
    O = O[np.random.default_rng(5).permutation(O.shape[0])]
O_d = torch.from_numpy(O).to(dev)
for _ in range(2):
    r = routines.pair_hist(1, O_d, None, box, 0.1, 150)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    r = routines.pair_hist(1, O_d, None, box, 0.1, 150)
e1.record()
torch.cuda.synchronize()
k = torch.arange(1, 151, device=dev, dtype=torch.int64)
print("radialdistsame 1M waters: %.3f ms, pairs %d, checksum %d, first bins with counts %s" % (
    e0.elapsed_time(e1) / 3, int(r.sum().item()), int((r * k * k).sum().item()), r[r > 0][:4].tolist()))
