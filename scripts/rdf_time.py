"""RadialDistSame of one 1M-water frame, a few calls (development aid; run it under ncu for the kernel's own time)."""
import sys

import torch

sys.path.insert(0, ".")
from waterorderlib_b200 import routines, synth  # noqa: E402

dev = torch.device("cuda")
O, box = synth.water_box(50, sigma=0.25, seed=1)
O_d = torch.from_numpy(O).to(dev)
for _ in range(3):
    r = routines.pair_hist(1, O_d, None, box, 0.1, 150)
torch.cuda.synchronize()
print("pairs", int(r.sum().item()))
