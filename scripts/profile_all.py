"""One pass over the kernels north_star asks evidence for, for ncu (no oracle, product calls only):
K1 cell build, K2 brick sweep (fp64), K2f (fp32 mode), queued-centre passes, K3 H-bond counts, K5 Willard-Chandler field and
interface search.

    python scripts/profile_all.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from waterorderlib_b200 import engine, routines, synth  # noqa: E402

dev = torch.device("cuda", 0)
pos, box = synth.device_frames(50, 0, 2, sigma=0.25, device=dev)
import os
REPS = int(os.environ.get("WOL_PROFILE_REPS", "2"))  # 1 under ncu --set full (the report must stay below 64 MB)
for prec in ("fp64", "fp32"):
    for _ in range(REPS):
        r = engine.q3b_frames(pos, box, precision=prec)
    torch.cuda.synchronize()
    print(prec, "angles", int(r["ang_hist"].sum()), "widened", r["n_widened"], "overflow", r["n_overflow"], flush=True)
# K3: H-bond counts of a 1M-water frame (acceptors = O, donors = O listed twice, their hydrogens)
o = pos[0].cpu().numpy()
h = synth.add_hydrogens(o, seed=3)
o_d, h_d = pos[:1], torch.from_numpy(h[None]).to(dev)
d_d = o_d.repeat_interleave(2, dim=1).contiguous()
for _ in range(REPS):
    hb = routines.hbond_counts(o_d, d_d, h_d, box, 3.5, 120.0)
torch.cuda.synchronize()
print("hbonds per water", float(hb["acc_count"].sum()) / o.shape[0], flush=True)
# same-sweep observables of the same frame: LSI, CSR neighbour list, materialised three-body angles, RadialDistSame
for _ in range(REPS):
    lsi = routines.lsi(None, o_d, box)
    csr = routines.neighbors_csr(None, o_d[0], box, 0.0, 3.5)
    ang = routines.three_body_angles(None, o_d, box)
    rdf = routines.pair_hist(1, o_d[0], None, box, 0.1, 150)
torch.cuda.synchronize()
print("lsi mean", float(lsi[0].mean()), "csr pairs", int(csr[1].numel()), "angles", int(ang[0].numel()), "rdf pairs", int(rdf.sum()), flush=True)
del pos, o_d, h_d, d_d, lsi, csr, ang, rdf
# K5: 65536-water slab, 80^3 Willard-Chandler field, interface search against the ideal faces
sp, sbox, z_lo, z_hi = synth.slab_box(32, 32, 8, sigma=0.3, seed=11)
gp, gn = synth.plane_interface(sbox, z_lo, z_hi, spacing=2.0)
sp_d, gp_d, gn_d = torch.from_numpy(sp).to(dev), torch.from_numpy(gp).to(dev), torch.from_numpy(gn).to(dev)
grid = [(np.arange(80) + 0.5) * (sbox[d] / 80) for d in range(3)]
for _ in range(REPS):
    dens, _ = routines.willard_density(sp_d, sbox, 2.4, grid=grid, want_normals=True)
    iw = routines.interface_water(sp_d, gp_d, gn_d, 0.0, sbox, want_surfclose=True)
torch.cuda.synchronize()
print("field max", float(dens.max()), "interface points", gp.shape[0], flush=True)
