"""BASELINE.json configs 1-4 on one B200, each checked against the CPU oracle on the same inputs and timed next to
the CPU path (the reference's compiled Fortran + Python loops where its dense matrices fit, else the C restatement).
Writes profiles/configs_<tag>.json.   usage: python tests/tools/config_times.py <tag>   (lives under tests/ because it runs the CPU oracle as the checker)"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import port, ref_driver, ref_fortran  # noqa: E402  (checker / CPU baseline only)
from waterorderlib_b200 import engine, routines, synth  # noqa: E402
from waterorderlib_b200.structureLibs import surface_library as sl  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
dev = torch.device("cuda", 0)
out = []


def gpu_time(fn, reps=5):
    r = None
    for _ in range(3):
        r = fn()  # keep the previous result alive, as the timed loop does: the allocator then holds both generations
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        r = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


have_ref = ref_fortran.reference_available()
wl = ref_fortran.RefWaterlib() if have_ref else None

# ---- config 1: single frame, 512 waters ---------------------------------------------------------------------
pos, box = synth.water_box(4, sigma=0.25, seed=1234)
pos_d = torch.from_numpy(pos).to(dev)
ms, r = gpu_time(lambda: engine.q3b_frames(pos_d, box, check_status=False))
t0 = time.perf_counter()
if have_ref:
    q_ref = ref_driver.get_order_param_q(wl, pos, pos, box)
    ang, _ = ref_driver.get_cos_angs(wl, pos, pos, box)
    h_ref = np.histogram(ang, bins=500, range=[0.0, 180.0])[0]
else:
    q_ref = port.getOrderParamq(pos, pos, box)
    h_ref = port.three_body(pos, pos, box, materialize=False)["hist"]
cpu = time.perf_counter() - t0
ok = bool(np.allclose(r.q.cpu().numpy()[0], q_ref, rtol=1e-6, atol=1e-9) and np.array_equal(r.ang_hist.cpu().numpy()[0], h_ref))
out.append({"config": 1, "what": "512 waters, 1 frame: q + 3-body histogram", "gpu_ms": ms, "cpu_s": cpu,
            "cpu_kind": "reference (1 core)" if have_ref else "port", "parity": ok})

# ---- config 2: 4096 waters x 100 frames: q, neighbour counts, H-bond counts ---------------------------------
F = 100
O = np.stack([synth.water_box(8, sigma=0.25, seed=1234 + f)[0] for f in range(F)])
box = synth.water_box(8, sigma=0.0, seed=0)[1]
H = np.stack([synth.add_hydrogens(O[f], seed=1234 + f) for f in range(F)])
O_d, H_d = torch.from_numpy(O).to(dev), torch.from_numpy(H).to(dev)
D_d = O_d.repeat_interleave(2, dim=1).contiguous()


def cfg2():
    a = engine.q3b_frames(O_d, box, check_status=False, want=("q", "n3", "frame_stats"))
    b = routines.hbond_counts(O_d, D_d, H_d, box, 3.5, 120.0)
    return a, b


ms, (a, b) = gpu_time(cfg2, reps=3)
t0 = time.perf_counter()
n_cpu = 2
ok = True
for f in range(n_cpu):
    if have_ref:
        q_ref = ref_driver.get_order_param_q(wl, O[f], O[f], box)
        _, num = ref_driver.get_cos_angs(wl, O[f], O[f], box)
        mat = wl.generalhbonds(O[f], np.repeat(O[f], 2, axis=0), H[f], box, 3.5, 120.0)
        acc, don = mat.sum(1), mat.sum(0)
    else:
        q_ref = port.getOrderParamq(O[f], O[f], box)
        num = port.three_body(O[f], O[f], box, materialize=False)["numAngs"]
        acc, don = port.hbonds(O[f], np.repeat(O[f], 2, axis=0), H[f], box, 3.5, 120.0)
    ok &= bool(np.allclose(a.q.cpu().numpy()[f], q_ref, rtol=1e-6, atol=1e-9) and np.array_equal(a.n3.cpu().numpy()[f], num.astype(np.int32))
               and np.array_equal(b["acc_count"].cpu().numpy()[f], acc) and np.array_equal(b["don_count"].cpu().numpy()[f], don))
cpu = (time.perf_counter() - t0) / n_cpu * F
out.append({"config": 2, "what": "4096 waters x 100 frames: q, neighbour counts, H-bond counts (3.5 A / 120 deg)", "gpu_ms": ms,
            "cpu_s": cpu, "cpu_kind": ("reference (1 core)" if have_ref else "port") + ", %d frames timed, scaled to 100" % n_cpu,
            "parity": ok, "mean_hbonds_per_water": float((b["acc_count"].sum() + b["don_count"].sum()).item()) / (F * 4096)})

# ---- config 3: 32768 waters + solute: hydration-shell three-body distribution ------------------------------
pos, box = synth.water_box(16, sigma=0.4, seed=7)
sol = synth.solute_grid(box)
pos_d, sol_d = torch.from_numpy(pos).to(dev), torch.from_numpy(sol).to(dev)


def cfg3():
    mask = routines.shell_mask(sol_d, pos_d, box, 4.0)
    idx = torch.nonzero(mask[0]).squeeze(1)
    shell = pos_d[idx]
    rs = engine.q3b_frames(pos_d, box, shell, do_q=False, want=("n3", "ang_hist", "frame_stats"), check_status=False)
    ra = engine.q3b_frames(pos_d, box, None, do_q=False, want=("n3", "ang_hist", "frame_stats"), check_status=False)
    return idx, rs, ra


ms, (idx, rs, ra) = gpu_time(cfg3, reps=3)
t0 = time.perf_counter()
m_ref = port.shell_mask(sol, pos, box, 4.0)
sh = pos[np.nonzero(m_ref)[0]]
tb_s = port.three_body(sh, pos, box, materialize=False)
tb_a = port.three_body(pos, pos, box, materialize=False)
cpu = time.perf_counter() - t0
ok = bool(np.array_equal(idx.cpu().numpy(), np.nonzero(m_ref)[0]) and np.array_equal(rs.ang_hist.cpu().numpy()[0], tb_s["hist"])
          and np.array_equal(ra.ang_hist.cpu().numpy()[0], tb_a["hist"]))
out.append({"config": 3, "what": "32768 waters + 64-atom solute: shell (4 A) selection, 3-body histogram of shell and of all waters",
            "gpu_ms": ms, "cpu_s": cpu, "cpu_kind": "port (C restatement, cell list, OpenMP; the reference needs a 4 GiB matrix per call)",
            "parity": ok, "shell_waters": int(idx.numel())})

# ---- config 4: air-water slab, 65536 waters: depth-binned q --------------------------------------------------
pos, box, z_lo, z_hi = synth.slab_box(32, 32, 8, sigma=0.3, seed=11)
gp, gn = synth.plane_interface(box, z_lo, z_hi, spacing=2.0)
pos_d, gp_d, gn_d = torch.from_numpy(pos).to(dev), torch.from_numpy(gp).to(dev), torch.from_numpy(gn).to(dev)
ms, prof = gpu_time(lambda: sl.depthBinnedQ(pos_d, box, gp_d, gn_d, binWidth=1.0, depthRange=(-30.0, 6.0)), reps=3)
t0 = time.perf_counter()
q_ref = port.getOrderParamq(pos, pos, box)
_, _, nw, depth = port.interface_water(pos, gp, gn, 0.0, box)
cpu = time.perf_counter() - t0
b = np.floor((depth + 30.0) / 1.0)
sel = (b >= 0) & (b < 36)
ok = bool(np.array_equal(prof["depth"].cpu().numpy(), depth) and np.allclose(prof["q"].cpu().numpy(), q_ref, rtol=1e-6, atol=1e-9)
          and np.array_equal(prof["count"], np.bincount(b[sel].astype(int), minlength=36)))
out.append({"config": 4, "what": "slab, 65536 waters, %d interface points: q, InterfaceWater depth, 1 A depth profile" % gp.shape[0],
            "gpu_ms": ms, "cpu_s": cpu, "cpu_kind": "port (C restatement, 1 core for the interface search)", "parity": ok,
            "q_mean_surface_vs_bulk": [float(np.nanmean(prof["q_mean"][-8:-4])), float(np.nanmean(prof["q_mean"][8:16]))]})

# ---- config 4, interface from the frame itself: 80^3 Willard-Chandler grid (sigma 2.4 A, level 0.016) -> iso-surface
# vertices -> normals -> InterfaceWater depth -> profile, all on the device ----------------------------------------------
grid = [(np.arange(80) + 0.5) * (box[d] / 80) for d in range(3)]
ms_if, (ipts, inrm) = gpu_time(lambda: sl.instantaneousInterface(pos_d, box, grid=grid), reps=3)
ms_all, prof2 = gpu_time(lambda: sl.depthBinnedQ(pos_d, box, binWidth=1.0, depthRange=(-30.0, 6.0), grid=grid), reps=3)
sub = [grid[0][:8], grid[1], grid[2]]  # CPU: one tenth of the grid, scaled
t0 = time.perf_counter()
d_ref, _ = port.willard_density_field(pos, sub[0], sub[1], sub[2], box, 2.4)
cpu_field = (time.perf_counter() - t0) * 10.0
d_gpu, _ = routines.willard_density(pos_d, box, 2.4, grid=sub, want_normals=False)
pts_ref = port.iso_points(routines.willard_density(pos_d, box, 2.4, grid=grid, want_normals=False)[0].cpu().numpy(), grid[0], grid[1], grid[2], 0.016)
t0 = time.perf_counter()
_, _, nw2, depth2 = port.interface_water(pos, ipts.cpu().numpy(), inrm.cpu().numpy(), 0.0, box)
cpu_iw = time.perf_counter() - t0
given = sl.depthBinnedQ(pos_d, box, ipts, inrm, binWidth=1.0, depthRange=(-30.0, 6.0))
ok = bool(np.allclose(d_gpu.cpu().numpy(), d_ref, rtol=1e-12, atol=1e-18) and pts_ref.shape == tuple(ipts.shape)
          and np.allclose(ipts.cpu().numpy(), pts_ref, rtol=0, atol=1e-9) and np.array_equal(given["depth"].cpu().numpy(), depth2))
zs = ipts[:, 2].cpu().numpy()
out.append({"config": "4 (instantaneous interface)", "what": "slab, 65536 waters: 80^3 Willard-Chandler field, %d iso-surface vertices + normals, "
            "q, InterfaceWater depth, 1 A depth profile" % ipts.shape[0], "gpu_ms": ms_all, "gpu_ms_interface_only": ms_if,
            "cpu_s": cpu_field + cpu_iw + cpu, "cpu_kind": "port (C restatement, 1 core; density field timed on a tenth of the grid and scaled)",
            "parity": ok, "surface_z_minus_ideal_faces": [float(np.mean(zs[zs < 0.5 * (z_lo + z_hi)]) - z_lo), float(np.mean(zs[zs > 0.5 * (z_lo + z_hi)]) - z_hi)],
            "q_mean_surface_vs_bulk": [float(np.nanmean(prof2["q_mean"][-8:-4])), float(np.nanmean(prof2["q_mean"][8:16]))]})

for o in out:
    o["speedup"] = o["cpu_s"] * 1e3 / o["gpu_ms"]
    print(json.dumps(o))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "configs_%s.json" % tag), "w"), indent=1)
