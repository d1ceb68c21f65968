"""Parity of the slab / interface routines (BASELINE config 4) and histrr3b: CUDA kernels through the C ABI
against golden fixtures from the reference's compiled Fortran (fortran/waterlib.f90:1286-1469, :1550-1593) and
against the CPU oracle at a larger size.

Integer outputs (nearest-point indices, counts, histogram bins) bit-exact; depths bit-exact (same fp64 operations
in the same order); densities within 1e-12 relative (device exp vs glibc's, terms summed in cell order);
unit normals within 1e-9 absolute.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import port  # noqa: E402  (the checker)
from waterorderlib_b200 import routines, synth  # noqa: E402
from waterorderlib_b200.structureLibs import surface_library as sl  # noqa: E402
from waterorderlib_b200.structureLibs import water_properties as wp  # noqa: E402
from waterorderlib_b200.structureLibs import waterlib as wl  # noqa: E402

DENS_RTOL, NORM_ATOL = 1e-12, 1e-9


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


def check_density(d, n, d_ref, n_ref):
    assert d.shape == d_ref.shape and n.shape == n_ref.shape
    assert np.allclose(d, d_ref, rtol=DENS_RTOL, atol=1e-18)
    fin = np.isfinite(n_ref)
    assert np.array_equal(np.isfinite(n), fin)  # NaN normals (0/0) exactly where the reference has them
    assert np.allclose(n[fin], n_ref[fin], rtol=0, atol=NORM_ATOL)


def test_willard_density_golden(golden_dir):
    g = load(golden_dir, "slab_n256")
    d, n = wl.willarddensityfield(g["pos"], g["gx"], g["gy"], g["gz"], g["box"], float(g["smoothlen"]))
    assert d.flags.f_contiguous
    check_density(d, n, g["dens"], g["norms"])
    assert (g["dens"] > 0.016).any() and (g["dens"] == 0.0).any()  # liquid and vacuum both sampled
    d, n = wl.willarddensitypoints(g["pos"], g["pts"], g["box"], float(g["smoothlen"]))
    check_density(d, n, g["pdens"], g["pnorms"])


def test_interface_water_golden(golden_dir):
    g = load(golden_dir, "slab_n256")
    wc, sc, nw, dist = wl.interfacewater(g["pos"], g["gridpos"], g["gridnorm"], float(g["cutoff"]), g["box"])
    assert np.array_equal(wc, g["watclose"]) and np.array_equal(sc, g["surfclose"])  # 1-based, like f2py
    assert nw == int(g["numwater"]) and np.array_equal(dist, g["allwatdists"])


def test_interface_water_vs_oracle_larger_and_far_points():
    pos, box, z_lo, z_hi = synth.slab_box(8, 8, 3, sigma=0.3, seed=2)  # 1536 waters
    gp, gn = synth.plane_interface(box, z_lo, z_hi, spacing=2.0)
    # a few waters evaporated far into the vacuum, beyond sqrt(1000) A of any surface point in z? (box z is 56 A:
    # not reachable here), so instead move the interface points of one face far away in a second call
    r = routines.interface_water(pos, gp, gn, 2.0, box)
    wc, sc, nw, dist = port.interface_water(pos, gp, gn, 2.0, box)
    assert np.array_equal(r["watclose"].cpu().numpy(), wc) and np.array_equal(r["surfclose"].cpu().numpy(), sc)
    assert int(r["numwater"].item()) == nw and np.array_equal(r["allwatdists"].cpu().numpy(), dist)
    # huge box: waters farther than sqrt(1000) from the only interface point keep watclose = -1, depth 0
    big = np.array([200.0, 200.0, 200.0])
    one_p, one_n = np.array([[100.0, 100.0, 100.0]]), np.array([[0.0, 0.0, 1.0]])
    w = np.array([[100.0, 100.0, 110.0], [10.0, 10.0, 10.0], [100.0, 131.0, 100.0]])
    r = routines.interface_water(w, one_p, one_n, 20.0, big)
    wc, sc, nw, dist = port.interface_water(w, one_p, one_n, 20.0, big)
    assert np.array_equal(r["watclose"].cpu().numpy(), wc) and list(wc) == [0, -1, 0]
    assert np.array_equal(r["allwatdists"].cpu().numpy(), dist) and int(r["numwater"].item()) == nw == 2


def test_histrr3b_golden(golden_dir):
    g = load(golden_dir, "histrr3b_n216")
    h = wl.histrr3b(g["pos"], g["box"], float(g["dwidth"]), int(g["dnum"]), float(g["awidth"]), int(g["anum"]))
    assert h.dtype == np.float64 and h.flags.f_contiguous and h.shape == (8, 8, 36)
    assert np.array_equal(h, g["hist"].astype(np.float64)) and h.sum() == g["hist"].sum() > 1000


def test_histrr3b_vs_oracle_liquid_and_collinear():
    pos, box = synth.water_box(4, sigma=0.6, seed=5)
    for dw, dn, aw, an in ((0.25, 14, 2.0, 90), (1.0, 5, 7.5, 24)):
        assert np.array_equal(routines.histrr3b(pos, box, dw, dn, aw, an).cpu().numpy(), port.histrr3b(pos, box, dw, dn, aw, an))
    # simple cubic lattice: exactly antiparallel pairs give -180 degrees -> no bin, in both
    g = np.arange(4) * 3.0
    cub = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    h = routines.histrr3b(cub, np.array([12.0, 12.0, 12.0]), 0.5, 7, 10.0, 18).cpu().numpy()
    assert np.array_equal(h, port.histrr3b(cub, np.array([12.0, 12.0, 12.0]), 0.5, 7, 10.0, 18)) and h.sum() == 64 * 12


def test_profile_bins_and_depth_binned_q():
    pos, box, z_lo, z_hi = synth.slab_box(6, 6, 3, sigma=0.3, seed=9)
    gp, gn = synth.plane_interface(box, z_lo, z_hi, spacing=2.0)
    out = sl.depthBinnedQ(pos, box, gp, gn, binWidth=1.0, depthRange=(-12.0, 4.0))
    q = port.getOrderParamq(pos, pos, box)
    _, _, nw, depth = port.interface_water(pos, gp, gn, 0.0, box)
    assert np.array_equal(out["depth"].cpu().numpy(), depth) and out["numwater"] == nw
    assert np.allclose(out["q"].cpu().numpy(), q, rtol=1e-6, atol=1e-9)
    b = np.floor((depth - (-12.0)) / 1.0)
    ok = (b >= 0) & (b < 16)
    cnt = np.bincount(b[ok].astype(int), minlength=16)
    assert np.array_equal(out["count"], cnt) and cnt.sum() > 0.9 * len(q)
    s1 = np.bincount(b[ok].astype(int), weights=q[ok], minlength=16)
    nz = cnt > 0
    assert np.allclose(out["q_mean"][nz], s1[nz] / cnt[nz], rtol=1e-9)
    # waters near the surface are less tetrahedral than the slab interior (they miss neighbours)
    assert out["q_mean"][nz][-2:].mean() < out["q_mean"][nz][:4].mean()


def test_density_field_of_densityGrid():
    pos, box, z_lo, z_hi = synth.slab_box(3, 3, 2, sigma=0.3, seed=4)
    heavy = pos[:5]
    dens, norms, spans, spaces = sl.densityField(heavy, pos, box, nBins=21)
    assert dens.shape == (20, 20, 20) and norms.shape == (20, 20, 20, 3)
    ref_d, ref_n = port.willard_density_field(pos, spans[0].ravel(), spans[1].ravel(), spans[2].ravel(), box, 2.4)
    check_density(dens, norms, ref_d, ref_n)


def test_density_field_golden_and_voxel_grid(golden_dir):
    """DensityField (fortran/waterlib.f90:1219-1268): counts per cube, bit-exact (integer counts / pow(binwidth, 3.0))."""
    g = load(golden_dir, "slab_n256")
    d = wl.densityfield(g["pos"], g["gx"] + 0.3, g["gy"], g["gz"], g["box"])
    assert d.shape == g["voxel"].shape and np.array_equal(d, g["voxel"]) and d.max() > 0
    pos, box, _, _ = synth.slab_box(6, 6, 3, sigma=0.3, seed=5)
    dv = sl.densityVoxel(pos[:40], pos, box)
    spans = []
    for k in range(3):
        s = np.linspace(0.8 * pos[:40, k].min(), 1.2 * pos[:40, k].max(), 11)
        spans.append(s[:-1] + (s[1] - s[0]))
    assert dv.shape == (10, 10, 10) and np.array_equal(dv, port.density_field(pos, spans[0], spans[1], spans[2], box))


def test_iso_points_vs_oracle_analytic_and_density():
    """wol_iso_points against the numpy restatement: bit-exact vertices, same order, on an analytic field with
    anisotropic grid spacing and on a Willard density field; count-only and truncated-capacity behaviour."""
    gx, gy, gz = np.linspace(-2.0, 2.0, 23), np.linspace(-2.1, 2.2, 17), np.linspace(-1.9, 2.0, 31)
    X, Y, Z = np.meshgrid(gx, gy, gz, indexing="ij")
    field = np.sqrt(X * X + 1.3 * Y * Y + 0.8 * Z * Z) + 0.05 * np.sin(3 * X) * np.cos(2 * Z)
    for level in (1.3, 0.4, 5.0):
        ref = port.iso_points(field, gx, gy, gz, level)
        got = routines.iso_points(field, (gx, gy, gz), level).cpu().numpy()
        assert got.shape == ref.shape and np.array_equal(got, ref)
    assert routines.iso_points(field, (gx, gy, gz), 5.0).shape == (0, 3)
    pos, box, z_lo, z_hi = synth.slab_box(4, 4, 2, sigma=0.3, seed=2)
    grid = [(np.arange(n) + 0.5) * (box[d] / n) for d, n in enumerate((12, 12, 19))]
    dens, _ = routines.willard_density(pos, box, 2.4, grid=grid)
    ref = port.iso_points(dens.cpu().numpy(), grid[0], grid[1], grid[2], 0.016)
    got = routines.iso_points(dens, grid, 0.016).cpu().numpy()
    assert np.array_equal(got, ref) and len(ref) > 200
    # a grid with a single node along an axis has no edges along it
    one = routines.iso_points(field[:, :1, :], (gx, gy[:1], gz), 1.3).cpu().numpy()
    assert np.array_equal(one, port.iso_points(field[:, :1, :], gx, gy[:1], gz, 1.3))


def test_iso_points_capacity_and_errors():
    import ctypes
    from waterorderlib_b200._capi import lib
    gx = np.linspace(0.0, 1.0, 9)
    field = np.add.outer(np.add.outer(gx, gx), gx)  # x + y + z
    ref = port.iso_points(field, gx, gx, gx, 1.45)
    d = torch.from_numpy(field).cuda()
    g = torch.from_numpy(gx).cuda()
    nbytes = lib().wol_iso_scratch_bytes(9, 9, 9)
    scratch = torch.empty(nbytes // 4 + 4, dtype=torch.int32, device="cuda")
    n_total = torch.zeros(1, dtype=torch.int32, device="cuda")
    cap = 10
    pts = torch.full((cap + 2, 3), -7.0, dtype=torch.float64, device="cuda")
    vp = ctypes.c_void_p
    s = vp(torch.cuda.current_stream().cuda_stream)
    rc = lib().wol_iso_points(vp(d.data_ptr()), vp(g.data_ptr()), vp(g.data_ptr()), vp(g.data_ptr()), 9, 9, 9, 1.45, vp(scratch.data_ptr()),
                              nbytes, vp(pts.data_ptr()), cap, vp(n_total.data_ptr()), s)
    assert rc == 0 and int(n_total.item()) == len(ref) > cap
    out = pts.cpu().numpy()
    assert np.array_equal(out[:cap], ref[:cap]) and np.all(out[cap:] == -7.0)  # nothing past the capacity
    rc = lib().wol_iso_points(vp(d.data_ptr()), vp(g.data_ptr()), vp(g.data_ptr()), vp(g.data_ptr()), 9, 9, 9, 1.45, vp(scratch.data_ptr()),
                              16, vp(pts.data_ptr()), cap, vp(n_total.data_ptr()), s)
    assert rc != 0 and b"scratch" in lib().wol_last_error()


def test_instantaneous_interface_slab_end_to_end():
    """cfg4 without the analytic planes: density -> iso-surface -> normals -> depth -> q profile on the device.  The
    Willard-Chandler surface of a slab sits near the ideal faces, its normals are close to +-z and point out of the
    liquid, and the depth profile agrees with the analytic-plane one up to the surface's roughness."""
    pos, box, z_lo, z_hi = synth.slab_box(6, 6, 3, sigma=0.3, seed=9)
    gp, gn = sl.instantaneousInterface(pos, box, spacing=1.5)
    assert gp.shape == gn.shape and gp.shape[0] > 500
    top = gp[:, 2] > 0.5 * (z_lo + z_hi)
    assert np.all(np.abs(gp[top, 2] - z_hi) < 3.0) and np.all(np.abs(gp[~top, 2] - z_lo) < 3.0)
    assert np.allclose(np.linalg.norm(gn, axis=1), 1.0, atol=1e-12)
    assert np.all(gn[top, 2] > 0.5) and np.all(gn[~top, 2] < -0.5)
    # the vertices lie on the iso-surface of the exact field up to the linear-interpolation error
    d_at, _ = routines.willard_density(pos, box, 2.4, points=gp)
    assert np.all(np.abs(d_at.cpu().numpy() - 0.016) < 2e-3)
    # vertices and normals are what the oracle's restatements give for the same grid
    grid = [(np.arange(n) + 0.5) * (box[d] / n) for d, n in enumerate(int(np.ceil(b / 1.5)) for b in box)]
    dens_ref, _ = port.willard_density_field(pos, grid[0], grid[1], grid[2], box, 2.4)
    dens_gpu, _ = routines.willard_density(pos, box, 2.4, grid=grid)
    ref_pts = port.iso_points(dens_gpu.cpu().numpy(), grid[0], grid[1], grid[2], 0.016)
    # (a second density evaluation: the cell list orders atoms within a cell by atomics, so sums differ by ulps)
    assert gp.shape == ref_pts.shape and np.allclose(gp, ref_pts, rtol=0, atol=1e-9)
    assert np.allclose(dens_gpu.cpu().numpy(), dens_ref, rtol=DENS_RTOL, atol=1e-18)
    _, n_ref = port.willard_density_points(pos, gp, box, 2.4)
    assert np.allclose(gn, -n_ref, rtol=0, atol=NORM_ATOL)
    out = sl.depthBinnedQ(pos, box, binWidth=1.0, depthRange=(-12.0, 4.0), spacing=1.5)
    assert out["n_surface"] == gp.shape[0]
    _, _, nw, depth = port.interface_water(pos, gp, gn, 0.0, box)
    given = sl.depthBinnedQ(pos, box, gp, gn, binWidth=1.0, depthRange=(-12.0, 4.0))
    assert np.array_equal(given["depth"].cpu().numpy(), depth) and given["numwater"] == nw
    # the self-computed interface is a fresh density evaluation (ulps apart, see above)
    assert np.mean(np.abs(out["depth"].cpu().numpy() - depth) < 1e-6) > 0.999 and abs(out["numwater"] - nw) <= 2
    gp0, gn0 = synth.plane_interface(box, z_lo, z_hi, spacing=2.0)
    flat = sl.depthBinnedQ(pos, box, gp0, gn0, binWidth=1.0, depthRange=(-12.0, 4.0))
    assert np.abs(np.median(out["depth"].cpu().numpy() - flat["depth"].cpu().numpy())) < 3.0
    assert out["count"].sum() > 0.9 * len(pos)


def test_iso_surface_faces_are_a_closed_oriented_mesh_over_the_iso_points():
    """wol_iso_faces (parity unpinned: skimage is not vendored) pinned by properties on fields with known surfaces: same
    vertices as iso_points; every mesh edge is shared by exactly two faces with opposite directions (watertight, consistently
    oriented); Euler characteristic 2 per closed component; area and enclosed volume of a sphere; normals towards lower
    values; the area rule of fortran/imagelib.f90:254-267."""
    g = np.linspace(-1.0, 1.0, 41)
    X, Y, Z = np.meshgrid(g, g, g, indexing="ij")
    r = np.sqrt(X * X + Y * Y + Z * Z)
    field = 1.0 - r                      # above the level inside the sphere
    R = 0.6
    verts, faces, areas = routines.iso_surface(field, (g, g, g), 1.0 - R)
    v, f, a = verts.cpu().numpy(), faces.cpu().numpy(), areas.cpu().numpy()
    assert np.array_equal(v, routines.iso_points(field, (g, g, g), 1.0 - R).cpu().numpy())
    assert f.min() >= 0 and f.max() < len(v) and len(np.unique(f)) == len(v)
    # directed edges: each appears once, and its reverse appears once
    de = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
    key = de[:, 0].astype(np.int64) * len(v) + de[:, 1]
    rev = de[:, 1].astype(np.int64) * len(v) + de[:, 0]
    assert len(np.unique(key)) == len(key) and np.array_equal(np.sort(key), np.sort(rev))
    assert len(v) - len(key) // 2 + len(f) == 2                                          # V - E + F of a sphere
    p0, p1, p2 = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    cr = np.cross(p1 - p0, p2 - p0)
    assert np.allclose(a, np.linalg.norm(cr, axis=1), rtol=1e-5, atol=1e-9)            # |v1 x v2| (imagelib.f90:254-267; 1 - cos^2 cancels for slivers)
    assert abs(0.5 * a.sum() - 4.0 * np.pi * R * R) < 0.01 * 4.0 * np.pi * R * R
    vol = np.einsum("ij,ij->i", p0, cr).sum() / 6.0
    assert abs(vol - 4.0 / 3.0 * np.pi * R ** 3) < 0.01 * 4.0 / 3.0 * np.pi * R ** 3      # positive: normals point outwards = to lower values
    # two separate blobs and an ambiguous saddle: still closed and oriented, Euler characteristic 2 per component
    blobs = np.maximum(1.0 - np.sqrt((X - 0.45) ** 2 + Y * Y + Z * Z), 1.0 - np.sqrt((X + 0.45) ** 2 + Y * Y + Z * Z))
    for fld, lvl, chi in ((blobs, 0.7, 4), (blobs, 0.56, None), (np.sin(3 * X) * np.sin(3 * Y) * np.sin(3 * Z), 0.3, None)):
        vv, ff, _ = routines.iso_surface(fld, (g, g, g), lvl, want_areas=False)
        ff = ff.cpu().numpy()
        n = vv.shape[0]
        de = np.concatenate([ff[:, [0, 1]], ff[:, [1, 2]], ff[:, [2, 0]]]).astype(np.int64)
        key, rev = de[:, 0] * n + de[:, 1], de[:, 1] * n + de[:, 0]
        inner = ~np.isin(key, rev)  # edges on the grid boundary have no partner; everything else is matched exactly once
        pts = vv.cpu().numpy()
        on_border = lambda idx: np.any(np.isclose(np.abs(pts[idx]), 1.0), axis=1)  # noqa: E731
        assert len(np.unique(key)) == len(key) and np.all(on_border(de[inner, 0]) & on_border(de[inner, 1]))
        if chi is not None:
            assert n - len(key) // 2 + len(ff) == chi
    e = routines.iso_surface(field, (g, g, g), 5.0)
    assert e[0].shape == (0, 3) and e[1].shape == (0, 3)


def test_interfaceMesh_of_a_slab():
    pos, box, z_lo, z_hi = synth.slab_box(6, 6, 3, sigma=0.3, seed=2)
    verts, faces, norms, values, areas = sl.interfaceMesh(pos, box, spacing=1.5)
    assert verts.shape[0] > 100 and faces.shape[1] == 3 and norms.shape == verts.shape and np.all(values == 0.016)
    # two sheets near the ideal faces; the triangulated area is about twice the box cross-section (0.5 * the imagelib rule)
    assert np.all((np.abs(verts[:, 2] - z_lo) < 3.0) | (np.abs(verts[:, 2] - z_hi) < 3.0))
    assert 0.9 < 0.5 * areas.sum() / (2.0 * box[0] * box[1]) < 1.6
