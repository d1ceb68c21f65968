"""Host side of the value-returning routines of the per-frame path (include/wol_capi.h, second half):
materialised three-body angles, np.histogram on the device, dense neighbour matrices, reimage /
tetracosang / lsidists, hydrogen-bond counting and hydration-shell selection.

Like engine.py this module only moves pointers: PyTorch owns device memory and streams, libwol.so's sm_100a
kernels compute.  Inputs may be numpy arrays or torch tensors (CUDA tensors are used in place); outputs are
torch CUDA tensors -- the numpy-in / numpy-out drop-in layer is waterorderlib_b200.structureLibs.
"""
import ctypes

import numpy as np
import torch

from . import engine
from ._capi import WOL_F64, WOL_PREC_FP64, HbondArgs, WolError, check, lib

_vp = ctypes.c_void_p


def _device(device=None, *tensors):
    if device is not None:
        return torch.device(device)
    for t in tensors:
        if isinstance(t, torch.Tensor) and t.is_cuda:
            return t.device
    if not torch.cuda.is_available():
        raise RuntimeError("waterorderlib_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _f64(a, device, shape_tail=None):
    """-> contiguous float64 CUDA tensor."""
    if isinstance(a, torch.Tensor):
        t = a.to(device=device, dtype=torch.float64)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(device)
    t = t.contiguous()
    if shape_tail is not None and tuple(t.shape[-len(shape_tail):]) != tuple(shape_tail):
        raise ValueError("expected trailing shape %s, got %s" % (shape_tail, tuple(t.shape)))
    return t


def _box3(box, device):
    """BoxDims as (3,) or (1,3) -> float64 CUDA tensor of 3 values (the reference accepts both,
    structureLibs/orderParam_lib.py:1315 vs :183)."""
    t = _f64(box, device).reshape(-1)
    if t.numel() != 3:
        raise ValueError("box must hold 3 edge lengths, got %d values" % t.numel())
    return t


def _stream():
    return _vp(torch.cuda.current_stream().cuda_stream)


class CellList:
    """A built cell list over one batch of frames (FP64 records), reusable by several queries."""

    def __init__(self, pos, box, r_cell, device=None, workspace=None, n_centres_max=0, others=None, reach=None):
        """others: the further position arrays (centres, hydrogens ...) the queries on this list measure distances
        to; reach: the largest distance any of them uses (default r_cell).  Both only matter for non-periodic
        (negative) box edges, which become equivalent periods (engine.effective_boxes); box_h / box_d hold those.
        A routine that has not declared its `others` does not take non-periodic axes."""
        self.device = _device(device, pos)
        self.pos = engine.as_device_positions(pos, self.device)
        self.F, self.N = int(self.pos.shape[0]), int(self.pos.shape[1])
        self.box_h = engine.as_host_boxes(box, self.F)
        if (self.box_h < 0.0).any():
            if others is None:
                raise ValueError("non-periodic (negative) box edges are not taken by this routine (Willard-Chandler "
                                 "density and binOnGrid need a periodic or explicitly bounded grid)")
            extra = [engine.as_device_positions(o, self.device) for o in others if o is not None]
            extra = [o for o in extra if o.shape[1] > 0]
            cen = None
            if extra:
                cen = torch.cat([o.to(torch.float64) for o in extra], dim=1).contiguous()
            self.box_h = engine.effective_boxes(self.box_h, self.pos, cen, float(reach if reach is not None else r_cell), self.device)
        self.box_d = torch.from_numpy(self.box_h.copy()).to(self.device)
        self.nc, self.edge_min, self.box_max = engine.plan_grid(self.box_h, r_cell)
        self.n_centres_max = max(int(n_centres_max), self.N)
        self.ws = workspace if workspace is not None else engine.Workspace(self.device)
        need = lib().wol_workspace_bytes(self.F, self.N, self.n_centres_max, ctypes.byref(self.nc))
        self.ws_ptr, self.ws_bytes = self.ws.get(need)
        with torch.cuda.device(self.device):
            check(lib().wol_cell_build(_vp(self.pos.data_ptr()), engine._dtype_code(self.pos), _vp(self.box_d.data_ptr()),
                                       self.F, self.N, ctypes.byref(self.nc), WOL_PREC_FP64, _vp(self.ws_ptr), self.ws_bytes,
                                       _stream()), "wol_cell_build")
        self.launches = lib().wol_last_launch_count()

    def status(self):
        st = (ctypes.c_int32 * 4)()
        with torch.cuda.device(self.device):
            check(lib().wol_status(_vp(self.ws_ptr), self.F, self.N, self.n_centres_max, ctypes.byref(self.nc), _stream(),
                                   ctypes.byref(st)), "wol_status")
        return tuple(st)


def _total_from_offsets(offsets, what):
    """Last entry of a 32-bit offsets array as a python int; the library marks a total beyond 32 bits with 0xFFFFFFFF."""
    total = int(offsets[-1].item()) & 0xFFFFFFFF
    if total == 0xFFFFFFFF:
        raise WolError("more than 2^32 - 2 %s in one call: materialise the frames in smaller batches" % what)
    return total


def three_body_angles(sub, pos, box, low=0.0, high=3.413, device=None):
    """getCosAngs (structureLibs/water_properties.py:210-250) for one or several frames.
    Returns (angles f64 (n_angles,), n3 int32 (F, M), offsets uint32-as-int64 (F*M+1,)): the flat angle
    array in the reference's order, the per-centre neighbour counts and where each centre's block starts."""
    device = _device(device, pos, sub)
    pos_d = engine.as_device_positions(pos, device)
    cen_d = pos_d if sub is None else engine.as_device_positions(sub, device)
    if cen_d.shape[0] != pos_d.shape[0]:
        raise ValueError("sub and pos must hold the same number of frames")
    F, N, M = int(pos_d.shape[0]), int(pos_d.shape[1]), int(cen_d.shape[1])
    if M == 0 or N == 0:
        return (torch.zeros(0, dtype=torch.float64, device=device), torch.zeros((F, M), dtype=torch.int32, device=device),
                torch.zeros(F * M + 1, dtype=torch.int64, device=device))
    ws = engine.Workspace(device)
    # (non-periodic axes become equivalent periods here, once, so that both passes see the same grid)
    box_h = engine.effective_boxes(engine.as_host_boxes(box, F), pos_d, None if sub is None else cen_d, max(float(high), 1e-3), device)
    r = engine.q3b_frames(pos_d, box_h, None if sub is None else cen_d, do_q=False, do_3body=True, low3=low, high3=high, want=("n3",),
                          workspace=ws, device=device, r_cell=max(float(high), 1e-3))
    n3 = r["n3"]
    L = lib()
    total = F * M
    offsets = torch.empty(total + 1, dtype=torch.int32, device=device)
    scratch = torch.empty(total // 2048 + 8, dtype=torch.int32, device=device)
    box_d = torch.from_numpy(box_h.copy()).to(device)
    nc, edge_min, _ = engine.plan_grid(box_h, max(float(high), 1e-3))
    ws_ptr, ws_bytes = ws.get(L.wol_workspace_bytes(F, N, M, ctypes.byref(nc)))
    with torch.cuda.device(device):
        check(L.wol_angle_offsets(_vp(n3.data_ptr()), total, _vp(offsets.data_ptr()), _vp(scratch.data_ptr()), _stream()),
              "wol_angle_offsets")
        n_angles = _total_from_offsets(offsets, "three-body angles")
        n3_max = int(n3.max().item()) if n3.numel() else 0
        if n3_max > 64:  # kMatCap of angles_fill_kernel (csrc/wol_aux.cu)
            raise WolError("getCosAngs materialises at most 64 neighbours per centre; a centre has %d inside the cutoff "
                           "(%g A).  The histogram path (tetOrderCalc / threeBodyCalc / q3b_frames) has no such limit."
                           % (n3_max, float(high)))
        angles = torch.empty(n_angles, dtype=torch.float64, device=device)
        if n_angles > 0:
            # (sub is None: NULL centres = every atom, walked in cell order)
            check(L.wol_angles_fill(None if sub is None else _vp(cen_d.data_ptr()), engine._dtype_code(cen_d), _vp(box_d.data_ptr()), F, N, M,
                                    ctypes.byref(nc), edge_min, float(low), float(high), _vp(ws_ptr), ws_bytes,
                                    _vp(offsets.data_ptr()), _vp(angles.data_ptr()), _stream()), "wol_angles_fill")
            st = (ctypes.c_int32 * 4)()
            check(L.wol_status(_vp(ws_ptr), F, N, M, ctypes.byref(nc), _stream(), ctypes.byref(st)), "wol_status")
    return angles, n3, offsets.to(torch.int64) & 0xFFFFFFFF


def neighbors_csr(sub, pos, box, low=0.0, high=3.413, device=None):
    """Neighbour lists of allNearNeighbors (sub is None) / nearNeighbors (fortran/waterlib.f90:830-862, :710-743) in CSR
    form instead of the dense logical matrix: returns (offsets int64 (F*M+1,), indices int32 (n_pairs,)); the neighbours
    of centre i of frame f are indices[offsets[f*M+i]:offsets[f*M+i+1]], frame-local atom indices, ascending."""
    device = _device(device, pos, sub)
    cells = CellList(pos, box, max(float(high), 1e-3), device=device,
                     n_centres_max=0 if sub is None else int(np.shape(sub)[-2]), others=(sub,))
    cen_d = cells.pos if sub is None else engine.as_device_positions(sub, device)
    if cen_d.shape[0] != cells.F:
        raise ValueError("sub and pos must hold the same number of frames")
    F, N, M = cells.F, cells.N, int(cen_d.shape[1])
    total = F * M
    offsets = torch.zeros(total + 1, dtype=torch.int32, device=device)
    if total == 0 or N == 0:
        return offsets.to(torch.int64), torch.zeros(0, dtype=torch.int32, device=device)
    scratch = torch.empty(total // 2048 + 8, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        def run(indices, cap):
            check(lib().wol_neighbors_csr(None if sub is None else _vp(cen_d.data_ptr()), engine._dtype_code(cen_d), _vp(cells.box_d.data_ptr()), F, N, M,
                                          ctypes.byref(cells.nc), cells.edge_min, float(low), float(high), _vp(cells.ws_ptr),
                                          cells.ws_bytes, _vp(offsets.data_ptr()), _vp(scratch.data_ptr()),
                                          _vp(indices.data_ptr()) if indices is not None else None, cap, _stream()), "wol_neighbors_csr")
        run(None, 0)
        n_pairs = _total_from_offsets(offsets, "neighbour pairs")  # sizes the output: one host read
        indices = torch.empty(n_pairs, dtype=torch.int32, device=device)
        if n_pairs:
            run(indices, n_pairs)
        torch.cuda.current_stream().synchronize()  # the cell list dies with this call
    return offsets.to(torch.int64) & 0xFFFFFFFF, indices


def histogram(x, nbins=500, bin_range=(0.0, 180.0), tet_window=(100.0, 120.0), device=None):
    """np.histogram(x, bins=nbins, range=bin_range) counts (int64 CUDA tensor) and the
    (count, sum cos, sum cos^2) of tetrahedralMetrics' window (water_properties.py:328-335)."""
    device = _device(device, x)
    xd = _f64(x, device).reshape(-1)
    hist = torch.zeros(nbins, dtype=torch.int64, device=device)
    tet = torch.zeros(3, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        check(lib().wol_histogram(_vp(xd.data_ptr()), xd.numel(), float(bin_range[0]), float(bin_range[1]), int(nbins),
                                  _vp(hist.data_ptr()), float(tet_window[0]), float(tet_window[1]), _vp(tet.data_ptr()),
                                  _stream()), "wol_histogram")
    return hist, tet


def neighbor_matrix(sub, pos, box, low, high, device=None):
    """nearNeighbors / allNearNeighbors (fortran/waterlib.f90:710-743, :830-862) -> (m, n) int32 CUDA tensor."""
    device = _device(device, pos, sub)
    pos_d = _f64(pos, device, (3,)).reshape(-1, 3)
    sub_d = pos_d if sub is None else _f64(sub, device, (3,)).reshape(-1, 3)
    box_d = _box3(box, device)
    m, n = int(sub_d.shape[0]), int(pos_d.shape[0])
    out = torch.empty((m, n), dtype=torch.int32, device=device)
    if m and n:
        with torch.cuda.device(device):
            check(lib().wol_neighbor_matrix(_vp(sub_d.data_ptr()), WOL_F64, m, _vp(pos_d.data_ptr()), WOL_F64, n,
                                            _vp(box_d.data_ptr()), float(low), float(high), _vp(out.data_ptr()), _stream()),
                  "wol_neighbor_matrix")
    return out


def _reimage(pos, ref, box, mode, device):
    device = _device(device, pos, ref)
    pos_d = _f64(pos, device, (3,)).reshape(-1, 3)
    ref_d = _f64(ref, device).reshape(-1)
    if ref_d.numel() != 3:
        raise ValueError("reference position must hold 3 values")
    box_d = _box3(box, device)
    n = int(pos_d.shape[0])
    out = torch.empty((n, 3) if mode == 0 else (n,), dtype=torch.float64, device=device)
    if n:
        with torch.cuda.device(device):
            check(lib().wol_reimage(_vp(pos_d.data_ptr()), n, _vp(ref_d.data_ptr()), _vp(box_d.data_ptr()),
                                    _vp(out.data_ptr()), mode, _stream()), "wol_reimage")
    return out


def reimage(pos, ref, box, device=None):
    """reimage (fortran/waterlib.f90:32-47): ref + minimum-image(pos - ref) -> (n,3)."""
    return _reimage(pos, ref, box, 0, device)


def lsidists(ref, neigh, box, device=None):
    """lsiDists (fortran/waterlib.f90:900-918): minimum-image distances ref -> neigh, (n,)."""
    return _reimage(neigh, ref, box, 1, device)


def tetracosang(ref, neigh, box, device=None):
    """tetraCosAng (fortran/waterlib.f90:867-895): (k,k) angles in degrees about ref, zero diagonal."""
    device = _device(device, neigh, ref)
    nd = _f64(neigh, device, (3,)).reshape(-1, 3)
    rd = _f64(ref, device).reshape(-1)
    box_d = _box3(box, device)
    k = int(nd.shape[0])
    out = torch.zeros((k, k), dtype=torch.float64, device=device)
    if k:
        with torch.cuda.device(device):
            check(lib().wol_tetracosang(_vp(rd.data_ptr()), _vp(nd.data_ptr()), k, _vp(box_d.data_ptr()), _vp(out.data_ptr()),
                                        _stream()), "wol_tetracosang")
    return out


def hbond_counts(acc, don, donh, box, dist_cut=3.5, ang_cut=150.0, dense=False, pairs=False, device=None,
                 cells=None):
    """generalHbonds (fortran/waterlib.f90:1156-1210) for one or several frames.
    acc (F,Na,3) acceptors, don (F,Nd,3) donor heavy atoms (one entry per hydrogen), donh (F,Nd,3) hydrogens.
    Returns dict(acc_count (F,Na) int32, don_count (F,Nd) int32 [, dense (F,Na,Nd) int32][, pairs (P,2) int32
    rows (frame*Na + acceptor, donor)]).  `cells`: a CellList already built over `don` with r_cell >= dist_cut."""
    device = _device(device, acc, don, donh)
    acc_d = engine.as_device_positions(acc, device)
    donh_d = engine.as_device_positions(donh, device)
    F, Na = int(acc_d.shape[0]), int(acc_d.shape[1])
    if cells is None:
        don_d = engine.as_device_positions(don, device)
        if don_d.shape != donh_d.shape:
            raise ValueError("Number of donor hydrogens and heavy-atoms do not match.")  # waterlib.f90:1171-1174
        Nd = int(don_d.shape[1])
    else:
        Nd = cells.N
        if int(donh_d.shape[1]) != Nd:
            raise ValueError("Number of donor hydrogens and heavy-atoms do not match.")
    res = {"acc_count": torch.zeros((F, Na), dtype=torch.int32, device=device),
           "don_count": torch.zeros((F, Nd), dtype=torch.int32, device=device)}
    if dense:
        res["dense"] = torch.zeros((F, Na, Nd), dtype=torch.int32, device=device)
    if Na == 0 or Nd == 0:
        if pairs:
            res["pairs"] = torch.zeros((0, 2), dtype=torch.int32, device=device)
        return res
    if cells is None:
        cells = CellList(don_d, box, max(float(dist_cut), 1e-3), device=device, others=(acc_d, donh_d))
    a = HbondArgs()
    a.struct_size = ctypes.sizeof(HbondArgs)
    a.n_frames, a.n_acc, a.n_don = F, Na, Nd
    a.acc_dtype, a.donh_dtype = engine._dtype_code(acc_d), engine._dtype_code(donh_d)
    a.acc, a.donh, a.box = acc_d.data_ptr(), donh_d.data_ptr(), cells.box_d.data_ptr()
    a.workspace, a.workspace_bytes = cells.ws_ptr, cells.ws_bytes
    a.nc = cells.nc
    a.edge_min = cells.edge_min
    a.dist_cut, a.ang_cut = float(dist_cut), float(ang_cut)
    a.acc_count, a.don_count = res["acc_count"].data_ptr(), res["don_count"].data_ptr()
    a.dense = res["dense"].data_ptr() if dense else None
    counter = torch.zeros(1, dtype=torch.int32, device=device)
    a.pair_counter = counter.data_ptr()
    with torch.cuda.device(device):
        if pairs:
            # pass 1 counts (the per-acceptor sums are what bounds the list), pass 2 fills
            check(lib().wol_hbond_counts(ctypes.byref(a), _stream()), "wol_hbond_counts")
            n_pairs = int(res["acc_count"].sum().item())
            plist = torch.empty((max(n_pairs, 1), 2), dtype=torch.int32, device=device)
            res["don_count"].zero_()
            a.pairs, a.pair_capacity = plist.data_ptr(), n_pairs
            check(lib().wol_hbond_counts(ctypes.byref(a), _stream()), "wol_hbond_counts")
            plist = plist[:n_pairs]
            # the kernel appends in no particular order: sort rows by (acceptor, donor), the order the
            # reference walks its matrix in (water_properties.py:705-713)
            key = plist[:, 0].to(torch.int64) * (Nd + 1) + plist[:, 1].to(torch.int64)
            res["pairs"] = plist[torch.argsort(key)]
        else:
            check(lib().wol_hbond_counts(ctypes.byref(a), _stream()), "wol_hbond_counts")
    res["_keep"] = (acc_d, donh_d, cells)
    return res


def shell_mask(sol, wat, box, cutoff=4.0, low=0.0, device=None, cells=None):
    """Hydration-shell selection (structureLibs/orderParam_lib.py:495-498): int32 mask (F, Nw), 1 where a
    water lies within (low, cutoff] of any solute atom."""
    device = _device(device, wat, sol)
    sol_d = engine.as_device_positions(sol, device)
    if cells is None:
        wat_d = engine.as_device_positions(wat, device)
        F, Nw = int(wat_d.shape[0]), int(wat_d.shape[1])
    else:
        F, Nw = cells.F, cells.N
    mask = torch.zeros((F, Nw), dtype=torch.int32, device=device)
    Ns = int(sol_d.shape[1])
    if Ns == 0 or Nw == 0:
        return mask
    if cells is None:
        cells = CellList(wat_d, box, max(float(cutoff), 1e-3), device=device, others=(sol_d,))
    with torch.cuda.device(device):
        check(lib().wol_shell_mask(_vp(sol_d.data_ptr()), engine._dtype_code(sol_d), Ns, _vp(cells.box_d.data_ptr()), F, Nw,
                                   ctypes.byref(cells.nc), cells.edge_min, float(low), float(cutoff), _vp(cells.ws_ptr),
                                   cells.ws_bytes, _vp(mask.data_ptr()), _stream()), "wol_shell_mask")
    return mask


def hbond_locations(pairs, acc, donh, box, device=None):
    """HBloc of HBondsGeneral (structureLibs/water_properties.py:709-714) for a pair list from hbond_counts."""
    device = _device(device, acc, donh)
    acc_d = engine.as_device_positions(acc, device)
    donh_d = engine.as_device_positions(donh, device)
    F = int(acc_d.shape[0])
    box_d = torch.from_numpy(engine.as_host_boxes(box, F).copy()).to(device)
    pairs = pairs.to(device=device, dtype=torch.int32).contiguous()
    n = int(pairs.shape[0])
    out = torch.empty((n, 3), dtype=torch.float64, device=device)
    if n:
        with torch.cuda.device(device):
            check(lib().wol_hbond_locations(_vp(pairs.data_ptr()), n, _vp(acc_d.data_ptr()), engine._dtype_code(acc_d),
                                            int(acc_d.shape[1]), _vp(donh_d.data_ptr()), engine._dtype_code(donh_d),
                                            int(donh_d.shape[1]), _vp(box_d.data_ptr()), _vp(out.data_ptr()), _stream()),
                  "wol_hbond_locations")
    return out


# ---- slab / interface routines (BASELINE config 4) and histrr3b ----------------------------------------

def willard_density(pos, box, smoothlen=2.4, grid=None, points=None, want_normals=True, device=None):
    """WillardDensityField (grid=(gridx, gridy, gridz)) or WillardDensityPoints (points=(n,3)),
    fortran/waterlib.f90:1286-1398.  Returns (densvals, densnorms) CUDA tensors shaped (nx,ny,nz) /
    (nx,ny,nz,3) or (n,) / (n,3)."""
    if (grid is None) == (points is None):
        raise ValueError("give exactly one of grid=(gridx, gridy, gridz) or points=")
    device = _device(device, pos)
    cells = CellList(pos, box, 3.0 * float(smoothlen) * (1.0 + 1e-9), device=device)
    if cells.F != 1:
        raise ValueError("the density field is evaluated one frame at a time")
    if grid is not None:
        gx, gy, gz = (_f64(np.asarray(g, dtype=np.float64).reshape(-1) if not isinstance(g, torch.Tensor) else g.reshape(-1), device)
                      for g in grid)
        nx, ny, nz = int(gx.numel()), int(gy.numel()), int(gz.numel())
        n = nx * ny * nz
        shape = (nx, ny, nz)
        pts_ptr, g_ptrs = None, (_vp(gx.data_ptr()), _vp(gy.data_ptr()), _vp(gz.data_ptr()))
    else:
        pts = _f64(points, device, (3,)).reshape(-1, 3)
        n = int(pts.shape[0])
        nx = ny = nz = 0
        shape = (n,)
        pts_ptr, g_ptrs = _vp(pts.data_ptr()), (None, None, None)
    dens = torch.empty(shape, dtype=torch.float64, device=device)
    norms = torch.empty(shape + (3,), dtype=torch.float64, device=device) if want_normals else None
    with torch.cuda.device(device):
        check(lib().wol_willard_density(pts_ptr, n, g_ptrs[0], g_ptrs[1], g_ptrs[2], nx, ny, nz, _vp(cells.box_d.data_ptr()), cells.N,
                                        ctypes.byref(cells.nc), cells.edge_min, float(smoothlen), _vp(cells.ws_ptr), cells.ws_bytes,
                                        _vp(dens.data_ptr()), _vp(norms.data_ptr()) if norms is not None else None, _stream()),
              "wol_willard_density")
        torch.cuda.current_stream().synchronize()  # the cell list (cells) dies with this frame
    return dens, norms


def iso_points(dens, grid, level, device=None):
    """Vertices of the iso-surface dens == level on the rectilinear grid (gridx, gridy, gridz): one per grid edge whose
    end values straddle the level, linearly interpolated -- the vertex set skimage.measure.marching_cubes gives
    densityGrid (structureLibs/surface_library.py:202).  Returns a CUDA tensor (n, 3) f64, ordered by node index then axis."""
    device = _device(device, dens)
    gx, gy, gz = (_f64(np.asarray(g, dtype=np.float64).reshape(-1) if not isinstance(g, torch.Tensor) else g.reshape(-1), device)
                  for g in grid)
    nx, ny, nz = int(gx.numel()), int(gy.numel()), int(gz.numel())
    d = _f64(dens, device).reshape(-1)
    if int(d.numel()) != nx * ny * nz:
        raise ValueError("dens has %d values, the grid %d x %d x %d nodes" % (d.numel(), nx, ny, nz))
    if nx * ny * nz == 0:
        return torch.zeros((0, 3), dtype=torch.float64, device=device)
    nbytes = int(lib().wol_iso_scratch_bytes(nx, ny, nz))
    scratch = torch.empty(nbytes // 4 + 4, dtype=torch.int32, device=device)
    n_total = torch.zeros(1, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        def run(points, cap):
            check(lib().wol_iso_points(_vp(d.data_ptr()), _vp(gx.data_ptr()), _vp(gy.data_ptr()), _vp(gz.data_ptr()), nx, ny, nz,
                                       float(level), _vp(scratch.data_ptr()), nbytes, _vp(points.data_ptr()) if points is not None else None,
                                       cap, _vp(n_total.data_ptr()), _stream()), "wol_iso_points")
        run(None, 0)
        n = int(n_total.item())  # the vertex count sizes the output: one host read
        pts = torch.empty((n, 3), dtype=torch.float64, device=device)
        if n:
            run(pts, n)
    return pts


def iso_surface(dens, grid, level, want_areas=True, device=None):
    """Triangulated iso-surface dens == level: (verts (n, 3) f64, faces (m, 3) int32 into verts, areas (m,) f64 or None), all
    CUDA tensors.  Vertices are iso_points' (same order); faces come from the library's own marching-cubes table
    (scripts/make_mc_table.py: watertight, normals towards lower values; PARITY UNPINNED against skimage, which the
    reference calls at structureLibs/surface_library.py:202 and which is not vendored); areas follow the reference's
    triangleArea (fortran/imagelib.f90:254-267 = |v1 x v2|, twice the geometric area)."""
    device = _device(device, dens)
    gx, gy, gz = (_f64(np.asarray(g, dtype=np.float64).reshape(-1) if not isinstance(g, torch.Tensor) else g.reshape(-1), device)
                  for g in grid)
    nx, ny, nz = int(gx.numel()), int(gy.numel()), int(gz.numel())
    d = _f64(dens, device).reshape(-1)
    if int(d.numel()) != nx * ny * nz:
        raise ValueError("dens has %d values, the grid %d x %d x %d nodes" % (d.numel(), nx, ny, nz))
    empty = (torch.zeros((0, 3), dtype=torch.float64, device=device), torch.zeros((0, 3), dtype=torch.int32, device=device),
             torch.zeros(0, dtype=torch.float64, device=device) if want_areas else None)
    if nx * ny * nz == 0:
        return empty
    L = lib()
    vbytes, fbytes = int(L.wol_iso_scratch_bytes(nx, ny, nz)), int(L.wol_iso_face_scratch_bytes(nx, ny, nz))
    vscr = torch.empty(vbytes // 4 + 4, dtype=torch.int32, device=device)
    fscr = torch.empty(fbytes // 4 + 4, dtype=torch.int32, device=device)
    n_total = torch.zeros(1, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        def verts(points, cap):
            check(L.wol_iso_points(_vp(d.data_ptr()), _vp(gx.data_ptr()), _vp(gy.data_ptr()), _vp(gz.data_ptr()), nx, ny, nz, float(level),
                                   _vp(vscr.data_ptr()), vbytes, _vp(points.data_ptr()) if points is not None else None, cap,
                                   _vp(n_total.data_ptr()), _stream()), "wol_iso_points")

        def tris(faces, cap, pts, areas):
            check(L.wol_iso_faces(_vp(d.data_ptr()), nx, ny, nz, float(level), _vp(vscr.data_ptr()), _vp(fscr.data_ptr()), fbytes,
                                  _vp(faces.data_ptr()) if faces is not None else None, cap, _vp(pts.data_ptr()) if pts is not None else None,
                                  _vp(areas.data_ptr()) if areas is not None else None, _vp(n_total.data_ptr()), _stream()), "wol_iso_faces")
        verts(None, 0)
        n = int(n_total.item())
        if n == 0:
            return empty
        pts = torch.empty((n, 3), dtype=torch.float64, device=device)
        verts(pts, n)
        tris(None, 0, None, None)
        m = int(n_total.item())
        faces = torch.empty((m, 3), dtype=torch.int32, device=device)
        areas = torch.empty(m, dtype=torch.float64, device=device) if want_areas else None
        if m:
            tris(faces, m, pts, areas)
    return pts, faces, areas


def interface_water(pos, gridpos, gridnorm, cutoff, box, want_surfclose=True, device=None):
    """InterfaceWater (fortran/waterlib.f90:1414-1469) -> dict(watclose int32 (n,) 0-based / -1, surfclose int32
    (ng,), numwater int, allwatdists f64 (n,))."""
    device = _device(device, pos, gridpos)
    pos_d = _f64(pos, device, (3,)).reshape(-1, 3)
    gp = _f64(gridpos, device, (3,)).reshape(-1, 3)
    gn = _f64(gridnorm, device, (3,)).reshape(-1, 3)
    if gp.shape != gn.shape:
        raise ValueError("gridpos and gridnorm differ in shape")
    box_d = _box3(box, device)
    n, ng = int(pos_d.shape[0]), int(gp.shape[0])
    watclose = torch.full((n,), -1, dtype=torch.int32, device=device)
    surfclose = torch.full((ng,), -1, dtype=torch.int32, device=device) if want_surfclose else None
    numwater = torch.zeros(1, dtype=torch.int32, device=device)
    dists = torch.zeros(n, dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        check(lib().wol_interface_water(_vp(pos_d.data_ptr()), n, _vp(gp.data_ptr()), _vp(gn.data_ptr()), ng, float(cutoff),
                                        _vp(box_d.data_ptr()), _vp(watclose.data_ptr()),
                                        _vp(surfclose.data_ptr()) if surfclose is not None else None, _vp(numwater.data_ptr()),
                                        _vp(dists.data_ptr()), _stream()), "wol_interface_water")
    return {"watclose": watclose, "surfclose": surfclose, "numwater": numwater, "allwatdists": dists, "_keep": (pos_d, gp, gn, box_d)}


def profile_bins(value, coord, lo, width, nbins, out=None, device=None):
    """Depth-binned profile: (count int64, sum f64, sumsq f64) per bin of floor((coord - lo) / width), accumulated
    into `out` when given (a tuple of three tensors)."""
    device = _device(device, value, coord)
    v = _f64(value, device).reshape(-1)
    c = _f64(coord, device).reshape(-1)
    if v.numel() != c.numel():
        raise ValueError("value and coord differ in length")
    if out is None:
        out = (torch.zeros(nbins, dtype=torch.int64, device=device), torch.zeros(nbins, dtype=torch.float64, device=device),
               torch.zeros(nbins, dtype=torch.float64, device=device))
    with torch.cuda.device(device):
        check(lib().wol_profile_bins(_vp(v.data_ptr()), _vp(c.data_ptr()), v.numel(), float(lo), float(width), int(nbins),
                                     _vp(out[0].data_ptr()), _vp(out[1].data_ptr()), _vp(out[2].data_ptr()), _stream()),
              "wol_profile_bins")
        torch.cuda.current_stream().synchronize()
    return out


_CEIL_TABLES = {}


def histrr3b(pos, box, dist_width, d_num, ang_width, a_num, device=None):
    """histrr3b (fortran/waterlib.f90:1550-1593) -> int64 counts (d_num, d_num, a_num) CUDA tensor."""
    device = _device(device, pos)
    key = (float(ang_width), int(a_num), str(device))
    table = _CEIL_TABLES.get(key)
    if table is None:
        from ._capi import WOL_TABLE_EXTRA
        host = np.zeros(a_num + 1 + WOL_TABLE_EXTRA, dtype=np.float64)
        check(lib().wol_angle_table_ceil(float(ang_width), int(a_num), host.ctypes.data_as(_vp)), "wol_angle_table_ceil")
        table = torch.from_numpy(host).to(device)
        _CEIL_TABLES[key] = table
    cells = CellList(pos, box, float(dist_width) * int(d_num) * (1.0 + 1e-9), device=device, others=())
    if cells.F != 1:
        raise ValueError("histrr3b takes one frame")
    hist = torch.zeros((d_num, d_num, a_num), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        check(lib().wol_histrr3b(_vp(cells.box_d.data_ptr()), cells.N, ctypes.byref(cells.nc), cells.edge_min, float(dist_width),
                                 int(d_num), float(ang_width), int(a_num), _vp(table.data_ptr()), _vp(cells.ws_ptr), cells.ws_bytes,
                                 _vp(hist.data_ptr()), _stream()), "wol_histrr3b")
        cells.status()  # raises if a neighbour list overflowed; also synchronises
    return hist


def lsi(sub, pos, box, low=0.0, high=3.7, device=None):
    """getLSI (structureLibs/water_properties.py:252-311) for one or several frames -> (lsi f64 (F, M), num int32
    (F, M)); num == 0 marks centres without a value (fewer than two neighbours or an empty next shell)."""
    device = _device(device, pos, sub)
    pos_d = engine.as_device_positions(pos, device)
    cen_d = pos_d if sub is None else engine.as_device_positions(sub, device)
    if cen_d.shape[0] != pos_d.shape[0]:
        raise ValueError("sub and pos must hold the same number of frames")
    F, N, M = int(pos_d.shape[0]), int(pos_d.shape[1]), int(cen_d.shape[1])
    out = torch.zeros((F, M), dtype=torch.float64, device=device)
    num = torch.zeros((F, M), dtype=torch.int32, device=device)
    if M == 0 or N == 0:
        return out, num
    cells = CellList(pos_d, box, (float(high) + 3.7) * (1.0 + 1e-9), device=device, n_centres_max=M,
                     others=() if sub is None else (cen_d,))
    with torch.cuda.device(device):
        check(lib().wol_lsi(None if sub is None else _vp(cen_d.data_ptr()), engine._dtype_code(cen_d), _vp(cells.box_d.data_ptr()), F, N, M,
                            ctypes.byref(cells.nc), cells.edge_min, float(low), float(high), _vp(cells.ws_ptr), cells.ws_bytes,
                            _vp(out.data_ptr()), _vp(num.data_ptr()), _stream()), "wol_lsi")
        cells.status()
    return out, num


# ---- same-sweep observables: pair-distance histograms, psi ------------------------------------------------------

_FOUR_THIRDS = float(np.float32(4.0) / np.float32(3.0))  # the Fortran literal (4./3.) is single precision (waterlib.f90:228)
_PI_RDF = 3.141592653589                                 # and its pi is truncated (waterlib.f90:204)


def pair_hist(mode, pos1, pos2, box, binwidth, totbins, device=None):
    """Counts of RadialDist (mode 0: pos1 = Pos1, pos2 = Pos2), RadialDistSame (mode 1: pos1 only) or
    PairDistanceHistogram (mode 2) -> int64 CUDA tensor (totbins,)."""
    device = _device(device, pos1, pos2)
    p1 = _f64(pos1, device, (3,)).reshape(-1, 3)
    p2 = p1 if pos2 is None else _f64(pos2, device, (3,)).reshape(-1, 3)
    outer, inner = (p2, p1) if mode == 0 else ((p1, p1) if mode == 1 else (p1, p2))
    counts = torch.zeros(int(totbins), dtype=torch.int64, device=device)
    if outer.shape[0] == 0 or inner.shape[0] == 0:
        return counts
    cells = CellList(inner, box, float(binwidth) * int(totbins) * (1.0 + 1e-9), device=device, others=(outer,))
    with torch.cuda.device(device):
        check(lib().wol_pair_hist(int(mode), _vp(outer.data_ptr()), WOL_F64, int(outer.shape[0]), _vp(cells.box_d.data_ptr()), cells.N,
                                  ctypes.byref(cells.nc), cells.edge_min, float(binwidth), int(totbins), _vp(cells.ws_ptr),
                                  cells.ws_bytes, _vp(counts.data_ptr()), _stream()), "wol_pair_hist")
        torch.cuda.current_stream().synchronize()
    return counts


def radial_dist_plane(pos1, pos2, box, binwidth, totbins, bulkdens, device=None):
    """RadialDistPlane (fortran/waterlib.f90:237-314) -> (counts int64 CUDA tensor (totbins, totbins), number of slab atoms
    whose bin index is <= 0 -- an out-of-bounds write in the Fortran, skipped here)."""
    device = _device(device, pos2, pos1)
    p1 = _f64(pos1, device, (3,)).reshape(-1, 3)
    if p1.shape[0] != 3:
        raise ValueError("pos1 must hold the three points that span the plane, shape (3, 3)")
    p2 = _f64(pos2, device, (3,)).reshape(-1, 3)
    b = _box3(box, device)
    counts = torch.zeros((int(totbins), int(totbins)), dtype=torch.int64, device=device)
    bad = torch.zeros(1, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(lib().wol_radial_dist_plane(_vp(p1.data_ptr()), _vp(p2.data_ptr()), int(p2.shape[0]), _vp(b.data_ptr()), float(binwidth),
                                          int(totbins), float(bulkdens), _vp(counts.data_ptr()), _vp(bad.data_ptr()), _stream()),
              "wol_radial_dist_plane")
    return counts, int(bad.item())


def histogram2d(x, y, xedges, yedges, out=None, device=None):
    """np.histogram2d(x, y, bins=(xedges, yedges))[0] as an int64 CUDA tensor (accumulated into `out` when given)."""
    device = _device(device, x, y)
    xd, yd = _f64(x, device).reshape(-1), _f64(y, device).reshape(-1)
    if xd.numel() != yd.numel():
        raise ValueError("x and y must have the same length")
    xe, ye = _f64(np.asarray(xedges, dtype=np.float64), device).reshape(-1), _f64(np.asarray(yedges, dtype=np.float64), device).reshape(-1)
    if out is None:
        out = torch.zeros((xe.numel() - 1, ye.numel() - 1), dtype=torch.int64, device=device)
    with torch.cuda.device(device):
        check(lib().wol_histogram2d(_vp(xd.data_ptr()), _vp(yd.data_ptr()), int(xd.numel()), _vp(xe.data_ptr()), int(xe.numel()),
                                    _vp(ye.data_ptr()), int(ye.numel()), _vp(out.data_ptr()), _stream()), "wol_histogram2d")
    return out


def rdf_normalise(counts, n_norm, binwidth, bulkdens):
    """counts(k) / (N * BulkDens * (4./3.) * pi * binwidth**3 * (k**3 - (k-1)**3)) in the Fortran's order of
    operations (waterlib.f90:227-229); O(totbins) on the host."""
    c = counts.cpu().numpy().astype(np.float64)
    k = np.arange(1, c.size + 1, dtype=np.int64)
    shell = (k ** 3 - (k - 1) ** 3).astype(np.float64)
    denom = ((((float(n_norm) * float(bulkdens)) * _FOUR_THIRDS) * _PI_RDF) * ((binwidth * binwidth) * binwidth)) * shell
    return c / denom


def psi(sub, pos, box, low=0.0, high=10.0, device=None):
    """getOrderParamPsi (structureLibs/water_properties.py:393-433) for one or several frames -> f64 (F, M)."""
    device = _device(device, pos, sub)
    pos_d = engine.as_device_positions(pos, device)
    cen_d = pos_d if sub is None else engine.as_device_positions(sub, device)
    if cen_d.shape[0] != pos_d.shape[0]:
        raise ValueError("sub and pos must hold the same number of frames")
    F, N, M = int(pos_d.shape[0]), int(pos_d.shape[1]), int(cen_d.shape[1])
    out = torch.zeros((F, M), dtype=torch.float64, device=device)
    if M == 0 or N == 0:
        return out
    cells = CellList(pos_d, box, max(float(high), 1e-3) * (1.0 + 1e-9), device=device, n_centres_max=M,
                     others=() if sub is None else (cen_d,))
    with torch.cuda.device(device):
        check(lib().wol_psi(_vp(cen_d.data_ptr()), engine._dtype_code(cen_d), _vp(cells.box_d.data_ptr()), F, N, M,
                            ctypes.byref(cells.nc), cells.edge_min, float(low), float(high), _vp(cells.ws_ptr), cells.ws_bytes,
                            _vp(out.data_ptr()), _stream()), "wol_psi")
        cells.status()
    return out


def density_field(pos, box, grid, device=None):
    """DensityField (fortran/waterlib.f90:1219-1268): waters per cube of edge gridx[1] - gridx[0] around each grid point,
    divided by the cube volume -> f64 CUDA tensor (nx, ny, nz)."""
    device = _device(device, pos)
    gx, gy, gz = (_f64(np.asarray(g, dtype=np.float64).reshape(-1) if not isinstance(g, torch.Tensor) else g.reshape(-1), device) for g in grid)
    if gx.numel() < 2:
        raise ValueError("gridx needs at least two points (the bin width is gridx[1] - gridx[0])")
    binwidth = float((gx[1] - gx[0]).item())
    if not binwidth > 0.0:
        raise ValueError("gridx must be ascending")
    cells = CellList(pos, box, 0.5 * binwidth * (1.0 + 1e-9), device=device)
    if cells.F != 1:
        raise ValueError("the density field is evaluated one frame at a time")
    nx, ny, nz = int(gx.numel()), int(gy.numel()), int(gz.numel())
    dens = torch.empty((nx, ny, nz), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        check(lib().wol_density_field(_vp(gx.data_ptr()), _vp(gy.data_ptr()), _vp(gz.data_ptr()), nx, ny, nz, binwidth,
                                      _vp(cells.box_d.data_ptr()), cells.N, ctypes.byref(cells.nc), cells.edge_min, _vp(cells.ws_ptr),
                                      cells.ws_bytes, _vp(dens.data_ptr()), _stream()), "wol_density_field")
        torch.cuda.current_stream().synchronize()
    return dens


def water_orient(opos, hpos, box, refvec=(0.0, 0.0, 1.0), device=None):
    """watOrient (fortran/waterlib.f90:973-1011) for one frame (N,3)/(2N,3) or several (F,N,3)/(F,2N,3):
    -> (angDip, angPlane) float64 CUDA tensors (F, N): angles in degrees of the water dipoles / molecular-plane normals
    with refvec.  hpos rows 2i, 2i+1 are the hydrogens of oxygen i."""
    device = _device(device, opos, hpos)
    o = _f64(opos, device, (3,))
    h = _f64(hpos, device, (3,))
    o = o.reshape(1, -1, 3) if o.dim() == 2 else o
    h = h.reshape(1, -1, 3) if h.dim() == 2 else h
    F, N = int(o.shape[0]), int(o.shape[1])
    if h.shape[0] != F or h.shape[1] != 2 * N:
        raise ValueError("Number of hydrogens must be two times number of oxygens.")  # the Fortran STOPs here (:988-991)
    box_d = torch.from_numpy(engine.as_host_boxes(box, F).copy()).to(device)
    ref = (ctypes.c_double * 3)(*[float(v) for v in np.asarray(refvec, dtype=np.float64).reshape(-1)[:3]])
    dip = torch.empty((F, N), dtype=torch.float64, device=device)
    plane = torch.empty((F, N), dtype=torch.float64, device=device)
    with torch.cuda.device(device):
        check(lib().wol_water_orient(_vp(o.data_ptr()), _vp(h.data_ptr()), _vp(box_d.data_ptr()), F, N, ref, _vp(dip.data_ptr()),
                                     _vp(plane.data_ptr()), _stream()), "wol_water_orient")
    return dip, plane


def bin_on_grid(opos, xbins, ybins, zbins, device=None):
    """binOnGrid (fortran/waterlib.f90:1047-1099): atoms per cubic bin, counted inside the bin's inscribed sphere only
    -> int32 CUDA tensor (nx-1, ny-1, nz-1).  The bins must be uniform cubes."""
    device = _device(device, opos)
    o = _f64(opos, device, (3,)).reshape(-1, 3)
    edges = [np.asarray(b.detach().cpu() if isinstance(b, torch.Tensor) else b, dtype=np.float64).reshape(-1) for b in (xbins, ybins, zbins)]
    if any(e.size < 2 for e in edges):
        raise ValueError("each axis needs at least two bin edges")
    binwidth = edges[0][1] - edges[0][0]
    if edges[1][1] - edges[1][0] != binwidth or edges[2][1] - edges[2][0] != binwidth:
        raise ValueError("Must break volume into CUBES. Currently, bin-widths do not match.")  # the Fortran STOPs here (:1060-1063)
    xb, yb, zb = (torch.from_numpy(np.ascontiguousarray(e)).to(device) for e in edges)
    nx, ny, nz = (int(e.size) for e in edges)
    out = torch.empty((nx - 1, ny - 1, nz - 1), dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(lib().wol_bin_on_grid(_vp(o.data_ptr()), int(o.shape[0]), _vp(xb.data_ptr()), _vp(yb.data_ptr()), _vp(zb.data_ptr()), nx, ny,
                                    nz, float(binwidth), _vp(out.data_ptr()), _stream()), "wol_bin_on_grid")
        torch.cuda.current_stream().synchronize()  # xb, yb, zb die with this call
    return out


def components(adj, device=None):
    """Connected components of a symmetric 0/1 matrix -> int32 CUDA tensor labels (n,), labels[i] = smallest member."""
    device = _device(device, adj)
    a = adj if isinstance(adj, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(adj)))
    a = (a == 1).to(device=device, dtype=torch.int32).contiguous()
    if a.dim() != 2 or a.shape[0] != a.shape[1]:
        raise ValueError("adjacency matrix must be square")
    n = int(a.shape[0])
    labels = torch.empty(n, dtype=torch.int32, device=device)
    flag = torch.zeros(1, dtype=torch.int32, device=device)
    with torch.cuda.device(device):
        check(lib().wol_components(_vp(a.data_ptr()), n, _vp(labels.data_ptr()), _vp(flag.data_ptr()), _stream()), "wol_components")
    return labels
