"""Drop-in for the hot-path pieces of the reference's ``structureLibs/surface_library.py``: the Willard-Chandler
density field of ``densityGrid`` (:170-210) and the slab composition BASELINE config 4 names
("instantaneous-interface depth-binned q profiles"), which the reference has the ingredients for
(``wl.willarddensityfield``, ``wl.interfacewater``, ``wp.getOrderParamq``) but no function.

``instantaneousInterface`` extracts the iso-surface on the device (the vertex set marching cubes would produce, with
the exact Willard-Chandler normals at the vertices) so that the slab pipeline density -> interface -> depth -> profile
never leaves the GPU.  Triangulation (``skimage.measure.marching_cubes``, surface_library.py:202), meshing and plotting
stay third-party work outside the hot path: ``densityGrid`` hands the density field to skimage only if it is installed.
"""
import numpy as np
import torch

from .. import engine, routines


def densityField(heavyPos, watPos, thisbox, nBins=81):
    """The grid and density field densityGrid builds before marching cubes (surface_library.py:170-199):
    returns (dens (n,n,n), dens_norm (n,n,n,3), (xSpan, ySpan, zSpan), (xSpace, ySpace, zSpace)), n = nBins - 1."""
    heavyPos = np.asarray(heavyPos, dtype=np.float64)
    thisbox = np.asarray(thisbox, dtype=np.float64).reshape(1, 3)
    allMin, allMax = np.min(heavyPos), np.max(heavyPos)
    span = np.linspace(allMin - thisbox[0, 0] / 2.0, allMax + thisbox[0, 0] / 2.0, nBins).reshape(1, nBins)
    space = span[:, 1] - span[:, 0]
    span = span[:, :-1] + space
    dens, norms = routines.willard_density(watPos, thisbox, 2.4, grid=(span, span, span))
    return dens.cpu().numpy(), norms.cpu().numpy(), (span, span, span), (space, space, space)


def densityGrid(heavyPos, watPos, thisbox, level=0.016, minFrac=0.7):
    """Instantaneous-interface mesh (verts, faces) around heavyPos (reference surface_library.py:170-210).  The
    density field comes from the CUDA kernel; the iso-surface needs scikit-image, exactly as in the reference."""
    dens, _norms, _spans, (xSpace, ySpace, zSpace) = densityField(heavyPos, watPos, thisbox)
    from skimage import measure  # third-party, not part of the hot path (ImportError if absent, as in the reference)
    verts, faces, _n, _v = measure.marching_cubes(dens, level, spacing=(xSpace, ySpace, zSpace))
    allMin = np.min(np.asarray(heavyPos))
    verts = verts - allMin
    verts = verts - 0.5 * np.max(verts)
    return verts, faces


def interfaceMesh(watPos, thisbox, grid=None, spacing=2.0, level=0.016, smoothlen=2.4):
    """The triangulated Willard-Chandler interface of one frame -- what densityGrid gets from
    skimage.measure.marching_cubes (structureLibs/surface_library.py:197-202): (verts (n,3), faces (m,3) int32,
    normals (n,3), values (n,), areas (m,)).  verts / normals as instantaneousInterface (normals = the density gradient
    at the vertices, pointing to HIGHER density as skimage's gradient direction does for this field), values = the level;
    faces from the library's own marching-cubes table, oriented towards lower density (parity unpinned: skimage is not
    vendored); areas by the reference's triangleArea rule (fortran/imagelib.f90:254-267).  CUDA tensors for tensor
    input, numpy arrays otherwise."""
    box = np.asarray(thisbox.detach().cpu() if isinstance(thisbox, torch.Tensor) else thisbox, dtype=np.float64).reshape(-1)[:3]
    if grid is None:
        grid = []
        for d in range(3):
            n = max(int(np.ceil(box[d] / float(spacing))), 1)
            grid.append((np.arange(n) + 0.5) * (box[d] / n))
    dens, _ = routines.willard_density(watPos, box, smoothlen, grid=grid, want_normals=False)
    verts, faces, areas = routines.iso_surface(dens, grid, level)
    norms = routines.willard_density(watPos, box, smoothlen, points=verts)[1] if verts.shape[0] else torch.zeros_like(verts)
    values = torch.full((verts.shape[0],), float(level), dtype=torch.float64, device=verts.device)
    out = (verts, faces, norms, values, areas)
    return out if isinstance(watPos, torch.Tensor) else tuple(t.cpu().numpy() for t in out)


def instantaneousInterface(watPos, thisbox, grid=None, spacing=2.0, level=0.016, smoothlen=2.4, outward=True):
    """Willard-Chandler instantaneous interface of one frame as the point set InterfaceWater consumes
    (fortran/waterlib.f90:1414-1469): density field on `grid` (default: the whole box at about `spacing` Angstrom,
    nodes at cell centres) -> vertices of the iso-surface dens == level (the marching-cubes vertex rule of
    surface_library.py:202) -> density gradient at the vertices (WillardDensityPoints, fortran/waterlib.f90:1351-1398).
    `outward` normals point towards lower density, so that the depth (water - point) . normal is negative inside the
    liquid.  Returns (gridpos (n,3), gridnorm (n,3)): CUDA tensors for tensor input, numpy arrays otherwise."""
    box = np.asarray(thisbox.detach().cpu() if isinstance(thisbox, torch.Tensor) else thisbox, dtype=np.float64).reshape(-1)[:3]
    if grid is None:
        grid = []
        for d in range(3):
            n = max(int(np.ceil(box[d] / float(spacing))), 1)
            grid.append((np.arange(n) + 0.5) * (box[d] / n))
    dens, _ = routines.willard_density(watPos, box, smoothlen, grid=grid, want_normals=False)
    pts = routines.iso_points(dens, grid, level)
    if pts.shape[0]:
        _, norms = routines.willard_density(watPos, box, smoothlen, points=pts)
        if outward:
            norms = -norms
    else:
        norms = torch.zeros_like(pts)
    if isinstance(watPos, torch.Tensor):
        return pts, norms
    return pts.cpu().numpy(), norms.cpu().numpy()


def depthBinnedQ(watPos, thisbox, gridpos=None, gridnorm=None, binWidth=1.0, depthRange=(-30.0, 10.0), lowCut=0.0, highCut=10.0,
                 cutoff=0.0, **interface_kw):
    """Config-4 composition for one frame: tetrahedral q of every water (getOrderParamq), its signed depth below
    the instantaneous interface (InterfaceWater: (water - nearest surface point) . normal, negative inside the
    liquid for outward normals), and the profile of q against depth.  Without gridpos / gridnorm the interface is
    the frame's own Willard-Chandler surface (instantaneousInterface(**interface_kw)).
    Returns dict(depth_edges, count, q_mean, q_var, q (n,), depth (n,), numwater, n_surface)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    pos = torch.as_tensor(np.ascontiguousarray(np.asarray(watPos, dtype=np.float64))).to(dev) if not isinstance(watPos, torch.Tensor) else watPos
    r = engine.q3b_frames(pos, thisbox, None, do_q=True, do_3body=False, lowq=lowCut, highq=highCut, want=("q",))
    q = r["q"][0]
    if gridpos is None:
        gridpos, gridnorm = instantaneousInterface(pos, thisbox, **interface_kw)
    iw = routines.interface_water(pos, gridpos, gridnorm, cutoff, thisbox, want_surfclose=False)
    depth = iw["allwatdists"]
    nb = int(np.ceil((depthRange[1] - depthRange[0]) / binWidth))
    count, s1, s2 = routines.profile_bins(q, depth, depthRange[0], binWidth, nb)
    count_h, s1_h, s2_h = count.cpu().numpy(), s1.cpu().numpy(), s2.cpu().numpy()
    with np.errstate(all="ignore"):
        mean = s1_h / count_h
        var = np.maximum(s2_h / count_h - mean * mean, 0.0)
    return {"depth_edges": depthRange[0] + binWidth * np.arange(nb + 1), "count": count_h, "q_mean": mean, "q_var": var,
            "q": q, "depth": depth, "numwater": int(iw["numwater"].item()), "n_surface": int(gridpos.shape[0])}


def densityVoxel(heavyPos, watPos, thisbox):
    """Plain voxel density of the waters on the reference's 10^3 grid around heavyPos (surface_library.py:213-241)."""
    heavyPos = np.asarray(heavyPos, dtype=np.float64)
    nBins = 11
    spans = []
    for d in range(3):
        span = np.linspace(0.8 * np.min(heavyPos[:, d]), 1.2 * np.max(heavyPos[:, d]), nBins).reshape(1, nBins)
        width = span[:, 1] - span[:, 0]
        spans.append(span[:, :-1] + width)
    from . import waterlib as wl
    return wl.densityfield(watPos, spans[0], spans[1], spans[2], thisbox)
