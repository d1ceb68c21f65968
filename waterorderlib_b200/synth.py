"""Seeded synthetic water boxes: jittered diamond-cubic ("ice Ic") oxygen
lattices at liquid density, optional rigid hydrogens, a lattice solute and a vacuum slab.

These stand in for the ``frame.xyz`` / ``frame.box.values[:3]`` arrays the reference reads from
pytraj frames (structureLibs/orderParam_lib.py:1314-1316).  Coordinates are *float32-representable*
doubles so the fp64 and fp32 device paths, the CPU oracle and the reference all see the same inputs.
"""
import numpy as np

# liquid-water number density used by the reference (structureLibs/water_properties.py:55)
RHO_WATER = 0.033456
A_CELL = (8.0 / RHO_WATER) ** (1.0 / 3.0)  # diamond-cubic cell edge holding 8 waters: 6.2069 A
R_OH = 0.9572
ANG_HOH = 104.52

_BASIS = np.array(
    [[0.0, 0.0, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0],
     [0.25, 0.25, 0.25], [0.25, 0.75, 0.75], [0.75, 0.25, 0.75], [0.75, 0.75, 0.25]])


def diamond_lattice(mx, my=None, mz=None):
    """Ideal O positions for an mx*my*mz block of diamond-cubic cells -> ((8*mx*my*mz, 3) f64, box (3,))."""
    my = mx if my is None else my
    mz = mx if mz is None else mz
    ix, iy, iz = np.meshgrid(np.arange(mx), np.arange(my), np.arange(mz), indexing="ij")
    cells = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1).astype(np.float64)
    pos = (cells[:, None, :] + _BASIS[None, :, :]).reshape(-1, 3) * A_CELL
    box = np.array([mx, my, mz], dtype=np.float64) * A_CELL
    return pos, box


def water_box(m, sigma=0.25, seed=1234, dims=None):
    """Jittered-ice O positions: m^3 cells (or dims=(mx,my,mz)), Gaussian jitter `sigma` (A), wrapped
    into [0, L), rounded to float32 and returned as float64.  Returns (pos (N,3), box (3,))."""
    mx, my, mz = dims if dims is not None else (m, m, m)
    pos, box = diamond_lattice(mx, my, mz)
    rng = np.random.Generator(np.random.PCG64(seed))
    if sigma > 0.0:
        pos = pos + rng.normal(0.0, sigma, size=pos.shape)
    box = box.astype(np.float32).astype(np.float64)
    pos = pos - box * np.floor(pos / box)
    pos = pos.astype(np.float32)
    # float32 rounding can land exactly on L; fold it back so every coordinate is in [0, L)
    boxf = box.astype(np.float32)
    pos = np.where(pos >= boxf, pos - boxf, pos)
    return pos.astype(np.float64), box


def add_hydrogens(opos, seed=1234):
    """Rigid H pair per O (0.9572 A, 104.52 deg) with a random orientation per molecule.
    Returns hpos (2N,3) ordered H1,H2 per water, float32-representable."""
    n = opos.shape[0]
    rng = np.random.Generator(np.random.PCG64(seed + 7919))
    # random orthonormal frame: u uniform on the sphere, v orthogonal to it
    u = rng.normal(size=(n, 3))
    u /= np.linalg.norm(u, axis=1, keepdims=True)
    w = rng.normal(size=(n, 3))
    v = w - np.sum(w * u, axis=1, keepdims=True) * u
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    half = np.deg2rad(ANG_HOH) / 2.0
    h1 = opos + R_OH * (np.cos(half) * u + np.sin(half) * v)
    h2 = opos + R_OH * (np.cos(half) * u - np.sin(half) * v)
    hpos = np.empty((2 * n, 3))
    hpos[0::2] = h1
    hpos[1::2] = h2
    return hpos.astype(np.float32).astype(np.float64)


def solute_grid(box, n_side=4, spacing=1.5):
    """cfg3 solute: n_side^3 heavy atoms on a cubic grid centred in the box."""
    g = (np.arange(n_side) - 0.5 * (n_side - 1)) * spacing
    gx, gy, gz = np.meshgrid(g, g, g, indexing="ij")
    sol = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1) + 0.5 * np.asarray(box)
    return sol.astype(np.float32).astype(np.float64)


def slab_box(mx, my, mz, sigma=0.25, seed=1234, vacuum_factor=3.0):
    """cfg4 air-water slab: mx*my*mz cells of water centred in a box whose z edge is
    vacuum_factor * slab thickness.  Returns (pos, box, z_lo, z_hi) with z_lo/z_hi the ideal faces."""
    pos, box = water_box(0, sigma=sigma, seed=seed, dims=(mx, my, mz))
    thick = box[2]
    lz = np.float64(np.float32(vacuum_factor * thick))
    z_lo = 0.5 * (lz - thick)
    pos = pos.copy()
    pos[:, 2] = (pos[:, 2] + z_lo).astype(np.float32).astype(np.float64)
    return pos, np.array([box[0], box[1], lz]), z_lo, z_lo + thick


def trajectory(m, n_frames, sigma=0.25, seed0=1234, dims=None):
    """(F,N,3) positions and (F,3) boxes; frame f uses seed0+f so results do not depend on sharding."""
    frames, boxes = [], []
    for f in range(n_frames):
        p, b = water_box(m, sigma=sigma, seed=seed0 + f, dims=dims)
        frames.append(p)
        boxes.append(b)
    return np.stack(frames), np.stack(boxes)


def plane_interface(box, z_lo, z_hi, spacing=2.0):
    """cfg4 parity interface: the two ideal faces of a slab sampled on a square grid, with outward unit
    normals (-z at z_lo, +z at z_hi).  Stands in for the marching-cubes surface of the reference
    (structureLibs/surface_library.py:202), which needs skimage.  Returns (gridpos (G,3), gridnorm (G,3))."""
    nx, ny = max(1, int(round(box[0] / spacing))), max(1, int(round(box[1] / spacing)))
    x = (np.arange(nx) + 0.5) * (box[0] / nx)
    y = (np.arange(ny) + 0.5) * (box[1] / ny)
    X, Y = np.meshgrid(x, y, indexing="ij")
    pts, nrm = [], []
    for z, s in ((z_lo, -1.0), (z_hi, 1.0)):
        pts.append(np.stack([X.ravel(), Y.ravel(), np.full(X.size, z)], axis=1))
        nrm.append(np.tile(np.array([0.0, 0.0, s]), (X.size, 1)))
    return (np.concatenate(pts).astype(np.float32).astype(np.float64), np.concatenate(nrm))


def device_frames(m, f0, f1, sigma=0.25, device="cuda", seed_base=17):
    """Jittered-ice frames f0 .. f1-1 of an m^3-cell box generated ON THE DEVICE (torch): (f1-f0, N, 3) float64 values that
    are float32-representable, wrapped into [0, L), plus the box (3,) as numpy.  Frame f depends only on its absolute
    index (generator seeded with 1000003 f + seed_base), so a sharded trajectory does not depend on the number of ranks.
    Not the same stream of numbers as water_box (numpy PCG64): use one or the other for a given comparison."""
    import torch
    lattice, box = diamond_lattice(m)
    box = box.astype(np.float32).astype(np.float64)
    lat_d = torch.from_numpy(lattice).to(device)
    box_d = torch.from_numpy(box).to(device)
    out = torch.empty((f1 - f0, lattice.shape[0], 3), dtype=torch.float64, device=device)
    for k, f in enumerate(range(f0, f1)):
        g = torch.Generator(device=device)
        g.manual_seed(1_000_003 * f + seed_base)
        p = lat_d + sigma * torch.randn(lattice.shape, generator=g, device=device, dtype=torch.float64)
        p = p - box_d * torch.floor(p / box_d)
        p = p.to(torch.float32)
        boxf = box_d.to(torch.float32)
        p = torch.where(p >= boxf, p - boxf, p)  # float32 rounding can land exactly on L
        out[k] = p.to(torch.float64)
    return out, box
