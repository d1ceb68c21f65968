"""Generates the committed golden fixtures from the LIVE reference (run in the build container only).

    python tests/golden/make_golden.py

Every value stored here comes out of the reference's own compiled Fortran
(/root/reference/fortran/waterlib.cpython-37m-x86_64-linux-gnu.so, called through ctypes) and its own
unmodified Python function bodies (AST-extracted from structureLibs/water_properties.py), see
oracle/ref_fortran.py.  Inputs are regenerated from seeds by waterorderlib_b200.synth, and are also
stored so a drift in the generator is caught.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import build_oracle, ref_fortran  # noqa: E402
from waterorderlib_b200 import synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def nn4_from_reference(wl, sub, pos, box, low, high):
    """The selection getOrderParamq performs (water_properties.py:372-374), with the build's tie rule
    (stable argsort == smaller atom index first).  Returns int32 (m,4) -1 padded and the neighbour count."""
    mat = wl.nearneighbors(sub, pos, box, low, high).astype(bool)
    m = sub.shape[0]
    nn4 = -np.ones((m, 4), dtype=np.int32)
    cnt = mat.sum(axis=1).astype(np.int32)
    for i in range(m):
        idx = np.nonzero(mat[i])[0]
        if idx.size == 0:
            continue
        this = wl.reimage(pos[idx], sub[i], box)
        d = np.linalg.norm(this - sub[i], axis=1)
        order = np.argsort(d, kind="stable")[:4]
        nn4[i, : order.size] = idx[order]
    return nn4, cnt


def case_q3b(wl, fn, name, sub, pos, box, low3=0.0, high3=3.413, lowq=0.0, highq=10.0, keep_angles=True):
    q = fn["getOrderParamq"](sub, pos, box, lowq, highq)
    nn4, nq = nn4_from_reference(wl, sub, pos, box, lowq, highq)
    ang, num = fn["getCosAngs"](sub, pos, box, low3, high3)
    out = dict(sub=sub, pos=pos, box=box, low3=low3, high3=high3, lowq=lowq, highq=highq, q=q, nn4=nn4, nq=nq,
               n3=num.astype(np.int32), n_angles=np.int64(ang.size))
    if ang.size:
        hist, _edges, frac, avg, var, ent = fn["tetrahedralMetrics"](ang)
        out.update(hist=hist.astype(np.int64), fracTet=frac, avgCos=avg, varCos=var, entropy=ent)
    if keep_angles:
        out["angVals"] = ang
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "m=%d n=%d angles=%d <q>=%.6f" % (sub.shape[0], pos.shape[0], ang.size, q.mean()))


def main():
    build_oracle.build(verbose=False)
    wl = ref_fortran.RefWaterlib()
    fn = ref_fortran.load_reference_functions(wl=wl)

    # cfg1: 512-water box, jittered ice and liquid-like
    for sigma, tag in ((0.25, "ice"), (0.6, "liq")):
        pos, box = synth.water_box(4, sigma=sigma, seed=1234)
        case_q3b(wl, fn, "cfg1_n512_%s" % tag, pos, pos, box)
    # perfect lattice: q == 1, all angles 109.47
    pos, box = synth.water_box(3, sigma=0.0, seed=1)
    case_q3b(wl, fn, "lattice_n216", pos, pos, box, highq=6.0)
    # non-cubic small random box (gas-like: many waters with < 4 neighbours -> padding rule), unwrapped coords
    rng = np.random.Generator(np.random.PCG64(99))
    box = np.array([14.0, 17.5, 21.25])
    pos = (rng.random((160, 3)) * box * 1.0 + box * rng.integers(-2, 3, size=(160, 3))).astype(np.float32).astype(np.float64)
    case_q3b(wl, fn, "random_n160_noncubic", pos, pos, box, highq=6.5)
    # sub-population centres that are not members of Pos, plus members
    pos, box = synth.water_box(4, sigma=0.4, seed=77)
    sub = np.concatenate([pos[::9], (rng.random((40, 3)) * box).astype(np.float32).astype(np.float64)])
    case_q3b(wl, fn, "subpop_n512_m97", sub, pos, box, highq=8.0)
    # cfg2 single frame, 4096 waters (angles not stored, histogram is)
    pos, box = synth.water_box(8, sigma=0.25, seed=1234)
    case_q3b(wl, fn, "cfg2_n4096_frame0", pos, pos, box, keep_angles=False)

    # H-bonds (hbCalc's wat-wat call, orderParam_lib.py:805-807: 3.5 A / 120 deg) and 3.0/150 (getBoundWrap)
    opos, box = synth.water_box(4, sigma=0.3, seed=4321)
    hpos = synth.add_hydrogens(opos, seed=4321)
    don = np.repeat(opos, 2, axis=0)
    for dc, ac, tag in ((3.5, 120.0, "35_120"), (3.0, 150.0, "30_150")):
        mat = wl.generalhbonds(opos, don, hpos, box, dc, ac)
        np.savez_compressed(os.path.join(OUT, "hbonds_n512_%s.npz" % tag), acc=opos, don=don, donh=hpos, box=box,
                            distcut=dc, angcut=ac, acc_count=mat.sum(axis=1).astype(np.int32),
                            don_count=mat.sum(axis=0).astype(np.int32), n_bonds=np.int64(mat.sum()))
        print("hbonds", tag, int(mat.sum()))

    # hydration shell: solute grid in a 4096-water box, cutoff 4.0 (orderParam_lib.py:421,495-498)
    pos, box = synth.water_box(8, sigma=0.25, seed=5)
    sol = synth.solute_grid(box)
    mat = wl.nearneighbors(sol, pos, box, 0.0, 4.0).astype(bool)
    shell = np.unique(np.nonzero(mat)[1]).astype(np.int32)
    np.savez_compressed(os.path.join(OUT, "shell_n4096.npz"), sol=sol, pos=pos, box=box, cutoff=4.0, shell=shell)
    print("shell", shell.size)

    # raw routine vectors: reimage / tetracosang / cosangle3 incl. the -180 quirk and the L/2 tie
    box = np.array([10.0, 12.0, 14.0])
    ref = np.array([1.0, 2.0, 3.0])
    neigh = np.array([[2.0, 2.0, 3.0], [0.0, 2.0, 3.0], [1.0, 8.0, 3.0], [6.0, 2.0, 3.0], [-4.0, 2.0, 3.0],
                      [1.5, 2.5, 9.9], [11.0, 14.0, 17.0]])
    np.savez_compressed(os.path.join(OUT, "routines.npz"), box=box, ref=ref, neigh=neigh,
                        reimaged=np.ascontiguousarray(wl.reimage(neigh, ref, box)),
                        angs=np.ascontiguousarray(wl.tetracosang(ref, neigh, box)),
                        lsid=wl.lsidists(ref, neigh, box),
                        antiparallel=wl.cosangle3([1, 0, 0], [0, 0, 0], [-1, 0, 0]),
                        parallel=wl.cosangle3([1, 0, 0], [0, 0, 0], [2, 0, 0]),
                        tetra=wl.cosangle3([1, 1, 1], [0, 0, 0], [1, -1, -1]))
    print("routines ok")

    # slab (cfg4 building blocks): Willard-Chandler density on a grid and at points, InterfaceWater against the
    # analytic faces, all from the reference's compiled Fortran (waterlib.f90:1286-1469)
    pos, box, z_lo, z_hi = synth.slab_box(4, 4, 2, sigma=0.3, seed=31)
    gx = np.linspace(0.0, box[0], 9, endpoint=False)
    gy = np.linspace(0.0, box[1], 8, endpoint=False)
    gz = np.linspace(0.0, box[2], 25, endpoint=False)
    dens, norms = wl.willarddensityfield(pos, gx, gy, gz, box, 2.4)
    pts = (rng.random((64, 3)) * box).astype(np.float32).astype(np.float64)
    pdens, pnorms = wl.willarddensitypoints(pos, pts, box, 2.4)
    gp, gn = synth.plane_interface(box, z_lo, z_hi, spacing=2.0)
    watclose, surfclose, numwater, dists = wl.interfacewater(pos, gp, gn, 3.0, box)
    voxel = wl.densityfield(pos, gx + 0.3, gy, gz, box)  # DensityField (waterlib.f90:1219-1268), cubes of edge gx[1] - gx[0]
    np.savez_compressed(os.path.join(OUT, "slab_n256.npz"), pos=pos, box=box, z_lo=z_lo, z_hi=z_hi, gx=gx, gy=gy, gz=gz, voxel=voxel,
                        smoothlen=2.4, dens=dens, norms=norms, pts=pts, pdens=pdens, pnorms=pnorms, gridpos=gp, gridnorm=gn,
                        cutoff=3.0, watclose=watclose, surfclose=surfclose, numwater=np.int64(numwater), allwatdists=dists)
    print("slab", pos.shape[0], "waters", gp.shape[0], "interface points", numwater, "within cutoff")

    # getLSI (water_properties.py:252-311), whole system and a sub-population with non-default cutoffs
    pos, box = synth.water_box(4, sigma=0.6, seed=1234)
    fn_lsi = ref_fortran.load_reference_functions(names=("getLSI",), wl=wl)["getLSI"]
    v, n = fn_lsi(pos, pos, box)
    sub = np.concatenate([pos[::9], (rng.random((30, 3)) * box).astype(np.float32).astype(np.float64)])
    v2, n2 = fn_lsi(sub, pos, box, 0.5, 3.5)
    np.savez_compressed(os.path.join(OUT, "lsi_n512.npz"), pos=pos, box=box, lsi=v, num=n, sub=sub, lsi_sub=v2, num_sub=n2,
                        low_sub=0.5, high_sub=3.5)
    print("lsi", v.size, "of", pos.shape[0], "values, mean %.5f" % v.mean())

    # pair-distance histograms (waterlib.f90:193-231, :316-353, :358-389) and psi (water_properties.py:393-433)
    pos, box = synth.water_box(4, sigma=0.5, seed=77)
    sol = (rng.random((40, 3)) * box).astype(np.float32).astype(np.float64)
    fn_psi = ref_fortran.load_reference_functions(names=("getOrderParamPsi",), wl=wl)["getOrderParamPsi"]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")  # the reference's complex -> real assignment (:428)
        psi_all = fn_psi(pos, pos, box, 0.0, 4.5)
        psi_sub = fn_psi(sol, pos, box, 1.0, 6.0)
    np.savez_compressed(os.path.join(OUT, "pairs_n512.npz"), pos=pos, sol=sol, box=box,
                        rdf_same=wl.radialdistsame(pos, 0.1, 120, 1.0, box), rdf_cross=wl.radialdist(sol, pos, 0.1, 120, 0.0334, box),
                        pdh=wl.pairdistancehistogram(sol, pos, 0.25, 40, box), psi_all=psi_all, psi_sub=psi_sub)
    print("pairs: rdf peak %.3f, <psi> %.4f" % (wl.radialdistsame(pos, 0.1, 120, 1.0, box).max(), psi_all.mean()))

    # histrr3b (waterlib.f90:1550-1593): triplet histogram, ceiling bins
    pos, box = synth.water_box(3, sigma=0.35, seed=8)
    h = wl.histrr3b(pos, box, 0.5, 8, 5.0, 36)
    np.savez_compressed(os.path.join(OUT, "histrr3b_n216.npz"), pos=pos, box=box, dwidth=0.5, dnum=8, awidth=5.0, anum=36,
                        hist=np.ascontiguousarray(h).astype(np.int64))
    print("histrr3b", int(h.sum()))

    # watOrient / binOnGrid (waterlib.f90:973-1011, :1047-1099): water orientation angles and sphere-in-cube occupancy
    o, box = synth.water_box(4, sigma=0.4, seed=3)
    h = synth.add_hydrogens(o, seed=3)
    h[::5] += box
    refvec = np.array([1.0, 2.0, -0.5])
    dip, plane = wl.watorient(o, h, refvec, box)
    edges = np.arange(0.0, 24.0, 3.0)
    occ = wl.binongrid(o, edges + 0.7, edges, edges - 1.0)
    np.savez_compressed(os.path.join(OUT, "orient_n512.npz"), opos=o, hpos=h, box=box, refvec=refvec, angdip=dip, angplane=plane,
                        xbins=edges + 0.7, ybins=edges, zbins=edges - 1.0, occupancy=np.ascontiguousarray(occ))
    print("orient: <dip> %.3f <plane> %.3f, %d atoms inside spheres" % (dip.mean(), plane.mean(), int(occ.sum())))


if __name__ == "__main__":
    main()
