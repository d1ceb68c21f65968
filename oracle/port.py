"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle (oracle/wol_oracle.c).

Function names and argument order follow the reference's per-frame API
(structureLibs/water_properties.py:210-391) so the parity tests read like calls into the reference.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this module; nothing under ``waterorderlib_b200/`` does.

Pinned by: tests/test_oracle_vs_reference.py (live reference, build container only) and
tests/test_oracle_golden.py (committed fixtures generated from the live reference by
tests/golden/make_golden.py).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int32)
_lp = ctypes.POINTER(ctypes.c_int64)


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libwol_oracle.so")
        if not os.path.exists(path):
            from . import build_oracle
            build_oracle.build(verbose=False)
        _LIB = ctypes.CDLL(path)
    return _LIB


def _c(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


def _pos(a):
    a = _c(a)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("positions must be (n,3)")
    return a


def _box(b):
    b = _c(b).reshape(-1)
    if b.size != 3:
        raise ValueError("box must hold 3 values")
    return b


def _ptr(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def histogram(x, nbins=500, lo=0.0, hi=180.0):
    """np.histogram(x, bins=nbins, range=[lo, hi]) counts, restated (numpy uniform-bin rule)."""
    x = _c(x).reshape(-1)
    h = np.zeros(nbins, dtype=np.int64)
    _lib().wol_oracle_histogram(_ptr(x, _dp), ctypes.c_int64(x.size), ctypes.c_double(lo), ctypes.c_double(hi),
                                ctypes.c_int(nbins), _ptr(h, _lp))
    return h


def three_body(subPos, Pos, BoxDims, lowCut=0.0, highCut=3.413, nBins=500, binRange=(0.0, 180.0),
               materialize=True):
    """getCosAngs + the integer/sum parts of tetrahedralMetrics in one pass.
    Returns dict(angVals, numAngs (int32 neighbour counts), hist (int64), tet (count, sum cos, sum cos^2),
    n_angles)."""
    sub, pos, box = _pos(subPos), _pos(Pos), _box(BoxDims)
    m, n = sub.shape[0], pos.shape[0]
    ncount = np.zeros(m, dtype=np.int32)
    hist = np.zeros(nBins, dtype=np.int64)
    tet = np.zeros(3, dtype=np.float64)
    n_ang = ctypes.c_int64(0)
    lib = _lib()

    def run(buf, cap):
        hist[:] = 0
        tet[:] = 0
        rc = lib.wol_oracle_three_body(
            _ptr(sub, _dp), m, _ptr(pos, _dp), n, _ptr(box, _dp), ctypes.c_double(lowCut),
            ctypes.c_double(highCut), _ptr(ncount, _ip), _ptr(buf, _dp), ctypes.c_int64(cap),
            ctypes.byref(n_ang), _ptr(hist, _lp), nBins, ctypes.c_double(binRange[0]),
            ctypes.c_double(binRange[1]), _ptr(tet, _dp))
        if rc != 0:
            raise RuntimeError("oracle three_body failed rc=%d" % rc)

    ang = None
    if materialize:
        cap = max(16 * m, 1024)
        ang = np.zeros(cap, dtype=np.float64)
        run(ang, cap)
        if n_ang.value > cap:
            cap = n_ang.value
            ang = np.zeros(cap, dtype=np.float64)
            run(ang, cap)
        ang = ang[: n_ang.value].copy()
    else:
        run(None, 0)
    return {"angVals": ang, "numAngs": ncount, "hist": hist, "tet": tet, "n_angles": int(n_ang.value)}


def getCosAngs(subPos, Pos, BoxDims, lowCut=0.0, highCut=3.413):
    """structureLibs/water_properties.py:210-250 -> (angVals f64, numAngs f64 neighbour counts)."""
    r = three_body(subPos, Pos, BoxDims, lowCut, highCut)
    return r["angVals"], r["numAngs"].astype(np.float64)


def order_param_q(subPos, Pos, BoxDims, lowCut=0.0, highCut=10.0):
    """getOrderParamq plus the selection it implies -> (q f64 (m,), nn4 int32 (m,4) -1 padded, ncount int32)."""
    sub, pos, box = _pos(subPos), _pos(Pos), _box(BoxDims)
    m, n = sub.shape[0], pos.shape[0]
    q = np.zeros(m, dtype=np.float64)
    nn4 = np.zeros((m, 4), dtype=np.int32)
    ncount = np.zeros(m, dtype=np.int32)
    rc = _lib().wol_oracle_order_param_q(_ptr(sub, _dp), m, _ptr(pos, _dp), n, _ptr(box, _dp),
                                         ctypes.c_double(lowCut), ctypes.c_double(highCut), _ptr(q, _dp),
                                         _ptr(nn4, _ip), _ptr(ncount, _ip))
    if rc != 0:
        raise RuntimeError("oracle order_param_q failed rc=%d" % rc)
    return q, nn4, ncount


def getOrderParamq(subPos, Pos, BoxDims, lowCut=0.0, highCut=10.0):
    """structureLibs/water_properties.py:344-391."""
    return order_param_q(subPos, Pos, BoxDims, lowCut, highCut)[0]


def tetrahedralMetrics(angVals, nBins=500, binRange=(0.0, 180.0)):
    """structureLibs/water_properties.py:314-342 with the histogram restated."""
    angVals = _c(angVals).reshape(-1)
    angDist = histogram(angVals, nBins, binRange[0], binRange[1])
    bins = np.linspace(binRange[0], binRange[1], nBins + 1)
    angTet = angVals[(angVals >= 100.0) & (angVals <= 120.0)]
    fracTet = float(len(angTet)) / float(len(angVals))
    c = np.cos(angTet * np.pi / 180.0)
    avgCos, varCos = np.mean(c), np.var(c)
    dens = angDist / float(np.sum(angDist))
    dens = dens[dens != 0]
    entropy = -np.sum(dens * np.log(dens))
    return angDist, bins, fracTet, avgCos, varCos, entropy


def neighbor_matrix(subPos, Pos, BoxDims, lowCut, highCut):
    """nearNeighbors / allNearNeighbors (fortran/waterlib.f90:710-743, :830-862) -> (m,n) int32."""
    sub, pos, box = _pos(subPos), _pos(Pos), _box(BoxDims)
    mat = np.zeros((sub.shape[0], pos.shape[0]), dtype=np.int32)
    _lib().wol_oracle_neighbor_matrix(_ptr(sub, _dp), sub.shape[0], _ptr(pos, _dp), pos.shape[0], _ptr(box, _dp),
                                      ctypes.c_double(lowCut), ctypes.c_double(highCut), _ptr(mat, _ip))
    return mat


def shell_mask(solPos, watPos, BoxDims, cutoff=4.0, lowCut=0.0):
    """structureLibs/orderParam_lib.py:495-498 -> int32 mask over waters."""
    sol, wat, box = _pos(solPos), _pos(watPos), _box(BoxDims)
    mask = np.zeros(wat.shape[0], dtype=np.int32)
    _lib().wol_oracle_shell_mask(_ptr(sol, _dp), sol.shape[0], _ptr(wat, _dp), wat.shape[0], _ptr(box, _dp),
                                 ctypes.c_double(lowCut), ctypes.c_double(cutoff), _ptr(mask, _ip))
    return mask


def hbonds(accPos, donPos, donHPos, BoxDims, distCut=3.5, angCut=150.0, dense=False):
    """generalHbonds (fortran/waterlib.f90:1156-1210) -> (acc_count, don_count[, matrix])."""
    acc, don, donh, box = _pos(accPos), _pos(donPos), _pos(donHPos), _box(BoxDims)
    if don.shape[0] != donh.shape[0]:
        raise ValueError("donor heavy atoms and hydrogens differ in number")
    na, nd = acc.shape[0], don.shape[0]
    ac = np.zeros(na, dtype=np.int32)
    dc = np.zeros(nd, dtype=np.int32)
    mat = np.zeros((na, nd), dtype=np.int32) if dense else None
    _lib().wol_oracle_hbonds(_ptr(acc, _dp), na, _ptr(don, _dp), _ptr(donh, _dp), nd, _ptr(box, _dp),
                             ctypes.c_double(distCut), ctypes.c_double(angCut), _ptr(ac, _ip), _ptr(dc, _ip),
                             _ptr(mat, _ip))
    return (ac, dc, mat) if dense else (ac, dc)


def angles_from_cos(c):
    """CosAngle3's clamp/acos/mod/degrees tail (fortran/waterlib.f90:698-702) for an array of cosines."""
    c = _c(c).reshape(-1)
    out = np.zeros_like(c)
    _lib().wol_oracle_angles_from_cos(_ptr(c, _dp), ctypes.c_int64(c.size), _ptr(out, _dp))
    return out


def willard_density_points(pos, denspts, BoxL, smoothlen):
    """WillardDensityPoints (fortran/waterlib.f90:1351-1398) -> (densvals (npts,), densnorms (npts,3))."""
    pos, pts, box = _pos(pos), _pos(denspts), _box(BoxL)
    dens = np.zeros(pts.shape[0], dtype=np.float64)
    norms = np.zeros((pts.shape[0], 3), dtype=np.float64)
    _lib().wol_oracle_willard_points(_ptr(pos, _dp), pos.shape[0], _ptr(pts, _dp), ctypes.c_int64(pts.shape[0]),
                                     _ptr(box, _dp), ctypes.c_double(smoothlen), _ptr(dens, _dp), _ptr(norms, _dp))
    return dens, norms


def willard_density_field(pos, gridx, gridy, gridz, BoxL, smoothlen):
    """WillardDensityField (fortran/waterlib.f90:1286-1341) -> (densvals (nx,ny,nz), densnorms (nx,ny,nz,3))."""
    gx, gy, gz = (np.asarray(g, dtype=np.float64) for g in (gridx, gridy, gridz))
    X, Y, Z = np.meshgrid(gx, gy, gz, indexing="ij")
    pts = np.stack([X.ravel(), Y.ravel(), Z.ravel()], axis=1)
    dens, norms = willard_density_points(pos, pts, BoxL, smoothlen)
    return dens.reshape(gx.size, gy.size, gz.size), norms.reshape(gx.size, gy.size, gz.size, 3)


def interface_water(pos, gridpos, gridnorm, cutoff, BoxL):
    """InterfaceWater (fortran/waterlib.f90:1414-1469) with 0-based indices, -1 = none within r^2 < 1000.
    -> (watclose int32 (n,), surfclose int32 (ng,), numwater, allwatdists f64 (n,))."""
    pos, gp, gn, box = _pos(pos), _pos(gridpos), _pos(gridnorm), _box(BoxL)
    n, ng = pos.shape[0], gp.shape[0]
    watclose = np.zeros(n, dtype=np.int32)
    surfclose = np.zeros(ng, dtype=np.int32)
    numwater = ctypes.c_int32(0)
    dists = np.zeros(n, dtype=np.float64)
    _lib().wol_oracle_interface_water(_ptr(pos, _dp), n, _ptr(gp, _dp), _ptr(gn, _dp), ng, ctypes.c_double(cutoff),
                                      _ptr(box, _dp), _ptr(watclose, _ip), _ptr(surfclose, _ip), ctypes.byref(numwater),
                                      _ptr(dists, _dp))
    return watclose, surfclose, int(numwater.value), dists


def histrr3b(Pos, BoxL, distWidth, dNum, angWidth, aNum):
    """histrr3b (fortran/waterlib.f90:1550-1593) -> int64 counts (dNum, dNum, aNum)."""
    pos, box = _pos(Pos), _box(BoxL)
    hist = np.zeros((dNum, dNum, aNum), dtype=np.int64)
    _lib().wol_oracle_histrr3b(_ptr(pos, _dp), pos.shape[0], _ptr(box, _dp), ctypes.c_double(distWidth), int(dNum),
                               ctypes.c_double(angWidth), int(aNum), _ptr(hist, _lp))
    return hist


def getLSI(subPos, Pos, BoxDims, lowCut=0.0, highCut=3.7):
    """structureLibs/water_properties.py:252-311 -> (lsiVals of the centres that have one, numLSI f64 (m,))."""
    sub, pos, box = _pos(subPos), _pos(Pos), _box(BoxDims)
    m = sub.shape[0]
    lsi = np.zeros(m, dtype=np.float64)
    num = np.zeros(m, dtype=np.int32)
    has = np.zeros(m, dtype=np.int32)
    _lib().wol_oracle_lsi(_ptr(sub, _dp), m, _ptr(pos, _dp), pos.shape[0], _ptr(box, _dp), ctypes.c_double(lowCut),
                          ctypes.c_double(highCut), _ptr(lsi, _dp), _ptr(num, _ip), _ptr(has, _ip))
    return lsi[has.astype(bool)], num.astype(np.float64)


def _pair_hist(mode, pos1, pos2, box, binwidth, totbins):
    p1, b = _pos(pos1), _box(box)
    p2 = _pos(pos2) if pos2 is not None else p1
    counts = np.zeros(totbins, dtype=np.int64)
    _lib().wol_oracle_pair_hist(mode, _ptr(p1, _dp), p1.shape[0], _ptr(p2, _dp), p2.shape[0], _ptr(b, _dp),
                                ctypes.c_double(binwidth), int(totbins), _ptr(counts, _lp))
    return counts


_FOUR_THIRDS = float(np.float32(4.0) / np.float32(3.0))  # the Fortran literal (4./3.) is single precision
_PI_RDF = 3.141592653589                                 # and its pi is truncated (waterlib.f90:204,327)


def rdf_normalise(counts, n_norm, binwidth, bulkdens):
    """counts(k) / (N * BulkDens * (4./3.) * pi * binwidth**3 * (k**3 - (k-1)**3)), in the Fortran's order."""
    k = np.arange(1, len(counts) + 1, dtype=np.int64)
    shell = (k ** 3 - (k - 1) ** 3).astype(np.float64)
    denom = ((((float(n_norm) * bulkdens) * _FOUR_THIRDS) * _PI_RDF) * ((binwidth * binwidth) * binwidth)) * shell
    return counts.astype(np.float64) / denom


def radialdist(pos1, pos2, binwidth, totbins, bulkdens, boxl):
    """RadialDist (fortran/waterlib.f90:193-231)."""
    return rdf_normalise(_pair_hist(0, pos1, pos2, boxl, binwidth, totbins), _pos(pos1).shape[0], binwidth, bulkdens)


def radialdistsame(pos, binwidth, totbins, bulkdens, boxl):
    """RadialDistSame (fortran/waterlib.f90:316-353)."""
    return rdf_normalise(_pair_hist(1, pos, None, boxl, binwidth, totbins), _pos(pos).shape[0], binwidth, bulkdens)


def pairdistancehistogram(pos1, pos2, binwidth, totbins, boxl):
    """PairDistanceHistogram (fortran/waterlib.f90:358-389), 3-D."""
    return _pair_hist(2, pos1, pos2, boxl, binwidth, totbins).astype(np.float64)


def getOrderParamPsi(subPos, Pos, BoxDims, lowCut=0.0, highCut=10.0):
    """structureLibs/water_properties.py:393-433."""
    sub, pos, box = _pos(subPos), _pos(Pos), _box(BoxDims)
    psi = np.zeros(sub.shape[0], dtype=np.float64)
    _lib().wol_oracle_psi(_ptr(sub, _dp), sub.shape[0], _ptr(pos, _dp), pos.shape[0], _ptr(box, _dp), ctypes.c_double(lowCut),
                          ctypes.c_double(highCut), _ptr(psi, _dp))
    return psi


def density_field(pos, gridx, gridy, gridz, BoxL):
    """DensityField (fortran/waterlib.f90:1219-1268) -> (nx, ny, nz) float64."""
    p, box = _pos(pos), _box(BoxL)
    gx, gy, gz = (_c(g).reshape(-1) for g in (gridx, gridy, gridz))
    out = np.zeros((gx.size, gy.size, gz.size), dtype=np.float64)
    _lib().wol_oracle_density_field(_ptr(p, _dp), p.shape[0], _ptr(gx, _dp), gx.size, _ptr(gy, _dp), gy.size, _ptr(gz, _dp), gz.size,
                                    _ptr(box, _dp), _ptr(out, _dp))
    return out


def iso_points(dens, gridx, gridy, gridz, level):
    """Iso-surface vertices: one per grid edge whose end values straddle the level ((va > level) != (vb > level)),
    at a + t (b - a), t = (level - va) / (vb - va); ordered by node index ((i * ny + j) * nz + k), then axis.  This is
    the rule marching cubes places its vertices with (skimage.measure.marching_cubes at
    structureLibs/surface_library.py:202 is un-vendored and not installed: PARITY UNPINNED for this step -- the
    restatement is pinned by its own properties in tests/test_oracle_golden.py)."""
    dens = np.asarray(dens, dtype=np.float64)
    gx, gy, gz = (np.asarray(g, dtype=np.float64).reshape(-1) for g in (gridx, gridy, gridz))
    nx, ny, nz = dens.shape
    above = dens > level
    node = np.arange(nx * ny * nz).reshape(nx, ny, nz)
    keys, pts = [], []
    for ax, g in enumerate((gx, gy, gz)):
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[ax], hi[ax] = slice(0, -1), slice(1, None)
        lo, hi = tuple(lo), tuple(hi)
        cross = above[lo] != above[hi]
        ijk = np.nonzero(cross)
        va, vb = dens[lo][cross], dens[hi][cross]
        t = (level - va) / (vb - va)
        p = np.stack([gx[ijk[0]], gy[ijk[1]], gz[ijk[2]]], axis=1)
        a = g[ijk[ax]]
        p[:, ax] = a + t * (g[ijk[ax] + 1] - a)
        keys.append(node[lo][cross] * 3 + ax)
        pts.append(p)
    keys = np.concatenate(keys)
    pts = np.concatenate(pts)
    order = np.argsort(keys, kind="stable")
    return pts[order]


def watorient(opos, hpos, refvec, boxl):
    """watOrient (fortran/waterlib.f90:973-1011) -> (angDip, angPlane) in degrees, one per water."""
    o, h, box = _pos(opos), _pos(hpos), _box(boxl)
    if h.shape[0] != 2 * o.shape[0]:
        raise ValueError("Number of hydrogens must be two times number of oxygens.")
    ref = _c(np.asarray(refvec, dtype=np.float64).reshape(-1))
    a, b = np.zeros(o.shape[0]), np.zeros(o.shape[0])
    _lib().wol_oracle_watorient(_ptr(o, _dp), o.shape[0], _ptr(h, _dp), _ptr(ref, _dp), _ptr(box, _dp), _ptr(a, _dp), _ptr(b, _dp))
    return a, b


def binongrid(opos, xbins, ybins, zbins):
    """binOnGrid (fortran/waterlib.f90:1047-1099) -> int32 (nx-1, ny-1, nz-1)."""
    o = _pos(opos)
    xb, yb, zb = (_c(np.asarray(g, dtype=np.float64).reshape(-1)) for g in (xbins, ybins, zbins))
    out = np.zeros((xb.size - 1, yb.size - 1, zb.size - 1), dtype=np.int32)
    rc = _lib().wol_oracle_binongrid(_ptr(o, _dp), o.shape[0], _ptr(xb, _dp), xb.size, _ptr(yb, _dp), yb.size, _ptr(zb, _dp), zb.size,
                                     _ptr(out, _ip))
    if rc != 0:
        raise ValueError("Must break volume into CUBES. Currently, bin-widths do not match.")
    return out


def _anint(x):
    """Fortran anint: round half away from zero."""
    t = np.trunc(x)
    f = x - t
    return t + (f >= 0.5) - (f <= -0.5)


def radialdistplane(Pos1, Pos2, binwidth, totbins, BulkDens, BoxL):
    """RadialDistPlane (fortran/waterlib.f90:237-314) operation by operation -> (rdf (totbins, totbins) f64, number of
    slab atoms whose bin index is <= 0: the Fortran writes out of bounds for those, they are skipped here)."""
    p1 = np.asarray(Pos1, dtype=np.float64).reshape(3, 3)
    p2 = np.asarray(Pos2, dtype=np.float64).reshape(-1, 3)
    L = np.asarray(BoxL, dtype=np.float64).reshape(3)
    with np.errstate(divide="ignore"):
        iL = np.where(L >= 0.0, 1.0 / L, 0.0)                      # :261
    v1 = p1[2] - p1[0]                                                # :264-266
    v2 = p1[1] - p1[0]
    v3 = np.array([v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]])
    v1 = v1 - L * _anint(v1 * iL)                                     # :268-270
    v2 = v2 - L * _anint(v2 * iL)
    v3 = v3 - L * _anint(v3 * iL)

    def dot(a, b):
        return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]
    v2 = v2 - (dot(v1, v2) / dot(v1, v1)) * v1                        # :272
    v1 = v1 / np.sqrt(dot(v1, v1))                                    # :274-276
    v2 = v2 / np.sqrt(dot(v2, v2))
    v3 = v3 / np.sqrt(dot(v3, v3))
    Q = np.stack([v1, v2, v3], axis=1)                                # Q(:, k) = v_k
    rdf = np.zeros((totbins, totbins))
    bad = 0
    for p in p2:
        q = p * BulkDens / BulkDens                                   # :291
        q = q - L * _anint(q * iL)
        r = np.zeros(3)
        for n in range(3):                                            # libgfortran's matmul: dest(x) += Q(x, n) * q(n)
            r = r + Q[:, n] * q[n]
        if -5.0 <= r[2] <= 5.0:                                       # newPos1(1, :) = matmul(Q, 0) = 0
            bx, by = np.ceil(r[0] / binwidth), np.ceil(r[1] / binwidth)
            if bx <= totbins and by <= totbins:
                if bx >= 1 and by >= 1:
                    rdf[int(bx) - 1, int(by) - 1] += 1.0
                else:
                    bad += 1
    return rdf, bad
