"""Drop-in for the hot-path functions of the reference's ``structureLibs/water_properties.py``.

Same names, positional order, keyword names, defaults and return values:

    getOrderParamq(subPos, Pos, BoxDims, lowCut=0.0, highCut=10.0)                   reference :344-391
    getCosAngs(subPos, Pos, BoxDims, lowCut=0.0, highCut=3.413)                      reference :210-250
    tetrahedralMetrics(angVals, nBins=500, binRange=[0.0, 180.0])                    reference :314-342
    HBondsGeneral(accPos, donPos, donHPos, boxL, accInds, donInds, donHInds,
                  distCut=3.5, angCut=150.0)                                         reference :681-719
    getLSI(subPos, Pos, BoxDims, lowCut=0.0, highCut=3.7)                            reference :252-311
    getOrderParamPsi(subPos, Pos, BoxDims, lowCut=0.0, highCut=10.0)                 reference :393-433
    waterOrientation(Opos, Hpos, boxDim, refVec=[0.0, 0.0, 1.0])                     reference :622-638
    waterOrientationBinZ(Opos, Hpos, boxDim, refVec=[0.0, 0.0, 1.0], refBins=None,
                         angBins=None)                                               reference :578-619
    binnedVolumePofN(Opos, volBins, numBins, binMask=None)                           reference :641-676

numpy arrays in -> numpy arrays out; torch CUDA tensors in -> torch CUDA tensors out (zero copy).  The
per-water Python loops and f2py calls of the reference are replaced by one pass of the cell-list kernels in
libwol.so (include/wol_capi.h).  There is no CPU fallback.
"""
import numpy as np
import torch

from .. import engine, routines


def _is_torch(*xs):
    return any(isinstance(x, torch.Tensor) for x in xs)


def _same(subPos, Pos):
    """np.array_equal(subPos, Pos) of the reference (water_properties.py:235, :363)."""
    if subPos is Pos:
        return True
    if _is_torch(subPos, Pos):
        a, b = torch.as_tensor(subPos), torch.as_tensor(Pos)
        return a.shape == b.shape and a.device == b.device and bool(torch.equal(a, b))
    a, b = np.asarray(subPos), np.asarray(Pos)
    return a.shape == b.shape and bool(np.array_equal(a, b))


def _out(t, like_torch, dtype=None):
    if like_torch:
        return t if dtype is None else t.to(dtype)
    a = t.detach().cpu().numpy()
    return a if dtype is None else a.astype(dtype)


def _pos2(a, name):
    shape = tuple(a.shape)
    if len(shape) != 2 or shape[1] != 3:
        raise ValueError("%s must have shape (n,3), got %s" % (name, shape))
    return a


def getOrderParamq(subPos, Pos, BoxDims, lowCut=0.0, highCut=10.0):
    """Tetrahedral order parameter q (Errington & Debenedetti 2001) of every position in subPos from its 4
    nearest neighbours in Pos within (lowCut, highCut]; 1/2/3 neighbours are padded with 180-degree angles,
    0 neighbours give q = 0 (reference water_properties.py:344-391)."""
    tor = _is_torch(subPos, Pos)
    subPos = subPos if tor else np.asarray(subPos)
    Pos = Pos if tor else np.asarray(Pos)
    _pos2(subPos, "subPos"); _pos2(Pos, "Pos")
    if len(subPos) == 0:
        return torch.zeros(0, dtype=torch.float64) if tor else np.zeros(0)
    if len(Pos) == 0:
        return _out(torch.zeros(len(subPos), dtype=torch.float64), tor)
    # a loop over frames of a few thousand waters is launch-bound: replay the captured launch sequence (the result is
    # copied out of the graph's buffer either way)
    call = engine.q3b_frames_graphed if len(Pos) <= engine.GRAPH_MAX_ATOMS else engine.q3b_frames
    r = call(Pos, BoxDims, None if _same(subPos, Pos) else subPos, do_q=True, do_3body=False, lowq=lowCut, highq=highCut,
             want=("q",))
    q = r["q"][0]
    return q.clone() if tor else _out(q, tor)


def getCosAngs(subPos, Pos, BoxDims, lowCut=0.0, highCut=3.413):
    """All three-body angles (degrees, despite the name) about each position of subPos among its neighbours in
    Pos within (lowCut, highCut], in the reference's order, and the neighbour count per centre as float64
    (reference water_properties.py:210-250).  Limit: at most 64 neighbours per centre inside the cutoff (about 7.5 A in
    liquid water) -- beyond that WolError names the count; the fused histogram path has no such limit."""
    tor = _is_torch(subPos, Pos)
    subPos = subPos if tor else np.asarray(subPos)
    Pos = Pos if tor else np.asarray(Pos)
    _pos2(subPos, "subPos"); _pos2(Pos, "Pos")
    ang, n3, _ = routines.three_body_angles(None if _same(subPos, Pos) else subPos, Pos, BoxDims, lowCut, highCut)
    return _out(ang, tor), _out(n3[0], tor, torch.float64 if tor else np.float64)


def tetrahedralMetrics(angVals, nBins=500, binRange=[0.0, 180.0]):  # noqa: B006 (the reference's default)
    """(angDist, bins, fracTet, avgCos, varCos, entropy) of a set of three-body angles
    (reference water_properties.py:314-342)."""
    tor = _is_torch(angVals)
    n = int(angVals.numel()) if tor else int(np.asarray(angVals).size)
    hist, tet = routines.histogram(angVals, nBins, (binRange[0], binRange[1]))
    angDist = hist.cpu().numpy()
    cnt, s1, s2 = (float(v) for v in tet.cpu().numpy())
    bins = np.linspace(binRange[0], binRange[1], nBins + 1)
    fracTet = float(cnt) / float(n)  # ZeroDivisionError on empty input, like the reference (:333)
    if cnt > 0:
        avgCos = s1 / cnt
        varCos = max(s2 / cnt - avgCos * avgCos, 0.0)
    else:
        avgCos = varCos = float("nan")  # np.mean / np.var of an empty array
    angDens = angDist / float(np.sum(angDist))
    angDens = angDens[np.where(angDens != 0)[0]]
    entropy = -np.sum(angDens * np.log(angDens))
    return angDist, bins, fracTet, avgCos, varCos, entropy


def HBondsGeneral(accPos, donPos, donHPos, boxL, accInds, donInds, donHInds, distCut=3.5, angCut=150.0):
    """Hydrogen bonds between acceptors and donor (heavy atom, hydrogen) pairs -> (NumHB, HBlist, HBloc)
    (reference water_properties.py:681-719; criterion fortran/waterlib.f90:1184-1206)."""
    tor = _is_torch(accPos, donPos, donHPos)
    r = routines.hbond_counts(accPos, donPos, donHPos, boxL, distCut, angCut, pairs=True)
    pairs = r["pairs"]
    NumHB = int(pairs.shape[0])
    loc = routines.hbond_locations(pairs, accPos, donHPos, boxL)
    p = pairs.cpu().numpy()
    HBlist = (-1) * np.ones((NumHB, 2))
    if NumHB:
        HBlist[:, 0] = np.asarray(accInds)[p[:, 0]]
        HBlist[:, 1] = np.asarray(donInds)[p[:, 1]]
    if tor:
        return NumHB, torch.from_numpy(HBlist).to(loc.device), loc
    return NumHB, HBlist, loc.cpu().numpy()



def getLSI(subPos, Pos, BoxDims, lowCut=0.0, highCut=3.7):
    """Local structure index (Shiratani & Sasai 1996) -> (lsiVals, numLSI) (reference water_properties.py:252-311):
    lsiVals holds one value per centre that has more than one neighbour inside highCut and a next-shell neighbour,
    in centre order; numLSI[i] is the number of distance gaps of centre i (0 where it has no value).  Includes the
    reference's choice of the next neighbour by NON-periodic distance (:289)."""
    tor = _is_torch(subPos, Pos)
    subPos = subPos if tor else np.asarray(subPos)
    Pos = Pos if tor else np.asarray(Pos)
    _pos2(subPos, "subPos"); _pos2(Pos, "Pos")
    vals, num = routines.lsi(None if _same(subPos, Pos) else subPos, Pos, BoxDims, lowCut, highCut)
    vals, num = vals[0], num[0]
    lsiVals = vals[num > 0]
    return _out(lsiVals, tor), _out(num, tor, torch.float64 if tor else np.float64)


def getOrderParamPsi(subPos, Pos, BoxDims, lowCut=0.0, highCut=10.0):
    """Hexagonal order parameter of Dallin & van Lehn (2019) as the reference computes it (water_properties.py:393-433).
    NOTE: the reference assigns the complex mean of exp(6 i theta) into a float array (:428), which drops the imaginary
    part, so its result -- reproduced here for drop-in parity -- is |mean cos(6 theta)| over the pairs of neighbours inside
    (lowCut, highCut], not |mean exp(6 i theta)|.  0 for centres with fewer than two neighbours."""
    tor = _is_torch(subPos, Pos)
    subPos = subPos if tor else np.asarray(subPos)
    Pos = Pos if tor else np.asarray(Pos)
    _pos2(subPos, "subPos"); _pos2(Pos, "Pos")
    r = routines.psi(None if _same(subPos, Pos) else subPos, Pos, BoxDims, lowCut, highCut)
    return _out(r[0], tor)


def waterOrientation(Opos, Hpos, boxDim, refVec=[0.0, 0.0, 1.0]):
    """(dipAngs, planeAngs): angles in degrees between refVec and each water's dipole / molecular-plane normal
    (reference water_properties.py:622-638 over wl.watorient)."""
    like_torch = _is_torch(Opos, Hpos)
    dip, plane = routines.water_orient(Opos, Hpos, boxDim, refVec)
    return _out(dip[0], like_torch), _out(plane[0], like_torch)


def waterOrientationBinZ(Opos, Hpos, boxDim, refVec=[0.0, 0.0, 1.0], refBins=None, angBins=None):
    """(plane2Dhist, dip2Dhist): 2-D histograms of the plane-normal and dipole angles against the oxygen coordinate along
    refVec (reference water_properties.py:578-619).  One deliberate difference: the reference bins the N plane angles
    against a 2N-long coordinate array (``zOposforH``, :600,:617), which ``np.histogram2d`` rejects (unequal lengths), so
    that line cannot run as written; here the plane angles are binned against the per-water coordinate like the dipole
    angles.  ``normed=False`` (removed from numpy) is dropped."""
    refVec = np.asarray(refVec, dtype=np.float64)
    refVec = refVec / np.linalg.norm(refVec)
    Opos_h = Opos.detach().cpu().numpy() if isinstance(Opos, torch.Tensor) else np.asarray(Opos, dtype=np.float64)
    zOpos = np.dot(Opos_h, refVec)
    angDip, angPlane = routines.water_orient(Opos, Hpos, boxDim, refVec)
    angDip, angPlane = angDip[0].cpu().numpy(), angPlane[0].cpu().numpy()
    if refBins is None:
        refBins = np.arange(np.min(zOpos), np.max(zOpos), 0.2)
    if angBins is None:
        angBins = np.arange(0.0, 180.001, 180.0 / 500.0)
    plane2Dhist, _a, _r = np.histogram2d(angPlane, zOpos, bins=[angBins, refBins])
    dip2Dhist, _a, _r = np.histogram2d(angDip, zOpos, bins=[angBins, refBins])
    return plane2Dhist, dip2Dhist


def binnedVolumePofN(Opos, volBins, numBins, binMask=None):
    """P(N) counts: histogram over the spatial bins of the number of oxygens inside each bin's inscribed sphere
    (reference water_properties.py:641-676 over wl.binongrid)."""
    shape = (len(volBins[0]) - 1, len(volBins[1]) - 1, len(volBins[2]) - 1)
    if binMask is None:
        binMask = np.ones(shape, dtype=bool)
    elif np.shape(binMask) != shape:
        raise ValueError("Dimensions of mask for spatial bins does not match dimensions of spatial bins.")  # reference: sys.exit(2)
    hist = routines.bin_on_grid(Opos, volBins[0], volBins[1], volBins[2]).cpu().numpy()
    numWatHist, _edges = np.histogram(hist[np.asarray(binMask, dtype=bool)].flatten(), bins=numBins)
    return numWatHist
