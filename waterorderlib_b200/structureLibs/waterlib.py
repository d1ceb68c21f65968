"""f2py-compatible face of the hot-path routines of ``waterlib`` (built in the reference by
fortran/buildWrappers.sh:3-5 and imported as ``wl`` at structureLibs/water_properties.py:42-43).

Lower-case names and argument order as f2py generates them (signatures recovered from the reference's
prebuilt module); logical outputs are int32 arrays, array outputs are Fortran-ordered like
f2py's.  numpy in -> numpy out.  These are the small-array routines: O(m n) dense matrices by construction.
Large systems go through water_properties / orderParam_lib, which call the fused cell-list kernels.
"""
import numpy as np
import torch

from .. import routines


def _np(t, fortran=True):
    a = t.detach().cpu().numpy()
    return np.asfortranarray(a) if fortran and a.ndim > 1 else a


def _check_pos(a, name):
    a = np.asarray(a, dtype=np.float64)
    if a.ndim != 2 or a.shape[1] != 3:
        raise ValueError("%s must have shape (n,3), got %s" % (name, a.shape))  # f2py raises waterlib.error here
    return a


def allnearneighbors(pos, boxl, lowcut, highcut):
    """nneighbors = allnearneighbors(pos,boxl,lowcut,highcut)   (fortran/waterlib.f90:830-862)"""
    pos = _check_pos(pos, "pos")
    return _np(routines.neighbor_matrix(None, pos, boxl, lowcut, highcut))


def nearneighbors(subpos, pos, boxl, lowcut, highcut):
    """nneighbors = nearneighbors(subpos,pos,boxl,lowcut,highcut)   (fortran/waterlib.f90:710-743)"""
    return _np(routines.neighbor_matrix(_check_pos(subpos, "subpos"), _check_pos(pos, "pos"), boxl, lowcut, highcut))


def reimage(pos, refpos, boxl):
    """reimagedpos = reimage(pos,refpos,boxl)   (fortran/waterlib.f90:32-47)"""
    return _np(routines.reimage(_check_pos(pos, "pos"), refpos, boxl))


def tetracosang(refpos, neighpos, boxl):
    """allangs = tetracosang(refpos,neighpos,boxl)   (fortran/waterlib.f90:867-895); the diagonal, which the
    Fortran never writes, is zero."""
    return _np(routines.tetracosang(refpos, _check_pos(neighpos, "neighpos"), boxl))


def lsidists(refpos, neighpos, boxl):
    """dist1 = lsidists(refpos,neighpos,boxl)   (fortran/waterlib.f90:900-918)"""
    return _np(routines.lsidists(refpos, _check_pos(neighpos, "neighpos"), boxl))


def generalhbonds(acceptorpos, donorpos, donorhpos, boxl, distcut, angcut):
    """bondbool = generalhbonds(acceptorpos,donorpos,donorhpos,boxl,distcut,angcut)
    (fortran/waterlib.f90:1156-1210) -> (nacc, ndon) int32."""
    acc, don, donh = _check_pos(acceptorpos, "acceptorpos"), _check_pos(donorpos, "donorpos"), _check_pos(donorhpos, "donorhpos")
    if acc.shape[0] == 0 or don.shape[0] == 0:
        if don.shape[0] != donh.shape[0]:
            raise ValueError("Number of donor hydrogens and heavy-atoms do not match.")
        return np.zeros((acc.shape[0], don.shape[0]), dtype=np.int32, order="F")
    r = routines.hbond_counts(acc, don, donh, boxl, distcut, angcut, dense=True)
    torch.cuda.current_stream().synchronize()
    return _np(r["dense"][0])


def willarddensityfield(pos, gridx, gridy, gridz, boxl, smoothlen):
    """densvals,densnorms = willarddensityfield(pos,gridx,gridy,gridz,boxl,smoothlen)
    (fortran/waterlib.f90:1286-1341) -> (nx,ny,nz) and (nx,ny,nz,3), Fortran-ordered like f2py's."""
    dens, norms = routines.willard_density(_check_pos(pos, "pos"), boxl, smoothlen, grid=(gridx, gridy, gridz))
    return _np(dens), _np(norms)


def willarddensitypoints(pos, denspts, boxl, smoothlen):
    """densvals,densnorms = willarddensitypoints(pos,denspts,boxl,smoothlen)   (fortran/waterlib.f90:1351-1398)"""
    dens, norms = routines.willard_density(_check_pos(pos, "pos"), boxl, smoothlen, points=_check_pos(denspts, "denspts"))
    return _np(dens), _np(norms)


def interfacewater(pos, gridpos, gridnorm, cutoff, boxl):
    """watclose,surfclose,numwater,allwatdists = interfacewater(pos,gridpos,gridnorm,cutoff,boxl)
    (fortran/waterlib.f90:1414-1469).  Indices are 1-BASED, as the Fortran returns them through f2py;
    0 marks an entry the Fortran would have left unwritten (nothing within distance^2 < 1000)."""
    r = routines.interface_water(_check_pos(pos, "pos"), _check_pos(gridpos, "gridpos"), _check_pos(gridnorm, "gridnorm"),
                                 cutoff, boxl)
    return (_np(r["watclose"]) + 1, _np(r["surfclose"]) + 1, int(r["numwater"].item()), _np(r["allwatdists"]))


def histrr3b(pos, boxl, distwidth, dnum, angwidth, anum):
    """histout = histrr3b(pos,boxl,distwidth,dnum,angwidth,anum)   (fortran/waterlib.f90:1550-1593) -> float64
    (dnum,dnum,anum), Fortran-ordered."""
    h = routines.histrr3b(_check_pos(pos, "pos"), boxl, distwidth, dnum, angwidth, anum)
    return np.asfortranarray(h.cpu().numpy().astype(np.float64))


def radialdist(pos1, pos2, binwidth, totbins, bulkdens, boxl):
    """rdf = radialdist(pos1,pos2,binwidth,totbins,bulkdens,boxl)   (fortran/waterlib.f90:193-231)"""
    p1 = _check_pos(pos1, "pos1")
    counts = routines.pair_hist(0, p1, _check_pos(pos2, "pos2"), boxl, binwidth, totbins)
    return routines.rdf_normalise(counts, p1.shape[0], binwidth, bulkdens)


def radialdistplane(pos1, pos2, binwidth, totbins, bulkdens, boxl):
    """rdf = radialdistplane(pos1,pos2,binwidth,totbins,bulkdens,boxl)   (fortran/waterlib.f90:237-314): counts of the atoms
    of pos2 within 5 A of the plane through the three points pos1, on a totbins x totbins grid of in-plane coordinates.
    The Fortran indexes bin <= 0 (an out-of-bounds write) for atoms with a non-positive in-plane coordinate; that raises
    ValueError here instead of corrupting memory."""
    counts, bad = routines.radial_dist_plane(np.asarray(pos1, dtype=np.float64), _check_pos(pos2, "pos2"), boxl, binwidth, totbins,
                                             bulkdens)
    if bad:
        raise ValueError("radialdistplane: %d atoms of the slab have a non-positive in-plane coordinate (bin <= 0, out of "
                         "bounds in the reference)" % bad)
    return np.asfortranarray(counts.cpu().numpy().astype(np.float64))


def radialdistsame(pos, binwidth, totbins, bulkdens, boxl):
    """rdf = radialdistsame(pos,binwidth,totbins,bulkdens,boxl)   (fortran/waterlib.f90:316-353)"""
    p = _check_pos(pos, "pos")
    return routines.rdf_normalise(routines.pair_hist(1, p, None, boxl, binwidth, totbins), p.shape[0], binwidth, bulkdens)


def pairdistancehistogram(pos1, pos2, binwidth, totbins, boxl):
    """hist = pairdistancehistogram(pos1,pos2,binwidth,totbins,boxl)   (fortran/waterlib.f90:358-389), 3-D positions"""
    counts = routines.pair_hist(2, _check_pos(pos1, "pos1"), _check_pos(pos2, "pos2"), boxl, binwidth, totbins)
    return counts.cpu().numpy().astype(np.float64)


def densityfield(pos, gridx, gridy, gridz, boxl):
    """densvals = densityfield(pos,gridx,gridy,gridz,boxl)   (fortran/waterlib.f90:1219-1268)"""
    return _np(routines.density_field(_check_pos(pos, "pos"), boxl, (gridx, gridy, gridz)))


def watorient(opos, hpos, refvec, boxl):
    """angdip,angplane = watorient(opos,hpos,refvec,boxl)   (fortran/waterlib.f90:973-1011)"""
    dip, plane = routines.water_orient(_check_pos(opos, "opos"), _check_pos(hpos, "hpos"), boxl, refvec)
    return _np(dip[0]), _np(plane[0])


def binongrid(opos, xbins, ybins, zbins):
    """outhist = binongrid(opos,xbins,ybins,zbins)   (fortran/waterlib.f90:1047-1099)"""
    return _np(routines.bin_on_grid(_check_pos(opos, "opos"), xbins, ybins, zbins))
