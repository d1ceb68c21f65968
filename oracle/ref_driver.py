"""TEST INFRASTRUCTURE ONLY -- the reference's per-frame Python loops restated over its compiled Fortran.

On a GPU box the reference's Python source is not present (only its prebuilt Fortran binary is staged
into oracle/_ref/), so the CPU baseline there runs these restatements of

    getCosAngs       structureLibs/water_properties.py:210-250
    getOrderParamq   structureLibs/water_properties.py:344-391
    getClusters      structureLibs/orderParam_lib.py:123-156 (over sortlib.depthfirstsort)

on top of ``RefWaterlib`` (the reference's own compiled ``allnearneighbors`` / ``nearneighbors`` /
``reimage`` / ``tetracosang``).  Same call sequence per water, same NumPy calls, so the timing has the
reference's cost structure (dense N x N matrix, one f2py-style call pair per molecule) and the results
are bit-identical to the live bodies (tests/test_oracle_vs_reference.py::test_ref_driver_matches_live).
"""
import numpy as np


def get_cos_angs(wl, subPos, Pos, BoxDims, lowCut=0.0, highCut=3.413):
    same = np.array_equal(subPos, Pos)
    mat = (wl.allnearneighbors(Pos, BoxDims, lowCut, highCut) if same
           else wl.nearneighbors(subPos, Pos, BoxDims, lowCut, highCut)).astype(bool)
    chunks = []
    num = np.zeros(len(subPos))
    for i, centre in enumerate(subPos):
        nb = Pos[mat[i]]
        if len(nb) > 0:
            ang = wl.tetracosang(centre, nb, BoxDims)
            chunks.append(ang[np.triu_indices(len(ang), k=1)])
            num[i] = ang.shape[0]
    return (np.concatenate(chunks) if chunks else np.array([])), num


def get_order_param_q(wl, subPos, Pos, BoxDims, lowCut=0.0, highCut=10.0):
    same = np.array_equal(subPos, Pos)
    mat = (wl.allnearneighbors(Pos, BoxDims, lowCut, highCut) if same
           else wl.nearneighbors(subPos, Pos, BoxDims, lowCut, highCut)).astype(bool)
    q = np.zeros(len(subPos))
    for i, centre in enumerate(subPos):
        k = int(np.sum(mat[i]))
        if k == 0:
            continue
        img = wl.reimage(Pos[mat[i]], centre, BoxDims)
        dist = np.linalg.norm(img - centre, axis=1)
        four = img[np.argsort(dist)][:4]
        ang = wl.tetracosang(centre, four, BoxDims)
        vals = ang[np.triu_indices(len(ang), k=1)]
        if k == 1:
            vals = np.full(6, 180.0)
        elif k == 2:
            vals = np.concatenate((vals, np.full(5, 180.0)))
        elif k == 3:
            vals = np.concatenate((vals, np.full(3, 180.0)))
        q[i] = 1.0 - (3.0 / 8.0) * np.sum((np.cos(vals * np.pi / 180.0) + (1.0 / 3.0)) ** 2)
    return q


def get_clusters(sortlib, hbMat):
    """Clusters of a residue-connectivity matrix through the reference's compiled depth-first search, in the order and
    with the conventions of the reference's loop: residues already placed are skipped, an unconnected residue is a
    cluster of one, a cluster that spans everything ends the search."""
    n = hbMat.shape[0]
    clusters = []
    placed = np.zeros(n, dtype=bool)
    for i in range(n):
        if placed[i]:
            continue
        seen = sortlib.depthfirstsort(i + 1, hbMat, np.zeros(n, dtype=int), np.int64(np.sum(hbMat[i, :])), n)
        members = np.where(seen == 1)[0]
        placed[members] = True
        if len(members) == n:
            clusters.append(members)
            break
        clusters.append(members if len(members) else np.array([i]))
    return clusters
