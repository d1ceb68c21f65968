"""Golden fixtures for the round-2 additions, from the LIVE reference (run in the build container only):

    python tests/golden/make_golden_extras.py

  clusters.npz   getClusters (structureLibs/orderParam_lib.py:123-156, AST-extracted, unmodified) over the reference's
                 compiled sortlib.depthfirstsort (fortran/sortlib.f90:26-72) on random symmetric 0/1 matrices
  rdfplane.npz   RadialDistPlane (fortran/waterlib.f90:237-314) from the reference's compiled waterlib; its matmul runs in
                 oracle/gfortran_stub.c (libgfortran 5's accumulation order).  Inputs keep every slab atom's in-plane
                 coordinates positive: the Fortran writes out of bounds otherwise.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import build_oracle, port, ref_fortran  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def main():
    build_oracle.build(verbose=False)
    sl = ref_fortran.RefSortlib()
    fn = ref_fortran.load_reference_driver_functions(["getClusters"], sl)
    rng = np.random.default_rng(2024)
    out = {}
    cases = [(30, 0.03), (50, 0.02), (64, 0.012), (12, 0.5), (8, 0.0), (40, 0.08), (1, 0.0)]
    for k, (n, p) in enumerate(cases):
        m = (rng.random((n, n)) < p).astype(int)
        m = np.triu(m, 1)
        m = m + m.T
        cl = fn["getClusters"](m)
        out["mat%d" % k] = m
        out["sizes%d" % k] = np.array([len(c) for c in cl], dtype=np.int64)
        out["members%d" % k] = np.concatenate(cl).astype(np.int64)
        print("clusters case %d: n=%d, %d clusters, largest %d" % (k, n, len(cl), max(len(c) for c in cl)))
    out["n_cases"] = len(cases)
    np.savez_compressed(os.path.join(OUT, "clusters.npz"), **out)

    wl = ref_fortran.RefWaterlib()
    L = np.array([40.0, 44.0, 36.0])
    # a tilted plane through three points; atoms in the positive octant of its frame, inside and outside the 5 A slab
    p1 = np.array([[1.0, 1.0, 1.0], [1.2, 4.0, 1.3], [4.0, 0.8, 1.2]])  # |v1 x v2| stays below L / 2: the Fortran min-images it
    p2 = rng.random((2000, 3)) * np.array([14.0, 14.0, 12.0]) + np.array([3.0, 3.0, 0.5])
    res = {}
    for tag, bw, nb in (("a", 0.5, 40), ("b", 0.37, 25)):
        rdf = wl.radialdistplane(p1, p2, bw, nb, 0.0334, L)
        mine, bad = port.radialdistplane(p1, p2, bw, nb, 0.0334, L)
        assert bad == 0, "pick inputs the reference can take: %d atoms index bin <= 0" % bad
        assert np.array_equal(rdf, mine), "restatement differs from the compiled Fortran"
        res["rdf_" + tag] = rdf
        res["binwidth_" + tag] = bw
        res["totbins_" + tag] = nb
        print("rdfplane %s: %d atoms counted" % (tag, int(rdf.sum())))
    np.savez_compressed(os.path.join(OUT, "rdfplane.npz"), pos1=p1, pos2=p2, box=L, bulkdens=0.0334, **res)


if __name__ == "__main__":
    main()
