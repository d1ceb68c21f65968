#!/usr/bin/env python
"""Generates waterorderlib_b200/csrc/wol_mc_table.h: the triangle table of a marching-cubes style surface extraction, derived
here from first principles (nothing is copied from a published table):

  corners   c = x | y << 1 | z << 2 of the unit cube; a corner is INSIDE when its value is above the level
  edges     12, edge id = 4 * axis + (the two other coordinates of its lower end, lower axis first), see EDGES
  faces     on each of the 6 cube faces the crossed edges are joined pairwise into segments.  Two crossings: one segment.
            Four crossings (the two inside corners sit on a diagonal): each segment cuts off ONE inside corner -- inside
            regions never connect across a face diagonal.  Both cubes that share a face see the same corner values and
            so draw the same segments: the surface is watertight.
  loops     every crossed edge lies on two faces, so the segments close into loops; each segment is directed with the inside
            corner on its left seen from outside the cube, which orients all loops alike (normals point to the outside =
            lower values); a loop is triangulated as a fan.

Output: kMcTri[256][16] int8, triples of edge ids, -1 terminated (at most 5 triangles per case occur with this rule... the
generator asserts it), and kMcEdge[12][2] corner pairs.
"""
import itertools
import os

AX = (1, 2, 4)
EDGES = []  # (corner a, corner b, axis)
for axis in range(3):
    others = [a for a in range(3) if a != axis]
    for hi in (0, 1):
        for lo in (0, 1):
            base = (lo << others[0]) | (hi << others[1])
            EDGES.append((base, base | AX[axis], axis))
EDGE_ID = {(a, b): i for i, (a, b, _) in enumerate(EDGES)}


def edge_between(c0, c1):
    return EDGE_ID[(min(c0, c1), max(c0, c1))]


def face_cycles():
    """For each of the 6 faces: its 4 corners in counter-clockwise order seen from OUTSIDE the cube."""
    out = []
    for axis in range(3):
        u, v = [a for a in range(3) if a != axis]
        for side in (0, 1):
            def corner(cu, cv, side=side, axis=axis, u=u, v=v):
                return (side << axis) | (cu << u) | (cv << v)
            cyc = [corner(0, 0), corner(1, 0), corner(1, 1), corner(0, 1)]
            # (u, v, axis) right-handed when (axis - u) % 3 == ... check orientation with a cross product
            import numpy as np
            p = [np.array([(c >> k) & 1 for k in range(3)], dtype=float) for c in cyc]
            n = np.cross(p[1] - p[0], p[2] - p[1])
            outward = np.zeros(3)
            outward[axis] = 1.0 if side else -1.0
            if np.dot(n, outward) < 0:
                cyc = cyc[::-1]
            out.append(cyc)
    return out


FACES = face_cycles()


def case_triangles(case):
    inside = [(case >> c) & 1 for c in range(8)]
    nxt = {}  # directed segments between crossed edges: edge -> next edge
    for cyc in FACES:
        # walk the face boundary counter-clockwise (seen from outside): corners cyc[0..3]; boundary edge k joins cyc[k], cyc[k+1]
        crossed = [k for k in range(4) if inside[cyc[k]] != inside[cyc[(k + 1) % 4]]]
        if not crossed:
            continue
        # A segment runs from the crossing where the boundary walk ENTERS the inside region to the crossing where it LEAVES
        # it ... seen from outside with the inside corner(s) on the left means: start at the leaving crossing of an inside
        # run, end at the entering crossing of the same run, going around the inside corners clockwise -- i.e. for every
        # maximal run of inside corners along the walk, connect (crossing after the run) -> (crossing before the run).
        # With four crossings the two inside corners are separate runs of length one: each is cut off on its own.
        runs = []
        for k in range(4):
            if inside[cyc[k]] and not inside[cyc[(k - 1) % 4]]:
                j = k
                while inside[cyc[(j + 1) % 4]] and (j + 1 - k) < 4:
                    j += 1
                runs.append((k, j))
        for (k, j) in runs:
            before = edge_between(cyc[(k - 1) % 4], cyc[k % 4])
            after = edge_between(cyc[j % 4], cyc[(j + 1) % 4])
            assert after not in nxt
            nxt[after] = before
    tris = []
    seen = set()
    for start in sorted(nxt):
        if start in seen:
            continue
        loop = [start]
        seen.add(start)
        e = nxt[start]
        while e != start:
            loop.append(e)
            seen.add(e)
            e = nxt[e]
        assert len(loop) >= 3
        for k in range(1, len(loop) - 1):
            tris.append((loop[0], loop[k + 1], loop[k]))  # this order makes the normal (right-hand rule) point to the OUTSIDE
    return tris


def check_orientation(rows):
    """Case 1 (only corner 0 inside): the triangle's normal must point away from corner 0."""
    import numpy as np
    mid = [0.5 * (np.array([(a >> k) & 1 for k in range(3)], float) + np.array([(b >> k) & 1 for k in range(3)], float)) for a, b, _ in EDGES]
    t = rows[1][:3]
    p0, p1, p2 = (mid[e] for e in t)
    n = np.cross(p1 - p0, p2 - p0)
    assert np.dot(n, (p0 + p1 + p2) / 3.0) > 0, "normals must point towards lower values"


def main():
    rows = []
    most = 0
    for case in range(256):
        t = case_triangles(case)
        most = max(most, len(t))
        flat = [e for tri in t for e in tri]
        assert len(flat) <= 15
        rows.append(flat + [-1] * (16 - len(flat)))
    # sanity: complementary cases use the same edges; the number of crossed edges matches
    for case in range(256):
        crossed = {i for i, (a, b, _) in enumerate(EDGES) if ((case >> a) & 1) != ((case >> b) & 1)}
        used = {e for e in rows[case] if e >= 0}
        assert used == crossed, (case, used, crossed)
    check_orientation(rows)
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "waterorderlib_b200", "csrc", "wol_mc_table.h")
    with open(path, "w") as f:
        f.write("// GENERATED by scripts/make_mc_table.py (derivation and conventions there) -- do not edit.\n")
        f.write("// Triangles of the iso-surface inside one grid cube, per corner configuration (bit c set = corner c above the level,\n")
        f.write("// c = x | y << 1 | z << 2): triples of cube-edge ids, -1 terminated; at most %d triangles.  kMcEdge: the two corners of\n" % most)
        f.write("// each edge and its axis.\n#pragma once\n\nnamespace wol {\n\n")
        f.write("__device__ __constant__ signed char kMcTri[256][16] = {\n")
        for r in rows:
            f.write("    {" + ", ".join("%d" % v for v in r) + "},\n")
        f.write("};\n\n__device__ __constant__ unsigned char kMcEdge[12][3] = {\n")
        for a, b, ax in EDGES:
            f.write("    {%d, %d, %d},\n" % (a, b, ax))
        f.write("};\n\n}  // namespace wol\n")
    print("wrote", path, "max triangles per cube", most)


if __name__ == "__main__":
    main()
