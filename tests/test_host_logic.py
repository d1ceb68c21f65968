"""CPU tests of the host-side logic around the kernels: Amber-mask selection and TrajObject, the H-bond index
bookkeeping, the bootstrap statistics, and the multi-rank frame sharding / reduction (world_size 2, gloo)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from waterorderlib_b200 import distributed as wdist
from waterorderlib_b200.structureLibs.TrajObject import ArrayTrajectory, Topology, TrajObject

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def small_top():
    names = ["C1", "O1", "HO1", "N1", "HN1", "S1"] + ["O", "H1", "H2", "EPW"] * 3
    resn = ["GLY"] * 6 + ["WAT"] * 12
    resid = [0] * 6 + [1] * 4 + [2] * 4 + [3] * 4
    bonds = [(1, 2), (3, 4), (0, 1), (0, 3)] + [(6 + 4 * k, 7 + 4 * k) for k in range(3)] + [(6 + 4 * k, 8 + 4 * k) for k in range(3)]
    return Topology(names, resn, resid, bonds)


def test_amber_masks():
    top = small_top()
    assert list(top.select("(:WAT)")) == list(range(6, 18))
    assert list(top.select("(!:WAT)")) == list(range(6))
    assert list(top.select("(:WAT)&(!@H=)&(!@EP=)")) == [6, 10, 14]
    assert list(top.select("(:WAT)&(@H=)")) == [7, 8, 11, 12, 15, 16]
    assert list(top.select("(@C=)|(@S=)")) == [0, 5]
    assert list(top.select("(!:WAT)&(!@H=)")) == [0, 1, 3, 5]
    assert list(top.select(":GLY & @O1,N1")) == [1, 3]
    with pytest.raises(ValueError):
        top.select("(:WAT")
    assert top.n_residues() == 4 and top.n_residues("(!:WAT)") == 1


def test_trajobject_indices_and_frames(tmp_path):
    top = small_top()
    xyz = np.arange(5 * 18 * 3, dtype=np.float64).reshape(5, 18, 3)
    traj = ArrayTrajectory(xyz, np.array([10.0, 11.0, 12.0]))
    obj = TrajObject(top, traj, stride=2)
    assert len(obj.traj) == 3 and np.array_equal(obj.traj[1].xyz, xyz[2])
    assert np.array_equal(obj.traj[0].box.values, [10.0, 11.0, 12.0, 90.0, 90.0, 90.0])
    watInds, watHInds, lenWat = obj.getWatInds()
    assert list(watInds) == [6, 10, 14] and lenWat == 4 and len(watHInds) == 6
    sol = obj.getSolInds()
    assert [list(s) for s in sol] == [[0, 1, 3, 5], [2, 4], [0], [3], [1], [5]]
    assert list(obj.getHeavyInds()) == [0, 1, 3, 5, 6, 10, 14]
    # .npz round trip
    top.save(tmp_path / "top.npz")
    traj.save(tmp_path / "traj.npz")
    obj2 = TrajObject(str(tmp_path / "top.npz"), str(tmp_path / "traj.npz"))
    assert len(obj2.traj) == 5 and list(obj2.getWatInds()[0]) == [6, 10, 14]
    assert [f.xyz[0, 0] for f in obj2.traj] == [0.0, 54.0, 108.0, 162.0, 216.0]


def test_getHBInds():
    from waterorderlib_b200.structureLibs import orderParam_lib as opl
    top = small_top()
    hbO, hbN = opl.getHBInds(top, None, [0, 1, 3, 5], [2, 4], [3], [1])
    assert [list(x) for x in hbO] == [[1], [1], [2]] and [list(x) for x in hbN] == [[3], [3], [4]]
    wat = [6, 10, 14]
    hbW, _ = opl.getHBInds(top, None, wat, [7, 8, 11, 12, 15, 16], [], wat)
    assert list(hbW[0]) == wat and list(hbW[1]) == [6, 6, 10, 10, 14, 14] and list(hbW[2]) == [7, 8, 11, 12, 15, 16]


def test_blockAverage_matches_reference_loop():
    from waterorderlib_b200.structureLibs import orderParam_lib as opl
    vals = np.random.default_rng(0).normal(size=173)

    def reference(vals, nBlocks=20):  # structureLibs/orderParam_lib.py:394-417, literally
        obsBlocks = np.zeros(nBlocks)
        lenBlock = len(vals) / nBlocks
        for i in range(nBlocks):
            obsBlocks[i] = np.mean(vals[int(i * lenBlock):int((i + 1) * lenBlock)])
        obsMeans = np.zeros(10000)
        for n in range(10000):
            obsMeans[n] = np.mean(np.random.choice(obsBlocks, nBlocks))
        return opl.getCI(np.sort(obsMeans))

    np.random.seed(11)
    want = reference(vals)
    np.random.seed(11)
    assert opl.blockAverage(vals) == want


def test_shard_frames_partition():
    for n in (0, 1, 7, 8, 1000):
        for ws in (1, 2, 3, 8):
            blocks = [wdist.shard_frames(n, r, ws) for r in range(ws)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(ws - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1 and sizes == wdist.shard_sizes(n, ws)


def _rank_main(rank, ws, port, n_frames, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        b, e = wdist.shard_frames(n_frames)
        # every frame contributes a known histogram and a known stats row
        g = torch.Generator().manual_seed(1)
        all_hist = torch.randint(0, 1000, (n_frames, 50), generator=g, dtype=torch.int64)
        all_rows = torch.rand((n_frames, 8), generator=g, dtype=torch.float64)
        h1 = all_hist[b:e].sum(dim=0)
        h2 = (all_hist[b:e] * 3).sum(dim=0).reshape(5, 10)
        wdist.reduce_histograms(h1, h2)
        rows = wdist.gather_frame_rows(all_rows[b:e].clone(), n_frames)
        mx = wdist.max_over_ranks(float(rank + 1), torch.device("cpu"))
        ok = (torch.equal(h1, all_hist.sum(dim=0)) and torch.equal(h2.reshape(-1), all_hist.sum(dim=0) * 3)
              and torch.equal(rows, all_rows) and mx == float(ws))
        open(os.path.join(out_dir, "ok%d" % rank), "w").write("1" if ok else "0")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [7, 8])
def test_two_rank_reduce_and_gather_gloo(tmp_path, n_frames):
    port = 29500 + (os.getpid() % 2000) + n_frames
    mp.spawn(_rank_main, args=(2, port, n_frames, str(tmp_path)), nprocs=2, join=True)
    assert [open(tmp_path / ("ok%d" % r)).read() for r in range(2)] == ["1", "1"]


def test_reduce_rejects_non_integer_histograms_single_rank():
    h = torch.zeros(4, dtype=torch.int64)
    assert wdist.reduce_histograms(h)[0] is h  # identity without a process group
    assert wdist.world() == (0, 1)
    rows = torch.ones((3, 2))
    assert wdist.gather_frame_rows(rows, 3) is rows


def test_amber_parm7_and_netcdf_readers(tmp_path):
    """TrajObject on AMBER files without parmed / pytraj: parm7 topology + NetCDF-3 trajectory round trip."""
    from waterorderlib_b200.structureLibs import amber_io
    from waterorderlib_b200.structureLibs import orderParam_lib as opl
    top = small_top()
    rng = np.random.default_rng(0)
    xyz = (rng.random((7, 18, 3)) * 20.0).astype(np.float32)
    boxes = np.array([[20.0, 21.0, 22.0]] * 7) + np.arange(7)[:, None] * 0.125
    amber_io.write_parm7(tmp_path / "sys.parm7", top)
    amber_io.write_netcdf(str(tmp_path / "sys.nc"), xyz, boxes)
    obj = TrajObject(str(tmp_path / "sys.parm7"), str(tmp_path / "sys.nc"), stride=2)
    assert list(obj.top.names) == list(top.names) and list(obj.top.resnames) == list(top.resnames)
    assert sorted(map(tuple, obj.top.bonds.tolist())) == sorted(map(tuple, top.bonds.tolist()))
    assert list(obj.getWatInds()[0]) == [6, 10, 14] and [list(s) for s in obj.getSolInds()][0] == [0, 1, 3, 5]
    assert len(obj.traj) == 4
    for k, frame in enumerate(obj.traj):
        assert frame.xyz.dtype == np.float32 and frame.xyz.dtype.isnative and np.array_equal(frame.xyz, xyz[2 * k])
        assert np.array_equal(frame.box.values[:3], boxes[2 * k]) and list(frame.box.values[3:]) == [90.0, 90.0, 90.0]
    block = obj.traj.xyz[1:3]
    assert block.shape == (2, 18, 3) and block.dtype.isnative and np.array_equal(block, xyz[[2, 4]])
    assert np.array_equal(obj.traj.xyz[1:3, [6, 10]], xyz[[2, 4]][:, [6, 10]])
    hbO, _ = opl.getHBInds(obj.top, obj.traj[0], [6, 10, 14], [7, 8, 11, 12, 15, 16], [], [6, 10, 14])
    assert list(hbO[1]) == [6, 6, 10, 10, 14, 14]
    obj.traj.close()


def test_legacy_simps_and_trajectory_slices(tmp_path):
    """rdfCalc's host pieces: the SciPy < 1.11 ``simps`` rule (even='avg') and pytraj-style traj[a:b] slicing."""
    from scipy.integrate import simpson
    from waterorderlib_b200.structureLibs.amber_io import NetCDFTrajectory, write_netcdf
    from waterorderlib_b200.structureLibs.orderParam_lib import simps
    from waterorderlib_b200.structureLibs.TrajObject import ArrayTrajectory
    rng = np.random.default_rng(3)
    for n in (3, 7, 151):
        x, y = np.cumsum(rng.uniform(0.5, 1.5, n)), rng.normal(size=n)
        assert np.isclose(simps(y, x), simpson(y, x=x), rtol=1e-12)
    for n in (2, 4, 150):
        x, y = np.cumsum(rng.uniform(0.5, 1.5, n)), rng.normal(size=n)
        first, last = 0.5 * (x[1] - x[0]) * (y[1] + y[0]), 0.5 * (x[-1] - x[-2]) * (y[-1] + y[-2])
        a = (simpson(y[:-1], x=x[:-1]) if n > 2 else 0.0) + last
        b = (simpson(y[1:], x=x[1:]) if n > 2 else 0.0) + first
        assert np.isclose(simps(y, x), 0.5 * (a + b), rtol=1e-12)
    assert np.isclose(simps(np.arange(6.0) ** 2, np.arange(6.0)), 125.0 / 3.0, rtol=0.01)
    xyz = rng.normal(size=(7, 5, 3)).astype(np.float32)
    boxes = np.tile(np.array([10.0, 11.0, 12.0]), (7, 1)) + np.arange(7)[:, None]
    t = ArrayTrajectory(xyz, boxes)
    sub = t[2:5]
    assert len(sub) == 3 and np.array_equal(sub[0].xyz, xyz[2]) and np.array_equal(sub[2].box.values[:3], boxes[4])
    assert [f.box.values[0] for f in t[5:99]] == [15.0, 16.0]
    path = str(tmp_path / "t.nc")
    write_netcdf(path, xyz, boxes)
    nc = NetCDFTrajectory(path)
    sub = nc[1:4]
    assert len(sub) == 3 and np.array_equal(sub[1].xyz, xyz[2]) and np.array_equal(sub[1].box.values[:3], boxes[2])
    nc.close()
    with pytest.raises(TypeError):
        t["a"]


def test_product_never_touches_the_oracle_and_has_no_cpu_fallback():
    """The oracle is test infrastructure: nothing under waterorderlib_b200/ may import it (only tests/, smoke() and
    bench.py's CPU legs do), and without a CUDA device the product path raises instead of computing on the CPU."""
    import ast
    import pathlib
    root = pathlib.Path(__file__).resolve().parents[1]
    offenders = []
    import itertools
    for path in itertools.chain((root / "waterorderlib_b200").rglob("*.py"), (root / "scripts").glob("*.py"),
                                (root / "examples").glob("*.py")):
        for node in ast.walk(ast.parse(path.read_text())):
            names = []
            if isinstance(node, ast.Import):
                names = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                names = [node.module or ""]
            if any(n == "oracle" or n.startswith("oracle.") for n in names):
                offenders.append(str(path))
    assert not offenders
    for path in (root / "waterorderlib_b200" / "csrc").glob("*"):
        assert "oracle" not in path.read_text(), path          # the kernels do not include or link it either
    import torch
    if not torch.cuda.is_available():
        from waterorderlib_b200 import synth
        from waterorderlib_b200.structureLibs import water_properties as wp
        pos, box = synth.water_box(3, sigma=0.3, seed=1)
        with pytest.raises(Exception):
            wp.getOrderParamq(pos, pos, box)
        with pytest.raises(Exception):
            wp.getCosAngs(pos, pos, box)


def test_vectorised_getHBInds_equals_the_per_atom_loop():
    """getHBInds over the built-in Topology avoids the reference's Python loop over every atom (orderParam_lib.py:46-120);
    the lists must come out identical, whatever the order and orientation of the bond list."""
    from waterorderlib_b200.structureLibs import orderParam_lib as opl
    from waterorderlib_b200.structureLibs.TrajObject import Topology
    rng = np.random.default_rng(2)
    n_sol, n_w = 6, 40
    names, resn, bonds = [], [], []
    for s in range(n_sol):                       # C, O-H, N-H2 cosolvent
        b = 6 * s
        names += ["C1", "O1", "HO1", "N1", "HN1", "HN2"]
        resn += ["SOL"] * 6
        bonds += [(b, b + 1), (b + 1, b + 2), (b, b + 3), (b + 3, b + 4), (b + 3, b + 5)]
    top0 = Topology.water_box(n_w)
    off = len(names)
    names += list(top0.names)
    resn += list(top0.resnames)
    bonds += [(int(a) + off, int(b) + off) for a, b in top0.bonds]
    bonds = np.array(bonds)
    perm = rng.permutation(len(bonds))
    bonds = bonds[perm]
    flip = rng.random(len(bonds)) < 0.5
    bonds[flip] = bonds[flip][:, ::-1]
    top = Topology(names, resn, None, bonds)
    solO, solN = top.select("(!:WAT)&(@O=)"), top.select("(!:WAT)&(@N=)")
    wat = top.select("(:WAT)&(!@H=)")

    class LoopTop:                               # anything without arrays takes the reference-style loop
        atoms, residues = top.atoms, ()

    for o_set, n_set in ((solO, solN), (wat, []), (np.concatenate([solO, wat]), solN)):
        fast = opl.getHBInds(top, None, None, None, n_set, o_set)
        slow = opl.getHBInds(LoopTop, None, None, None, n_set, o_set)
        for a, b in zip(fast, slow):
            for x, y in zip(a, b):
                assert np.array_equal(x, y)
    assert len(fast[1][1]) == 2 * n_sol and len(fast[0][1]) == n_sol + 2 * n_w


def test_threaded_staging_copy_is_a_plain_copy():
    """The frame drivers' host staging copy (orderParam_lib._host_copy) over several threads = np.copyto."""
    from waterorderlib_b200.structureLibs import orderParam_lib as opl
    rng = np.random.default_rng(5)
    for n in (0, 1, 4097, (8 << 20) - 1, (8 << 20) + 12345, 3 * (8 << 20) + 7):
        src = rng.integers(0, 256, n, dtype=np.uint8)
        dst = np.zeros(n, dtype=np.uint8)
        opl._host_copy(dst, src)
        assert np.array_equal(dst, src)
