"""Parity of the frame drivers (waterorderlib_b200.structureLibs.orderParam_lib: tetOrderCalc, threeBodyCalc,
hbCalc, getBoundWrap, getNeighborStats) against the reference's driver arithmetic restated here on top of
the CPU oracle (oracle/port.py), frame by frame, the way structureLibs/orderParam_lib.py:1426-1503, :1269-1424,
:729-917, :419-572 loop.  Small synthetic trajectories; the bootstrap CIs use numpy's global RNG exactly as
the reference does, so seeding it makes them comparable too.
"""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import port  # noqa: E402  (the checker)
from waterorderlib_b200 import synth  # noqa: E402
from waterorderlib_b200.structureLibs import orderParam_lib as opl  # noqa: E402
from waterorderlib_b200.structureLibs.TrajObject import ArrayTrajectory, Topology, TrajObject  # noqa: E402

SOL_NAMES = ["C1", "O1", "HO1", "N1", "HN1"]
SOL_BONDS = [(0, 1), (1, 2), (0, 3), (3, 4)]


def make_system(m, n_frames, n_sol=0, sigma=0.3, seed0=100):
    """Waters (O,H1,H2 contiguous) preceded by n_sol five-atom cosolvent molecules (C, O-H, N-H)."""
    n_w = 8 * m ** 3
    names, resn, resid, bonds = [], [], [], []
    for s in range(n_sol):
        names += SOL_NAMES
        resn += ["SOL"] * 5
        resid += [s] * 5
        bonds += [(5 * s + a, 5 * s + b) for a, b in SOL_BONDS]
    o0 = 5 * n_sol
    names += ["O", "H1", "H2"] * n_w
    resn += ["WAT"] * (3 * n_w)
    resid += list(np.repeat(np.arange(n_w) + n_sol, 3))
    oi = o0 + 3 * np.arange(n_w)
    bonds += [(int(o), int(o) + 1) for o in oi] + [(int(o), int(o) + 2) for o in oi]
    top = Topology(names, resn, resid, bonds)
    xyz = np.zeros((n_frames, len(names), 3))
    boxes = np.zeros((n_frames, 3))
    rng = np.random.default_rng(seed0)
    for f in range(n_frames):
        o, box = synth.water_box(m, sigma=sigma, seed=seed0 + f)
        h = synth.add_hydrogens(o, seed=seed0 + f)
        xyz[f, oi] = o
        xyz[f, oi + 1] = h[0::2]
        xyz[f, oi + 2] = h[1::2]
        boxes[f] = box
        for s in range(n_sol):
            c = rng.random(3) * box
            u = rng.normal(size=(2, 3))
            u /= np.linalg.norm(u, axis=1, keepdims=True)
            xyz[f, 5 * s + 0] = c
            xyz[f, 5 * s + 1] = c + 1.43 * u[0]
            xyz[f, 5 * s + 2] = c + 1.43 * u[0] + 0.96 * u[1]
            xyz[f, 5 * s + 3] = c - 1.47 * u[0]
            xyz[f, 5 * s + 4] = c - 1.47 * u[0] - 1.01 * u[1]
    xyz = xyz.astype(np.float32).astype(np.float64)
    return top, ArrayTrajectory(xyz, boxes, top=top)


@pytest.fixture()
def in_tmp(tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)
    return tmp_path


def test_tetOrderCalc_matches_reference_loop(in_tmp):
    T = 40
    top, traj = make_system(3, T)
    obj = TrajObject(top, traj)
    watInds, _, lenWat = obj.getWatInds()
    assert lenWat == 3 and len(watInds) == 216
    rng = np.random.default_rng(3)
    subInds = [[np.sort(rng.choice(watInds, 30, replace=False)), watInds[rng.random(216) < 0.1]] for _ in range(T)]
    # reference loop (orderParam_lib.py:1458-1480) on the oracle
    avg = np.zeros((3, T)); var = np.zeros((3, T)); qall = [[], [], []]
    for t in range(T):
        pos, box = traj.xyz[t], traj.boxes[t]
        watPos = pos[watInds]
        for j, sub in enumerate([watPos, pos[subInds[t][0]], pos[subInds[t][1]]]):
            q = port.getOrderParamq(sub, watPos, box)
            qall[j].append(q)
            avg[j, t], var[j, t] = np.mean(q), np.var(q)
    np.random.seed(7)
    ref_ci = [opl.blockAverage(avg[j]) for j in range(3)]
    np.random.seed(7)
    avgQ, varQ = opl.tetOrderCalc(top, traj, subInds=subInds, nPops=2)
    assert np.allclose(avgQ[0], avg.mean(axis=1), rtol=1e-9) and np.allclose(varQ[0], var.mean(axis=1), rtol=1e-7)
    assert avgQ[1].shape == (3,) and abs(avgQ[1][0] - ref_ci[0]) < 1e-9  # first CI drawn from the same RNG state
    for j in range(3):
        got = np.loadtxt("qDistribution_%d.txt" % j)
        want, edges = np.histogram(np.concatenate(qall[j]), bins=500, range=[0.0, 1.0])
        assert got.shape == (500, 2) and np.array_equal(got[:, 1], np.array([float("%.3e" % v) for v in want]))
        assert np.allclose(got[:, 0], 0.5 * (edges[:-1] + edges[1:]), rtol=2e-3)


def test_threeBodyCalc_matches_reference_loop(in_tmp):
    T = 24
    top, traj = make_system(3, T, sigma=0.35, seed0=300)
    obj = TrajObject(top, traj)
    watInds, _, _ = obj.getWatInds()
    rng = np.random.default_rng(5)
    subInds = [[np.sort(rng.choice(watInds, 40, replace=False))] for _ in range(T)]
    series = np.zeros((2, 4, T)); pooled = [[], []]
    for t in range(T):
        pos, box = traj.xyz[t], traj.boxes[t]
        watPos = pos[watInds]
        for j, sub in enumerate([watPos, pos[subInds[t][0]]]):
            ang, _ = port.getCosAngs(sub, watPos, box)
            pooled[j].append(ang)
            _, _, a, b, c, d = port.tetrahedralMetrics(ang)
            series[j, :, t] = (a, b, c, d)
    pTet, avgCos, varCos, entropy, nWats = opl.threeBodyCalc(top, traj, subInds=subInds, nPops=1)
    for j in range(2):
        assert abs(pTet[0][j] - series[j, 0].mean()) < 1e-12
        assert abs(avgCos[0][j] - series[j, 1].mean()) < 1e-12
        assert abs(varCos[0][j] - series[j, 2].mean()) < 1e-12
        assert abs(entropy[0][j] - series[j, 3].mean()) < 1e-12
        got = np.loadtxt("3bDistribution_%d.txt" % j)
        want = port.histogram(np.concatenate(pooled[j]), 500, 0.0, 180.0)
        assert np.array_equal(got[:, 1], np.array([float("%.3e" % v) for v in want]))
    assert nWats[0][0] == 216 and nWats[0][1] == 40


def ref_hbcalc(top, traj, obj):
    """hbCalc's arithmetic (orderParam_lib.py:742-884) with oracle matrices."""
    watInds, watHInds, _ = obj.getWatInds()
    solInds, solHInds, _c, solNInds, solOInds, _s = obj.getSolInds()
    (sAccO, sDonO, sDonHO), (sAccN, sDonN, sDonHN) = opl.getHBInds(top, traj[0], solInds, solHInds, solNInds, solOInds)
    (wAcc, wDon, wDonH), _ = opl.getHBInds(top, traj[0], watInds, watHInds, [], watInds)
    nSol = top.n_residues('(!:WAT)')
    nAccO, nAccN, nDonO, nDonN = (int(len(x) / nSol) for x in (sAccO, sAccN, sDonO, sDonN))
    numWat, numSol = [], []
    for t in range(len(traj)):
        pos, box = traj.xyz[t], traj.boxes[t]
        hb = lambda a, d, h: port.hbonds(pos[a], pos[d], pos[h], box, 3.5, 120.0, dense=True)[2]  # noqa: E731
        ww, wsO, swO = hb(wAcc, wDon, wDonH), hb(wAcc, sDonO, sDonHO), hb(sAccO, wDon, wDonH)
        wsN, swN = hb(wAcc, sDonN, sDonHN), hb(sAccN, wDon, wDonH)
        OO, ON, NO, NN = hb(sAccO, sDonO, sDonHO), hb(sAccO, sDonN, sDonHN), hb(sAccN, sDonO, sDonHO), hb(sAccN, sDonN, sDonHN)
        solOAcc = swO.sum(1) + OO.sum(1) + ON.sum(1)
        solODon = wsO.sum(0) + OO.sum(0) + NO.sum(0)
        solOAcc = sum(solOAcc[i::nAccO] for i in range(nAccO))
        solODon = sum(solODon[i::nDonO] for i in range(nDonO))
        solNAcc = swN.sum(1) + NN.sum(1) + NO.sum(1)
        solNDon = wsN.sum(0) + NN.sum(0) + ON.sum(0)
        solNAcc = sum(solNAcc[i::nAccN] for i in range(nAccN))
        solNDon = sum(solNDon[i::nDonN] for i in range(nDonN))
        numSol.append(solNAcc + solNDon + solOAcc + solODon)
        d = ww.sum(0); dO = swO.sum(0); dN = swN.sum(0)
        numWat.append(ww.sum(1) + d[::2] + d[1::2] + wsO.sum(1) + dO[::2] + dO[1::2] + wsN.sum(1) + dN[::2] + dN[1::2])
    return np.concatenate(numWat), np.concatenate(numSol)


def test_hbCalc_matches_reference_loop(in_tmp):
    top, traj = make_system(3, 6, n_sol=5, seed0=500)
    obj = TrajObject(top, traj)
    numWat, numSol = ref_hbcalc(top, traj, obj)
    avgW, avgS = opl.hbCalc(top, traj)
    assert avgW == np.mean(numWat) and avgS == np.mean(numSol)
    assert numSol.sum() > 0  # the cosolvent really bonds in this system
    got = np.loadtxt("hbDistribution_water.txt")
    assert np.array_equal(got[:, 1], np.histogram(numWat, bins=np.arange(11))[0].astype(float))
    got = np.loadtxt("hbDistribution_cosolv.txt")
    assert np.array_equal(got[:, 1], np.histogram(numSol, bins=np.arange(11))[0].astype(float))


def test_getBoundWrap_matches_reference(in_tmp):
    top, traj = make_system(4, 1, n_sol=6, seed0=700)
    obj = TrajObject(top, traj)
    watInds, watHInds, _ = obj.getWatInds()
    solInds, solHInds, solCInds, solNInds, solOInds, solSInds = obj.getSolInds()
    frame = traj[0]
    bound, wrap, shell, non = opl.getBoundWrap(top, frame, watInds, watHInds, solInds, solHInds, solCInds, solOInds,
                                               solNInds, solSInds)
    # reference arithmetic (orderParam_lib.py:495-570) with oracle matrices
    pos, box = traj.xyz[0], traj.boxes[0]
    nb = port.neighbor_matrix(pos[solInds], pos[watInds], box, 0.0, 4.0)
    mask = np.unique(np.where(nb == 1)[1])
    shell_ref = watInds[mask]
    (sAccO, sDonO, sDonHO), _ = opl.getHBInds(top, frame, solInds, solHInds, solNInds, solOInds)
    (wAcc, wDon, wDonH), _ = opl.getHBInds(top, frame, shell_ref, watHInds, solNInds, shell_ref)
    hb = lambda a, d, h: port.hbonds(pos[a], pos[d], pos[h], box, 3.0, 150.0, dense=True)[2]  # noqa: E731
    watSol, solWat = hb(wAcc, sDonO, sDonHO), hb(sAccO, wDon, wDonH)
    b_wat = np.unique(np.where(watSol == 1)[0])
    dummy = np.zeros(len(wDon)); dummy[np.unique(np.where(solWat == 1)[1])] = 1
    b_sol = np.where(np.ceil(0.5 * (dummy[0::2] + dummy[1::2])))[0]
    bmask = np.sort(np.unique(np.concatenate([b_wat, b_sol]))).astype(int)
    keep = np.ones(len(shell_ref), dtype=bool); keep[bmask] = False
    assert np.array_equal(shell, shell_ref) and np.array_equal(non, np.delete(watInds, mask))
    assert np.array_equal(bound, shell_ref[bmask]) and np.array_equal(wrap, shell_ref[keep])
    assert len(shell) > 0 and len(bound) > 0 and len(wrap) > 0


def test_getNeighborStats(in_tmp):
    top, traj = make_system(3, 3, n_sol=4, seed0=900)
    obj = TrajObject(top, traj)
    watInds, _, _ = obj.getWatInds()
    solInds = obj.getSolInds()[0]
    got = opl.getNeighborStats(top, traj, solInds, watInds, 3, 1, distCut=4.0)
    vals = []
    for t in range(3):
        nb = port.neighbor_matrix(traj.xyz[t][solInds], traj.xyz[t][watInds], traj.boxes[t], 0.0, 4.0)
        for n in range(len(solInds) // 3):
            vals.append(len(np.unique(np.where(nb[3 * n:3 * n + 3] == 1)[1])))
    assert got == np.mean(vals) and os.path.exists("coordDistribution.txt")


def test_lsiCalc_and_hexOrderCalc_match_reference_loops(in_tmp):
    T = 20
    top, traj = make_system(3, T, sigma=0.45, seed0=1100)
    obj = TrajObject(top, traj)
    watInds, _, _ = obj.getWatInds()
    rng = np.random.default_rng(9)
    subInds = [[np.sort(rng.choice(watInds, 50, replace=False))] for _ in range(T)]
    # lsiCalc (orderParam_lib.py:1617-1661)
    avg = np.zeros((2, T)); var = np.zeros((2, T)); pooled = [[], []]
    for t in range(T):
        pos, box = traj.xyz[t], traj.boxes[t]
        for j, sub in enumerate([pos[watInds], pos[subInds[t][0]]]):
            v, _ = port.getLSI(sub, pos[watInds], box)
            pooled[j].append(v); avg[j, t], var[j, t] = np.mean(v), np.var(v)
    avgL, varL = opl.lsiCalc(top, traj, subInds=subInds, nPops=1)
    assert np.allclose(avgL[0], avg.mean(axis=1), rtol=1e-10) and np.allclose(varL[0], var.mean(axis=1), rtol=1e-9)
    for j in range(2):
        got = np.loadtxt("lsiDistribution_%d.txt" % j)
        want = np.histogram(np.concatenate(pooled[j]), bins=500, range=[0.0, 0.3])[0]
        assert np.array_equal(got[:, 1], np.array([float("%.3e" % v) for v in want]))
    # hexOrderCalc (orderParam_lib.py:1537-1582): every second end atom, 0 / 7.0 for the whole set, 0 / 10 for sub-populations
    ends = watInds[1::2]
    subE = [[np.sort(rng.choice(ends, 20, replace=False))] for _ in range(T)]
    avg = np.zeros((2, T)); pooled = [[], []]
    for t in range(T):
        pos, box = traj.xyz[t], traj.boxes[t]
        a = port.getOrderParamPsi(pos[ends], pos[ends], box, 0.0, 7.0)
        b = port.getOrderParamPsi(pos[subE[t][0]], pos[ends], box, 0.0, 10.0)
        avg[0, t], avg[1, t] = a.mean(), b.mean()
        pooled[0].append(a); pooled[1].append(b)
    avgP, varP = opl.hexOrderCalc(top, traj, subInds=subE, nPops=1)
    assert np.allclose(avgP[0], avg.mean(axis=1), rtol=1e-8)
    got = np.loadtxt("psiDistribution_0.txt")
    want = np.histogram(np.concatenate(pooled[0]), bins=500, range=[0.0, 1.0])[0]
    assert abs(got[:, 1] - want).sum() <= 2  # psi values agree to ~1e-12: a value on a bin edge may move one bin


def test_drivers_on_amber_files(in_tmp):
    """tetOrderCalc / hbCalc fed from parm7 + NetCDF files through the built-in readers give what the in-memory
    trajectory gives (frames are float32 in the file; the synthetic coordinates are float32-representable)."""
    from waterorderlib_b200.structureLibs import amber_io
    top, traj = make_system(3, 20, n_sol=3, seed0=1300)
    amber_io.write_parm7("sys.parm7", top)
    amber_io.write_netcdf("sys.nc", traj.xyz, traj.boxes)
    np.random.seed(3)
    a_mem = opl.tetOrderCalc(top, traj)
    hb_mem = opl.hbCalc(top, traj)
    q_mem = np.loadtxt("qDistribution_0.txt")
    np.random.seed(3)
    a_file = opl.tetOrderCalc("sys.parm7", "sys.nc")
    hb_file = opl.hbCalc("sys.parm7", "sys.nc")
    # (double atomics make the per-frame sums order-dependent in the last bits)
    assert np.allclose(a_mem[0][0], a_file[0][0], rtol=1e-12) and np.allclose(a_mem[1][0], a_file[1][0], rtol=1e-10)
    assert hb_mem == hb_file and np.array_equal(q_mem, np.loadtxt("qDistribution_0.txt"))


def _components_ref(mat):
    """Plain union-find reference for getClusters' output convention."""
    n = mat.shape[0]
    parent = list(range(n))

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    for i, j in zip(*np.nonzero(mat == 1)):
        a, b = find(i), find(j)
        if a != b:
            parent[max(a, b)] = min(a, b)
    roots = np.array([find(i) for i in range(n)])
    out = []
    for r in np.unique(roots):
        out.append(np.nonzero(roots == r)[0])
        if len(out[-1]) == n:
            break
    return out


def test_getClusters_and_cluster_stats(in_tmp):
    rng = np.random.default_rng(2)
    for n, p in ((1, 0.0), (7, 0.0), (40, 0.03), (200, 0.004), (60, 0.5)):
        m = (rng.random((n, n)) < p).astype(int)
        m = np.triu(m, 1); m = m + m.T
        got, want = opl.getClusters(m), _components_ref(m)
        assert len(got) == len(want) and all(np.array_equal(a, b) for a, b in zip(got, want))
    chain = np.zeros((300, 300), dtype=int)  # a path graph: the slowest case for label propagation
    idx = np.arange(299); chain[idx, idx + 1] = 1; chain[idx + 1, idx] = 1
    c = opl.getClusters(chain)
    assert len(c) == 1 and len(c[0]) == 300
    # drivers: ion contacts and residue H-bond clusters against matrices from the oracle
    top, traj = make_system(3, 3, n_sol=4, seed0=1500)
    obj = TrajObject(top, traj)
    watInds, watHInds, _ = obj.getWatInds()
    got = opl.getIonClusterStats(top, traj, watInds, np.ones(len(watInds)), distCut=3.0)
    sizes = []
    for t in range(3):
        sizes += [len(x) for x in _components_ref(port.neighbor_matrix(traj.xyz[t][watInds], traj.xyz[t][watInds], traj.boxes[t], 0.0, 3.0))]
    assert got == np.mean(sizes) and os.path.exists("clusterDistribution.txt")
    (wAcc, wDon, wDonH), _ = opl.getHBInds(top, traj[0], watInds, watHInds, [], watInds)
    got = opl.getHBClusterStats(top, traj, wAcc, wDon, wDonH, distCut=3.5, angCut=120.0)
    sizes = []
    res = np.asarray(top.resids)
    for t in range(3):
        pos, box = traj.xyz[t], traj.boxes[t]
        mat = port.hbonds(pos[wAcc], pos[wDon], pos[wDonH], box, 3.5, 120.0, dense=True)[2]
        hb = np.zeros((top.n_residues(), top.n_residues()), dtype=int)
        ai, dj = np.nonzero(mat)
        hb[res[wAcc][ai], res[wDonH][dj]] = 1; hb[res[wDonH][dj], res[wAcc][ai]] = 1
        sizes += [len(x) for x in _components_ref(hb) if len(x) != 1]
    assert got == np.mean(sizes)


def test_rdfCalc_matches_reference_golden(in_tmp, golden_dir):
    """rdfCalc against the fixture made by running the reference's own rdfCalc body over the same trajectory with its
    compiled Fortran (tests/golden/make_golden_rdfcalc.py): return values to 1e-10 relative (the g(r) per frame are
    bit-exact; the Simpson sums are restated), the two written tables to the 3 digits they are printed with."""
    g = np.load(os.path.join(golden_dir, "rdfcalc_n512.npz"))
    xyz, boxes, n_sol = g["xyz"], g["boxes"], int(g["n_sol"])
    n_wat = xyz.shape[1] - n_sol
    top = Topology(["C1"] * n_sol + ["O"] * n_wat, ["SOL"] * n_sol + ["WAT"] * n_wat)
    ret = opl.rdfCalc(top, (xyz, boxes), binwidth=0.1, totbins=110)
    ref = g["ret_sol"]
    assert np.allclose(np.asarray(ret, dtype=np.float64), ref, rtol=1e-10, atol=0)
    for name in ("rdf", "coord"):
        tab, tab_ref = np.loadtxt(name + ".txt"), g[name + "_txt_sol"]
        assert tab.shape == tab_ref.shape and np.allclose(tab, tab_ref, rtol=2.1e-3, atol=1e-12)
        assert np.mean(tab == tab_ref) > 0.999
    # no solute selected: the reference returns (n1_OwOw, index of the last frame of a chunk)
    top_w = Topology(["O"] * n_wat, ["WAT"] * n_wat)
    n1, t = opl.rdfCalc(top_w, (xyz[:, n_sol:], boxes), binwidth=0.1, totbins=110)
    assert np.isclose(n1, g["ret_nosol"][0], rtol=1e-10) and t == int(g["ret_nosol"][1])
    assert np.allclose(np.loadtxt("rdf.txt"), g["rdf_txt_nosol"], rtol=2.1e-3, atol=1e-12)


def test_drivers_accept_pinned_and_cuda_trajectories_and_small_staging_batches(in_tmp, monkeypatch):
    """The batched drivers stage numpy frames through page-locked buffers (several batches, one staged ahead) and take
    torch tensors (pinned host or CUDA) as they are: same numbers whichever way the frames arrive, sub-populations included."""
    T = 23
    top, traj = make_system(4, T)
    obj = TrajObject(top, traj)
    watInds, _, _ = obj.getWatInds()
    rng = np.random.default_rng(5)
    subInds = [[np.sort(rng.choice(watInds, 40, replace=False))] for _ in range(T)]
    monkeypatch.setattr(opl, "_MAX_ATOMS_PER_BATCH", 5 * len(watInds))   # five frames per batch: 5 batches, the last ragged
    results = []
    for kind in ("numpy64", "numpy32", "pinned", "cuda"):
        xyz = traj.xyz
        if kind == "numpy32":
            xyz = xyz.astype(np.float32)
        elif kind == "pinned":
            xyz = torch.from_numpy(xyz).pin_memory()
        elif kind == "cuda":
            xyz = torch.from_numpy(xyz).cuda()
        tr = ArrayTrajectory(xyz, traj.boxes, top=top)
        np.random.seed(11)
        q = opl.tetOrderCalc(top, tr, subInds=subInds, nPops=1)
        qd = np.loadtxt("qDistribution_1.txt")
        np.random.seed(11)
        tb = opl.threeBodyCalc(top, tr, subInds=subInds, nPops=1)
        results.append((q, qd, tb, np.loadtxt("3bDistribution_0.txt")))
    for r in results[1:]:
        for a, b in zip(results[0][0] + results[0][2], r[0] + r[2]):
            # per-frame sums are accumulated with double atomics: equal to rounding, not bit for bit
            assert np.allclose(a[0], b[0], rtol=1e-12, atol=0) and np.allclose(a[1], b[1], rtol=1e-9, atol=1e-15)
        assert np.array_equal(results[0][1], r[1]) and np.array_equal(results[0][3], r[3])


def test_chemPotCalc_matches_reference_loop(in_tmp):
    """Hard-sphere insertion statistics: same np.random call sequence as the reference (orderParam_lib.py:1757-1772 and
    the shell variant :1709-1731), overlap counts from the kernel = row sums of the dense nearNeighbors matrix."""
    T = 3
    top, traj = make_system(3, T, n_sol=4)
    obj = TrajObject(top, traj)
    heavy = top.select('(!@H=)&(!@EPW)')
    sol = obj.getSolInds()[0]
    np.random.seed(21)
    want = np.zeros(100)
    for t in range(T):
        pos, box = traj.xyz[t], traj.boxes[t]
        hs = np.zeros((10000, 3))
        for k in range(3):
            hs[:, k] = np.random.random(10000) * box[k]
        tot = port.neighbor_matrix(hs, pos[heavy], box, 0.0, 3.3).sum(axis=1).astype(int)
        want[np.arange(tot.max() + 1)] += np.bincount(tot)
    np.random.seed(21)
    mu, avgN, avgN2 = opl.chemPotCalc(top, traj)
    got = np.loadtxt("HS-solute_overlap_hist.txt")
    assert np.array_equal(got[:, 1], want) and np.array_equal(got[:, 0], np.arange(100)) and want.sum() == T * 10000
    with np.errstate(divide="ignore"):   # no empty 3.3 A cavity among 30 000 insertions in dense water: mu = inf, as in the reference
        assert np.isclose(mu, -np.log(want[0] / want.sum())) and np.isclose(avgN, np.dot(np.arange(100), want) / want.sum())
    assert np.isclose(avgN2, np.dot(np.arange(100) ** 2.0, want) / want.sum()) and 0.5 < avgN < 6.0
    # shell variant on one frame: the reference's rejection loop, call for call
    one = ArrayTrajectory(traj.xyz[:1], traj.boxes[:1], top=top)
    np.random.seed(5)
    pos, box = traj.xyz[0], traj.boxes[0]
    hs, count = np.zeros((100000, 3)), 0
    while count < 100000:
        rx, ry, rz = (2.0 * (np.random.random(1) - 0.5) * 4.2 for _ in range(3))
        if np.sqrt(rx[0] ** 2.0 + ry[0] ** 2.0 + rz[0] ** 2.0) > 4.2:
            continue
        hs[count] = pos[np.random.choice(sol)] + np.array([rx[0], ry[0], rz[0]])
        count += 1
    tot = port.neighbor_matrix(hs, pos[heavy], box, 0.0, 3.3).sum(axis=1).astype(int)
    want = np.zeros(100)
    want[np.arange(tot.max() + 1)] += np.bincount(tot)
    np.random.seed(5)
    mu_s, _, _ = opl.chemPotCalc(top, one, keyword=True)
    assert np.array_equal(np.loadtxt("HS-solute_overlap_hist_Shell.txt")[:, 1], want) and want[0] < want.sum()


def test_boundWrapPopulations_cache_and_use_as_subInds(in_tmp, monkeypatch):
    """The reference script's subInds workflow (orderParam_lib.py:2010-2036): populations from getBoundWrap per frame,
    cached in boundFile.npy, reused when the cache fits the trajectory, rebuilt when it does not; usable as subInds."""
    T = 6
    top, traj = make_system(4, T, n_sol=3)
    sub = opl.boundWrapPopulations(top, traj)
    assert len(sub) == T and all(len(row) == 4 for row in sub) and os.path.exists("boundFile.npy")
    obj = TrajObject(top, traj)
    watInds = obj.getWatInds()[0]
    for row in sub:
        bound, wrap, shell, rest = row
        assert np.array_equal(np.sort(np.concatenate([bound, wrap])), np.sort(shell))
        assert np.array_equal(np.sort(np.concatenate([shell, rest])), watInds) and len(shell) > 0
    calls = []
    real = opl.getBoundWrap
    monkeypatch.setattr(opl, "getBoundWrap", lambda *a, **k: calls.append(1) or real(*a, **k))
    again = opl.boundWrapPopulations(top, traj)                       # served from the cache
    assert not calls and all(np.array_equal(a, b) for ra, rb in zip(sub, again) for a, b in zip(ra, rb))
    shorter = ArrayTrajectory(traj.xyz[:4], traj.boxes[:4], top=top)  # cache does not fit: rebuilt
    assert len(opl.boundWrapPopulations(top, shorter)) == 4 and len(calls) == 4
    assert len(np.load("boundFile.npy", allow_pickle=True)) == 4
    np.random.seed(2)
    avgQ, _ = opl.tetOrderCalc(top, traj, subInds=sub, nPops=4)
    # (a frame without bound waters gives that population a NaN mean, as in the reference)
    assert avgQ[0].shape == (5,) and np.all(np.isfinite(avgQ[0][[0, 3, 4]]))


def test_threeBodyCalc_output2D_is_numpy_histogram2d_of_the_reference_lists(in_tmp):
    """output2D=True: the (N_c - 1, theta) histogram of reference orderParam_lib.py:1329-1335 / :1385-1393, rebuilt here
    with the reference's own list construction and np.histogram2d over the oracle's angles."""
    T = 6
    top, traj = make_system(3, T, sigma=0.45)
    obj = TrajObject(top, traj)
    watInds, _, _ = obj.getWatInds()
    numbers, angles = [], []
    for t in range(T):
        watPos = traj.xyz[t][watInds]
        tb = port.three_body(watPos, watPos, traj.boxes[t])
        angles.append(tb["angVals"])
        for n in tb["numAngs"]:
            count = int(n - 1)
            while count > 0:
                numbers.append([int(n - 1) for _ in range(count)])
                count -= 1
    numbers = np.concatenate(numbers).astype(float)
    angles = np.concatenate(angles)
    H, xe, ye = np.histogram2d(numbers, angles, bins=(np.arange(-1.5, 13.5, 1), np.linspace(0, 180, 500)))
    H = H / np.sum(H)
    plain = opl.threeBodyCalc(top, traj)
    with2d = opl.threeBodyCalc(top, traj, output2D=True)
    got, gx, gy = opl.threeBodyCalc.last_2d
    assert got.shape == (14, 499) and np.array_equal(gx, xe) and np.array_equal(gy, ye)
    assert np.array_equal(got, H) and abs(got.sum() - 1.0) < 1e-12
    for a, b in zip(plain[:4], with2d[:4]):  # the statistics do not change (the CIs are bootstrap draws)
        assert np.allclose(a[0], b[0], rtol=1e-12, atol=0.0)  # (sums of doubles accumulated by atomics: the order is not fixed)
    assert np.loadtxt("3bDistribution_2D.txt").shape == (14, 499)
