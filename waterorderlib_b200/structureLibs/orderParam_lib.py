"""Drop-in for the hot-path frame drivers of the reference's ``structureLibs/orderParam_lib.py``.

Same names, positional order, keyword names, defaults, return values and output files:

    tetOrderCalc(topFile, trajFile, subInds=None, nPops=0, solResName, watResName, stride=1)       reference :1426-1503
    threeBodyCalc(topFile, trajFile, subInds=None, nPops=0, solResName, watResName, nBins=500,
                  stride=1, output2D=False)                                                          reference :1269-1424
    hbCalc(topFile, trajFile, solResName, watResName, stride=1)                                     reference :729-917
    getBoundWrap(topFile, frame, watInds, watHInds, solInds, solHInds, solCInds, solOInds, solNInds,
                 solSInds, cutoff=4.0, hbDist=3.0, hbAng=150.0)                                      reference :419-572
    getNeighborStats(topFile, trajFile, Inds1, Inds2, nAtoms1, nAtoms2, stride=1, distCut=3.4,
                     switch=False)                                                                   reference :313-384
    getHBInds(top, frame, solInds, solHInds, solNInds, solOInds)                                     reference :46-120
    boundWrapPopulations(topFile, trajFile, ..., cutoff=4.6, nPops=4, cacheFile='boundFile.npy')     reference :2010-2036 (script)
    lsiCalc(topFile, trajFile, subInds=None, nPops=0, solResName, watResName, stride=1)            reference :1586-1663
    hexOrderCalc(topFile, trajFile, subInds=None, nPops=0, solResName, endResName, stride=1,
                 lowCut=0.0, highCut=7.0)                                                            reference :1505-1584
    rdfCalc(topFile, trajFile, solResName, watResName, binwidth=0.1, totbins=150, stride=1)         reference :575-727
    chemPotCalc(topFile, trajFile, solResName, watResName, probeRadius=3.3, keyword=False, stride=1)  reference :1666-1791
    blockAverage(vals, nBlocks=20), getCI(means)                                                     reference :386-417
    getClusters(hbMat), getHBClusterStats(...), getIonClusterStats(...)                              reference :123-311

``topFile`` / ``trajFile`` are whatever ``TrajObject`` accepts (in-memory objects, .npz, or AMBER files when
parmed/pytraj are installed).  Where the reference loops over frames calling f2py routines per water, these
drivers push BATCHES of frames through the fused cell-list kernels of libwol.so and keep histograms and
per-frame sums on the device; under ``torch.distributed`` (one process per GPU) frames are sharded across
ranks and combined with one all-reduce + one all-gather (waterorderlib_b200.distributed).
"""
import os

import numpy as np
import torch

from .. import distributed as wdist
from .. import engine, routines
from .._capi import STAT_NAMES
from . import waterlib as wl
from .TrajObject import TrajObject

_S = {n: i for i, n in enumerate(STAT_NAMES)}
_MAX_ATOMS_PER_BATCH = 8_000_000


# ---- statistics helpers (host side, O(frames)) ------------------------------------------------------

def getCI(means):
    """reference orderParam_lib.py:386-391"""
    meanCI = means[int(0.5 * len(means))]
    upperCI = means[int(0.975 * len(means))] - meanCI
    lowerCI = meanCI - means[int(0.025 * len(means))]
    return max(upperCI, lowerCI)


def blockAverage(vals, nBlocks=20):
    """Bootstrap confidence interval over block means (reference orderParam_lib.py:394-417): 20 blocks,
    10 000 resamples with np.random.choice (unseeded there and here; seed numpy's global RNG to reproduce)."""
    vals = np.asarray(vals)
    obsBlocks = np.zeros(nBlocks)
    lenBlock = len(vals) / nBlocks
    for i in range(nBlocks):
        obsBlocks[i] = np.mean(vals[int(i * lenBlock):int((i + 1) * lenBlock)])
    nSamp = nBlocks
    nResamp = 10000
    obsMeans = np.mean(np.random.choice(obsBlocks, (nResamp, nSamp)), axis=1)
    obsMeans = np.sort(obsMeans)
    return getCI(obsMeans)


def _mean_ci(series):
    with np.errstate(all="ignore"):
        return np.mean(series), blockAverage(series)


def getHBInds(top, frame, solInds, solHInds, solNInds, solOInds):
    """Acceptor / donor / donor-hydrogen index lists from the bonded topology (reference :46-120):
    every O (N) in solOInds (solNInds) is an acceptor and appears once per bonded hydrogen as a donor."""
    if hasattr(top, "bonds") and hasattr(top, "names") and not hasattr(top, "residues"):
        return _hb_inds_from_arrays(top, solOInds, solNInds)   # same lists without a Python loop over every atom
    solO, solN = set(int(i) for i in solOInds), set(int(i) for i in solNInds)
    acc = {"O": [], "N": []}
    don = {"O": [], "N": []}
    donH = {"O": [], "N": []}
    for i, atom in enumerate(top.atoms):
        kind = "O" if i in solO else ("N" if i in solN else None)
        if kind is None:
            continue
        count = 0
        for jatom in atom.bond_partners:
            if 'H' in jatom.name:
                donH[kind].append(jatom.idx)
                count += 1
        acc[kind].append(i)
        don[kind].extend([i] * count)
    hbOInds = [np.array(acc["O"], dtype=int), np.array(don["O"], dtype=int), np.array(donH["O"], dtype=int)]
    hbNInds = [np.array(acc["N"], dtype=int), np.array(don["N"], dtype=int), np.array(donH["N"], dtype=int)]
    return hbOInds, hbNInds


def _hb_inds_from_arrays(top, solOInds, solNInds):
    """getHBInds for the built-in Topology (arrays of names and bonds): acceptors ascending; a donor entry per bonded atom
    whose name contains 'H', in the order the reference meets them -- atoms ascending, each atom's partners in bond-list
    order (Atom.bond_partners is filled bond by bond)."""
    is_h = getattr(top, "_name_has_h", None)
    if is_h is None:
        is_h = np.char.find(top.names, "H") >= 0
        top._name_has_h = is_h
    bonds = np.asarray(top.bonds, dtype=np.int64).reshape(-1, 2)
    k = np.arange(bonds.shape[0])
    heavy = np.concatenate([bonds[:, 0], bonds[:, 1]])
    other = np.concatenate([bonds[:, 1], bonds[:, 0]])
    when = np.concatenate([k, k])
    out = []
    for inds in (solOInds, solNInds):
        member = np.zeros(top.n_atoms, dtype=bool)
        inds = np.asarray(inds, dtype=np.int64)
        member[inds] = True
        sel = member[heavy] & is_h[other]
        h, o, w = heavy[sel], other[sel], when[sel]
        order = np.lexsort((w, h))
        out.append([np.nonzero(member)[0].astype(int), h[order].astype(int), o[order].astype(int)])
    return out[0], out[1]


# ---- frame batching ----------------------------------------------------------------------------------

def _frame_arrays(traj, begin, end):
    """(xyz (f, natom, 3), box (f, 3)) of frames [begin, end) of an ArrayTrajectory or any pytraj-like iterable."""
    if hasattr(traj, "boxes") and hasattr(traj, "xyz") and getattr(traj.xyz, "ndim", 0) == 3:
        if hasattr(traj.xyz, "raw"):   # file-backed frames in the file's own byte order (amber_io.NetCDFTrajectory)
            return traj.xyz.raw(begin, end), traj.boxes[begin:end, :3]
        return traj.xyz[begin:end], traj.boxes[begin:end, :3]
    xyz, box = [], []
    for t in range(begin, end):
        frame = traj[t]
        xyz.append(np.array(frame.xyz))
        box.append(np.array(frame.box.values[:3]))
    return np.stack(xyz), np.stack(box)


_COPY_POOL = None
_COPY_THREADS = max(1, int(os.environ.get("WOL_STAGE_THREADS", "4")))


def _host_copy(dst, src):
    """dst[:] = src for 1-D uint8 numpy arrays, large copies split over a few threads (numpy releases the GIL for a
    contiguous copy; one core moves ~10 GB/s, which would be the slowest stage of the frame drivers)."""
    global _COPY_POOL
    n = src.size
    if n < (8 << 20):
        np.copyto(dst, src)
        return
    if _COPY_POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _COPY_POOL = ThreadPoolExecutor(max_workers=_COPY_THREADS, thread_name_prefix="wol-stage")
    step = ((n + _COPY_THREADS - 1) // _COPY_THREADS + 4095) // 4096 * 4096
    jobs = [_COPY_POOL.submit(np.copyto, dst[i:i + step], src[i:i + step]) for i in range(0, n, step)]
    for j in jobs:
        j.result()


class _FrameStager:
    """Moves batches of whole frames to the device.  numpy frames live in pageable memory: each batch is copied into one
    of two page-locked buffers (a plain, multi-threaded host memcpy) and sent from there on a side stream, so the host copy
    of batch k+1 runs while the kernels of batch k do, and the transfer itself runs at full PCIe rate.  torch tensors are
    passed through: CUDA tensors as they are, (pinned) CPU tensors with one asynchronous copy."""

    def __init__(self, dev):
        self.dev = dev
        self.stream = torch.cuda.Stream(device=dev)
        self.bufs = [None, None]
        self.sent = [None, None]
        self.k = 0

    def put(self, xyz):
        """-> (device tensor (f, natom, 3), event recorded when it is complete)."""
        done = torch.cuda.Event()
        if isinstance(xyz, torch.Tensor):
            if xyz.is_cuda:
                done.record(torch.cuda.current_stream(self.dev))
                return xyz, done
            src = xyz
        else:
            a = np.ascontiguousarray(np.asarray(xyz))
            swap = not a.dtype.isnative         # e.g. the big-endian floats of a NetCDF-3 file: bytes go up as they are
            native = a.dtype.newbyteorder("=")  # and are swapped on the device, which saves a pass over them on the host
            raw = a.reshape(-1).view(np.uint8)
            i = self.k & 1
            self.k += 1
            if self.sent[i] is not None:
                self.sent[i].synchronize()          # the previous transfer out of this buffer
            if self.bufs[i] is None or self.bufs[i].numel() < raw.size:
                self.bufs[i] = torch.empty(raw.size, dtype=torch.uint8, pin_memory=True)
            src = self.bufs[i][:raw.size]
            _host_copy(src.numpy(), raw)
            self.sent[i] = done
            with torch.cuda.stream(self.stream):
                up = src.to(self.dev, non_blocking=True)
                if swap:
                    up = up.view(-1, native.itemsize).flip(1).contiguous().view(-1)
                out = up.view(torch.from_numpy(np.zeros(0, dtype=native)).dtype).view(a.shape)
                done.record(self.stream)
            return out, done
        with torch.cuda.stream(self.stream):
            out = src.to(self.dev, non_blocking=True)
            done.record(self.stream)
        return out, done


def _batches(begin, end, n_atoms):
    per = max(1, _MAX_ATOMS_PER_BATCH // max(n_atoms, 1))
    for b in range(begin, end, per):
        yield b, min(end, b + per)


def _run_populations(obj, subInds, nPops, do_q, do_3body, nBins):
    """Shared core of tetOrderCalc / threeBodyCalc: every frame of this rank's shard through the fused kernel, a whole
    batch of frames per call: population 0 (all waters) builds the cell list, each sub-population (ragged: its members
    change every frame, structureLibs/orderParam_lib.py:1343-1346, :1475-1478) reuses it with padded centres.
    Returns numpy arrays holding ALL frames on every rank: stats (T, P, NSTATS) f64, per-frame angle
    histograms (T, P, nBins) i64 or None, pooled q histograms (P, 500) i64 or None, members (T, P)."""
    traj = obj.traj
    watInds, _watHInds, _lenWat = obj.getWatInds()
    T, P, NS = len(traj), nPops + 1, len(STAT_NAMES)
    dev = torch.device("cuda", torch.cuda.current_device())
    begin, end = wdist.shard_frames(T)
    Tl = end - begin
    stats = torch.zeros((P, Tl, NS), dtype=torch.float64, device=dev)
    ang_hist = torch.zeros((P, Tl, nBins), dtype=torch.int64, device=dev) if do_3body else None
    q_hist = torch.zeros((P, 1, 500), dtype=torch.int64, device=dev) if do_q else None
    members = torch.zeros((Tl, P), dtype=torch.float64, device=dev)
    ws = engine.Workspace(dev)
    # angle histograms are kept per frame (the drivers report per-frame entropies), q histograms are pooled;
    # the two never run together here
    kw = dict(do_q=do_q, do_3body=do_3body, nbins=nBins, device=dev, hist_per_frame=do_3body)

    def outputs(j, l0, l1):
        o = {"frame_stats": stats[j, l0:l1]}
        if do_3body:
            o["ang_hist"] = ang_hist[j, l0:l1]
        if do_q:
            o["q_hist"] = q_hist[j]
        return o

    wat_d = torch.from_numpy(np.ascontiguousarray(np.asarray(watInds, dtype=np.int64))).to(dev)
    stager = _FrameStager(dev)
    batches = list(_batches(begin, end, len(watInds)))
    if not (isinstance(getattr(traj, "xyz", None), torch.Tensor) and traj.xyz.is_cuda) and Tl > 0:
        # host frames: batches of at most ~64 MB, so that little of the staging is exposed before the first kernels start
        # and the page-locked buffers stay small
        first = _frame_arrays(traj, begin, begin + 1)[0]
        frame_bytes = int(np.prod(first.shape[1:])) * (first.element_size() if isinstance(first, torch.Tensor) else first.dtype.itemsize)
        per = max(1, min(batches[0][1] - batches[0][0], (64 << 20) // max(frame_bytes, 1)))
        batches = [(b, min(end, b + per)) for b in range(begin, end, per)]
    # whole frames go to the device as they are and the water oxygens are gathered THERE: a host-side fancy-index
    # gather of 10^6 waters costs 40-90 ms per frame, a hundred times the kernels.  One batch is staged ahead.
    ahead = None
    last_shape = None
    box_pin = torch.empty((max(Tl, 1), 3), dtype=torch.float64, pin_memory=True)
    for n, (b0, b1) in enumerate(batches):
        if ahead is None:
            xyz, box = _frame_arrays(traj, b0, b1)
            ahead = (stager.put(xyz), box)
        (xyz_d, ready), box = ahead
        torch.cuda.current_stream(dev).wait_event(ready)
        xyz_d.record_stream(torch.cuda.current_stream(dev))
        watPos = xyz_d.index_select(1, wat_d)
        l0, l1 = b0 - begin, b1 - begin
        o = outputs(0, l0, l1)
        # the boxes go up from page-locked memory: a pageable copy would block the host until the previous batch's kernels
        # are done, and the staging of the next batch could no longer overlap them
        box_h = engine.as_host_boxes(box, b1 - b0)
        bkw = {}
        if (box_h > 0.0).all():
            box_pin[l0:l1] = torch.from_numpy(box_h)  # (every batch has its own rows: nothing in flight is overwritten)
            bkw["box_device"] = box_pin[l0:l1].to(dev, non_blocking=True)
        engine.q3b_frames(watPos, box_h, None, out=o, want=tuple(o.keys()), workspace=ws, check_status=False, **bkw, **kw)
        last_shape = (b1 - b0, box_h)
        members[l0:l1, 0] = float(len(watInds))
        # sub-populations: their centres change from frame to frame, so each population is padded to the batch's
        # largest member count and evaluated in ONE call against the cell list population 0 just built
        for j in range(1, P):
            inds = [np.asarray(subInds[t][j - 1], dtype=np.int64) for t in range(b0, b1)]
            counts = np.array([len(i) for i in inds], dtype=np.int32)
            members[l0:l1, j] = torch.from_numpy(counts.astype(np.float64)).to(dev)
            m_max = int(counts.max()) if len(counts) else 0
            if m_max == 0:
                continue
            pad = np.zeros((b1 - b0, m_max), dtype=np.int64)   # padded member indices; rows beyond n_valid are ignored
            for k, i in enumerate(inds):
                pad[k, :len(i)] = i
            cen = torch.gather(xyz_d, 1, torch.from_numpy(pad).to(dev)[:, :, None].expand(-1, -1, 3))
            o = outputs(j, l0, l1)
            engine.q3b_frames(watPos, box_h, cen, out=o, want=tuple(o.keys()), workspace=ws, n_valid=counts, reuse_cells=True,
                              check_status=False, **bkw, **kw)
        ahead = None
        if n + 1 < len(batches):   # the kernels of this batch are queued: stage the next one while they run
            nxt, nbox = _frame_arrays(traj, *batches[n + 1])
            ahead = (stager.put(nxt), nbox)
    # no batch synchronised the host (that is what lets the staging of batch k+1 overlap the kernels of batch k): the
    # sticky list-capacity flag of the workspace is read once, here
    if last_shape is not None:
        engine.workspace_status(ws, last_shape[0], len(watInds), len(watInds), engine.default_r_cell(do_q, do_3body, 3.413, 10.0),
                                last_shape[1])
    # ---- combine ranks: one all-reduce of the integer histograms, one all-gather of the per-frame rows ----
    rows = [stats.permute(1, 0, 2).reshape(Tl, P * NS), members]
    if do_3body:
        rows.append(ang_hist.permute(1, 0, 2).reshape(Tl, P * nBins).to(torch.float64))  # counts < 2^53: exact
    rows = wdist.gather_frame_rows(torch.cat(rows, dim=1).contiguous(), T)
    if do_q:
        wdist.reduce_histograms(q_hist)
    rows = rows.cpu().numpy()
    stats_np = rows[:, :P * NS].reshape(T, P, NS)
    members_np = rows[:, P * NS:P * NS + P]
    hist_np = np.rint(rows[:, P * NS + P:]).astype(np.int64).reshape(T, P, nBins) if do_3body else None
    return stats_np, hist_np, (q_hist[:, 0].cpu().numpy() if do_q else None), members_np


# ---- drivers -----------------------------------------------------------------------------------------

def tetOrderCalc(topFile, trajFile, subInds=None, nPops=0, solResName='(!:WAT)', watResName='(:WAT)', stride=1):
    """Tetrahedral order parameter statistics and distribution for all waters and nPops sub-populations
    (reference orderParam_lib.py:1426-1503).  Returns (avgQ, varQ), each [means, CIs] over populations, and
    writes qDistribution_<j>.txt (two columns, "%.3e")."""
    obj = TrajObject(topFile, trajFile, stride, solResName, watResName)
    if subInds is None:
        nPops = 0
    stats, _, q_hist, _ = _run_populations(obj, subInds, nPops, True, False, 500)
    P = nPops + 1
    with np.errstate(all="ignore"):
        n = stats[:, :, _S["n_centres"]]
        mean_t = stats[:, :, _S["q_sum"]] / n
        var_t = np.maximum(stats[:, :, _S["q_sumsq"]] / n - mean_t * mean_t, 0.0)
    avgQ_mean, avgQ_CI, varQ_mean, varQ_CI = (np.zeros(P) for _ in range(4))
    for j in range(P):
        avgQ_mean[j], avgQ_CI[j] = _mean_ci(mean_t[:, j])
        varQ_mean[j], varQ_CI[j] = _mean_ci(var_t[:, j])
    if wdist.world()[0] == 0:
        bins = np.linspace(0.0, 1.0, 501)
        for j in range(P):
            np.savetxt('qDistribution_' + str(j) + '.txt', np.stack([0.5 * (bins[:-1] + bins[1:]), q_hist[j]], axis=1),
                       header='qVal    frequency', fmt="%.3e")
    return [avgQ_mean, avgQ_CI], [varQ_mean, varQ_CI]


def _entropy(counts):
    tot = float(np.sum(counts))
    if tot == 0.0:
        return 0.0
    dens = counts / tot
    dens = dens[dens != 0]
    return float(-np.sum(dens * np.log(dens)))


def threeBodyCalc(topFile, trajFile, subInds=None, nPops=0, solResName='(!:WAT)', watResName='(:WAT)', nBins=500, stride=1,
                  output2D=False):
    """Three-body angle distribution statistics for all waters and nPops sub-populations (reference
    orderParam_lib.py:1269-1424).  Returns (pTet, avgCos, varCos, entropy, nWats), each [means, CIs], and writes
    3bDistribution_<j>.txt.  output2D=True also accumulates the (N_c, theta) histogram the reference builds with
    np.histogram2d (:1329-1335, :1385-1393: water-water coordination number minus one against the angle, edges
    arange(-1.5, 13.5, 1) x linspace(0, 180, 500), normalised to 1): it is written to 3bDistribution_2D.txt and kept
    in threeBodyCalc.last_2d = (H, xedges, yedges); the matplotlib figure the reference draws from it is not
    produced."""
    obj = TrajObject(topFile, trajFile, stride, solResName, watResName)
    if subInds is None:
        nPops = 0
    stats, hist, _, members = _run_populations(obj, subInds, nPops, False, True, nBins)
    T, P = stats.shape[0], nPops + 1
    n_ang = stats[:, :, _S["n_angles"]]
    cnt = stats[:, :, _S["tet_count"]]
    pTet_t, avgCos_t, varCos_t, ent_t = (np.zeros((T, P)) for _ in range(4))
    with np.errstate(all="ignore"):
        has = n_ang > 0
        pTet_t[has] = cnt[has] / n_ang[has]
        mean = stats[:, :, _S["tet_cos"]] / cnt
        var = np.maximum(stats[:, :, _S["tet_cossq"]] / cnt - mean * mean, 0.0)
        avgCos_t[has] = mean[has]
        varCos_t[has] = var[has]
    for t in range(T):
        for j in range(P):
            ent_t[t, j] = _entropy(hist[t, j]) if has[t, j] else 0.0
    out = []
    for series in (pTet_t, avgCos_t, varCos_t, ent_t, members):
        m, ci = np.zeros(P), np.zeros(P)
        for j in range(P):
            m[j], ci[j] = _mean_ci(series[:, j])
        out.append([m, ci])
    if wdist.world()[0] == 0:
        bins = np.linspace(0.0, 180.0, nBins + 1)
        for j in range(P):
            if n_ang[:, j].sum() != 0:
                np.savetxt('3bDistribution_' + str(j) + '.txt',
                           np.stack([0.5 * (bins[:-1] + bins[1:]), hist[:, j].sum(axis=0)], axis=1),
                           header='3-body angle (deg)    frequency', fmt="%.3e")
    if output2D:
        threeBodyCalc.last_2d = _theta_nc_histogram(obj)
    pTet, avgCos, varCos, entropy, nWats = out
    return pTet, avgCos, varCos, entropy, nWats


def _theta_nc_histogram(obj):
    """The 2-D histogram of threeBodyCalc(output2D=True): for every three-body angle of every water, (number of
    neighbours of its central water - 1, angle) -- reference orderParam_lib.py:1329-1335 builds the first coordinate by
    repeating n - 1 once per angle, :1385-1393 bins the pairs.  Angles are materialised in the reference's order by the
    getCosAngs path and binned on the device with numpy's histogramdd rule; frames are sharded over ranks."""
    traj = obj.traj
    watInds, _, _ = obj.getWatInds()
    dev = torch.device("cuda", torch.cuda.current_device())
    xedges, yedges = np.arange(-1.5, 13.5, 1), np.linspace(0, 180, 500)
    H = torch.zeros((len(xedges) - 1, len(yedges) - 1), dtype=torch.int64, device=dev)
    wat_d = torch.from_numpy(np.ascontiguousarray(np.asarray(watInds, dtype=np.int64))).to(dev)
    begin, end = wdist.shard_frames(len(traj))
    stager = _FrameStager(dev)
    for t in range(begin, end):
        xyz, box = _frame_arrays(traj, t, t + 1)
        xyz_d, ready = stager.put(xyz)
        torch.cuda.current_stream(dev).wait_event(ready)
        xyz_d.record_stream(torch.cuda.current_stream(dev))
        watPos = xyz_d.index_select(1, wat_d).to(torch.float64)
        angles, n3, _ = routines.three_body_angles(None, watPos, box)
        k = n3.reshape(-1).to(torch.int64)
        numbers = torch.repeat_interleave((k - 1).to(torch.float64), k * (k - 1) // 2)
        routines.histogram2d(numbers, angles, xedges, yedges, out=H)
    wdist.reduce_histograms(H)
    Hn = H.cpu().numpy().astype(np.float64)
    tot = Hn.sum()
    if tot > 0:
        Hn = Hn / tot
    if wdist.world()[0] == 0:
        np.savetxt('3bDistribution_2D.txt', Hn, fmt="%.3e",
                   header='rows: N_c - 1 bins, edges arange(-1.5, 13.5, 1); columns: angle bins, edges linspace(0, 180, 500); sum = 1')
    return Hn, xedges, yedges


def _hb_sums(acc, don, donh, box, dist, ang, cells=None):
    """(row sums, column sums) of generalhbonds(acc, don, donh) for a batch of frames, as int64 numpy (F, n).
    cells: a routines.CellList already built over `don` (the same donors serve several acceptor sets)."""
    F = acc.shape[0]
    if acc.shape[1] == 0 or don.shape[1] == 0:
        return np.zeros((F, acc.shape[1]), dtype=np.int64), np.zeros((F, don.shape[1]), dtype=np.int64)
    r = routines.hbond_counts(acc, don, donh, box, dist, ang, cells=cells)
    return r["acc_count"].cpu().numpy().astype(np.int64), r["don_count"].cpu().numpy().astype(np.int64)


def _donor_cells(don, box, dist):
    return routines.CellList(don, box, max(float(dist), 1e-3)) if don.shape[1] else None


def hbCalc(topFile, trajFile, solResName='(!:WAT)', watResName='(:WAT)', stride=1):
    """Average hydrogen bonds per water and per cosolvent molecule, 3.5 A / 120 deg (reference
    orderParam_lib.py:729-917).  Returns (avgWatHBs, avgSolHBs) and writes hbDistribution_water.txt /
    hbDistribution_cosolv.txt."""
    obj = TrajObject(topFile, trajFile, stride, solResName, watResName)
    top, traj = obj.top, obj.traj
    watInds, watHInds, _lenWat = obj.getWatInds()
    solInds, solHInds, _solC, solNInds, solOInds, _solS = obj.getSolInds()
    hbO, hbN = getHBInds(top, traj[0], solInds, solHInds, solNInds, solOInds)
    sAccO, sDonO, sDonHO = hbO
    sAccN, sDonN, sDonHN = hbN
    hbW, _ = getHBInds(top, traj[0], watInds, watHInds, [], watInds)
    wAcc, wDon, wDonH = hbW
    nSol = top.n_residues(solResName) if hasattr(top, "n_residues") else traj[:1, solResName].topology.n_residues
    T = len(traj)
    begin, end = wdist.shard_frames(T)
    dev = torch.device("cuda", torch.cuda.current_device())
    wat_rows, sol_rows = [], []
    stager = _FrameStager(dev)
    on_dev = {}

    def dev_index(idx):   # each index list goes to the device once; the gathers run there (see _run_populations)
        key = id(idx)
        if key not in on_dev:
            on_dev[key] = (idx, torch.from_numpy(np.ascontiguousarray(np.asarray(idx, dtype=np.int64))).to(dev))
        return on_dev[key][1]

    for b0, b1 in _batches(begin, end, 3 * len(watInds) + len(solInds)):
        xyz, box = _frame_arrays(traj, b0, b1)
        xyz_d, ready = stager.put(xyz)
        torch.cuda.current_stream(dev).wait_event(ready)
        xyz_d.record_stream(torch.cuda.current_stream(dev))
        g = lambda idx: xyz_d.index_select(1, dev_index(idx))  # noqa: E731
        D, A = 3.5, 120.0
        # the nine generalhbonds calls of the reference (:805-834) share three donor sets: one cell list each
        pw, pO, pN = g(wDon), g(sDonO), g(sDonN)
        cw, cO, cN = _donor_cells(pw, box, D), _donor_cells(pO, box, D), _donor_cells(pN, box, D)
        ww_a, ww_d = _hb_sums(g(wAcc), pw, g(wDonH), box, D, A, cw)
        wsO_a, wsO_d = _hb_sums(g(wAcc), pO, g(sDonHO), box, D, A, cO)
        swO_a, swO_d = _hb_sums(g(sAccO), pw, g(wDonH), box, D, A, cw)
        wsN_a, wsN_d = _hb_sums(g(wAcc), pN, g(sDonHN), box, D, A, cN)
        swN_a, swN_d = _hb_sums(g(sAccN), pw, g(wDonH), box, D, A, cw)
        OO_a, OO_d = _hb_sums(g(sAccO), pO, g(sDonHO), box, D, A, cO)
        ON_a, ON_d = _hb_sums(g(sAccO), pN, g(sDonHN), box, D, A, cN)
        NO_a, NO_d = _hb_sums(g(sAccN), pO, g(sDonHO), box, D, A, cO)
        NN_a, NN_d = _hb_sums(g(sAccN), pN, g(sDonHN), box, D, A, cN)
        # per water molecule (reference :867-884); donors are listed once per hydrogen, two per water
        fold = lambda d: d[:, ::2] + d[:, 1::2]  # noqa: E731
        wat_rows.append(ww_a + fold(ww_d) + wsO_a + fold(swO_d) + wsN_a + fold(swN_d))
        if nSol > 0:
            def per_mol(v, n_per):  # sum( [ v[i::n] for i in range(n) ] )  (reference :850-851)
                return sum(v[:, i::n_per] for i in range(n_per)) if n_per > 0 else np.zeros((v.shape[0], nSol), dtype=np.int64)
            nAccO, nAccN = int(len(sAccO) / nSol), int(len(sAccN) / nSol)
            nDonO, nDonN = int(len(sDonO) / nSol), int(len(sDonN) / nSol)
            solOAcc = per_mol(swO_a + OO_a + ON_a, nAccO)
            solODon = per_mol(wsO_d + OO_d + NO_d, nDonO)
            solNAcc = per_mol(swN_a + NN_a + NO_a, nAccN)
            solNDon = per_mol(wsN_d + NN_d + ON_d, nDonN)
            sol_rows.append(solNAcc + solNDon + solOAcc + solODon)
    numWat = np.concatenate(wat_rows) if wat_rows else np.zeros((0, len(wAcc)), dtype=np.int64)
    numSol = np.concatenate(sol_rows) if sol_rows else np.zeros((end - begin, 0), dtype=np.int64)
    # combine ranks: per-frame rows gathered in frame order
    numWat = wdist.gather_frame_rows(torch.from_numpy(numWat).to(dev), T).cpu().numpy()
    if nSol > 0:
        numSol = wdist.gather_frame_rows(torch.from_numpy(numSol).to(dev), T).cpu().numpy()
    numWatHBs, numSolHBs = numWat.reshape(-1), numSol.reshape(-1)
    with np.errstate(all="ignore"):
        avgWatHBs = np.mean(numWatHBs)
        avgSolHBs = np.mean(numSolHBs) if numSolHBs.size else float("nan")
    if wdist.world()[0] == 0:
        edges = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10]
        for vals, name in ((numWatHBs, 'hbDistribution_water.txt'), (numSolHBs, 'hbDistribution_cosolv.txt')):
            hbDist, bins = np.histogram(vals, bins=edges, density=False)
            np.savetxt(name, np.stack([0.5 * (bins[:-1] + bins[1:]), hbDist], axis=1), header='# hbs    frequency', fmt="%.3e")
    return avgWatHBs, avgSolHBs


def getBoundWrap(topFile, frame, watInds, watHInds, solInds, solHInds, solCInds, solOInds, solNInds, solSInds,
                 cutoff=4.0, hbDist=3.0, hbAng=150.0):
    """Hydration-shell waters of a solute split into "bound" (hydrogen-bonded to it) and "wrap" (the rest)
    (reference orderParam_lib.py:419-572).  Returns (boundInds, wrapInds, shellInds, nonShellInds)."""
    obj = TrajObject(topFile, trajFile=None, stride=1, solResName=None, watResName=None)
    top = obj.top
    hbOInds, _hbNInds = getHBInds(top, frame, solInds, solHInds, solNInds, solOInds)
    sAccO, sDonO, sDonHO = hbOInds
    pos = np.array(frame.xyz)
    thisbox = np.array(frame.box.values[:3])
    watInds = np.asarray(watInds)
    watPos, solPos = pos[watInds], pos[np.asarray(solInds, dtype=int)]
    # waters within `cutoff` of any solute heavy atom (reference :495-498)
    mask = routines.shell_mask(solPos, watPos, thisbox, cutoff).cpu().numpy()[0].astype(bool)
    shellInds = watInds[mask]
    nonShellInds = watInds[~mask]
    hbW, _ = getHBInds(top, frame, shellInds, watHInds, solNInds, shellInds)
    wAcc, wDon, wDonH = hbW

    def counts(acc, don, donh):
        if len(acc) == 0 or len(don) == 0:
            return np.zeros(len(acc), dtype=np.int32), np.zeros(len(don), dtype=np.int32)
        r = routines.hbond_counts(pos[acc], pos[don], pos[donh], thisbox, hbDist, hbAng)
        return r["acc_count"].cpu().numpy()[0], r["don_count"].cpu().numpy()[0]

    # water accepts from the solute (rows = shell waters), solute accepts from water (columns = water donors)
    watSol_a, _ = counts(wAcc, sDonO, sDonHO)
    boundMask_wat = np.nonzero(watSol_a > 0)[0]
    _, solWat_d = counts(sAccO, wDon, wDonH)
    dummy = (solWat_d > 0).astype(np.float64)
    boundMask_sol = np.where(np.ceil(0.5 * (dummy[0::2] + dummy[1::2])))[0]
    boundMask = np.sort(np.unique(np.concatenate([boundMask_wat, boundMask_sol]))).astype(int)
    keep = np.ones(len(shellInds), dtype=bool)
    keep[boundMask] = False
    return shellInds[boundMask], shellInds[keep], shellInds, nonShellInds


def getNeighborStats(topFile, trajFile, Inds1, Inds2, nAtoms1, nAtoms2, stride=1, distCut=3.4, switch=False):
    """Mean number of distinct neighbouring atoms per molecule of type 1 (reference orderParam_lib.py:313-384);
    writes coordDistribution.txt."""
    obj = TrajObject(topFile, trajFile=trajFile, stride=stride, solResName=None, watResName=None)
    numberCoord = []
    Inds1, Inds2 = np.asarray(Inds1, dtype=int), np.asarray(Inds2, dtype=int)
    for frame in obj.traj:
        thisbox = np.reshape(np.array(frame.box.values[:3]), (1, 3))
        thispos = np.array(frame.xyz)
        subPos1, subPos2 = thispos[Inds1], thispos[Inds2]
        nRes = int(len(Inds1) / nAtoms1)
        resNumbers = np.zeros(nRes, dtype=int)
        if switch:
            neighbors = wl.allnearneighbors(subPos1, thisbox, 0.0, distCut)
        else:
            neighbors = wl.nearneighbors(subPos1, subPos2, thisbox, 0.0, distCut)
        for n in range(nRes):
            nNeighbors = neighbors[int(n * nAtoms1):int((n + 1) * nAtoms1), :]
            if switch:
                # the reference's own-molecule exclusion indexes a slice of the slice (:349-350); for molecules
                # after the first it is empty, so only molecule 0 has its own atoms removed.  Reproduced.
                nNeighbors[int(n * nAtoms1):int((n + 1) * nAtoms1), int(n * nAtoms1):int((n + 1) * nAtoms1)] = 0
            resNumbers[n] = len(np.unique(np.where(nNeighbors == 1)[1]))
        numberCoord.append(resNumbers)
    numberCoord = np.concatenate(numberCoord)
    meanCoord = np.mean(numberCoord)
    coordDist, bins = np.histogram(numberCoord, bins=[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10], density=False)
    np.savetxt('coordDistribution.txt', np.stack([0.5 * (bins[:-1] + bins[1:]), coordDist], axis=1),
               header='# coords    frequency', fmt="%.3e")
    return meanCoord


def _value_driver(obj, centreInds, subInds, nPops, per_frame, hist_range, fname, header):
    """Shared frame loop of lsiCalc / hexOrderCalc (reference orderParam_lib.py:1617-1661, :1537-1582): a per-centre
    observable for all centres and for nPops sub-populations, per-frame mean / variance, bootstrap CIs, pooled 500-bin
    histograms written as two-column text.  per_frame(j, sub, pos, box) -> 1-D array of values (j = population); it is
    handed CUDA tensors."""
    traj = obj.traj
    T, P = len(traj), nPops + 1
    dev = torch.device("cuda", torch.cuda.current_device())
    begin, end = wdist.shard_frames(T)
    rows = np.zeros((end - begin, 2 * P))
    hists = torch.zeros((P, 500), dtype=torch.int64, device=dev)
    cen_d = torch.from_numpy(np.ascontiguousarray(np.asarray(centreInds, dtype=np.int64))).to(dev)
    for t in range(begin, end):
        frame = traj[t]
        thisbox = np.array(frame.box.values[:3])
        # the frame goes to the device once; centres and sub-populations are gathered there (a host-side gather of 10^6
        # atoms costs more than the kernels, see _run_populations)
        xyz = frame.xyz if isinstance(frame.xyz, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(np.asarray(frame.xyz)))
        pos_d = xyz.to(dev)
        cenPos = pos_d.index_select(0, cen_d)
        for j in range(P):
            sub = cenPos if j == 0 else pos_d.index_select(
                0, torch.from_numpy(np.ascontiguousarray(np.asarray(subInds[t][j - 1], dtype=np.int64))).to(dev))
            vals_d = per_frame(j, sub, cenPos, thisbox)
            vals = vals_d.cpu().numpy() if isinstance(vals_d, torch.Tensor) else np.asarray(vals_d)
            with np.errstate(all="ignore"):
                rows[t - begin, 2 * j], rows[t - begin, 2 * j + 1] = np.mean(vals), np.var(vals)
            if vals.size:
                h, _ = routines.histogram(vals_d, 500, hist_range)
                hists[j] += h
    rows = wdist.gather_frame_rows(torch.from_numpy(rows).to(dev), T).cpu().numpy()
    wdist.reduce_histograms(hists)
    avg_mean, avg_ci, var_mean, var_ci = (np.zeros(P) for _ in range(4))
    for j in range(P):
        avg_mean[j], avg_ci[j] = _mean_ci(rows[:, 2 * j])
        var_mean[j], var_ci[j] = _mean_ci(rows[:, 2 * j + 1])
    if wdist.world()[0] == 0:
        bins = np.linspace(hist_range[0], hist_range[1], 501)
        counts = hists.cpu().numpy()
        for j in range(P):
            np.savetxt(fname % j, np.stack([0.5 * (bins[:-1] + bins[1:]), counts[j]], axis=1), header=header, fmt="%.3e")
    return [avg_mean, avg_ci], [var_mean, var_ci]


def lsiCalc(topFile, trajFile, subInds=None, nPops=0, solResName='(!:WAT)', watResName='(:WAT)', stride=1):
    """Local structure index statistics and distribution for all waters and nPops sub-populations (reference
    orderParam_lib.py:1586-1663).  Returns (avgLSI, varLSI), each [means, CIs]; writes lsiDistribution_<j>.txt
    (500 bins on [0, 0.3] A^2)."""
    from . import water_properties as wp
    obj = TrajObject(topFile, trajFile, stride, solResName, watResName)
    watInds, _watHInds, _lenWat = obj.getWatInds()
    if subInds is None:
        nPops = 0
    return _value_driver(obj, watInds, subInds, nPops, lambda j, sub, pos, box: wp.getLSI(sub, pos, box, lowCut=0.0, highCut=3.7)[0],
                         (0.0, 0.3), 'lsiDistribution_%d.txt', 'lsiVal [A^2]    frequency')


def hexOrderCalc(topFile, trajFile, subInds=None, nPops=0, solResName='(!:WAT)', endResName='(:WAT)', stride=1, lowCut=0.0,
                 highCut=7.0):
    """Hexagonal order parameter statistics and distribution for chain-end atoms (reference orderParam_lib.py:1505-1584).
    As in the reference: every second selected end atom is used (:1526), the all-ends population is evaluated with
    the hard-coded cutoffs 0 / 7.0 (:1549) and the sub-populations with getOrderParamPsi's defaults 0 / 10 (:1556);
    the lowCut / highCut arguments are accepted and, like there, unused.  Writes psiDistribution_<j>.txt."""
    from . import water_properties as wp
    obj = TrajObject(topFile, trajFile, stride, solResName, endResName)
    endInds, _endHInds, _lenEnd = obj.getWatInds()
    endInds = endInds[1::2]
    if subInds is None:
        nPops = 0

    def per_frame(j, sub, pos, box):
        if j == 0:
            return wp.getOrderParamPsi(sub, pos, box, lowCut=0.0, highCut=7.0)
        return wp.getOrderParamPsi(sub, pos, box)

    return _value_driver(obj, endInds, subInds, nPops, per_frame, (0.0, 1.0), 'psiDistribution_%d.txt', 'psiVal    frequency')


def boundWrapPopulations(topFile, trajFile, solResName='(!:WAT)', watResName='(:WAT)', stride=1, cutoff=4.6, nPops=4,
                         cacheFile='boundFile.npy'):
    """Per-frame water populations [bound, wrap, shell, non-shell] for the subInds argument of the frame drivers, cached
    in cacheFile: the workflow of the reference's driver script (orderParam_lib.py:2010-2036) -- an existing cache is
    reused when it holds nPops populations for every frame of the trajectory and discarded otherwise; a fresh one is
    built with getBoundWrap (cutoff 4.6 A there) and saved as a pickled object array.  Returns the list of per-frame lists."""
    import os
    obj = TrajObject(topFile, trajFile, stride, solResName, watResName)
    traj = obj.traj
    watInds, watHInds, _lenWat = obj.getWatInds()
    solInds, solHInds, solCInds, solNInds, solOInds, solSInds = obj.getSolInds()
    if cacheFile and os.path.exists(cacheFile):
        trial = np.load(cacheFile, allow_pickle=True)
        if len(trial) != len(traj) or len(trial) == 0 or len(trial[0]) != nPops:
            os.remove(cacheFile)
        else:
            return [list(row) for row in trial]
    subInds = []
    for frame in traj:
        boundInds, wrapInds, shellInds, nonShellInds = getBoundWrap(obj.top, frame, watInds, watHInds, solInds, solHInds, solCInds,
                                                                    solOInds, solNInds, solSInds, cutoff=cutoff)
        subInds.append([boundInds, wrapInds, shellInds, nonShellInds][:nPops])
    if cacheFile and wdist.world()[0] == 0:
        arr = np.empty((len(subInds), nPops), dtype=object)
        for t, row in enumerate(subInds):
            for j, v in enumerate(row):
                arr[t, j] = v
        np.save(cacheFile, arr, allow_pickle=True)
    return subInds


# ---- radial distribution functions (reference orderParam_lib.py:575-727) --------------------------------------------------

def _basic_simps(y, start, stop, x):
    h = np.diff(x)
    h0, h1 = h[start:stop:2], h[start + 1:stop + 1:2]
    hsum, hprod, h0divh1 = h0 + h1, h0 * h1, h0 / h1
    return np.sum(hsum / 6.0 * (y[start:stop:2] * (2.0 - 1.0 / h0divh1) + y[start + 1:stop + 1:2] * hsum * hsum / hprod
                                + y[start + 2:stop + 2:2] * (2.0 - h0divh1)))


def simps(y, x):
    """Composite Simpson rule over samples y(x) as ``scipy.integrate.simps`` computed it when the reference was written
    (orderParam_lib.py:19; SciPy < 1.11, default even='avg'): with an even number of samples, the mean of (Simpson over
    the first N-1 samples + a trapezoid on the last interval) and (a trapezoid on the first interval + Simpson over the
    last N-1 samples).  Today's ``scipy.integrate.simpson`` treats that case differently and has no ``simps`` name."""
    y, x = np.asarray(y, dtype=np.float64), np.asarray(x, dtype=np.float64)
    N = y.shape[0]
    if N % 2 == 1:
        return _basic_simps(y, 0, N - 2, x)
    val = 0.5 * (x[-1] - x[-2]) * (y[-1] + y[-2])
    result = _basic_simps(y, 0, N - 3, x)
    val += 0.5 * (x[1] - x[0]) * (y[1] + y[0])
    result += _basic_simps(y, 1, N - 2, x)
    return result / 2.0 + val / 2.0


def rdfCalc(topFile, trajFile, solResName='(!:WAT)', watResName='(:WAT)', binwidth=0.1, totbins=150, stride=1):
    """Ow-Ow, solute-solute and solute-Ow radial distribution functions over five trajectory chunks, their running
    coordination numbers (Simpson), the first-shell coordination n1 at the first RDF minimum and the translational
    parameter (reference orderParam_lib.py:575-727).  Writes rdf.txt and coord.txt; returns
    ([n1_OwOw, se], [n1_SolOw, se], [tParam, se]) with a solute, (n1_OwOw, last frame index of a chunk) without, as the
    reference does.  The pair histograms come from wol_pair_hist (bit-exact g(r) per frame, summed in frame order).
    The reference's ``solInds==[]`` tests (:633,:660,:673,:724) are read as "no solute atoms selected"."""
    from scipy.signal import argrelmin
    obj = TrajObject(topFile, trajFile, stride, solResName, watResName)
    traj = obj.traj
    watInds, _watHInds, _lenWat = obj.getWatInds()
    solInds = obj.getSolInds()[0]
    has_sol = len(solInds) > 0
    tot_rdf = {k: [] for k in ("OwOw", "SolOw", "SolSol")}
    tot_coord = {k: [] for k in ("OwOw", "SolOw", "SolSol")}
    tot_n1_OwOw, tot_n1_SolOw, tot_tParam = [], [], []
    nChunks = 5
    chunkSize = int(len(traj) / nChunks)
    dist = np.linspace(0, (totbins - 1) * binwidth, totbins) + binwidth
    bulkdens = 1.0  # local densities, as in the reference (:630)
    t = -1
    for c in range(nChunks):
        rdf_OwOw, rdf_SolOw, rdf_SolSol = np.zeros(totbins), np.zeros(totbins), np.zeros(totbins)
        for t, frame in enumerate(traj[int(c * chunkSize):int((c + 1) * chunkSize)]):
            thispos = np.asarray(frame.xyz)
            thisbox = np.reshape(np.array(frame.box.values[:3]), (1, 3))
            thisWat, thisSol = thispos[watInds], thispos[solInds]
            rdf_OwOw = rdf_OwOw + wl.radialdistsame(thisWat, binwidth, totbins, bulkdens, thisbox)
            if has_sol:
                rdf_SolSol = rdf_SolSol + wl.radialdistsame(thisSol, binwidth, totbins, bulkdens, thisbox)
                rdf_SolOw = rdf_SolOw + wl.radialdist(thisSol, thisWat, binwidth, totbins, bulkdens, thisbox)
        rdf_OwOw, rdf_SolSol, rdf_SolOw = rdf_OwOw / (1.0 + t), rdf_SolSol / (1.0 + t), rdf_SolOw / (1.0 + t)
        tot_rdf["OwOw"].append(rdf_OwOw); tot_rdf["SolSol"].append(rdf_SolSol); tot_rdf["SolOw"].append(rdf_SolOw)
        coord_OwOw, coord_SolOw, coord_SolSol = np.zeros(len(dist) - 2), np.zeros(len(dist) - 2), np.zeros(len(dist) - 2)
        for j in range(2, len(dist)):
            coord_OwOw[j - 2] = 8.0 * np.pi * simps(rdf_OwOw[:j] * (dist[:j]) ** 2.0, dist[:j])
            if has_sol:
                coord_SolOw[j - 2] = 4.0 * np.pi * simps(rdf_SolOw[:j] * (dist[:j]) ** 2.0, dist[:j])
                coord_SolSol[j - 2] = 8.0 * np.pi * simps(rdf_SolSol[:j] * (dist[:j]) ** 2.0, dist[:j])
        tot_coord["OwOw"].append(coord_OwOw); tot_coord["SolSol"].append(coord_SolSol); tot_coord["SolOw"].append(coord_SolOw)
        if has_sol:
            tot_n1_SolOw.append(coord_SolOw[argrelmin(rdf_SolOw)[0][0] - 2])
        first_min = argrelmin(rdf_OwOw)[0][0]
        n1_OwOw = coord_OwOw[first_min - 2]
        rdf = rdf_OwOw[:first_min] / rdf_OwOw[-1]
        tParam = simps(rdf, dist[:first_min]) / dist[first_min]
        tot_n1_OwOw.append(n1_OwOw)
        tot_tParam.append(tParam)

    def se(a, axis=None):
        return np.std(np.array(a), axis=axis, ddof=1) / np.sqrt(nChunks - 1)

    rdf_se = {k: se(v, 0) for k, v in tot_rdf.items()}
    coord_se = {k: se(v, 0) for k, v in tot_coord.items()}
    # as in the reference, the value columns are the LAST chunk's curves, the error columns the spread over chunks
    np.savetxt('rdf.txt', np.stack([dist, rdf_OwOw, rdf_se["OwOw"], rdf_SolSol, rdf_se["SolSol"], rdf_SolOw, rdf_se["SolOw"]], axis=1),
               header='pair distance (A)     Ow-Ow rdf     err     Sol-Sol rdf     err     Sol-Ow rdf     err', fmt="%.3e")
    np.savetxt('coord.txt', np.stack([dist[2:], coord_OwOw, coord_se["OwOw"], coord_SolSol, coord_se["SolSol"], coord_SolOw,
                                      coord_se["SolOw"]], axis=1),
               header='pair distance (A)     Ow-Ow n1     err     Sol-Sol n1     err     Sol-Ow n1     err', fmt="%.3e")
    n1_OwOw, tParam = np.mean(tot_n1_OwOw), np.mean(tot_tParam)
    if has_sol:
        return [n1_OwOw, se(tot_n1_OwOw)], [np.mean(tot_n1_SolOw), se(tot_n1_SolOw)], [tParam, se(tot_tParam)]
    return n1_OwOw, t


# ---- hard-sphere insertion (reference orderParam_lib.py:1666-1791) -----------------------------------------------------

def chemPotCalc(topFile, trajFile, solResName='(!:WAT)', watResName='(:WAT)', probeRadius=3.3, keyword=False, stride=1):
    """Hard-sphere solute insertion statistics: the distribution of the number of heavy atoms overlapping a probe sphere
    of radius probeRadius placed at random -- anywhere in the box, or (keyword=True) inside the 4.2 A shell of a random
    solute atom -- and from it mu = -ln P(0), <N>, <N^2> (reference orderParam_lib.py:1666-1791).  Writes
    HS-solute_overlap_hist.txt / HS-solute_overlap_hist_Shell.txt.  The random insertions are drawn with the same
    np.random calls in the same order as the reference (so a seeded run reproduces it); the overlap counts -- the row sums
    of wl.nearneighbors(hsPos, heavyPos, thisbox, 0.0, probeRadius) (:1727,:1770) -- come from the cell-list kernel."""
    obj = TrajObject(topFile, trajFile, stride, solResName, watResName)
    traj = obj.traj
    solInds = obj.getSolInds()[0]
    heavyInds = traj.top.select('(!@H=)&(!@EPW)')
    cutoff = 4.2
    numOverlap = np.arange(100)
    countOverlap = np.zeros(len(numOverlap))
    dev = torch.device("cuda", torch.cuda.current_device())
    heavy_d = torch.from_numpy(np.ascontiguousarray(np.asarray(heavyInds, dtype=np.int64))).to(dev)
    ws = engine.Workspace(dev)
    for t, frame in enumerate(traj):
        pos = np.array(frame.xyz)
        thisbox = np.array(frame.box.values[:3])
        if keyword:
            count, numIns = 0, 100000
            hsPos = np.zeros((numIns, 3))
            while count < numIns:   # the reference's rejection loop, call for call (:1709-1722)
                randX = 2.0 * (np.random.random(1) - 0.5) * cutoff
                randY = 2.0 * (np.random.random(1) - 0.5) * cutoff
                randZ = 2.0 * (np.random.random(1) - 0.5) * cutoff
                randSq = np.sqrt(randX[0] ** 2.0 + randY[0] ** 2.0 + randZ[0] ** 2.0)
                if randSq > cutoff:
                    continue
                randSolPos = pos[np.random.choice(solInds)]
                hsPos[count, 0] = randSolPos[0] + randX[0]
                hsPos[count, 1] = randSolPos[1] + randY[0]
                hsPos[count, 2] = randSolPos[2] + randZ[0]
                count += 1
        else:
            numIns = 10000
            hsPos = np.zeros((numIns, 3))
            hsPos[:, 0] = np.random.random(numIns) * thisbox[0]
            hsPos[:, 1] = np.random.random(numIns) * thisbox[1]
            hsPos[:, 2] = np.random.random(numIns) * thisbox[2]
        heavyPos = torch.from_numpy(np.ascontiguousarray(pos)).to(dev).index_select(0, heavy_d)
        r = engine.q3b_frames(heavyPos[None], thisbox, torch.from_numpy(hsPos).to(dev)[None], do_q=False, low3=0.0,
                              high3=float(probeRadius), want=("n3",), workspace=ws, device=dev)
        thisTotOverlap = r["n3"][0].cpu().numpy().astype(int)
        thisBins = np.arange(np.max(thisTotOverlap) + 1)
        countOverlap[thisBins] += np.bincount(thisTotOverlap)
    np.savetxt('HS-solute_overlap_hist_Shell.txt' if keyword else 'HS-solute_overlap_hist.txt', np.vstack((numOverlap, countOverlap)).T,
               header='Number of non-solute atoms overlapping           Histogram count')
    with np.errstate(all="ignore"):
        muHS = -np.log(countOverlap[0] / np.sum(countOverlap))
    avgN = np.dot(numOverlap, countOverlap) / np.sum(countOverlap)
    avgN2 = np.dot(numOverlap ** 2.0, countOverlap) / np.sum(countOverlap)
    return muHS, avgN, avgN2


# ---- cluster analysis (reference orderParam_lib.py:123-311 over sortlib.depthfirstsort) -------------------------------

def getClusters(hbMat):
    """Clusters (connected components) of a symmetric residue-connectivity matrix, as a list of index arrays in the
    reference's order: by smallest member, members ascending, an unconnected residue as a cluster of one; stops after
    a cluster that spans every residue (reference orderParam_lib.py:123-156)."""
    hbMat = np.asarray(hbMat)
    n = hbMat.shape[0]
    if n == 0:
        return []
    labels = routines.components(hbMat).cpu().numpy()
    clusters = []
    for root in np.unique(labels):  # ascending smallest member = the order the reference's loop meets them in
        members = np.nonzero(labels == root)[0]
        clusters.append(members)
        if len(members) == n:
            break
    return clusters


def _residue_index(top, atom_indices):
    if hasattr(top, "resids"):
        return np.asarray(top.resids)[np.asarray(atom_indices, dtype=int)]
    return np.array([top.atoms[int(i)].residue.idx for i in atom_indices], dtype=int)  # parmed topology


def _n_residues(top):
    return top.n_residues() if hasattr(top, "n_residues") else len(top.residues)


def getHBClusterStats(topFile, trajFile, acceptorInds, donorInds, donorHInds, stride=1, distCut=3.0, angCut=150.0):
    """Mean size of the hydrogen-bonded residue clusters with more than one member (reference
    orderParam_lib.py:158-233)."""
    obj = TrajObject(topFile, trajFile=trajFile, stride=stride, solResName=None, watResName=None)
    top = obj.top
    acceptorInds, donorInds, donorHInds = (np.asarray(x, dtype=int) for x in (acceptorInds, donorInds, donorHInds))
    resAccept, resDonorH = _residue_index(top, acceptorInds), _residue_index(top, donorHInds)
    nRes = _n_residues(top)
    clusters = []
    for frame in obj.traj:
        thisbox = np.reshape(np.array(frame.box.values[:3]), (1, 3))
        thispos = np.array(frame.xyz)
        allHB = wl.generalhbonds(thispos[acceptorInds], thispos[donorInds], thispos[donorHInds], thisbox, distCut, angCut)
        # residue i is linked to every residue it accepts from or donates to (reference :209-224)
        ai, dj = np.nonzero(allHB == 1)
        hbMat = np.zeros((nRes, nRes))
        hbMat[resAccept[ai], resDonorH[dj]] = 1
        hbMat[resDonorH[dj], resAccept[ai]] = 1
        iClusters = getClusters(hbMat)
        clusters.append(np.array([len(c) for c in iClusters if len(c) != 1]))
    clusters = np.concatenate(clusters)
    return np.mean(clusters)


def getIonClusterStats(topFile, trajFile, Inds, chargeAssign, stride=1, distCut=3.4):
    """Mean size of the contact clusters among the atoms Inds (reference orderParam_lib.py:235-311); writes
    clusterDistribution.txt."""
    obj = TrajObject(topFile, trajFile=trajFile, stride=stride, solResName=None, watResName=None)
    Inds = np.asarray(Inds, dtype=int)
    clusters = []
    for frame in obj.traj:
        thisbox = np.reshape(np.array(frame.box.values[:3]), (1, 3))
        subPos = np.array(frame.xyz)[Inds]
        pairMat = wl.allnearneighbors(subPos, thisbox, 0.0, distCut)
        tClusters = getClusters(pairMat)
        clusters.append(np.array([len(c) for c in tClusters]))
    clusters = np.concatenate(clusters)
    meanCluster = np.mean(clusters)
    clusterDist, bins = np.histogram(clusters, bins=[0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10], density=False)
    np.savetxt('clusterDistribution.txt', np.stack([0.5 * (bins[:-1] + bins[1:]), clusterDist], axis=1),
               header='# clusters    frequency', fmt="%.3e")
    return meanCluster
