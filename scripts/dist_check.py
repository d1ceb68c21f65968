"""Multi-GPU consistency check (run under torchrun on N GPUs): the frame drivers give the same histograms and
per-frame series on N ranks as on one rank (integer sums are exact, rows are gathered in frame order).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
"""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from waterorderlib_b200 import synth  # noqa: E402
from waterorderlib_b200.structureLibs import orderParam_lib as opl  # noqa: E402
from waterorderlib_b200.structureLibs.TrajObject import ArrayTrajectory, Topology  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
T, m = 21, 6
n_w = 8 * m ** 3
top = Topology.water_box(n_w)
xyz = np.zeros((T, 3 * n_w, 3))
boxes = np.zeros((T, 3))
for f in range(T):
    o, box = synth.water_box(m, sigma=0.35, seed=40 + f)
    h = synth.add_hydrogens(o, seed=40 + f)
    xyz[f, 0::3], xyz[f, 1::3], xyz[f, 2::3], boxes[f] = o, h[0::2], h[1::2], box
traj = ArrayTrajectory(xyz, boxes, top=top)
work = tempfile.mkdtemp()
os.chdir(work)

# single-rank result first (no process group yet: the drivers see world == 1)
np.random.seed(1)
ref_q = opl.tetOrderCalc(top, traj)
ref_3b = opl.threeBodyCalc(top, traj)
ref_hb = opl.hbCalc(top, traj)
ref_files = {n: np.loadtxt(n) for n in ("qDistribution_0.txt", "3bDistribution_0.txt", "hbDistribution_water.txt")}

dist.init_process_group("nccl", device_id=torch.device("cuda", local))
np.random.seed(1)
got_q = opl.tetOrderCalc(top, traj)
got_3b = opl.threeBodyCalc(top, traj)
got_hb = opl.hbCalc(top, traj)
dist.barrier()
ok = True
if rank == 0:
    for a, b in zip(ref_q + ref_3b, got_q + got_3b):
        ok &= np.allclose(a[0], b[0], rtol=1e-12, atol=0) and np.allclose(a[1], b[1], rtol=1e-9, atol=1e-15)
    ok &= ref_hb[0] == got_hb[0]
    for n, want in ref_files.items():
        ok &= np.array_equal(np.loadtxt(n), want)
    print("dist_check world=%d: %s  <q>=%.6f  pTet=%.6f  HB/water=%.4f" % (world, "OK" if ok else "MISMATCH", got_q[0][0][0], got_3b[0][0][0], got_hb[0]))

# the C-ABI combine over a communicator the host owns (what a native, non-torch host would do): same sums as NCCL via torch
import ctypes  # noqa: E402
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "tools"))
import raw_nccl  # noqa: E402
from waterorderlib_b200._capi import lib  # noqa: E402
nccl = raw_nccl.load()
uid = raw_nccl.unique_id(nccl) if rank == 0 else raw_nccl.UniqueId()
t = torch.tensor(list(bytes(uid)) if rank == 0 else [0] * 128, dtype=torch.uint8, device="cuda")
dist.broadcast(t, 0)
uid = raw_nccl.UniqueId.from_buffer_copy(bytes(t.cpu().tolist()))
comm = raw_nccl.comm_init(nccl, uid, world, rank)
hist = (torch.arange(1011, dtype=torch.int64, device="cuda") + rank) * 1_000_003
via_torch = hist.clone()
dist.all_reduce(via_torch)
rc = lib().wol_hist_allreduce(comm, ctypes.c_void_p(hist.data_ptr()), hist.numel(), 0, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
ok2 = rc == 0 and torch.equal(hist, via_torch)
nccl.ncclCommDestroy(comm)
flag = torch.tensor([1 if ok2 else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("wol_hist_allreduce over a raw ncclComm_t, world=%d: %s" % (world, "OK" if flag.item() == 1 else "MISMATCH"))
ok = ok and flag.item() == 1
dist.destroy_process_group()
sys.exit(0 if ok else 1)
