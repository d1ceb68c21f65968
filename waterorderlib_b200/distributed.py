"""Multi-GPU layer: trajectories shard by FRAME, one process per GPU, no data-path collective.

Frames are independent units in the reference's drivers (structureLibs/orderParam_lib.py:1312-1353,
:1458-1480 carry no state between iterations except list appends), so each rank analyses a contiguous block
of frames with its own device-resident int64 histograms and per-frame statistics rows.  At the end:

  * ONE all-reduce (sum, int64) of the concatenated histograms -- integer sums, so the result is
    bit-identical for any number of ranks;
  * ONE all-gather of the per-frame rows (the drivers return per-frame series for blockAverage,
    orderParam_lib.py:1355-1365), reassembled in frame order.

Both run over whatever backend the process group was created with: NCCL over NVLink 5 / NVSwitch on the
B200 box (tensors stay on the device), gloo in the CPU tests.  With no process group everything degrades to
the single-rank identity, so the drivers need no separate code path.
"""
import torch
import torch.distributed as dist


def world():
    """(rank, world_size) of the default process group, (0, 1) when none is initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_frames(n_frames, rank=None, world_size=None):
    """Contiguous block [begin, end) of frames owned by `rank`: blocks differ by at most one frame and their
    concatenation in rank order is the trajectory, so gathered per-frame rows need no permutation."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, extra = divmod(int(n_frames), int(world_size))
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def shard_sizes(n_frames, world_size):
    return [shard_frames(n_frames, r, world_size)[1] - shard_frames(n_frames, r, world_size)[0] for r in range(world_size)]


def reduce_histograms(*hists):
    """Sum integer histograms over ranks with ONE all-reduce (they are packed into a single buffer).
    In place; returns the tensors."""
    if world()[1] == 1 or not hists:
        return hists
    for h in hists:
        if h.dtype != torch.int64:
            raise TypeError("histograms must be int64 so that the reduction is exact")
    flat = torch.cat([h.reshape(-1) for h in hists])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    o = 0
    for h in hists:
        n = h.numel()
        h.copy_(flat[o:o + n].view_as(h))
        o += n
    return hists


def gather_frame_rows(local_rows, n_frames_total):
    """All-gather per-frame rows: local_rows (f_local, ...) on every rank -> (n_frames_total, ...) in frame
    order on every rank.  ONE all-gather of equal-sized padded blocks."""
    rank, ws = world()
    if ws == 1:
        return local_rows
    sizes = shard_sizes(n_frames_total, ws)
    if local_rows.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d rows, its shard has %d frames" % (rank, local_rows.shape[0], sizes[rank]))
    pad = max(sizes)
    tail = tuple(local_rows.shape[1:])
    block = torch.zeros((pad,) + tail, dtype=local_rows.dtype, device=local_rows.device)
    block[: sizes[rank]] = local_rows
    out = torch.empty((ws * pad,) + tail, dtype=local_rows.dtype, device=local_rows.device)
    dist.all_gather_into_tensor(out, block)
    out = out.view((ws, pad) + tail)
    return torch.cat([out[r, : sizes[r]] for r in range(ws)], dim=0)


def max_over_ranks(x, device):
    """Max of a python float over ranks (device-timed durations are reported as the slowest rank's)."""
    if world()[1] == 1:
        return float(x)
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
