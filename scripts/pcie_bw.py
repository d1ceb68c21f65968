"""Host <-> device bandwidth of every GPU of the box at once (the ceiling of the host-fed legs of bench.py):

    python scripts/pcie_bw.py                                   one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 scripts/pcie_bw.py

Each rank copies from / to its own page-locked buffers; all ranks start together; rank 0 prints one JSON line with the
per-rank and aggregate GB/s for host->device alone, device->host alone and both directions at once.
"""
import json
import os

import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = 192_000_000  # one bench step of float32 frames
h_in = torch.empty(n, dtype=torch.uint8, pin_memory=True)
h_out = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
res = {}
for name, up, down in (("h2d", True, False), ("d2h", False, True), ("both", True, True)):
    best = 0.0
    for rep in range(4):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s1.wait_stream(torch.cuda.current_stream())
        s2.wait_stream(torch.cuda.current_stream())
        if up:
            with torch.cuda.stream(s1):
                for k in range(8):
                    d_in[k * n // 8:(k + 1) * n // 8].copy_(h_in[k * n // 8:(k + 1) * n // 8], non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                for k in range(8):
                    h_out[k * n // 8:(k + 1) * n // 8].copy_(d_out[k * n // 8:(k + 1) * n // 8], non_blocking=True)
        torch.cuda.current_stream().wait_stream(s1)
        torch.cuda.current_stream().wait_stream(s2)
        e1.record()
        torch.cuda.synchronize()
        if rep:
            best = max(best, n * (int(up) + int(down)) / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    t = torch.tensor([best], dtype=torch.float64, device=dev)
    allr = [torch.zeros_like(t) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, t)
    else:
        allr = [t]
    res[name] = {"per_rank_GBps": [round(float(x.item()), 1) for x in allr], "aggregate_GBps": round(sum(float(x.item()) for x in allr), 1)}
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_direction_per_rank": n, "host_cpus": os.cpu_count(), **res}))
if world > 1:
    dist.destroy_process_group()
