#!/usr/bin/env python
"""Aggregate pinned host<->device copy rate of a box with all GPUs copying at once (what bounds the end-to-end SMEM scaling):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/pcie_bw_multi.py
Every rank copies 512 MB device->host and 512 MB host->device concurrently (two streams), REPS times between barriers; rank 0
prints per-GPU and aggregate GB/s per direction, first for one rank alone (the others idle), then for all ranks together."""
import json, os, time
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, REPS = 1 << 29, 6
h_a = torch.empty(n, dtype=torch.uint8).pin_memory(); h_b = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda"); d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
flag = torch.zeros(1, device="cuda")


def run(active):
    dist.all_reduce(flag); torch.cuda.synchronize(); t = time.perf_counter()
    if active:
        for _ in range(REPS):
            with torch.cuda.stream(s1): h_a.copy_(d_a, non_blocking=True)
            with torch.cuda.stream(s2): d_b.copy_(h_b, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    out = torch.tensor([dt], device="cuda"); dist.all_reduce(out, op=dist.ReduceOp.MAX)
    return float(out[0])


run(True)
alone = run(rank == 0)
together = run(True)
if rank == 0:
    per = REPS * n / 1e9
    print(json.dumps({"n_gpus": world, "bytes_per_direction_per_rank": REPS * n,
                      "one_rank_alone_GBs_per_direction": per / alone,
                      "all_ranks_GBs_per_direction_per_rank": per / together,
                      "all_ranks_GBs_per_direction_aggregate": world * per / together,
                      "all_ranks_GBs_both_directions_aggregate": 2 * world * per / together}))
dist.barrier(); dist.destroy_process_group()
