#!/usr/bin/env python
"""Concurrent host<->device copy ceiling of the box: every rank (one per GPU, torchrun) copies pinned host memory to its
GPU at the same time.  Prints one JSON line (rank 0): per-rank and aggregate GB/s for H2D alone, D2H alone and both
directions at once, for the step-sized transfer of the bench (64 x 3.1 MB) as one copy and as 64 copies.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/pcie_probe_multi.py"""
import json, os, sys
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
FRAME = 3110400
N = 64
host = torch.empty(N * FRAME, dtype=torch.uint8).pin_memory()
host_out = torch.empty(N * FRAME, dtype=torch.uint8).pin_memory()
dev = torch.empty(N * FRAME, dtype=torch.uint8, device="cuda")
dev2 = torch.empty(N * FRAME, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def barrier():
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()


def timed(fn, reps=20):
    fn()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def h2d_one():
    dev.copy_(host, non_blocking=True)


def h2d_many():
    for i in range(N):
        dev[i * FRAME:(i + 1) * FRAME].copy_(host[i * FRAME:(i + 1) * FRAME], non_blocking=True)


def d2h_one():
    host_out.copy_(dev2, non_blocking=True)


def both():
    with torch.cuda.stream(s1):
        dev.copy_(host, non_blocking=True)
    with torch.cuda.stream(s2):
        host_out.copy_(dev2, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


res = {}
gb = N * FRAME / 1e9
for name, fn, mult in (("h2d_one_copy", h2d_one, 1), ("h2d_64_copies", h2d_many, 1), ("d2h_one_copy", d2h_one, 1), ("both_directions", both, 2)):
    ms = timed(fn)
    res[name] = {"ms_max_over_ranks": ms, "per_gpu_GBps": mult * gb / (ms / 1e3), "aggregate_GBps": world * mult * gb / (ms / 1e3)}
if rank == 0:
    print(json.dumps({"n_gpus": world, "bytes_per_copy": N * FRAME, "host_cores": len(os.sched_getaffinity(0)), "results": res}))
if world > 1:
    dist.destroy_process_group()
