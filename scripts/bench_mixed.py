#!/usr/bin/env python
"""BASELINE.json configs[4]: mixed-mode box, 16 x 2160p60 (4 rungs) + 48 x 1080i (yadif + 4 rungs) channels with thumbnails,
channel-sharded over N GPUs (STRONG scaling: the 64 channels are fixed, `ott-packager_b200/sharding.py` places them).

  python scripts/bench_mixed.py [--steps K] [--warmup W]                      (1 GPU)
  python -m torch.distributed.run --nproc-per-node N ... scripts/bench_mixed.py --gpus N

One step = one frame of every channel of the rank.  Prints one JSON line (rank 0): frames/s over all ranks, the
"channel seconds per second" (how many times real time the box runs the 64-channel line-up), device-resident and
end-to-end (pinned host frames in, rungs in device surfaces, thumbnails to host)."""
import argparse
import ctypes as C
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402

I420 = 0
SPECS = [dict(w=3840, h=2160, fps=60000 / 1001.0, il=0, rungs=[(3840, 2160), (1920, 1080), (1280, 720), (640, 360)])] * 16 + \
        [dict(w=1920, h=1080, fps=30000 / 1001.0, il=1, rungs=[(1920, 1080), (1280, 720), (960, 540), (640, 360)])] * 48


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=8)
    args = ap.parse_args()
    B.claim_stdout()
    import torch
    og = importlib.import_module("ott-packager_b200")
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = og.Library()
    stream = torch.cuda.Stream()
    mine = og.my_channels(SPECS, rank, world)
    specs = [SPECS[i] for i in mine]
    n = len(specs)
    frames = {(3840, 2160): B.synth_frames(3840, 2160, 2, 0), (1920, 1080): B.synth_frames(1920, 1080, 2, 1)}

    def make():
        return [og.Channel(lib, og.make_cfg(s["w"], s["h"], [(dw, dh, I420, 0) for dw, dh in s["rungs"]],
                                            deint=og.DEINT_ALL if s["il"] else og.DEINT_FLAGGED, gpu=local_rank, thumb=B.THUMB,
                                            stream=stream.cuda_stream, thumb_phase=(k * B.THUMB[2]) // max(n, 1)))
                for k, s in enumerate(specs)]

    def region(host_inputs):
        chans = make()
        batch = og.Batch(lib, chans)
        pools, ios = [], []
        RING = 4
        src, dst = [], []
        for c, s in zip(chans, specs):
            sb = B.frame_bytes(s["w"], s["h"])
            p = og.Pool(lib, local_rank, og.MEM_HOST if host_inputs else og.MEM_DEVICE, RING, sb)
            pools.append(p)
            bufs = [p.take() for _ in range(RING)]
            for r, b in enumerate(bufs):
                f = frames[(s["w"], s["h"])][r % 2]
                if host_inputs:
                    C.memmove(b, f.ctypes.data, sb)
                else:
                    lib.to_device(b, f, gpu=local_rank)
            src.append(bufs)
            dp = [og.Pool(lib, local_rank, og.MEM_DEVICE, 2, rb) for rb in c.rung_bytes]
            pools += dp
            dst.append([[q.take() for q in dp] for _ in range(2)])
        for phase in range(RING):
            fin, fout = (og.FrameIn * n)(), (og.FrameOut * n)()
            for k, s in enumerate(specs):
                fin[k].data, fin[k].mem = src[k][phase], og.MEM_HOST if host_inputs else og.MEM_DEVICE
                fin[k].interlaced, fin[k].tff = s["il"], 1
                for r in range(len(s["rungs"])):
                    fout[k].dst[r], fout[k].dst_mem[r] = dst[k][phase % 2][r], og.MEM_DEVICE
            ios.append((fin, fout))
        libc = C.CDLL(None)
        launches = 0

        def steps(count):
            nonlocal launches
            for i in range(count):
                fin, fout = ios[i % RING]
                for k in range(n):
                    if fout[k].thumbnail:
                        libc.free(C.c_void_p(fout[k].thumbnail))
                        fout[k].thumbnail = None
                lib.check(lib.dll.ottgpu_batch_submit(batch.handle, fin, fout, og.SUBMIT_ASYNC))
                launches += batch.last_launches

        with torch.cuda.stream(stream):
            steps(args.warmup)
            batch.wait()
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
                torch.cuda.synchronize()
            launches = 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            steps(args.steps)
            e1.record(stream)
            batch.wait()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        batch.close()
        for c in chans:
            c.close()
        for p in pools:
            p.close()
        return ms, launches

    ms, launches = region(False)
    ems, _ = region(True)
    if rank == 0:
        total = len(SPECS)
        # every rank advances all of its channels by one frame per step; the slowest rank sets the pace
        fps = total * args.steps / (ms / 1000.0)
        efps = total * args.steps / (ems / 1000.0)
        # real-time factor: one step must not take longer than the fastest channel's frame period (4K60: 16.68 ms)
        rt = (1000.0 / (60000 / 1001.0)) / (ms / args.steps)
        ert = (1000.0 / (60000 / 1001.0)) / (ems / args.steps)
        B.emit({"metric": "ladder_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                          "dtype": "u8", "data": "synthetic",
                          "config": {"workload": "cfg5: 16 x 2160p60 (4 rungs) + 48 x 1080i (yadif + 4 rungs), 176x144 thumbnails /150, "
                                                 "channel-sharded, stepped in lock-step (the 1080i channels are advanced at the 4K rate)",
                                     "channels": total, "channels_rank0": n, "realtime_factor": rt},
                          "e2e": {"value": efps, "unit": "frames/s", "ms_per_step": ems / args.steps, "realtime_factor": ert},
                          "gpu_launches": launches})
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
