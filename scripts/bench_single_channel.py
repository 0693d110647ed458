#!/usr/bin/env python
"""One channel driven the way ONE fillet process drives it (the reference runs one process per channel): synchronous
ottgpu_channel_submit per frame, pinned host frame in, rungs to host memory (x264/x265 mode) or left in device surfaces
(NVENC mode).  Prints per-frame latency.  The CPU figure to hold beside it is bench.py's cpu_baseline.single_thread (one
libswscale thread doing the same ladder, how the reference's video_scale_thread works, transvideo.c:1504-1683)."""
import argparse, importlib, json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--frames", type=int, default=300)
    args = ap.parse_args()
    B.claim_stdout()
    import torch
    og = importlib.import_module("ott-packager_b200")
    lib = og.Library()
    wl = B.WORKLOADS[args.workload]
    w, h = wl["w"], wl["h"]
    fmt = wl["out_fmt"]
    rungs = [(dw, dh, fmt, wl["pitch_align"]) for dw, dh in wl["rungs"]]
    frames = B.synth_frames(w, h, 4, wl["interlaced"])
    pinned = [torch.from_numpy(f).pin_memory() for f in frames]
    out = {}
    for mode in ("host_rungs", "device_rungs"):
        ch = og.Channel(lib, og.make_cfg(w, h, rungs, deint=wl["deint"], thumb=B.THUMB))
        fin, fout = og.FrameIn(), og.FrameOut()
        host_dst = [torch.empty(n, dtype=torch.uint8).pin_memory() for n in ch.rung_bytes]
        dev_dst = [torch.empty(n, dtype=torch.uint8, device="cuda") for n in ch.rung_bytes]
        for r in range(len(rungs)):
            if mode == "host_rungs":
                fout.dst[r], fout.dst_mem[r] = host_dst[r].data_ptr(), og.MEM_HOST
            else:
                fout.dst[r], fout.dst_mem[r] = dev_dst[r].data_ptr(), og.MEM_DEVICE
        libc = __import__("ctypes").CDLL(None)
        lat = []
        for i in range(args.frames + 20):
            fin.data, fin.mem = pinned[i % 4].data_ptr(), og.MEM_HOST
            fin.interlaced, fin.tff = wl["interlaced"], 1
            t0 = time.perf_counter()
            lib.check(lib.dll.ottgpu_channel_submit(ch.handle, fin, fout))
            t1 = time.perf_counter()
            if fout.thumbnail:
                libc.free(__import__("ctypes").c_void_p(fout.thumbnail))
                fout.thumbnail = None
            if i >= 20:
                lat.append((t1 - t0) * 1000.0)
        ch.close()
        lat.sort()
        out[mode] = {"ms_median": lat[len(lat) // 2], "ms_p99": lat[int(len(lat) * 0.99)], "frames_per_sec": 1000.0 / (sum(lat) / len(lat))}
    B.emit({"metric": "single_channel_frame_latency", "workload": wl["desc"], "gpu": out})


if __name__ == "__main__":
    main()
