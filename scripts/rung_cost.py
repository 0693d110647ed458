#!/usr/bin/env python
"""Per-rung cost of ladder_kernel: times single-rung (and whole-ladder) channels, frames resident in HBM.
usage: python scripts/rung_cost.py [--channels 64] [--steps 60] [--src 1920x1080] [--fmt nv12|i420] "1280x720" "960x540,640x360" ..."""
import argparse, importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
og = importlib.import_module("ott-packager_b200")
ap = argparse.ArgumentParser()
ap.add_argument("--channels", type=int, default=64)
ap.add_argument("--steps", type=int, default=60)
ap.add_argument("--src", default="1920x1080")
ap.add_argument("--fmt", default="nv12")
ap.add_argument("--p010", action="store_true", help="P010 source and rungs, lanczos (cfg3 style)")
ap.add_argument("sets", nargs="+")
a = ap.parse_args()
lib = og.Library()
w, h = map(int, a.src.split("x"))
fmt = og.FMT_P010 if a.p010 else (og.FMT_NV12 if a.fmt == "nv12" else og.FMT_I420)
src_fmt = og.FMT_P010 if a.p010 else og.FMT_I420
src_bytes = lib.frame_bytes(src_fmt, w, h)
rng = np.random.default_rng(3)
frame = ((rng.integers(0, 1024, src_bytes // 2, dtype=np.uint16) << 6).astype(np.uint16).view(np.uint8) if a.p010
         else rng.integers(16, 236, src_bytes, dtype=np.uint8))
Cn = a.channels
src_pool = og.Pool(lib, 0, og.MEM_DEVICE, 2 * Cn, src_bytes)
srcs = [[src_pool.take() for _ in range(2)] for _ in range(Cn)]
for c in range(Cn):
    for k in range(2):
        lib.to_device(srcs[c][k], np.roll(frame, 98 * (2 * c + k)))
for spec in a.sets:
    rungs = [tuple(map(int, r.split("x"))) + (fmt, 256 if fmt == og.FMT_NV12 else 0) for r in spec.split(",")]
    chans = [og.Channel(lib, og.make_cfg(w, h, rungs, src_fmt=src_fmt, filt=og.FILTER_LANCZOS if a.p010 else og.FILTER_BICUBIC,
                                         deint=og.DEINT_BYPASS)) for _ in range(Cn)]
    batch = og.Batch(lib, chans)
    pools = [og.Pool(lib, 0, og.MEM_DEVICE, Cn, rb) for rb in chans[0].rung_bytes]
    dst = [[p.take() for p in pools] for _ in range(Cn)]
    for c in range(Cn):
        batch.fin[c].mem = og.MEM_DEVICE
        batch.fin[c].interlaced, batch.fin[c].tff = 0, 1
        for r in range(len(rungs)):
            batch.fout[c].dst[r], batch.fout[c].dst_mem[r] = dst[c][r], og.MEM_DEVICE
    def step(k):
        for c in range(Cn):
            batch.fin[c].data = srcs[c][k & 1]
        batch.submit_raw(og.SUBMIT_ASYNC)
    for k in range(5):
        step(k)
    batch.wait()
    batch.timing(True)
    for k in range(a.steps):
        step(k)
    batch.wait()
    ms, n = batch.timing_read()
    out_bytes = sum(chans[0].rung_bytes)
    us = 1000.0 * ms / max(n, 1) / Cn
    print("%-44s %7.3f us/frame  %6.0f GB/s  (%d items)" % (spec, us, (src_bytes + out_bytes) / us / 1e3, batch.last_items), flush=True)
    batch.close()
    for ch in chans:
        ch.close()
    for p in pools:
        p.close()
