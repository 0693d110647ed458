#!/usr/bin/env python
"""Host-side cost of one ottgpu_batch_submit (ASYNC) for N channels, measured with frames so small that the GPU is never
the bottleneck: what a step costs the submitting thread."""
import argparse, importlib, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--channels", type=int, default=32)
    ap.add_argument("--deint", type=int, default=0)
    ap.add_argument("--steps", type=int, default=2000)
    args = ap.parse_args()
    og = importlib.import_module("ott-packager_b200")
    lib = og.Library()
    w, h = 320, 180
    rungs = [(320, 180, og.FMT_I420, 0), (212, 120, og.FMT_I420, 0), (106, 60, og.FMT_I420, 0)]
    n = args.channels
    chans = [og.Channel(lib, og.make_cfg(w, h, rungs, deint=args.deint, thumb=(44, 36, 150), thumb_phase=k)) for k in range(n)]
    batch = og.Batch(lib, chans)
    nb = lib.frame_bytes(og.FMT_I420, w, h)
    src = [[torch.randint(0, 255, (nb,), dtype=torch.uint8, device="cuda") for _ in range(4)] for _ in range(n)]
    dst = [[[torch.empty(b, dtype=torch.uint8, device="cuda") for b in c.rung_bytes] for _ in range(2)] for c in chans]
    ios = []
    for ph in range(4):
        fin, fout = (og.FrameIn * n)(), (og.FrameOut * n)()
        for k in range(n):
            fin[k].data, fin[k].mem, fin[k].interlaced, fin[k].tff = src[k][ph].data_ptr(), og.MEM_DEVICE, 1, 1
            for r in range(3):
                fout[k].dst[r], fout[k].dst_mem[r] = dst[k][ph % 2][r].data_ptr(), og.MEM_DEVICE
        ios.append((fin, fout))
    import ctypes
    libc = ctypes.CDLL(None)
    def step(i):
        fin, fout = ios[i % 4]
        for k in range(n):
            if fout[k].thumbnail:
                libc.free(ctypes.c_void_p(fout[k].thumbnail)); fout[k].thumbnail = None
        lib.check(lib.dll.ottgpu_batch_submit(batch.handle, fin, fout, og.SUBMIT_ASYNC))
    for i in range(50): step(i)
    batch.wait()
    t0 = time.perf_counter()
    tc = 0.0
    for i in range(args.steps):
        fin, fout = ios[i % 4]
        c0 = time.perf_counter()
        lib.check(lib.dll.ottgpu_batch_submit(batch.handle, fin, fout, og.SUBMIT_ASYNC))
        tc += time.perf_counter() - c0
    batch.wait()
    t1 = time.perf_counter()
    print("channels %d deint %d: %.1f us per step wall, %.1f us inside ottgpu_batch_submit" % (n, args.deint, (t1 - t0) / args.steps * 1e6, tc / args.steps * 1e6))

if __name__ == "__main__":
    main()
