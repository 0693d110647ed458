#!/usr/bin/env python
"""Thumbnail JPEG: time per picture on the GPU (ottgpu_jpeg_encode, host planes in, file bytes out, synchronous:
what host/thumbnail_gpu.c pays per thumbnail) beside the reference's way of doing it on this box's CPU (the real libavcodec
MJPEG encoder with a fresh context per picture, transvideo.c:164-198) and the sequential C restatement.  One JSON line per size.
usage: python scripts/bench_jpeg.py [--n 200]"""
import argparse, importlib, json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import natural_i420  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200)
    args = ap.parse_args()
    og = importlib.import_module("ott-packager_b200")
    lib = og.Library(os.environ.get("OTTGPU_LIB"))
    from oracle import oracle as O, avcodec_so as A       # CPU legs only (bench-side use of oracle/, like bench.py's cpu_baseline)
    for (w, h, n) in ((176, 144, args.n), (1920, 1080, max(10, args.n // 10))):
        frames = [natural_i420(w, h, 3 + i, shift=7.0 * i) for i in range(4)]
        qs = 3
        if A.available():                                  # the qscale libavcodec's rate control picks for this content
            qs = A.qscale_of(A.mjpeg_encode(frames[(n - 1) & 3], w, h))[0] or 3
        enc = og.Jpeg(lib, w, h, qs)
        for f in frames:
            enc.encode(f)
        t0 = time.perf_counter()
        for i in range(n):
            out = enc.encode(frames[i & 3])
        gpu_us = (time.perf_counter() - t0) / n * 1e6
        pool = og.Pool(lib, 0, og.MEM_DEVICE, 1, frames[0].size)
        d = pool.take()
        lib.to_device(d, frames[0])
        t0 = time.perf_counter()
        for i in range(n):
            enc.encode(device_ptr=d)
        gpu_dev_us = (time.perf_counter() - t0) / n * 1e6
        same = out == O.jpeg_encode(frames[(n - 1) & 3], w, h, qs)
        t0 = time.perf_counter()
        for i in range(max(3, n // 10)):
            O.jpeg_encode(frames[i & 3], w, h, qs)
        port_us = (time.perf_counter() - t0) / max(3, n // 10) * 1e6
        ref_us, ref_bytes = None, None
        if A.available():
            t0 = time.perf_counter()
            for i in range(max(3, n // 10)):
                A.mjpeg_encode(frames[i & 3], w, h)
            r = A.mjpeg_encode(frames[(n - 1) & 3], w, h)
            ref_us = (time.perf_counter() - t0) / max(3, n // 10) * 1e6
            ref_bytes = len(r)
        print(json.dumps({"what": "thumbnail JPEG", "size": "%dx%d" % (w, h), "blocks": 6 * ((w + 15) // 16) * ((h + 15) // 16),
                          "gpu_us_per_picture_host_in": round(gpu_us, 1), "gpu_us_per_picture_device_in": round(gpu_dev_us, 1),
                          "qscale": qs, "file_bytes": len(out), "equals_cpu_restatement": bool(same),
                          "cpu_libavcodec_us_per_picture": None if ref_us is None else round(ref_us, 1), "libavcodec_file_bytes": ref_bytes,
                          "cpu_restatement_us_per_picture": round(port_us, 1),
                          "note": "GPU figure = synchronous call: H2D of the planes, 8 small launches, two D2H; libavcodec = fresh context per picture as save_frame_as_jpeg"}))
        enc.close()
        pool.close()


if __name__ == "__main__":
    main()
