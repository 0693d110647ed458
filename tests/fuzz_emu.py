#!/usr/bin/env python
"""TEST INFRASTRUCTURE: seeded random differential check of the kernel source (CPU emulation build, tests/emu) against the
oracle - ladder geometry (I420/NV12, P010), yadif frame sizes, decode-side converter.  Lives in tests/ (it uses the oracle) but is not collected by pytest
(minutes, not seconds); run it after touching the planner or a kernel core:

    python tests/fuzz_emu.py [--cases 600] [--seed 77] [--lib path/to/libottgpu(.so|_emu.so)]"""
import argparse, importlib, os, subprocess, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import noise_i420, noise_p010, decoded_planes, DECODED_FORMATS  # noqa: E402
from oracle import oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=600)
    ap.add_argument("--seed", type=int, default=77)
    ap.add_argument("--lib", default=None)
    args = ap.parse_args()
    og = importlib.import_module("ott-packager_b200")
    if args.lib is None:
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
        args.lib = os.path.join(ROOT, "tests", "emu", "libottgpu_emu.so")
    lib = og.Library(args.lib)
    rng = np.random.default_rng(args.seed)
    bad = 0

    def report(kind, what):
        nonlocal bad
        bad += 1
        print("MISMATCH", kind, what)

    for it in range(args.cases):
        sw, sh = int(rng.integers(8, 200)) * 2, int(rng.integers(4, 120)) * 2
        dw = max(8, int(sw * rng.uniform(0.08, 3.0))) & ~1
        dh = max(8, int(sh * rng.uniform(0.08, 3.0))) & ~1
        lanczos, tb, nv = (int(rng.integers(0, 2)) for _ in range(3))
        src = noise_i420(sw, sh, it)
        if it % 3 == 1:                                          # video-range content: no negative intermediates, so the V
            src = (16 + (src.astype(np.uint16) * 220 >> 8)).astype(np.uint8)   # pass runs as FMAs (noise mostly takes the integer path)
        elif it % 3 == 2:                                        # both paths in one picture: half the rows in video range
            half = src[: sw * sh].reshape(sh, sw)
            half[: sh // 2] = 16 + (half[: sh // 2].astype(np.uint16) * 220 >> 8)
        got = lib.scale_frame(src, og.FMT_I420, sw, sh, dw, dh, dst_fmt=og.FMT_NV12 if nv else og.FMT_I420,
                              filt=og.FILTER_LANCZOS if lanczos else og.FILTER_BICUBIC, target=tb)
        ref = oracle.scale_i420(src, sw, sh, dw, dh, oracle.LANCZOS if lanczos else oracle.BICUBIC, tb)
        if nv:
            ref = oracle.i420_to_nv12(ref, dw, dh)
        if not np.array_equal(got, ref):
            report("ladder", (sw, sh, dw, dh, lanczos, tb, nv))
    for it in range(args.cases // 4):
        sw, sh = int(rng.integers(8, 120)) * 2, int(rng.integers(4, 80)) * 2
        dw = max(8, int(sw * rng.uniform(0.15, 2.5))) & ~1
        dh = max(8, int(sh * rng.uniform(0.15, 2.5))) & ~1
        lanczos, tb = int(rng.integers(0, 2)), int(rng.integers(0, 2))
        src = noise_p010(sw, sh, it)
        got = lib.scale_frame(src, og.FMT_P010, sw, sh, dw, dh, filt=og.FILTER_LANCZOS if lanczos else og.FILTER_BICUBIC, target=tb)
        if not np.array_equal(got, oracle.scale_p010(src, sw, sh, dw, dh, oracle.LANCZOS if lanczos else oracle.BICUBIC, tb)):
            report("p010", (sw, sh, dw, dh, lanczos, tb))
    for it in range(args.cases // 4):
        w, h = int(rng.integers(4, 150)) * 2, int(rng.integers(3, 90)) * 2
        if rng.integers(0, 2):
            w = (w + 15) // 16 * 16                             # the 8-samples-per-thread path needs w % 16 == 0
        fr = [noise_i420(w, h, 3 * it + k) for k in range(3)]
        tff = int(rng.integers(0, 2))
        if not np.array_equal(lib.yadif_frame(fr[0], fr[1], fr[2], og.FMT_I420, w, h, 1, tff),
                              oracle.yadif_frame(fr[0], fr[1], fr[2], w, h, 1, tff)):
            report("yadif", (w, h, tff))
    pix = {"yuv420p": og.PIX_YUV420P, "yuv422p": og.PIX_YUV422P, "yuv420p10le": og.PIX_YUV420P10, "yuv422p10le": og.PIX_YUV422P10}
    for it in range(args.cases // 3):
        fmt = list(DECODED_FORMATS)[it % 4]
        w, h = int(rng.integers(4, 150)) * 2, int(rng.integers(2, 90)) * 2
        if rng.integers(0, 2):
            w = (w + 31) // 32 * 32                             # vector path
        planes = decoded_planes(fmt, w, h, it, "noise" if it % 3 else "extreme", pad=int(rng.integers(0, 3)) * 8)
        tb = int(rng.integers(0, 2))
        cv = og.Converter(lib, pix[fmt], w, h, tb)
        got = cv.convert(planes)
        cv.close()
        if not np.array_equal(got, oracle.convert_to_i420(planes, pix[fmt], w, h, tb)):
            report("convert", (fmt, w, h, tb))
    print("fuzz done: %d mismatches" % bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
