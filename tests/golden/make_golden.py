#!/usr/bin/env python
"""Generates tests/golden/swscale_small.npz and swscale_digests.json by running the REAL libswscale 8.0.1 that ships in
this image (oracle/swscale_so.py -> opencv_python_headless.libs/libswscale-*.so.9.1.100) exactly the way the reference
calls it (/root/reference/source/transvideo.c:1563-1581: sws_getContext(..., SWS_BICUBIC, NULL, NULL, NULL) +
full-frame sws_scale).  Flags alone = parity target A (x86 default); | SWS_BITEXACT = target B.

Run from the repo root:   python tests/golden/make_golden.py
The fixtures pin oracle/ref_swscale.c (and through it the CUDA kernels) to the real library even on a machine where
the bundled .so is missing.
"""
import hashlib, json, os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import swscale_so as S          # noqa: E402
from conftest import noise_i420, noise_p010, split_i420, chroma_dims  # noqa: E402

KERNELS = {"bicubic": S.SWS_BICUBIC, "lanczos": S.SWS_LANCZOS}

# (src_w, src_h, dst_w, dst_h, seed)
SMALL_I420 = [
    (96, 64, 64, 48, 11), (96, 64, 48, 32, 12), (96, 64, 40, 24, 13), (96, 64, 30, 18, 14),
    (160, 90, 106, 60, 15), (64, 48, 96, 72, 16), (96, 64, 96, 40, 17), (96, 64, 50, 64, 18),
    (128, 96, 22, 18, 19), (70, 46, 34, 22, 20),
]
SMALL_P010 = [(128, 72, 64, 36, 31), (128, 72, 96, 54, 32), (128, 72, 42, 24, 33), (64, 36, 128, 72, 34)]
BIG_I420 = [(1920, 1080, dw, dh, 1234) for dw, dh in
            [(1280, 720), (960, 540), (640, 360), (416, 234), (176, 144), (854, 480)]]
BIG_P010 = [(3840, 2160, dw, dh, 99) for dw, dh in [(1920, 1080), (640, 360)]]


def run_i420(sw, sh, dw, dh, seed, flags):
    src = noise_i420(sw, sh, seed)
    planes = [np.ascontiguousarray(p) for p in split_i420(src, sw, sh)]
    out = S.scale(planes, sw, sh, "yuv420p", dw, dh, "yuv420p", flags)
    return np.concatenate([p.ravel() for p in out])


def run_p010(sw, sh, dw, dh, seed, flags):
    src = noise_p010(sw, sh, seed)
    cw, ch = chroma_dims(sw, sh)
    planes = [np.ascontiguousarray(src[: sw * sh].reshape(sh, sw)),
              np.ascontiguousarray(src[sw * sh:].reshape(ch, cw * 2))]
    out = S.scale(planes, sw, sh, "p010le", dw, dh, "p010le", flags)
    return np.concatenate([p.ravel() for p in out])


def main():
    small, digests = {}, {}
    for kname, kflag in KERNELS.items():
        for tname, tflag in (("A", 0), ("B", S.SWS_BITEXACT)):
            for c in SMALL_I420:
                small["i420_%s_%s_%dx%d_%dx%d_s%d" % ((kname, tname) + c)] = run_i420(*c, kflag | tflag)
            for c in SMALL_P010:
                small["p010_%s_%s_%dx%d_%dx%d_s%d" % ((kname, tname) + c)] = run_p010(*c, kflag | tflag)
            for c in BIG_I420:
                digests["i420_%s_%s_%dx%d_%dx%d_s%d" % ((kname, tname) + c)] = \
                    hashlib.md5(run_i420(*c, kflag | tflag).tobytes()).hexdigest()
    for tname, tflag in (("A", 0), ("B", S.SWS_BITEXACT)):
        for c in BIG_P010:
            digests["p010_lanczos_%s_%dx%d_%dx%d_s%d" % ((tname,) + c)] = \
                hashlib.md5(run_p010(*c, S.SWS_LANCZOS | tflag).tobytes()).hexdigest()
    np.savez_compressed(os.path.join(HERE, "swscale_small.npz"), **small)
    meta = {"ffmpeg": S.version()[0], "swscale_version": S.version()[1], "digests": digests}
    with open(os.path.join(HERE, "swscale_digests.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote %d small cases, %d digests (libswscale %s)" % (len(small), len(digests), meta["ffmpeg"]))


if __name__ == "__main__":
    main()
