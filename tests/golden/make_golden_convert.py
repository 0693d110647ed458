#!/usr/bin/env python
"""Generates tests/golden/convert_small.npz and convert_digests.json by running the REAL libswscale 8.0.1 of this image
(oracle/swscale_so.py) the way the reference's decode thread calls it (/root/reference/source/transvideo.c:2797-2808:
sws_getContext(w, h, source_format, w, h, AV_PIX_FMT_YUV420P, SWS_BICUBIC, NULL, NULL, NULL) + full-frame sws_scale).
Flags alone = parity target A (x86 default); | SWS_BITEXACT = target B.

Run from the repo root:   python tests/golden/make_golden_convert.py
The fixtures pin oracle/ref_convert.c (and through it the CUDA kernel) to the real library."""
import hashlib, json, os, sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import swscale_so as S          # noqa: E402
from conftest import decoded_planes, DECODED_FORMATS  # noqa: E402

SMALL = [(64, 48, 41, "noise"), (66, 38, 42, "noise"), (130, 74, 43, "extreme"), (36, 22, 44, "ramp"), (70, 46, 45, "noise")]
BIG = [(1920, 1080, 1234, "noise"), (854, 480, 77, "noise")]


def run(fmt, w, h, seed, kind, flags):
    planes = [np.ascontiguousarray(p) for p in decoded_planes(fmt, w, h, seed, kind)]
    out = S.scale(planes, w, h, fmt, w, h, "yuv420p", flags)
    return np.concatenate([p.ravel() for p in out])


def main():
    small, digests = {}, {}
    for fmt in DECODED_FORMATS:
        for tname, tflag in (("A", 0), ("B", S.SWS_BITEXACT)):
            for c in SMALL:
                small["%s_%s_%dx%d_s%d_%s" % ((fmt, tname) + c)] = run(fmt, *c, S.SWS_BICUBIC | tflag)
            for c in BIG:
                digests["%s_%s_%dx%d_s%d_%s" % ((fmt, tname) + c)] = hashlib.md5(run(fmt, *c, S.SWS_BICUBIC | tflag).tobytes()).hexdigest()
    np.savez_compressed(os.path.join(HERE, "convert_small.npz"), **small)
    meta = {"ffmpeg": S.version()[0], "swscale_version": S.version()[1], "digests": digests}
    with open(os.path.join(HERE, "convert_digests.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote %d small cases, %d digests (libswscale %s)" % (len(small), len(digests), meta["ffmpeg"]))


if __name__ == "__main__":
    main()
