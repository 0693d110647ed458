"""Pins oracle/ref_convert.c (decode-side converter, transvideo.c:2797-2840) to the real libswscale: against the
committed fixtures (tests/golden/convert_small.npz, convert_digests.json, produced by make_golden_convert.py) and, when
the bundled library is present, against the library itself on fresh inputs."""
import hashlib, json, os, re
import numpy as np
import pytest
from conftest import decoded_planes, DECODED_FORMATS

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = np.load(os.path.join(HERE, "golden", "convert_small.npz"))
DIGESTS = json.load(open(os.path.join(HERE, "golden", "convert_digests.json")))["digests"]
KEY = re.compile(r"(yuv4\d\dp(?:10le)?)_([AB])_(\d+)x(\d+)_s(\d+)_(\w+)$")


def parse(key):
    fmt, target, w, h, seed, kind = KEY.match(key).groups()
    return fmt, target, int(w), int(h), int(seed), kind


def pix(oracle, fmt):
    return {"yuv420p": oracle.PIX_YUV420P, "yuv422p": oracle.PIX_YUV422P, "yuv420p10le": oracle.PIX_YUV420P10,
            "yuv422p10le": oracle.PIX_YUV422P10}[fmt]


@pytest.mark.parametrize("key", sorted(SMALL.files))
def test_fixture(oracle, key):
    fmt, target, w, h, seed, kind = parse(key)
    got = oracle.convert_to_i420(decoded_planes(fmt, w, h, seed, kind), pix(oracle, fmt), w, h, oracle.TARGET_A if target == "A" else oracle.TARGET_B)
    assert np.array_equal(got, SMALL[key])


@pytest.mark.parametrize("key", sorted(DIGESTS))
def test_digest_full_size(oracle, key):
    fmt, target, w, h, seed, kind = parse(key)
    got = oracle.convert_to_i420(decoded_planes(fmt, w, h, seed, kind), pix(oracle, fmt), w, h, oracle.TARGET_A if target == "A" else oracle.TARGET_B)
    assert hashlib.md5(got.tobytes()).hexdigest() == DIGESTS[key]


@pytest.mark.parametrize("fmt", sorted(DECODED_FORMATS))
def test_row_padding_is_ignored(oracle, fmt):
    a = oracle.convert_to_i420(decoded_planes(fmt, 70, 46, 9, "noise", pad=0), pix(oracle, fmt), 70, 46)
    # same samples, padded rows (decoder linesize > width): the generator draws pad columns too, so rebuild the planes
    base = decoded_planes(fmt, 70, 46, 9, "noise", pad=0)
    padded = []
    for p in base:
        q = np.full((p.shape[0], p.shape[1] + 26), 0xAB, p.dtype)
        q[:, :p.shape[1]] = p
        padded.append(q[:, :p.shape[1]])
    b = oracle.convert_to_i420(padded, pix(oracle, fmt), 70, 46)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("fmt", sorted(DECODED_FORMATS))
@pytest.mark.parametrize("size", [(192, 108), (66, 38), (34, 18)])
def test_live_against_bundled_libswscale(oracle, sws, fmt, size):
    w, h = size
    for kind in ("noise", "extreme"):
        planes = decoded_planes(fmt, w, h, 100 + w, kind)
        for flags, target in ((sws.SWS_BICUBIC, oracle.TARGET_A), (sws.SWS_BICUBIC | sws.SWS_BITEXACT, oracle.TARGET_B)):
            ref = np.concatenate([p.ravel() for p in sws.scale([np.ascontiguousarray(p) for p in planes], w, h, fmt, w, h, "yuv420p", flags)])
            assert np.array_equal(oracle.convert_to_i420(planes, pix(oracle, fmt), w, h, target), ref)


def test_10bit_420_is_the_2x2_ordered_dither(oracle):
    """Known answer: yuv420p10le -> yuv420p is clip_u8((s + d[y&1][x&1]) >> 2) with d = {{1,2},{3,0}}."""
    w, h = 8, 4
    y = np.arange(w * h, dtype=np.uint16).reshape(h, w) * 31 % 1024
    y[0, 0], y[1, 0] = 1023, 1021                        # 1023+1 -> 256 clips to 255; 1021+3 -> 256 clips too
    c = np.full((2, 4), 512, np.uint16)
    got = oracle.convert_to_i420([y, c, c], oracle.PIX_YUV420P10, w, h)
    d = np.array([[1, 2], [3, 0]])[np.arange(h)[:, None] & 1, np.arange(w)[None, :] & 1]
    assert np.array_equal(got[: w * h].reshape(h, w), np.clip((y.astype(int) + d) >> 2, 0, 255))
    assert got[0] == 255 and got[w] == 255
    assert np.array_equal(got[w * h:], np.clip((c.astype(int) + d[:2, :4]) >> 2, 0, 255).ravel().tolist() * 2)
