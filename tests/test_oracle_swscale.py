"""Pins oracle/ref_swscale.c (the CPU restatement the CUDA kernels are checked against) to the REAL libswscale 8.0.1:
 (1) against committed fixtures produced by tests/golden/make_golden.py (always runs),
 (2) live against the bundled .so on fresh inputs when it is present in the image.
Reference call sites being restated: /root/reference/source/transvideo.c:1563-1581 (rungs), 2155-2174 (thumbnail)."""
import hashlib, json, os, re
import numpy as np
import pytest
from conftest import noise_i420, noise_p010, natural_i420, split_i420, chroma_dims

G = os.path.join(os.path.dirname(__file__), "golden")
SMALL = np.load(os.path.join(G, "swscale_small.npz"))
DIGESTS = json.load(open(os.path.join(G, "swscale_digests.json")))["digests"]
KEY = re.compile(r"(i420|p010)_(bicubic|lanczos)_([AB])_(\d+)x(\d+)_(\d+)x(\d+)_s(\d+)")


def _run(O, key):
    fmt, kern, tgt, sw, sh, dw, dh, seed = KEY.fullmatch(key).groups()
    sw, sh, dw, dh, seed = map(int, (sw, sh, dw, dh, seed))
    k = O.BICUBIC if kern == "bicubic" else O.LANCZOS
    t = O.TARGET_A if tgt == "A" else O.TARGET_B
    if fmt == "i420":
        return O.scale_i420(noise_i420(sw, sh, seed), sw, sh, dw, dh, k, t)
    return O.scale_p010(noise_p010(sw, sh, seed), sw, sh, dw, dh, k, t)


@pytest.mark.parametrize("key", sorted(SMALL.files))
def test_oracle_matches_golden_small(oracle, key):
    got = _run(oracle, key)
    assert np.array_equal(got, SMALL[key]), "%s: %d differing samples" % (key, int((got != SMALL[key]).sum()))


@pytest.mark.parametrize("key", sorted(DIGESTS))
def test_oracle_matches_golden_digest(oracle, key):
    assert hashlib.md5(_run(oracle, key).tobytes()).hexdigest() == DIGESTS[key]


def test_survey_fingerprints(oracle):
    # md5 prefixes measured with the real library during the survey
    src = noise_i420(1920, 1080, 1234)
    a = oracle.scale_i420(src, 1920, 1080, 1280, 720, oracle.BICUBIC, oracle.TARGET_A)
    b = oracle.scale_i420(src, 1920, 1080, 1280, 720, oracle.BICUBIC, oracle.TARGET_B)
    assert hashlib.md5(a.tobytes()).hexdigest().startswith("d73124d3c7")
    assert hashlib.md5(b.tobytes()).hexdigest().startswith("850558d0fa")
    d = a.astype(int) - b.astype(int)
    assert np.abs(d).max() == 1 and int((d[:1280 * 720] != 0).sum()) == 76035


def test_same_size_is_identity(oracle):
    src = noise_i420(320, 180, 3)
    assert np.array_equal(oracle.scale_i420(src, 320, 180, 320, 180), src)
    p = noise_p010(320, 180, 3)
    assert np.array_equal(oracle.scale_p010(p, 320, 180, 320, 180), p)


def test_tables_sum_to_one(oracle):
    for (s, d) in [(1920, 1280), (1920, 640), (1080, 234), (1080, 144), (720, 1080), (1920, 854)]:
        for kern in (oracle.BICUBIC, oracle.LANCZOS):
            for align, zero in ((4, 0), (1, 1)):
                taps, pos, coef = oracle.make_filter(s, d, kern, align, 1 << 14, zero)
                assert (coef.astype(int).sum(axis=1) == 1 << 14).all()
                assert pos.min() >= 0 and (pos + taps).max() <= s


def test_tap_counts_match_survey_table(oracle):
    # SURVEY.md Appendix A "Resulting tap counts (H,V)"
    want = {(1920, 1280, 0, "A"): 8, (1080, 720, 0, "A"): 6, (1920, 1280, 0, "B"): 6, (1920, 640, 0, "A"): 12,
            (1080, 360, 0, "B"): 11, (1920, 416, 0, "A"): 20, (1920, 176, 0, "A"): 44, (1080, 144, 0, "A"): 30,
            (1920, 1280, 1, "A"): 12, (1080, 720, 1, "A"): 10, (3840, 640, 1, "A"): 36, (2160, 360, 1, "A"): 34}
    for (s, d, kern, tgt), taps in want.items():
        vertical = s in (1080, 2160)
        align = (2 if vertical else 4) if tgt == "A" else 1
        got, _, _ = oracle.make_filter(s, d, kern, align, (1 << 12) if vertical else (1 << 14), tgt == "B")
        assert got == taps, (s, d, kern, tgt, got, taps)


def test_nv12_is_interleaved_i420(oracle):
    w, h = 64, 36
    i420 = oracle.scale_i420(noise_i420(128, 72, 2), 128, 72, w, h)
    nv = oracle.i420_to_nv12(i420, w, h)
    y, u, v = split_i420(i420, w, h)
    assert np.array_equal(nv[: w * h], y.ravel())
    uv = nv[w * h:].reshape(h // 2, w)
    assert np.array_equal(uv[:, 0::2], u) and np.array_equal(uv[:, 1::2], v)


# ------------------------------------------------------------------ live checks against the bundled library
LIVE = [(1920, 1080, 1280, 720), (1920, 1080, 416, 234), (1280, 720, 640, 360), (720, 480, 320, 240),
        (640, 360, 1920, 1080), (1920, 1080, 1916, 1076), (854, 480, 426, 240)]


@pytest.mark.parametrize("case", LIVE)
@pytest.mark.parametrize("kern", ["bicubic", "lanczos"])
def test_oracle_vs_live_libswscale(oracle, sws, case, kern):
    sw, sh, dw, dh = case
    src = natural_i420(sw, sh, 7)
    planes = [np.ascontiguousarray(p) for p in split_i420(src, sw, sh)]
    kflag = sws.SWS_BICUBIC if kern == "bicubic" else sws.SWS_LANCZOS
    k = oracle.BICUBIC if kern == "bicubic" else oracle.LANCZOS
    for tgt, tflag in ((oracle.TARGET_A, 0), (oracle.TARGET_B, sws.SWS_BITEXACT)):
        ref = np.concatenate([p.ravel() for p in sws.scale(planes, sw, sh, "yuv420p", dw, dh, "yuv420p", kflag | tflag)])
        got = oracle.scale_i420(src, sw, sh, dw, dh, k, tgt)
        assert np.array_equal(got, ref), (case, kern, tgt, int((got != ref).sum()))


def test_oracle_vs_live_libswscale_p010(oracle, sws):
    sw, sh = 1280, 720
    src = noise_p010(sw, sh, 5)
    cw, ch = chroma_dims(sw, sh)
    planes = [np.ascontiguousarray(src[: sw * sh].reshape(sh, sw)), np.ascontiguousarray(src[sw * sh:].reshape(ch, cw * 2))]
    for dw, dh in ((640, 360), (854, 480), (214, 120)):
        for tgt, tflag in ((oracle.TARGET_A, 0), (oracle.TARGET_B, sws.SWS_BITEXACT)):
            ref = np.concatenate([p.ravel() for p in
                                  sws.scale(planes, sw, sh, "p010le", dw, dh, "p010le", sws.SWS_LANCZOS | tflag)])
            got = oracle.scale_p010(src, sw, sh, dw, dh, oracle.LANCZOS, tgt)
            assert np.array_equal(got, ref), (dw, dh, tgt)
