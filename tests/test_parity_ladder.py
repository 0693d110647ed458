"""Parity of the ladder scaler (the sws_scale replacement, transvideo.c:1576-1581 / 2169-2174) through the C ABI:
bit-exact against the CPU oracle (itself pinned to libswscale 8.0.1) and against the committed libswscale fixtures.
Every test runs on the CPU emulation of the kernel source (tier `emu`) and on the B200 (tier `gpu`)."""
import hashlib, json, os, re
import numpy as np
import pytest
from conftest import noise_i420, noise_p010, natural_i420, split_i420, chroma_dims

G = os.path.join(os.path.dirname(__file__), "golden")
SMALL = np.load(os.path.join(G, "swscale_small.npz"))
DIGESTS = json.load(open(os.path.join(G, "swscale_digests.json")))["digests"]
KEY = re.compile(r"(i420|p010)_(bicubic|lanczos)_([AB])_(\d+)x(\d+)_(\d+)x(\d+)_s(\d+)")


def _run_key(lib, og, key):
    fmt, kern, tgt, sw, sh, dw, dh, seed = KEY.fullmatch(key).groups()
    sw, sh, dw, dh, seed = map(int, (sw, sh, dw, dh, seed))
    filt = og.FILTER_BICUBIC if kern == "bicubic" else og.FILTER_LANCZOS
    target = og.TARGET_A if tgt == "A" else og.TARGET_B
    if fmt == "i420":
        return lib.scale_frame(noise_i420(sw, sh, seed), og.FMT_I420, sw, sh, dw, dh, filt=filt, target=target)
    return lib.scale_frame(noise_p010(sw, sh, seed), og.FMT_P010, sw, sh, dw, dh, filt=filt, target=target)


@pytest.mark.parametrize("key", sorted(SMALL.files))
def test_matches_libswscale_fixture(lib, og, key):
    got = _run_key(lib, og, key)
    assert np.array_equal(got, SMALL[key]), "%s: %d differing samples" % (key, int((got != SMALL[key]).sum()))


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(DIGESTS))
def test_matches_libswscale_digest_full_size(real_lib, og, key):
    # BASELINE full sizes (1080p noise seed 1234, 2160p P010 seed 99): md5 of the real library's output
    assert hashlib.md5(_run_key(real_lib, og, key).tobytes()).hexdigest() == DIGESTS[key]


CASES = [  # sw, sh, dw, dh : ragged / odd / upscale / one-axis / near-unity
    (192, 108, 128, 72), (192, 108, 96, 54), (192, 108, 64, 36), (854, 480, 426, 240), (427, 241, 213, 120),
    (320, 180, 480, 270), (320, 180, 320, 120), (320, 180, 212, 180), (320, 180, 318, 178), (640, 360, 58, 46),
    (130, 74, 66, 38),
]


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("filt", ["bicubic", "lanczos"])
def test_i420_vs_oracle(lib, og, oracle, case, filt):
    sw, sh, dw, dh = case
    src = natural_i420(sw, sh, 7) if (sw + dw) % 3 else noise_i420(sw, sh, 21)
    f, k = (og.FILTER_BICUBIC, oracle.BICUBIC) if filt == "bicubic" else (og.FILTER_LANCZOS, oracle.LANCZOS)
    for target in (og.TARGET_A, og.TARGET_B):
        ref = oracle.scale_i420(src, sw, sh, dw, dh, k, target)
        got = lib.scale_frame(src, og.FMT_I420, sw, sh, dw, dh, filt=f, target=target)
        assert np.array_equal(got, ref), (case, filt, target, int((got != ref).sum()))


@pytest.mark.parametrize("case", [(192, 108, 128, 72), (854, 480, 426, 240), (320, 180, 320, 180), (130, 74, 66, 38)])
def test_nv12_is_interleaved_target_a_i420(lib, og, oracle, case):
    # transvideo.c:543-557: the reference's NV12 = byte interleave of the I420 result
    sw, sh, dw, dh = case
    src = noise_i420(sw, sh, 33)
    ref = oracle.i420_to_nv12(oracle.scale_i420(src, sw, sh, dw, dh), dw, dh)
    got = lib.scale_frame(src, og.FMT_I420, sw, sh, dw, dh, dst_fmt=og.FMT_NV12)
    assert np.array_equal(got, ref)


def test_pitch_aligned_surfaces(lib, og, oracle):
    sw, sh, dw, dh = 320, 180, 208, 118
    src = noise_i420(sw, sh, 34)
    ref = oracle.i420_to_nv12(oracle.scale_i420(src, sw, sh, dw, dh), dw, dh)
    got = lib.scale_frame(src, og.FMT_I420, sw, sh, dw, dh, dst_fmt=og.FMT_NV12, pitch_align=256)
    (oy, py), (ouv, puv) = lib.rung_layout(og.RungCfg(dw, dh, og.FMT_NV12, 256))
    assert py == 256 and puv == 256
    y = got[oy: oy + py * dh].reshape(dh, py)[:, :dw]
    uv = got[ouv: ouv + puv * (dh // 2)].reshape(dh // 2, puv)[:, :dw]
    assert np.array_equal(y.ravel(), ref[: dw * dh]) and np.array_equal(uv.ravel(), ref[dw * dh:])


@pytest.mark.parametrize("case", [(256, 144, 128, 72), (256, 144, 170, 96), (256, 144, 86, 48), (128, 72, 256, 144),
                                  (256, 144, 256, 144), (258, 146, 130, 74)])
def test_p010_vs_oracle(lib, og, oracle, case):
    sw, sh, dw, dh = case
    src = noise_p010(sw, sh, 40)
    for f, k in ((og.FILTER_LANCZOS, oracle.LANCZOS), (og.FILTER_BICUBIC, oracle.BICUBIC)):
        for target in (og.TARGET_A, og.TARGET_B):
            ref = oracle.scale_p010(src, sw, sh, dw, dh, k, target)
            got = lib.scale_frame(src, og.FMT_P010, sw, sh, dw, dh, filt=f, target=target)
            assert np.array_equal(got, ref), (case, f, target, int((got != ref).sum()))


def test_extreme_values(lib, og, oracle):
    # checkerboard 0/255, all-0, all-255: exercises the 15-bit clip of the H pass and the 16-bit wraparound of target A
    sw, sh, dw, dh = 192, 108, 128, 72
    cw, ch = chroma_dims(sw, sh)
    yy, xx = np.mgrid[0:sh, 0:sw]
    checker = (((yy + xx) & 1) * 255).astype(np.uint8)
    for plane in (checker, np.zeros((sh, sw), np.uint8), np.full((sh, sw), 255, np.uint8), (((yy // 3 + xx // 5) & 1) * 255).astype(np.uint8)):
        src = np.concatenate([plane.ravel(), plane[:ch, :cw].ravel(), plane[:ch, :cw][::-1].ravel()])
        for target in (og.TARGET_A, og.TARGET_B):
            for f, k in ((og.FILTER_BICUBIC, oracle.BICUBIC), (og.FILTER_LANCZOS, oracle.LANCZOS)):
                ref = oracle.scale_i420(src, sw, sh, dw, dh, k, target)
                got = lib.scale_frame(src, og.FMT_I420, sw, sh, dw, dh, filt=f, target=target)
                assert np.array_equal(got, ref)


@pytest.mark.parametrize("size", [(960, 540), (1920, 1080)])
def test_thumbnail_size_filter(lib, og, oracle, size):
    # the 176x144 thumbnail (transvideo.c:2155-2174): long-filter code path (44/30 taps) and the tile-size fallback
    sw, sh = size
    src = noise_i420(sw, sh, 41)
    ref = oracle.scale_i420(src, sw, sh, 176, 144)
    got = lib.scale_frame(src, og.FMT_I420, sw, sh, 176, 144)
    assert np.array_equal(got, ref)


def test_full_size_planning_of_baseline_ladders(lib, og):
    # every BASELINE.json ladder must plan (tiles fit shared memory) -- channel creation is where that is decided
    nv = og.FMT_NV12
    og.Channel(lib, og.make_cfg(1920, 1080, [(1920, 1080, nv), (1280, 720, nv), (960, 540, nv), (640, 360, nv), (416, 234, nv)],
                                deint=og.DEINT_FLAGGED, thumb=(176, 144, 150))).close()
    p = og.FMT_P010
    og.Channel(lib, og.make_cfg(3840, 2160, [(3840, 2160, p), (2560, 1440, p), (1920, 1080, p), (1280, 720, p), (960, 540, p), (640, 360, p)],
                                src_fmt=p, filt=og.FILTER_LANCZOS, deint=og.DEINT_FLAGGED, thumb=(176, 144, 150))).close()
    og.Channel(lib, og.make_cfg(3840, 2160, [(3840, 2160, 0), (1920, 1080, 0), (1280, 720, 0), (640, 360, 0)], thumb=(176, 144, 150))).close()


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["cfg1", "cfg2", "cfg3", "thumb"])
def test_baseline_full_size_ladders(real_lib, og, oracle, cfg):
    """BASELINE.json configs at full size, every rung against the oracle."""
    if cfg == "cfg3":
        sw, sh = 3840, 2160
        src = noise_p010(sw, sh, 99)
        rungs = [(3840, 2160), (2560, 1440), (1920, 1080), (1280, 720), (960, 540), (640, 360)]
        for dw, dh in rungs:
            ref = oracle.scale_p010(src, sw, sh, dw, dh, oracle.LANCZOS, oracle.TARGET_A)
            got = real_lib.scale_frame(src, og.FMT_P010, sw, sh, dw, dh, filt=og.FILTER_LANCZOS)
            assert np.array_equal(got, ref), (dw, dh)
        return
    sw, sh = 1920, 1080
    src = natural_i420(sw, sh, 7)
    rungs = {"cfg1": [(1920, 1080), (1280, 720), (640, 360)],
             "cfg2": [(1920, 1080), (1280, 720), (960, 540), (640, 360), (416, 234)], "thumb": [(176, 144)]}[cfg]
    for dw, dh in rungs:
        i420 = oracle.scale_i420(src, sw, sh, dw, dh)
        if cfg == "cfg2":
            got = real_lib.scale_frame(src, og.FMT_I420, sw, sh, dw, dh, dst_fmt=og.FMT_NV12)
            assert np.array_equal(got, oracle.i420_to_nv12(i420, dw, dh)), (dw, dh)
        else:
            assert np.array_equal(real_lib.scale_frame(src, og.FMT_I420, sw, sh, dw, dh), i420), (dw, dh)


# ------------------------------------------------------------------------------------- size-independent properties
def _const_i420(w, h, y, u, v):
    cw, ch = (w + 1) // 2, (h + 1) // 2
    return np.concatenate([np.full(w * h, y, np.uint8), np.full(cw * ch, u, np.uint8), np.full(cw * ch, v, np.uint8)])


@pytest.mark.parametrize("levels", [(16, 128, 128), (235, 16, 240), (0, 0, 0), (255, 255, 255), (81, 90, 240)])
def test_flat_frame_stays_flat_small(lib, og, oracle, levels):
    """A flat frame scales to a flat frame whose level does not depend on the frame size: the small case is checked
    against the oracle here and gives the expected level for the full-size run below."""
    src = _const_i420(192, 108, *levels)
    got = lib.scale_frame(src, og.FMT_I420, 192, 108, 128, 72)
    assert np.array_equal(got, oracle.scale_i420(src, 192, 108, 128, 72))


@pytest.mark.gpu
@pytest.mark.parametrize("levels", [(16, 128, 128), (235, 16, 240), (255, 255, 255), (81, 90, 240)])
def test_flat_frame_stays_flat_full_size(real_lib, og, oracle, levels):
    """BASELINE.json cfg2 ladder at full size on a flat frame: every rung is flat, and its level is the one the oracle
    produces for the same scale ratio on a frame 1/10th the size (the polyphase phases repeat, so the level is the same)."""
    sw, sh = 1920, 1080
    src = _const_i420(sw, sh, *levels)
    for dw, dh in [(1280, 720), (960, 540), (640, 360)]:
        got = real_lib.scale_frame(src, og.FMT_I420, sw, sh, dw, dh)
        small = oracle.scale_i420(_const_i420(sw // 10, sh // 10, *levels), sw // 10, sh // 10, dw // 10, dh // 10)
        y, u, v = split_i420(got, dw, dh)
        ys, us, vs = split_i420(small, dw // 10, dh // 10)
        for plane, ref in ((y, ys), (u, us), (v, vs)):
            # interior rows/columns (the last two luma rows / last chroma row take libswscale's exact vertical path)
            assert np.unique(plane[4:-4, 4:-4]).tolist() == np.unique(ref[4:-4, 4:-4]).tolist(), (dw, dh)
            assert np.array_equal(plane[-4:, 8:16], np.broadcast_to(ref[-4:, 8:9], (4, 8)))


@pytest.mark.gpu
def test_full_size_crop_independence(real_lib, og, oracle):
    """Tile seams must not show: the top-left region of a full-size 2:1 downscale equals the oracle's result on the
    matching top-left crop of the source (outputs only see a finite window of source rows/columns)."""
    sw, sh = 1920, 1080
    src = natural_i420(sw, sh, 31)
    got = real_lib.scale_frame(src, og.FMT_I420, sw, sh, 960, 540)
    y, u, v = split_i420(src, sw, sh)
    cw, chh = 512, 288                                     # crop of the source, same 2:1 ratio -> 256x144 output
    crop = np.concatenate([y[:chh, :cw].ravel(), u[:chh // 2, :cw // 2].ravel(), v[:chh // 2, :cw // 2].ravel()])
    ref = oracle.scale_i420(crop, cw, chh, cw // 2, chh // 2)
    gy, gu, gv = split_i420(got, 960, 540)
    ry, ru, rv = split_i420(ref, cw // 2, chh // 2)
    m = 8                                                  # keep clear of the crop's right/bottom border handling
    assert np.array_equal(gy[:chh // 2 - m, :cw // 2 - m], ry[:-m, :-m])
    assert np.array_equal(gu[:chh // 4 - m, :cw // 4 - m], ru[:-m, :-m])
    assert np.array_equal(gv[:chh // 4 - m, :cw // 4 - m], rv[:-m, :-m])


@pytest.mark.gpu
def test_full_size_runs_are_deterministic(real_lib, og):
    """The dynamic tile scheduler hands tiles to CTAs in a different order every launch; the bytes must not depend on it."""
    import hashlib
    sw, sh = 1920, 1080
    nv = og.FMT_NV12
    rungs = [(1920, 1080, nv), (1280, 720, nv), (960, 540, nv), (640, 360, nv), (416, 234, nv)]
    src = natural_i420(sw, sh, 5)
    digests = set()
    for _ in range(4):
        ch = og.Channel(real_lib, og.make_cfg(sw, sh, rungs, deint=og.DEINT_BYPASS))
        r = ch.submit(src)
        digests.add(hashlib.md5(b"".join(np.ascontiguousarray(x).tobytes() for x in r["rungs"])).hexdigest())
        ch.close()
    assert len(digests) == 1


# ------------------------------------------------------------------------------------- randomised geometry sweep
def _random_cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        sw, sh = int(rng.integers(8, 120)) * 2, int(rng.integers(6, 80)) * 2
        ratio_w, ratio_h = rng.uniform(0.2, 2.2), rng.uniform(0.2, 2.2)
        dw, dh = max(8, int(sw * ratio_w)) & ~1, max(8, int(sh * ratio_h)) & ~1
        out.append((sw, sh, dw, dh, int(rng.integers(0, 2)), int(rng.integers(0, 2)), int(rng.integers(0, 2))))
    return out


@pytest.mark.parametrize("case", _random_cases(48, 2024))
def test_random_geometry_vs_oracle(lib, og, oracle, case):
    """Seeded random source/destination sizes (down- and up-scaling, independent ratios per axis), both filters, both
    targets, I420 and NV12 writers: planner corner cases (narrow last tiles, 1-tap axes, long filters) against the oracle."""
    sw, sh, dw, dh, lanczos, target_b, nv12 = case
    src = noise_i420(sw, sh, sw * 7 + dh)
    f, k = (og.FILTER_LANCZOS, oracle.LANCZOS) if lanczos else (og.FILTER_BICUBIC, oracle.BICUBIC)
    t, kt = (og.TARGET_B, oracle.TARGET_B) if target_b else (og.TARGET_A, oracle.TARGET_A)
    ref = oracle.scale_i420(src, sw, sh, dw, dh, k, kt)
    if nv12:
        ref = oracle.i420_to_nv12(ref, dw, dh)
    got = lib.scale_frame(src, og.FMT_I420, sw, sh, dw, dh, dst_fmt=og.FMT_NV12 if nv12 else og.FMT_I420, filt=f, target=t)
    assert np.array_equal(got, ref), (case, int((got != ref).sum()))


# ---- round 3: the target-A vertical pass as round-down FP32 FMAs (vrow_fma8) and its integer fall-back -------------
def _stripes_i420(w, h, period, lo=0, hi=255, vertical=True):
    """hard black / white stripes: the bicubic / lanczos overshoot next to them drives H-pass intermediates below zero"""
    cw, ch = chroma_dims(w, h)
    def plane(pw, ph):
        idx = (np.arange(pw)[None, :] // period) if vertical else (np.arange(ph)[:, None] // period)
        return np.broadcast_to(np.where(idx & 1, hi, lo), (ph, pw)).astype(np.uint8)
    return np.concatenate([plane(w, h).ravel(), plane(cw, ch).ravel(), plane(cw, ch).ravel()])


@pytest.mark.parametrize("case", [(640, 360, 426, 240), (640, 360, 320, 180), (640, 360, 212, 120), (1280, 720, 278, 156)])
@pytest.mark.parametrize("content", ["video-range", "stripes", "mixed", "noise"])
def test_vertical_fma_path_and_fallback(lib, og, oracle, case, content):
    """Tiles whose intermediates are all >= 0 take the FMA path, tiles with a negative one the integer path: both must be
    the reference's bytes, whatever the content (the decision is per tile, so `mixed` exercises both in one picture)."""
    sw, sh, dw, dh = case
    if content == "video-range":
        src = np.clip(natural_i420(sw, sh, 5), 16, 235).astype(np.uint8)
    elif content == "stripes":
        src = _stripes_i420(sw, sh, 3)
    elif content == "mixed":
        src = np.clip(natural_i420(sw, sh, 6), 16, 235).astype(np.uint8)
        y = src[: sw * sh].reshape(sh, sw)
        y[sh // 3: sh // 2, sw // 4: sw // 2] = _stripes_i420(sw, sh, 2)[: sw * sh].reshape(sh, sw)[sh // 3: sh // 2, sw // 4: sw // 2]
    else:
        src = noise_i420(sw, sh, 77)
    for filt, k in ((og.FILTER_BICUBIC, oracle.BICUBIC), (og.FILTER_LANCZOS, oracle.LANCZOS)):
        ref = oracle.scale_i420(src, sw, sh, dw, dh, k, og.TARGET_A)
        got = lib.scale_frame(src, og.FMT_I420, sw, sh, dw, dh, filt=filt, target=og.TARGET_A)
        assert np.array_equal(got, ref), (case, content, int((got != ref).sum()))


def test_vertical_fma_path_is_taken(emu_lib, og):
    """On the emulation the kernel source counts which path its fast rows took: video-range content must run on FMAs only,
    hard 0/255 stripes must send tiles to the integer path (otherwise the sign flag of the H pass is broken)."""
    import ctypes
    emu_lib.dll.ottgpu_emu_vrows.argtypes = [ctypes.POINTER(ctypes.c_longlong)]
    out = (ctypes.c_longlong * 2)()
    sw, sh, dw, dh = 640, 360, 426, 240
    emu_lib.dll.ottgpu_emu_vrows(out)
    emu_lib.scale_frame(np.clip(natural_i420(sw, sh, 5), 16, 235).astype(np.uint8), og.FMT_I420, sw, sh, dw, dh)
    emu_lib.dll.ottgpu_emu_vrows(out)
    assert out[0] > 0 and out[1] == 0, (out[0], out[1])
    emu_lib.scale_frame(_stripes_i420(sw, sh, 3), og.FMT_I420, sw, sh, dw, dh)
    emu_lib.dll.ottgpu_emu_vrows(out)
    assert out[1] > 0, (out[0], out[1])
