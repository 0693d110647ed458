"""Decode-side converter through the C ABI (ottgpu_converter_*): the bytes must equal what the reference's
`decode_converter` + packing loops produce (transvideo.c:2797-2840), i.e. libswscale's same-size conversion to yuv420p.
Checked against the libswscale fixtures/digests (tests/golden/convert_*) and the oracle restatement (oracle/ref_convert.c)."""
import hashlib, json, os, re
import numpy as np
import pytest
from conftest import decoded_planes, DECODED_FORMATS, interlaced_sequence

HERE = os.path.dirname(os.path.abspath(__file__))
SMALL = np.load(os.path.join(HERE, "golden", "convert_small.npz"))
DIGESTS = json.load(open(os.path.join(HERE, "golden", "convert_digests.json")))["digests"]
KEY = re.compile(r"(yuv4\d\dp(?:10le)?)_([AB])_(\d+)x(\d+)_s(\d+)_(\w+)$")


def parse(key):
    fmt, target, w, h, seed, kind = KEY.match(key).groups()
    return fmt, target, int(w), int(h), int(seed), kind


def pix(mod, fmt):
    return {"yuv420p": mod.PIX_YUV420P, "yuv422p": mod.PIX_YUV422P, "yuv420p10le": mod.PIX_YUV420P10,
            "yuv422p10le": mod.PIX_YUV422P10}[fmt]


@pytest.mark.parametrize("key", sorted(SMALL.files))
def test_matches_libswscale_fixture(lib, og, key):
    fmt, target, w, h, seed, kind = parse(key)
    cv = og.Converter(lib, pix(og, fmt), w, h, og.TARGET_A if target == "A" else og.TARGET_B)
    assert np.array_equal(cv.convert(decoded_planes(fmt, w, h, seed, kind)), SMALL[key])
    cv.close()


@pytest.mark.gpu
@pytest.mark.parametrize("key", sorted(DIGESTS))
def test_matches_libswscale_digest_full_size(real_lib, og, key):
    fmt, target, w, h, seed, kind = parse(key)
    cv = og.Converter(real_lib, pix(og, fmt), w, h, og.TARGET_A if target == "A" else og.TARGET_B)
    got = cv.convert(decoded_planes(fmt, w, h, seed, kind))
    cv.close()
    assert hashlib.md5(got.tobytes()).hexdigest() == DIGESTS[key]


@pytest.mark.parametrize("fmt", sorted(DECODED_FORMATS))
@pytest.mark.parametrize("size", [(192, 108), (64, 48), (66, 38), (130, 74), (24, 10), (328, 46)])
def test_vs_oracle_with_padded_rows(lib, og, oracle, fmt, size):
    """Decoder pictures have linesize > width; odd sizes take the scalar path, 16-aligned ones the vector path."""
    w, h = size
    for kind, pad in (("noise", 0), ("extreme", 32), ("noise", 5)):
        planes = decoded_planes(fmt, w, h, 3 + w + pad, kind, pad=pad)
        cv = og.Converter(lib, pix(og, fmt), w, h)
        assert np.array_equal(cv.convert(planes), oracle.convert_to_i420(planes, pix(oracle, fmt), w, h)), (kind, pad)
        cv.close()


@pytest.mark.parametrize("fmt", ["yuv422p", "yuv422p10le"])
def test_one_converter_many_pictures(lib, og, oracle, fmt):
    """The two staging sets alternate; results must not depend on which set a picture went through."""
    w, h = 96, 64
    cv = og.Converter(lib, pix(og, fmt), w, h)
    for i in range(5):
        planes = decoded_planes(fmt, w, h, 200 + i, "noise", pad=16)
        assert np.array_equal(cv.convert(planes), oracle.convert_to_i420(planes, pix(oracle, fmt), w, h))
    cv.close()


def test_device_result_feeds_a_channel(lib, og, oracle):
    """Zero-copy chain: 4:2:2 10-bit decoder pictures -> converter (device frame) -> channel (yadif + rungs)."""
    w, h = 192, 108
    rungs = [(192, 108, og.FMT_I420), (128, 72, og.FMT_NV12)]
    cv = og.Converter(lib, og.PIX_YUV422P10, w, h)
    ch = og.Channel(lib, og.make_cfg(w, h, rungs, deint=og.DEINT_ALL))
    pool = og.Pool(lib, 0, og.MEM_DEVICE, 6, lib.frame_bytes(og.FMT_I420, w, h))
    ref = oracle.YadifStream(w, h)
    held = []
    for i in range(5):
        planes = decoded_planes("yuv422p10le", w, h, 300 + i, "noise", pad=8)
        frame = oracle.convert_to_i420(planes, oracle.PIX_YUV422P10, w, h)
        buf = pool.take()
        assert buf
        cv.convert(planes, device_dst=buf)
        held.append(buf)
        r = ch.submit(buf, interlaced=1, tff=1, opaque=500 + i)
        want = ref.push(frame, 1, 1, tag=i)
        assert (r is None) == (want is None)
        if want is not None:
            deint = want[0]
            assert np.array_equal(r["rungs"][0], deint)
            assert np.array_equal(r["rungs"][1], oracle.i420_to_nv12(oracle.scale_i420(deint, w, h, 128, 72), 128, 72))
        for b in (r or {}).get("released", []):
            pool.give_back(b)
    ch.close()
    cv.close()
    pool.close()


def test_bad_arguments_are_refused(lib, og):
    with pytest.raises(og.OttGpuError):
        og.Converter(lib, 9, 64, 48)
    with pytest.raises(og.OttGpuError):
        og.Converter(lib, og.PIX_YUV422P, 4, 2)
    cv = og.Converter(lib, og.PIX_YUV422P, 64, 48)
    planes = decoded_planes("yuv422p", 64, 48, 1)
    pic = cv.picture(planes)
    pic.stride[1] = 8                                        # smaller than the 32-byte chroma row
    out = np.empty(cv.out_bytes, np.uint8)
    assert lib.dll.ottgpu_converter_convert(cv.handle, pic, out.ctypes.data, og.MEM_HOST, 0) == -1
    pic = cv.picture(planes)
    pic.plane[2] = None
    assert lib.dll.ottgpu_converter_convert(cv.handle, pic, out.ctypes.data, og.MEM_HOST, 0) == -1
    cv.close()


@pytest.mark.gpu
def test_full_size_1080_422p10_device_picture(real_lib, og, oracle):
    """Picture already in device memory (a hardware decoder's output): planes with 256-byte pitches, result on device."""
    w, h = 1920, 1080
    planes = decoded_planes("yuv422p10le", w, h, 77, "noise")
    want = oracle.convert_to_i420(planes, oracle.PIX_YUV422P10, w, h)
    pitches = [((p.shape[1] * 2 + 255) // 256) * 256 for p in planes]
    pools = [og.Pool(real_lib, 0, og.MEM_DEVICE, 1, pitch * p.shape[0]) for pitch, p in zip(pitches, planes)]
    dev = []
    for pool, pitch, p in zip(pools, pitches, planes):
        padded = np.zeros((p.shape[0], pitch // 2), np.uint16)
        padded[:, :p.shape[1]] = p
        buf = pool.take()
        real_lib.to_device(buf, padded)
        dev.append(buf)
    out_pool = og.Pool(real_lib, 0, og.MEM_DEVICE, 1, real_lib.frame_bytes(og.FMT_I420, w, h))
    dst = out_pool.take()
    cv = og.Converter(real_lib, og.PIX_YUV422P10, w, h)
    cv.convert(dev, device_dst=dst, mem=og.MEM_DEVICE, strides=pitches)
    got = real_lib.from_device(dst, want.size)
    assert np.array_equal(got, want)
    cv.close()
    for p in pools + [out_pool]:
        p.close()


def test_batched_conversion_one_launch(lib, og, oracle):
    """Several channels' pictures (different formats and sizes) converted by ONE launch; twice, so both staging sets of
    every converter are used; results identical to the per-picture path."""
    specs = [("yuv422p10le", 192, 108), ("yuv420p10le", 130, 74), ("yuv422p", 96, 64), ("yuv420p", 66, 38), ("yuv422p10le", 64, 48)]
    cvs = [og.Converter(lib, pix(og, f), w, h) for f, w, h in specs]
    for rnd in range(3):
        planes = [decoded_planes(f, w, h, 700 + 10 * rnd + k, "noise" if (k + rnd) % 2 else "extreme", pad=8 * (k % 3)) for k, (f, w, h) in enumerate(specs)]
        outs = og.convert_batch(lib, cvs, planes)
        for (f, w, h), p, got in zip(specs, planes, outs):
            assert np.array_equal(got, oracle.convert_to_i420(p, pix(oracle, f), w, h)), (f, w, h, rnd)
    # a converter may not appear twice in one batch
    with pytest.raises(og.OttGpuError):
        og.convert_batch(lib, [cvs[0], cvs[0]], [planes[0], planes[0]])
    for c in cvs:
        c.close()
