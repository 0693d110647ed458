"""Thumbnail JPEG on the GPU (reference: save_frame_as_jpeg, transvideo.c:164-198 -> libavcodec MJPEG).

 * bit-exact: the parallel GPU formulation == the sequential CPU restatement (oracle/ref_jpeg.c), byte for byte;
 * decodes: libjpeg (cv2.imdecode) accepts the file and returns the picture;
 * pinned to the REAL reference encoder (oracle/avcodec_so.py drives the bundled libavcodec the way the reference does):
   same quantisation table for the qscale its rate control picked, the same quantised levels (every DC, > 99.5 % of the AC
   levels: the quantiser is libavcodec's, the rest is its SIMD DCT), luma PSNR within 0.1 dB, the same segment sequence
   (COM DQT DHT SOF0 SOS, one DHT segment with the picture's four optimal tables) and the same size within 2 %.
   libavcodec's BYTES are not reproduced (see jpeg_core.cuh).
"""
import numpy as np
import pytest
from conftest import natural_i420, noise_i420


def _psnr(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)


def _decode_luma(jpeg):
    import cv2
    img = cv2.imdecode(np.frombuffer(jpeg, np.uint8), cv2.IMREAD_GRAYSCALE)
    assert img is not None, "libjpeg rejected the file"
    return img


@pytest.mark.parametrize("w,h", [(176, 144), (88, 72), (178, 146), (33, 17), (640, 360)])
@pytest.mark.parametrize("qscale", [2, 3, 8])
def test_gpu_bytes_equal_sequential_restatement(lib, og, oracle, w, h, qscale):
    enc = og.Jpeg(lib, w, h, qscale)
    for frame in (natural_i420(w, h, 3), noise_i420(w, h, 9), np.full(w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2), 255, np.uint8)):
        got = enc.encode(frame)
        want = oracle.jpeg_encode(frame, w, h, qscale)
        assert got == want, "JPEG bytes differ (%dx%d q%d: %d vs %d bytes)" % (w, h, qscale, len(got), len(want))
        luma = _decode_luma(got)
        assert luma.shape == (h, w)
    enc.close()


def test_strided_planes_and_small_destination(lib, og, oracle):
    w, h = 176, 144
    enc = og.Jpeg(lib, w, h, 3)
    fr = natural_i420(w, h, 5)
    y = np.zeros((h, 192), np.uint8); y[:, :w] = fr[:w * h].reshape(h, w)
    u = np.zeros((72, 96), np.uint8); u[:, :88] = fr[w * h:w * h + 88 * 72].reshape(72, 88)
    v = np.zeros((72, 96), np.uint8); v[:, :88] = fr[w * h + 88 * 72:].reshape(72, 88)
    assert enc.encode(planes=[y, u, v], strides=[192, 96, 96]) == oracle.jpeg_encode(fr, w, h, 3)
    with pytest.raises(og.OttGpuError):
        enc.encode(fr, capacity=1000)
    enc.close()


def test_device_resident_thumbnail(lib, og, oracle):
    w, h = 176, 144
    fr = natural_i420(w, h, 8)
    pool = og.Pool(lib, 0, og.MEM_DEVICE, 1, fr.size)
    d = pool.take()
    lib.to_device(d, fr)
    enc = og.Jpeg(lib, w, h, 3)
    assert enc.encode(device_ptr=d) == oracle.jpeg_encode(fr, w, h, 3)
    enc.close()
    pool.close()


def test_against_the_real_libavcodec_encoder(lib, og, oracle):
    """What the reference writes for the same thumbnail vs what the GPU writes, both decoded by libjpeg."""
    from oracle import avcodec_so as A
    if not A.available():
        pytest.skip("bundled libavcodec not present")
    w, h = 176, 144                                              # THUMBNAIL_WIDTH x THUMBNAIL_HEIGHT
    for seed in (3, 4, 5):
        fr = natural_i420(w, h, seed, shift=10.0 * seed)
        ref = A.mjpeg_encode(fr, w, h)
        q, tab = A.qscale_of(ref)
        assert q is not None, "libavcodec's DQT is not the scaled MPEG-1 intra matrix"
        enc = og.Jpeg(lib, w, h, q)
        ours = enc.encode(fr)
        enc.close()
        seg_ref, seg_our = dict(A.segments(ref)), dict(A.segments(ours))
        assert seg_our[0xDB] == seg_ref[0xDB]                    # identical quantisation table segment
        assert seg_our[0xC0] == seg_ref[0xC0]                    # identical frame header: 8 bit, size, 3 components 2x2/1x1/1x1, table 0
        # coefficient level: both files entropy-decoded (oracle/jpeg_parse.py).  Every DC level is identical; so are the AC
        # levels except where libavcodec's SIMD forward DCT rounds differently from the integer DCT used here (~0.2 %)
        from oracle import jpeg_parse as P
        cr, co = P.parse(ref)["blocks"], P.parse(ours)["blocks"]
        for comp in (1, 2, 3):
            assert cr[comp].shape == co[comp].shape
            assert np.array_equal(cr[comp][:, 0], co[comp][:, 0]), "DC levels differ from libavcodec's"
            same = np.mean(cr[comp][:, 1:] == co[comp][:, 1:])
            assert same > 0.995 and np.abs(cr[comp] - co[comp]).max() <= 1, (comp, same)
        src = fr[:w * h].reshape(h, w)
        p_ref, p_our = _psnr(_decode_luma(ref), src), _psnr(_decode_luma(ours), src)
        assert abs(p_ref - p_our) < 0.1, (p_ref, p_our)          # same table, same quantiser: the same quality
        # both files carry the picture's own optimal Huffman tables (built by different algorithms from statistics that
        # differ in 0.2 % of the symbols): the same segment sequence and, within 2 %, the same size
        assert [m for m, _ in A.segments(ours)] == [m for m, _ in A.segments(ref)] == [0xFE, 0xDB, 0xC4, 0xC0, 0xDA]
        assert abs(len(ours) - len(ref)) < 0.02 * len(ref) + 16, (len(ours), len(ref))
