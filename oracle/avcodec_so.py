"""TEST INFRASTRUCTURE ONLY - the real libavcodec MJPEG encoder (FFmpeg 8.0.1, libavcodec.so.62.11.100 bundled with the
opencv_python_headless wheel), driven exactly the way the reference's save_frame_as_jpeg drives it
(/root/reference/source/transvideo.c:164-198: bit_rate 250000, time_base 1001/30000, AV_PIX_FMT_YUVJ420P, one fresh context
per picture, one frame in, one packet out).  Used by tests/test_parity_jpeg.py to pin what CAN be pinned of the thumbnail
JPEG: the file structure, the quantisation table and the decoded quality.  The product never imports this.
"""
import ctypes as C, glob, os
import numpy as np
from . import swscale_so as S

_avc = None
AV_CODEC_ID_MJPEG = 7
AV_PIX_FMT_YUVJ420P = 12


class _FrameHead(C.Structure):           # the public head of AVFrame (stable across FFmpeg 4..8)
    _fields_ = [("data", C.c_void_p * 8), ("linesize", C.c_int * 8), ("extended_data", C.c_void_p),
                ("width", C.c_int), ("height", C.c_int), ("nb_samples", C.c_int), ("format", C.c_int)]


class _PacketHead(C.Structure):
    _fields_ = [("buf", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64), ("data", C.c_void_p), ("size", C.c_int)]


def available():
    d = S._find_dir()
    return bool(d and glob.glob(os.path.join(d, "libavcodec-*.so*")))


def load():
    global _avc
    if _avc is None:
        _, avutil = S.load()
        d = S._find_dir()
        C.CDLL(glob.glob(os.path.join(d, "libswresample-*.so*"))[0], mode=C.RTLD_GLOBAL)
        avc = C.CDLL(glob.glob(os.path.join(d, "libavcodec-*.so*"))[0], mode=C.RTLD_GLOBAL)
        vp = C.c_void_p
        avc.avcodec_find_encoder.restype, avc.avcodec_find_encoder.argtypes = vp, [C.c_int]
        avc.avcodec_alloc_context3.restype, avc.avcodec_alloc_context3.argtypes = vp, [vp]
        avc.avcodec_open2.restype, avc.avcodec_open2.argtypes = C.c_int, [vp, vp, vp]
        avc.avcodec_send_frame.restype, avc.avcodec_send_frame.argtypes = C.c_int, [vp, vp]
        avc.avcodec_receive_packet.restype, avc.avcodec_receive_packet.argtypes = C.c_int, [vp, vp]
        avc.av_packet_alloc.restype = vp
        avc.av_packet_free.argtypes = [C.POINTER(vp)]
        avc.avcodec_free_context.argtypes = [C.POINTER(vp)]
        avutil.av_opt_set.restype, avutil.av_opt_set.argtypes = C.c_int, [vp, C.c_char_p, C.c_char_p, C.c_int]
        avutil.av_frame_alloc.restype = vp
        avutil.av_frame_free.argtypes = [C.POINTER(vp)]
        _avc = (avc, avutil)
    return _avc


def mjpeg_encode(i420, w, h):
    """packed I420 (taken as YUVJ420P like the reference does) -> the bytes save_frame_as_jpeg would write"""
    avc, avutil = load()
    enc = avc.avcodec_find_encoder(AV_CODEC_ID_MJPEG)
    if not enc:
        raise OSError("the bundled libavcodec has no MJPEG encoder")
    ctx = avc.avcodec_alloc_context3(enc)
    for k, v in ((b"video_size", b"%dx%d" % (w, h)), (b"pixel_format", b"yuvj420p"), (b"time_base", b"1001/30000"), (b"b", b"250000")):
        if avutil.av_opt_set(ctx, k, v, 0) != 0:
            raise OSError("av_opt_set(%s) failed" % k.decode())
    if avc.avcodec_open2(ctx, enc, None) != 0:
        raise OSError("avcodec_open2(mjpeg) failed")
    buf = np.ascontiguousarray(i420, np.uint8)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    fr = avutil.av_frame_alloc()
    f = _FrameHead.from_address(fr)
    f.data[0], f.data[1], f.data[2] = buf.ctypes.data, buf.ctypes.data + w * h, buf.ctypes.data + w * h + cw * ch
    f.linesize[0], f.linesize[1], f.linesize[2] = w, cw, cw
    f.width, f.height, f.format = w, h, AV_PIX_FMT_YUVJ420P
    pkt = avc.av_packet_alloc()
    try:
        if avc.avcodec_send_frame(ctx, fr) != 0 or avc.avcodec_receive_packet(ctx, pkt) != 0:
            raise OSError("libavcodec did not return a packet")
        p = _PacketHead.from_address(pkt)
        return C.string_at(p.data, p.size)
    finally:
        f.data[0] = f.data[1] = f.data[2] = None              # the planes are ours
        a, b, c = C.c_void_p(pkt), C.c_void_p(fr), C.c_void_p(ctx)
        avc.av_packet_free(C.byref(a))
        avutil.av_frame_free(C.byref(b))
        avc.avcodec_free_context(C.byref(c))


def segments(jpeg):
    """[(marker, payload bytes)] up to and including SOS"""
    out, i = [], 2
    while i + 4 <= len(jpeg) and jpeg[i] == 0xFF:
        m, n = jpeg[i + 1], (jpeg[i + 2] << 8) | jpeg[i + 3]
        out.append((m, jpeg[i + 4:i + 2 + n]))
        i += 2 + n
        if m == 0xDA:
            break
    return out


def qscale_of(jpeg):
    """the qscale libavcodec's rate control picked: the DQT is (ff_mpeg1_default_intra_matrix * qscale) >> 3"""
    dqt = [p for m, p in segments(jpeg) if m == 0xDB][0]
    tab = list(dqt[1:65])
    for q in range(1, 32):
        if tab[1] == max(1, (16 * q) >> 3) and tab[3] == max(1, (19 * q) >> 3) and tab[63] == min(255, (83 * q) >> 3):
            return q, tab
    return None, tab
