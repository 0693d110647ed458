"""TEST INFRASTRUCTURE ONLY - ctypes front-end of oracle/liboracle.so (ref_swscale.c + ref_yadif.c).

Only tests/, bench.py (cpu_baseline / --impl reference) and __graft_entry__.smoke() may import this module.
"""
import ctypes, os, subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
BICUBIC, LANCZOS = 0, 1
TARGET_A, TARGET_B = 0, 1
_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE])


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.ora_filter_make.restype = vp
        L.ora_filter_make.argtypes = [ci] * 6
        L.ora_filter_free.argtypes = [vp]
        L.ora_filter_taps.argtypes = [vp]
        L.ora_filter_n.argtypes = [vp]
        L.ora_filter_pos.restype = ctypes.POINTER(ctypes.c_int32)
        L.ora_filter_pos.argtypes = [vp]
        L.ora_filter_coef.restype = ctypes.POINTER(ctypes.c_int16)
        L.ora_filter_coef.argtypes = [vp]
        L.ora_scale_plane_u8.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci, ci, ci, ci]
        L.ora_scale_plane_u10.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci, ci, ci]
        L.ora_scale_i420.argtypes = [vp, ci, ci, vp, ci, ci, ci, ci]
        L.ora_i420_to_nv12.argtypes = [vp, ci, ci, vp]
        L.ora_scale_p010.argtypes = [vp, ci, ci, vp, ci, ci, ci, ci]
        L.ora_convert_to_i420.argtypes = [vp, vp, ci, ci, ci, vp, ci]
        L.ora_convert_to_i420.restype = ci
        L.ora_yadif_plane.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, ci]
        L.ora_yadif_frame_420.argtypes = [vp, vp, vp, vp, ci, ci, ci, ci, ci]
        L.ora_jpeg_encode.argtypes = [vp, ci, vp, vp, ci, ci, ci, ci, vp, ctypes.c_size_t]
        L.ora_jpeg_encode.restype = ctypes.c_size_t
        for f in ("ora_scale_plane_u8", "ora_scale_plane_u10", "ora_scale_i420", "ora_i420_to_nv12",
                  "ora_scale_p010", "ora_yadif_plane", "ora_yadif_frame_420", "ora_filter_free"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def make_filter(src_n, dst_n, kernel, align, one, zero_tail):
    """-> (taps, pos[int32 n], coef[int16 n x taps])"""
    L = lib()
    f = L.ora_filter_make(src_n, dst_n, kernel, align, one, zero_tail)
    taps, n = L.ora_filter_taps(f), L.ora_filter_n(f)
    pos = np.ctypeslib.as_array(L.ora_filter_pos(f), (n,)).copy()
    coef = np.ctypeslib.as_array(L.ora_filter_coef(f), (n * taps,)).copy().reshape(n, taps)
    L.ora_filter_free(f)
    return taps, pos, coef


def i420_size(w, h):
    return w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)


def scale_i420(src, sw, sh, dw, dh, kernel=BICUBIC, target=TARGET_A):
    src = np.ascontiguousarray(src, np.uint8).ravel()
    assert src.size == i420_size(sw, sh)
    dst = np.empty(i420_size(dw, dh), np.uint8)
    lib().ora_scale_i420(src.ctypes.data, sw, sh, dst.ctypes.data, dw, dh, kernel, target)
    return dst


def i420_to_nv12(i420, w, h):
    i420 = np.ascontiguousarray(i420, np.uint8).ravel()
    out = np.empty_like(i420)
    lib().ora_i420_to_nv12(i420.ctypes.data, w, h, out.ctypes.data)
    return out


def scale_p010(src, sw, sh, dw, dh, kernel=LANCZOS, target=TARGET_A):
    src = np.ascontiguousarray(src, np.uint16).ravel()
    assert src.size == i420_size(sw, sh)
    dst = np.empty(i420_size(dw, dh), np.uint16)
    lib().ora_scale_p010(src.ctypes.data, sw, sh, dst.ctypes.data, dw, dh, kernel, target)
    return dst


PIX_YUV420P, PIX_YUV422P, PIX_YUV420P10, PIX_YUV422P10 = 0, 1, 2, 3


def convert_to_i420(planes, fmt, w, h, target=TARGET_A):
    """planes: three 2-D numpy arrays as a decoder hands them over (uint8 or uint16, any row stride) -> packed I420."""
    L = lib()
    pp = (ctypes.c_void_p * 3)(*[p.ctypes.data for p in planes])
    ss = (ctypes.c_int * 3)(*[p.strides[0] for p in planes])
    out = np.empty(i420_size(w, h), np.uint8)
    if L.ora_convert_to_i420(pp, ss, fmt, w, h, out.ctypes.data, target) != 0:
        raise ValueError("unknown pixel format %r" % (fmt,))
    return out


def yadif_frame(prev, cur, nxt, w, h, interlaced=1, tff=1, bps=1):
    dt = np.uint8 if bps == 1 else np.uint16
    prev, cur, nxt = (np.ascontiguousarray(a, dt).ravel() for a in (prev, cur, nxt))
    assert cur.size == i420_size(w, h)
    dst = np.empty_like(cur)
    lib().ora_yadif_frame_420(prev.ctypes.data, cur.ctypes.data, nxt.ctypes.data, dst.ctypes.data,
                              w, h, interlaced, tff, bps)
    return dst


class YadifStream:
    """Frame scheduling of yadif mode 0 (yadif_common.c ff_yadif_filter_frame): one-frame latency, first call emits
    nothing, frame 0 is filtered with prev = itself.  deint_flagged_only mirrors "yadif=0:-1:1"
    (transvideo.c:2036-2044): frames not flagged interlaced are passed through untouched (still delayed by one)."""

    def __init__(self, w, h, bps=1, deint_flagged_only=False):
        self.w, self.h, self.bps, self.flagged = w, h, bps, deint_flagged_only
        self.prev = self.cur = self.next = None

    def reset(self):
        self.prev = self.cur = self.next = None

    def push(self, frame, interlaced=1, tff=1, tag=None):
        """-> None or (output_frame, tag_of_that_frame)"""
        self.prev, self.cur, self.next = self.cur, self.next, (np.array(frame, copy=True), interlaced, tff, tag)
        if self.cur is None:
            self.cur = self.next
        if self.prev is None:
            return None
        f, il, tf, tg = self.cur
        if self.flagged and not il:
            return f.copy(), tg
        return yadif_frame(self.prev[0], f, self.next[0], self.w, self.h, il, tf, self.bps), tg


def jpeg_encode(i420, w, h, qscale=3):
    """packed I420 -> baseline JPEG bytes (ref_jpeg.c: the sequential restatement of the thumbnail encoder)"""
    L = lib()
    buf = np.ascontiguousarray(i420, np.uint8)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    out = np.zeros(w * h * 4 + 4096, np.uint8)
    base = buf.ctypes.data
    n = L.ora_jpeg_encode(base, w, base + w * h, base + w * h + cw * ch, cw, w, h, qscale, out.ctypes.data, out.size)
    assert n <= out.size
    return out[:n].tobytes()
