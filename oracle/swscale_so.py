"""TEST INFRASTRUCTURE ONLY - executable swscale reference (FFmpeg 8.0.1 libswscale.so.9.1.100).

The reference (cannonbeach/ott-packager) calls sws_getContext()/sws_scale() from its un-vendored
FFmpeg fork (/root/reference/source/transvideo.c:1563-1581, 2155-2174; setuptranscode.sh:97).  No FFmpeg
source is under /root/reference, but the opencv_python_headless wheel in this image bundles a real
libswscale (x86 asm enabled).  This module loads it with ctypes so that tests / golden-vector generation /
the CPU baseline can run the REAL library.  Only tests/, bench.py (cpu_baseline / --impl reference) and
__graft_entry__.smoke() may import this file; the product path never does.
"""
import ctypes, glob, os, re, sys
import numpy as np

SWS_BICUBIC = 4
SWS_LANCZOS = 0x200
SWS_ACCURATE_RND = 0x40000
SWS_BITEXACT = 0x80000

_libs = None


def _find_dir():
    import importlib.util
    cands = []
    spec = importlib.util.find_spec("cv2")
    if spec and spec.origin:
        cands.append(os.path.join(os.path.dirname(os.path.dirname(spec.origin)), "opencv_python_headless.libs"))
    for p in sys.path:
        cands.append(os.path.join(p, "opencv_python_headless.libs"))
    for c in cands:
        if glob.glob(os.path.join(c, "libswscale-*.so*")):
            return c
    return None


def available():
    return _find_dir() is not None


def load():
    """Returns (libswscale, libavutil) CDLL handles; raises OSError if the wheel is absent."""
    global _libs
    if _libs is not None:
        return _libs
    d = _find_dir()
    if d is None:
        raise OSError("bundled libswscale not found (opencv_python_headless.libs)")
    # libavutil's dependencies (libdrm, libcrypto, ...) live in the same dir but carry no RPATH: preload them by
    # absolute path (RTLD_GLOBAL), retrying until the set stops shrinking, so no LD_LIBRARY_PATH is needed.
    pending = [f for f in sorted(glob.glob(os.path.join(d, "*.so*")))
               if not re.search(r"lib(avutil|swscale|avcodec|avformat|swresample)-", f)]
    for _ in range(4):
        left = []
        for f in pending:
            try:
                ctypes.CDLL(f, mode=ctypes.RTLD_GLOBAL)
            except OSError:
                left.append(f)
        if len(left) == len(pending):
            break
        pending = left
    avutil = ctypes.CDLL(glob.glob(os.path.join(d, "libavutil-*.so*"))[0], mode=ctypes.RTLD_GLOBAL)
    sws = ctypes.CDLL(glob.glob(os.path.join(d, "libswscale-*.so*"))[0], mode=ctypes.RTLD_GLOBAL)
    sws.sws_getContext.restype = ctypes.c_void_p
    sws.sws_getContext.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                   ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    sws.sws_scale.restype = ctypes.c_int
    sws.sws_scale.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                              ctypes.c_void_p, ctypes.c_void_p]
    sws.sws_freeContext.argtypes = [ctypes.c_void_p]
    sws.sws_freeContext.restype = None
    avutil.av_get_pix_fmt.restype = ctypes.c_int
    avutil.av_get_pix_fmt.argtypes = [ctypes.c_char_p]
    avutil.av_force_cpu_flags.argtypes = [ctypes.c_int]
    avutil.av_force_cpu_flags.restype = None
    avutil.av_get_cpu_flags.restype = ctypes.c_int
    avutil.av_version_info.restype = ctypes.c_char_p
    _libs = (sws, avutil)
    return _libs


def version():
    sws, avu = load()
    return avu.av_version_info().decode(), sws.swscale_version()


def pix_fmt(name):
    _, avu = load()
    v = avu.av_get_pix_fmt(name.encode())
    if v < 0:
        raise ValueError(name)
    return v


class Scaler:
    """One SwsContext, used exactly like video_scale_thread uses it (transvideo.c:1562-1581):
    full-frame sws_scale(ctx, planes, strides, 0, srcH, out, outstrides)."""

    def __init__(self, sw, sh, sfmt, dw, dh, dfmt, flags):
        sws, _ = load()
        self.sws = sws
        self.sw, self.sh, self.dw, self.dh = sw, sh, dw, dh
        self.sfmt, self.dfmt = sfmt, dfmt
        self.ctx = sws.sws_getContext(sw, sh, pix_fmt(sfmt), dw, dh, pix_fmt(dfmt), flags, None, None, None)
        if not self.ctx:
            raise RuntimeError("sws_getContext failed")

    def close(self):
        if self.ctx:
            self.sws.sws_freeContext(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def scale_planes(self, src_planes, dst_planes):
        """src_planes / dst_planes: lists of C-contiguous 2-D numpy arrays (uint8 or uint16), one per plane.
        Strides are taken from the arrays (bytes)."""
        sp = (ctypes.c_void_p * 4)()
        ss = (ctypes.c_int * 4)()
        dp = (ctypes.c_void_p * 4)()
        ds = (ctypes.c_int * 4)()
        for i, a in enumerate(src_planes):
            assert a.flags["C_CONTIGUOUS"]
            sp[i] = a.ctypes.data
            ss[i] = a.strides[0]
        for i, a in enumerate(dst_planes):
            assert a.flags["C_CONTIGUOUS"] and a.flags["WRITEABLE"]
            dp[i] = a.ctypes.data
            ds[i] = a.strides[0]
        r = self.sws.sws_scale(self.ctx, sp, ss, 0, self.sh, dp, ds)
        return r


def _alloc_planes(fmt, w, h):
    cw, ch = (w + 1) // 2, (h + 1) // 2
    if fmt == "yuv420p":
        return [np.zeros((h, w), np.uint8), np.zeros((ch, cw), np.uint8), np.zeros((ch, cw), np.uint8)]
    if fmt == "nv12":
        return [np.zeros((h, w), np.uint8), np.zeros((ch, cw * 2), np.uint8)]
    if fmt == "p010le":
        return [np.zeros((h, w), np.uint16), np.zeros((ch, cw * 2), np.uint16)]
    if fmt == "yuv420p10le":
        return [np.zeros((h, w), np.uint16), np.zeros((ch, cw), np.uint16), np.zeros((ch, cw), np.uint16)]
    raise ValueError(fmt)


def scale(src_planes, sw, sh, sfmt, dw, dh, dfmt, flags):
    """Convenience: one-shot scale, returns list of destination planes."""
    sc = Scaler(sw, sh, sfmt, dw, dh, dfmt, flags)
    out = _alloc_planes(dfmt, dw, dh)
    sc.scale_planes(src_planes, out)
    sc.close()
    return out


def force_cpu_flags(flags):
    _, avu = load()
    avu.av_force_cpu_flags(flags)
