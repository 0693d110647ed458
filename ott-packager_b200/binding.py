"""ctypes binding of include/ottgpu.h (libottgpu.so).

Host-side mirror, in Python, of the calls a maintainer makes from transvideo.c: a Channel stands
for the avfilter graph + SwsContexts of one fillet process (transvideo.c:1562-1569, 2003-2054), Batch advances many
channels with one fused launch per stage, Pool follows the mempool.h take/return discipline.

There is no CPU fallback: loading fails loudly if libottgpu.so has not been built, and every compute call fails with
OTTGPU_ERR_CUDA when no sm_100 device is present.
"""
import ctypes as C
import os
import numpy as np

MAX_RUNGS = 8
FMT_I420, FMT_NV12, FMT_P010, FMT_I420P10 = 0, 1, 2, 3
FILTER_BICUBIC, FILTER_LANCZOS = 0, 1
TARGET_A, TARGET_B = 0, 1
DEINT_ALL, DEINT_FLAGGED, DEINT_BYPASS = 0, 1, 2
MEM_HOST, MEM_DEVICE, MEM_HOST_WC = 0, 1, 2
SUBMIT_SYNC, SUBMIT_ASYNC = 0, 1
OK, ERR_ARG, ERR_CUDA, ERR_NOMEM, ERR_STATE = 0, -1, -2, -3, -4
PIX_YUV420P, PIX_YUV422P, PIX_YUV420P10, PIX_YUV422P10 = 0, 1, 2, 3

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.environ.get("OTTGPU_LIB") or os.path.join(HERE, "libottgpu.so")     # OTTGPU_LIB: A/B builds in profiling sessions


class RungCfg(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fmt", C.c_int), ("pitch_align", C.c_int)]


class Cfg(C.Structure):
    _fields_ = [("gpu", C.c_int), ("src_width", C.c_int), ("src_height", C.c_int), ("src_fmt", C.c_int),
                ("n_rungs", C.c_int), ("rung", RungCfg * MAX_RUNGS), ("filter", C.c_int), ("target", C.c_int),
                ("deint", C.c_int), ("thumb_width", C.c_int), ("thumb_height", C.c_int), ("thumb_period", C.c_int),
                ("thumb_phase", C.c_int), ("stream", C.c_void_p)]


class FrameIn(C.Structure):
    _fields_ = [("data", C.c_void_p), ("mem", C.c_int), ("interlaced", C.c_int), ("tff", C.c_int),
                ("pts", C.c_int64), ("dts", C.c_int64), ("opaque", C.c_void_p)]


class FrameOut(C.Structure):
    _fields_ = [("dst", C.c_void_p * MAX_RUNGS), ("dst_mem", C.c_int * MAX_RUNGS), ("produced", C.c_int),
                ("opaque", C.c_void_p), ("pts", C.c_int64), ("dts", C.c_int64), ("interlaced", C.c_int),
                ("tff", C.c_int), ("thumbnail", C.c_void_p), ("n_released", C.c_int), ("released", C.c_void_p * 4)]


class Picture(C.Structure):
    _fields_ = [("plane", C.c_void_p * 3), ("stride", C.c_int * 3), ("mem", C.c_int)]


class OttGpuError(RuntimeError):
    def __init__(self, code, detail):
        RuntimeError.__init__(self, "ottgpu error %d: %s" % (code, detail))
        self.code = code


_SIGS = {
    "ottgpu_abi_version": (C.c_int, []),
    "ottgpu_build_info": (C.c_char_p, []),
    "ottgpu_device_count": (C.c_int, []),
    "ottgpu_strerror": (C.c_char_p, [C.c_int]),
    "ottgpu_last_error": (C.c_char_p, []),
    "ottgpu_frame_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ottgpu_rung_bytes": (C.c_size_t, [C.POINTER(RungCfg)]),
    "ottgpu_rung_layout": (C.c_int, [C.POINTER(RungCfg), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]),
    "ottgpu_channel_create": (C.c_int, [C.POINTER(Cfg), C.POINTER(C.c_void_p)]),
    "ottgpu_channel_submit": (C.c_int, [C.c_void_p, C.POINTER(FrameIn), C.POINTER(FrameOut)]),
    "ottgpu_channel_reset": (C.c_int, [C.c_void_p]),
    "ottgpu_channel_destroy": (None, [C.c_void_p]),
    "ottgpu_channel_gpu": (C.c_int, [C.c_void_p]),
    "ottgpu_channel_describe": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "ottgpu_channel_pending_opaque": (C.c_void_p, [C.c_void_p]),
    "ottgpu_batch_last_uploads": (C.c_int, [C.c_void_p]),
    "ottgpu_batch_last_readbacks": (C.c_int, [C.c_void_p]),
    "ottgpu_batch_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]),
    "ottgpu_batch_submit": (C.c_int, [C.c_void_p, C.POINTER(FrameIn), C.POINTER(FrameOut), C.c_int]),
    "ottgpu_batch_wait": (C.c_int, [C.c_void_p]),
    "ottgpu_batch_destroy": (None, [C.c_void_p]),
    "ottgpu_batch_last_launches": (C.c_int, [C.c_void_p]),
    "ottgpu_batch_last_items": (C.c_int, [C.c_void_p]),
    "ottgpu_batch_timing": (C.c_int, [C.c_void_p, C.c_int]),
    "ottgpu_batch_timing_read": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "ottgpu_batch_timing_read_yadif": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]),
    "ottgpu_pool_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "ottgpu_pool_take": (C.c_void_p, [C.c_void_p]),
    "ottgpu_pool_return": (C.c_int, [C.c_void_p, C.c_void_p]),
    "ottgpu_pool_unused": (C.c_int, [C.c_void_p]),
    "ottgpu_pool_destroy": (None, [C.c_void_p]),
    "ottgpu_memcpy": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_size_t]),
    "ottgpu_converter_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]),
    "ottgpu_converter_convert": (C.c_int, [C.c_void_p, C.POINTER(Picture), C.c_void_p, C.c_int, C.c_int]),
    "ottgpu_converter_convert_batch": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(Picture), C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int]),
    "ottgpu_converter_wait": (C.c_int, [C.c_void_p]),
    "ottgpu_converter_destroy": (None, [C.c_void_p]),
    "ottgpu_jpeg_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "ottgpu_jpeg_encode": (C.c_int, [C.c_void_p, C.POINTER(Picture), C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "ottgpu_jpeg_max_bytes": (C.c_size_t, [C.c_void_p]),
    "ottgpu_jpeg_destroy": (None, [C.c_void_p]),
    "ottgpu_scale_frame": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(RungCfg),
                                     C.c_int, C.c_int]),
    "ottgpu_yadif_frame": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.c_int]),
    "ottgpu_filter_table": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int16)]),
}
EXPORTS = sorted(_SIGS)


class Library:
    """One loaded libottgpu (the real one by default; tests may point it at the CPU emulation build)."""

    def __init__(self, path=None):
        self.path = path or DEFAULT_LIB
        if not os.path.exists(self.path):
            raise OttGpuError(-2, "%s is missing: build it with `make -C ott-packager_b200/csrc` "
                                  "(there is no CPU fallback)" % self.path)
        self.dll = C.CDLL(self.path)
        for name, (res, args) in _SIGS.items():
            try:
                fn = getattr(self.dll, name)
            except AttributeError:
                if name in ("ottgpu_build_info", "ottgpu_channel_pending_opaque", "ottgpu_batch_last_uploads", "ottgpu_batch_last_readbacks"):     # absent from older A/B builds
                    continue
                raise
            fn.restype, fn.argtypes = res, args

    def check(self, rc):
        if rc < 0:
            raise OttGpuError(rc, (self.dll.ottgpu_last_error() or b"").decode() or
                              self.dll.ottgpu_strerror(rc).decode())
        return rc

    # ---------------------------------------------------------------- geometry
    def frame_bytes(self, fmt, w, h):
        return self.dll.ottgpu_frame_bytes(fmt, w, h)

    def rung_bytes(self, rung):
        return self.dll.ottgpu_rung_bytes(C.byref(rung))

    def rung_layout(self, rung):
        off = (C.c_size_t * 3)()
        pitch = (C.c_int * 3)()
        n = self.check(self.dll.ottgpu_rung_layout(C.byref(rung), off, pitch))
        return [(off[i], pitch[i]) for i in range(n)]

    def filter_table(self, src_n, dst_n, filt, target, vertical):
        pos = np.empty(dst_n, np.int32)
        coef = np.empty(dst_n * 256, np.int16)
        taps = self.check(self.dll.ottgpu_filter_table(src_n, dst_n, filt, target, int(vertical),
                                                       pos.ctypes.data_as(C.POINTER(C.c_int32)),
                                                       coef.ctypes.data_as(C.POINTER(C.c_int16))))
        return taps, pos, coef[: dst_n * taps].reshape(dst_n, taps).copy()

    def to_device(self, dev_ptr, arr, gpu=0):
        arr = np.ascontiguousarray(arr)
        self.check(self.dll.ottgpu_memcpy(gpu, dev_ptr, MEM_DEVICE, arr.ctypes.data, MEM_HOST, arr.nbytes))

    def from_device(self, dev_ptr, nbytes, dtype=np.uint8, gpu=0):
        out = np.empty(nbytes, np.uint8)
        self.check(self.dll.ottgpu_memcpy(gpu, out.ctypes.data, MEM_HOST, dev_ptr, MEM_DEVICE, nbytes))
        return out.view(dtype)

    # ---------------------------------------------------------------- single shots (host numpy in / out)
    def scale_frame(self, src, src_fmt, sw, sh, dw, dh, dst_fmt=None, filt=FILTER_BICUBIC, target=TARGET_A,
                    pitch_align=0, gpu=0):
        dst_fmt = src_fmt if dst_fmt is None else dst_fmt
        rung = RungCfg(dw, dh, dst_fmt, pitch_align)
        src = np.ascontiguousarray(src)
        assert src.nbytes == self.frame_bytes(src_fmt, sw, sh), (src.nbytes, self.frame_bytes(src_fmt, sw, sh))
        out = np.zeros(self.rung_bytes(rung), np.uint8)
        self.check(self.dll.ottgpu_scale_frame(gpu, src.ctypes.data, src_fmt, sw, sh, out.ctypes.data, C.byref(rung),
                                               filt, target))
        return out if src.dtype == np.uint8 else out.view(np.uint16)

    def yadif_frame(self, prev, cur, nxt, fmt, w, h, interlaced=1, tff=1, gpu=0):
        prev, cur, nxt = (np.ascontiguousarray(a) for a in (prev, cur, nxt))
        out = np.zeros_like(cur)
        self.check(self.dll.ottgpu_yadif_frame(gpu, prev.ctypes.data, cur.ctypes.data, nxt.ctypes.data, out.ctypes.data,
                                               fmt, w, h, interlaced, tff))
        return out


def make_cfg(src_w, src_h, rungs, src_fmt=FMT_I420, filt=FILTER_BICUBIC, target=TARGET_A, deint=DEINT_ALL, gpu=0,
             thumb=(0, 0, 0), stream=None, thumb_phase=0):
    """rungs: list of (w, h, fmt[, pitch_align])"""
    cfg = Cfg()
    cfg.gpu, cfg.src_width, cfg.src_height, cfg.src_fmt = gpu, src_w, src_h, src_fmt
    cfg.n_rungs = len(rungs)
    for i, r in enumerate(rungs[:MAX_RUNGS]):
        cfg.rung[i] = RungCfg(r[0], r[1], r[2], r[3] if len(r) > 3 else 0)
    cfg.filter, cfg.target, cfg.deint = filt, target, deint
    cfg.thumb_width, cfg.thumb_height, cfg.thumb_period = thumb
    cfg.thumb_phase = thumb_phase
    cfg.stream = stream
    return cfg


class Channel:
    """One transvideo pipeline.  submit() takes one packed host frame (numpy) or a device pointer (int) and returns
    None while yadif primes, else a dict {rungs: [np arrays or device ptrs], opaque, pts, dts, thumbnail}."""

    def __init__(self, lib, cfg):
        self.lib, self.cfg = lib, cfg
        h = C.c_void_p()
        lib.check(lib.dll.ottgpu_channel_create(C.byref(cfg), C.byref(h)))
        self.handle = h
        self.rung_bytes = [lib.rung_bytes(cfg.rung[i]) for i in range(cfg.n_rungs)]
        self.bps = 1 if cfg.src_fmt == FMT_I420 else 2
        self._keep = []

    def close(self):
        if self.handle:
            self.lib.dll.ottgpu_channel_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self.lib.check(self.lib.dll.ottgpu_channel_reset(self.handle))

    def describe(self):
        buf = C.create_string_buffer(16384)
        self.lib.dll.ottgpu_channel_describe(self.handle, buf, len(buf))
        return buf.value.decode()

    def fill_in(self, fin, frame, interlaced, tff, pts, dts, opaque):
        if isinstance(frame, np.ndarray):
            frame = np.ascontiguousarray(frame)
            self._keep.append(frame)
            fin.data, fin.mem = frame.ctypes.data, MEM_HOST
        else:
            fin.data, fin.mem = int(frame), MEM_DEVICE
        fin.interlaced, fin.tff, fin.pts, fin.dts = int(interlaced), int(tff), pts, dts
        fin.opaque = opaque

    def alloc_out(self, fout, device_dst=None):
        bufs = []
        for r in range(self.cfg.n_rungs):
            if device_dst is not None:
                fout.dst[r], fout.dst_mem[r] = int(device_dst[r]), MEM_DEVICE
                bufs.append(int(device_dst[r]))
            else:
                a = np.zeros(self.rung_bytes[r], np.uint8)
                fout.dst[r], fout.dst_mem[r] = a.ctypes.data, MEM_HOST
                bufs.append(a)
        return bufs

    def collect(self, fout, bufs):
        if not fout.produced:
            return None
        thumb = None
        if fout.thumbnail:
            n = self.lib.frame_bytes(FMT_I420 if self.bps == 1 else FMT_I420P10, self.cfg.thumb_width,
                                     self.cfg.thumb_height)
            thumb = np.ctypeslib.as_array(C.cast(fout.thumbnail, C.POINTER(C.c_uint8)), (n,)).copy()
            C.CDLL(None).free(C.c_void_p(fout.thumbnail))
        rungs = [b if not isinstance(b, np.ndarray) or self.bps == 1 else b.view(np.uint16) for b in bufs]
        return {"rungs": rungs, "opaque": fout.opaque, "pts": fout.pts, "dts": fout.dts, "thumbnail": thumb,
                "released": [fout.released[i] for i in range(fout.n_released)]}

    def submit(self, frame, interlaced=1, tff=1, pts=0, dts=0, opaque=None, device_dst=None):
        fin, fout = FrameIn(), FrameOut()
        self._keep = []
        self.fill_in(fin, frame, interlaced, tff, pts, dts, opaque)
        bufs = self.alloc_out(fout, device_dst)
        self.lib.check(self.lib.dll.ottgpu_channel_submit(self.handle, C.byref(fin), C.byref(fout)))
        return self.collect(fout, bufs)


class Converter:
    """The decode thread's `decode_converter` + packing loops (transvideo.c:2797-2840): decoder planes -> packed I420."""

    def __init__(self, lib, pix_fmt, w, h, target=TARGET_A, gpu=0, stream=None):
        self.lib, self.pix_fmt, self.w, self.h, self.gpu = lib, pix_fmt, w, h, gpu
        h_ = C.c_void_p()
        lib.check(lib.dll.ottgpu_converter_create(gpu, pix_fmt, w, h, target, stream, C.byref(h_)))
        self.handle = h_
        self.out_bytes = lib.frame_bytes(FMT_I420, w, h)

    def close(self):
        if self.handle:
            self.lib.dll.ottgpu_converter_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def picture(self, planes, mem=MEM_HOST, strides=None):
        """planes: numpy arrays (host) or integer device addresses with `strides` in bytes."""
        pic = Picture()
        pic.mem = mem
        for i, p in enumerate(planes):
            if mem == MEM_HOST:
                pic.plane[i], pic.stride[i] = p.ctypes.data, p.strides[0]
            else:
                pic.plane[i], pic.stride[i] = p, strides[i]
        return pic

    def convert(self, planes, device_dst=None, flags=SUBMIT_SYNC, mem=MEM_HOST, strides=None):
        """-> packed I420 numpy array, or None when the frame is written to `device_dst`."""
        pic = self.picture(planes, mem, strides)
        self._keep = (planes, pic)
        if device_dst is not None:
            self.lib.check(self.lib.dll.ottgpu_converter_convert(self.handle, C.byref(pic), device_dst, MEM_DEVICE, flags))
            return None
        out = np.empty(self.out_bytes, np.uint8)
        self.lib.check(self.lib.dll.ottgpu_converter_convert(self.handle, C.byref(pic), out.ctypes.data, MEM_HOST, SUBMIT_SYNC))
        return out

    def wait(self):
        self.lib.check(self.lib.dll.ottgpu_converter_wait(self.handle))


class Jpeg:
    """The thumbnail JPEG encoder on the GPU (save_frame_as_jpeg, transvideo.c:164-198)."""

    def __init__(self, lib, w, h, qscale=3, gpu=0):
        self.lib, self.w, self.h = lib, w, h
        h_ = C.c_void_p()
        lib.check(lib.dll.ottgpu_jpeg_create(gpu, w, h, qscale, C.byref(h_)))
        self.handle = h_
        self.max_bytes = lib.dll.ottgpu_jpeg_max_bytes(h_)

    def encode(self, i420=None, device_ptr=None, planes=None, strides=None, capacity=None):
        """packed I420 numpy frame, a device pointer to one, or explicit host planes + strides -> JPEG bytes"""
        w, h = self.w, self.h
        cw, ch = (w + 1) // 2, (h + 1) // 2
        pic = Picture()
        if planes is not None:
            pic.mem = MEM_HOST
            for k in range(3):
                pic.plane[k], pic.stride[k] = planes[k].ctypes.data, strides[k]
        else:
            base = device_ptr if device_ptr is not None else np.ascontiguousarray(i420, np.uint8).ctypes.data
            pic.mem = MEM_DEVICE if device_ptr is not None else MEM_HOST
            for k, (off, st) in enumerate(((0, w), (w * h, cw), (w * h + cw * ch, cw))):
                pic.plane[k], pic.stride[k] = base + off, st
        out = np.empty(self.max_bytes if capacity is None else capacity, np.uint8)
        n = C.c_size_t(0)
        self.lib.check(self.lib.dll.ottgpu_jpeg_encode(self.handle, C.byref(pic), out.ctypes.data, out.size, C.byref(n)))
        return out[:n.value].tobytes()

    def close(self):
        if self.handle:
            self.lib.dll.ottgpu_jpeg_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def convert_batch(lib, converters, planes_list, device_dsts=None, flags=SUBMIT_SYNC, mem=MEM_HOST, strides_list=None):
    """One launch for the pictures of several converters (one per channel).  Returns the packed frames (host) or None."""
    n = len(converters)
    cvs = (C.c_void_p * n)(*[c.handle for c in converters])
    pics = (Picture * n)()
    keep = []
    for i, (c, planes) in enumerate(zip(converters, planes_list)):
        p = c.picture(planes, mem, None if strides_list is None else strides_list[i])
        pics[i].mem = p.mem
        for k in range(3):
            pics[i].plane[k], pics[i].stride[k] = p.plane[k], p.stride[k]
        keep.append(planes)
    if device_dsts is not None:
        dsts = (C.c_void_p * n)(*device_dsts)
        lib.check(lib.dll.ottgpu_converter_convert_batch(cvs, pics, dsts, n, MEM_DEVICE, flags))
        return None
    outs = [np.empty(c.out_bytes, np.uint8) for c in converters]
    dsts = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
    lib.check(lib.dll.ottgpu_converter_convert_batch(cvs, pics, dsts, n, MEM_HOST, SUBMIT_SYNC))
    return outs


class Batch:
    def __init__(self, lib, channels):
        self.lib, self.channels = lib, list(channels)
        arr = (C.c_void_p * len(channels))(*[c.handle for c in channels])
        h = C.c_void_p()
        lib.check(lib.dll.ottgpu_batch_create(arr, len(channels), C.byref(h)))
        self.handle = h
        self.fin = (FrameIn * len(channels))()
        self.fout = (FrameOut * len(channels))()

    def close(self):
        if self.handle:
            self.lib.dll.ottgpu_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def submit_raw(self, flags=SUBMIT_SYNC):
        """fin/fout arrays already filled by the caller (bench.py hot loop)."""
        self.lib.check(self.lib.dll.ottgpu_batch_submit(self.handle, self.fin, self.fout, flags))

    def wait(self):
        self.lib.check(self.lib.dll.ottgpu_batch_wait(self.handle))

    def submit(self, frames, interlaced=1, tff=1, device_dst=None):
        """frames: one per channel (None skips the channel).  Returns list of per-channel results (None if nothing)."""
        bufs = []
        for i, (c, f) in enumerate(zip(self.channels, frames)):
            c._keep = []
            if f is None:
                self.fin[i].data = None
                bufs.append(None)
                continue
            c.fill_in(self.fin[i], f, interlaced, tff, 0, 0, None)
            bufs.append(c.alloc_out(self.fout[i], None if device_dst is None else device_dst[i]))
        self.submit_raw(SUBMIT_SYNC)
        return [None if b is None else c.collect(self.fout[i], b) for i, (c, b) in enumerate(zip(self.channels, bufs))]

    def timing(self, enable=True):
        self.lib.check(self.lib.dll.ottgpu_batch_timing(self.handle, int(enable)))

    def timing_read(self):
        ms, n = C.c_double(), C.c_int()
        self.lib.check(self.lib.dll.ottgpu_batch_timing_read(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def timing_read_yadif(self):
        ms, n = C.c_double(), C.c_int()
        self.lib.check(self.lib.dll.ottgpu_batch_timing_read_yadif(self.handle, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    @property
    def last_launches(self):
        return self.lib.dll.ottgpu_batch_last_launches(self.handle)

    @property
    def last_items(self):
        return self.lib.dll.ottgpu_batch_last_items(self.handle)

    @property
    def last_uploads(self):
        return self.lib.dll.ottgpu_batch_last_uploads(self.handle)

    @property
    def last_readbacks(self):
        return self.lib.dll.ottgpu_batch_last_readbacks(self.handle)


class Pool:
    def __init__(self, lib, gpu, mem, count, size):
        self.lib = lib
        h = C.c_void_p()
        lib.check(lib.dll.ottgpu_pool_create(gpu, mem, count, size, C.byref(h)))
        self.handle = h

    def take(self):
        return self.lib.dll.ottgpu_pool_take(self.handle)

    def give_back(self, buf):
        return self.lib.dll.ottgpu_pool_return(self.handle, buf)

    def unused(self):
        return self.lib.dll.ottgpu_pool_unused(self.handle)

    def close(self):
        if self.handle:
            self.lib.dll.ottgpu_pool_destroy(self.handle)
            self.handle = None
