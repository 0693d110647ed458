"""Host-side C glue (ott-packager_b200/host/transvideo_gpu.c): the thread body that replaces video_prepare_thread +
video_scale_thread (source/transvideo.c:1896-2429, 1504-1683), driven through the REFERENCE's own dataqueue.c / mempool.c
(oracle/_ref/libfillet_rt.so, compiled unmodified from /root/reference) so the queue/pool hand-off is the real one.
Pixels are checked against the CPU oracle.  Tier `emu` runs the kernel source on the CPU, tier `gpu` on the B200."""
import ctypes as C
import os
import subprocess
import time
import numpy as np
import pytest
from conftest import ROOT, interlaced_sequence, natural_i420

MAXR = 8


class Msg(C.Structure):                       # include/dataqueue.h:32-59
    _fields_ = [("buffer_type", C.c_int), ("buffer_size", C.c_ulong), ("pts", C.c_int64), ("dts", C.c_int64),
                ("flags", C.c_uint32), ("source_discontinuity", C.c_int), ("tff", C.c_uint8), ("interlaced", C.c_uint8),
                ("fps_num", C.c_int), ("fps_den", C.c_int), ("aspect_num", C.c_int), ("aspect_den", C.c_int),
                ("width", C.c_int), ("height", C.c_int), ("stream_index", C.c_uint8), ("channels", C.c_int),
                ("sample_rate", C.c_int), ("first_pts", C.c_int64), ("caption_size", C.c_int),
                ("caption_timestamp", C.c_int64), ("splice_point", C.c_int), ("splice_duration", C.c_int64),
                ("splice_duration_remaining", C.c_int64), ("smallbuf", C.c_char * 256),
                ("caption_buffer", C.c_void_p), ("buffer", C.c_void_p)]


SIGNAL_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_char_p)


class TV(C.Structure):                        # ott-packager_b200/host/transvideo_gpu.h
    _fields_ = [("prepare_queue", C.c_void_p), ("encode_queue", C.c_void_p * MAXR), ("thumbnail_queue", C.c_void_p),
                ("raw_video_pool", C.c_void_p), ("msg_pool", C.c_void_p), ("num_outputs", C.c_int),
                ("out_width", C.c_int * MAXR), ("out_height", C.c_int * MAXR), ("gpu", C.c_int),
                ("first_timestamp", C.POINTER(C.c_int64)), ("running", C.POINTER(C.c_int)), ("out_fmt", C.c_int),
                ("device_outputs", C.c_int), ("surface_pool", C.c_void_p * MAXR), ("target", C.c_int),
                ("signal", SIGNAL_FN), ("direct_error", SIGNAL_FN), ("user", C.c_void_p), ("channel", C.c_void_p),
                ("chan_width", C.c_int), ("chan_height", C.c_int), ("chan_deint", C.c_int),
                ("deinterlaced_frame_count", C.c_int64), ("av_sync_compromised", C.c_int), ("frames_in", C.c_int64),
                ("frames_out", C.c_int64), ("thumbnails_out", C.c_int64), ("input_pool", C.c_void_p),
                ("held_inputs", C.c_void_p * 8), ("n_held", C.c_int)]


@pytest.fixture(scope="module")
def rt():
    """the reference's dataqueue.c + mempool.c, unmodified"""
    so = os.path.join(ROOT, "oracle", "_ref", "libfillet_rt.so")
    if not os.path.exists(so):
        subprocess.call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not os.path.exists(so):
        pytest.skip("oracle/_ref/libfillet_rt.so not built (reference tree absent)")
    L = C.CDLL(so, mode=C.RTLD_GLOBAL)
    L.dataqueue_create.restype = C.c_void_p
    L.dataqueue_take_back.restype = C.POINTER(Msg)
    L.dataqueue_take_back.argtypes = [C.c_void_p]
    L.dataqueue_put_front.argtypes = [C.c_void_p, C.POINTER(Msg)]
    L.dataqueue_get_size.argtypes = [C.c_void_p]
    L.memory_create.restype = C.c_void_p
    L.memory_create.argtypes = [C.c_int, C.c_int]
    L.memory_take.restype = C.c_void_p
    L.memory_take.argtypes = [C.c_void_p, C.c_int]
    L.memory_return.argtypes = [C.c_void_p, C.c_void_p]
    L.memory_unused.argtypes = [C.c_void_p]
    return L


@pytest.fixture()
def host(lib, rt):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "ott-packager_b200", "host")])
    C.CDLL(lib.path, mode=C.RTLD_GLOBAL)          # libottgpu (emulated or real) provides the ottgpu_* symbols
    H = C.CDLL(os.path.join(ROOT, "ott-packager_b200", "libottgpu_host.so"))
    H.ottgpu_transvideo_process.argtypes = [C.POINTER(TV), C.POINTER(Msg)]
    H.ottgpu_video_prepare_scale_thread.argtypes = [C.c_void_p]
    H.ottgpu_video_prepare_scale_thread.restype = C.c_void_p
    H.ottgpu_transvideo_close.argtypes = [C.POINTER(TV)]
    return H


def test_message_struct_matches_reference_header():
    hdr = "/root/reference/include/dataqueue.h"
    if not os.path.exists(hdr):
        pytest.skip("reference tree absent")
    src = r'''
#include <stdio.h>
#include <stddef.h>
#define P(f) printf("%zu ", offsetof(dataqueue_message_struct, f))
int main(void){ P(buffer_size);P(pts);P(flags);P(tff);P(interlaced);P(fps_num);P(width);P(stream_index);P(first_pts);P(caption_size);
 P(splice_point);P(splice_duration_remaining);P(smallbuf);P(caption_buffer);P(buffer); printf("%zu\n", sizeof(dataqueue_message_struct)); return 0; }
'''
    outs = []
    for flags in (["-I/root/reference/include", "-include", "dataqueue.h"],
                  ["-include", os.path.join(ROOT, "ott-packager_b200", "host", "fillet_abi.h")]):
        exe = "/tmp/ottgpu_abi_check"
        subprocess.run(["gcc", "-x", "c", "-o", exe] + flags + ["-"], input=src.encode(), check=True)
        outs.append(subprocess.check_output([exe]).decode())
    assert outs[0] == outs[1], outs
    assert C.sizeof(Msg) == int(outs[0].split()[-1]) == 400


def _make_tv(rt, w_h_list, errors, signals, first_ts=None):
    tv = TV()
    tv.prepare_queue = rt.dataqueue_create()
    for i in range(len(w_h_list)):
        tv.encode_queue[i] = rt.dataqueue_create()
        tv.out_width[i], tv.out_height[i] = w_h_list[i]
    tv.thumbnail_queue = rt.dataqueue_create()
    tv.raw_video_pool = rt.memory_create(512, 0)          # fillet.c:355: size 0 => memory_take mallocs the request
    tv.msg_pool = rt.memory_create(256, C.sizeof(Msg))    # fillet.c:339 (8192 in production)
    tv.num_outputs = len(w_h_list)
    tv.gpu, tv.out_fmt, tv.device_outputs, tv.target = 0, 0, 0, 0
    tv._running = C.c_int(1)
    tv.running = C.pointer(tv._running)
    if first_ts is not None:
        tv._first = C.c_int64(first_ts)
        tv.first_timestamp = C.pointer(tv._first)
    tv._err = SIGNAL_FN(lambda u, code, m: errors.append((code, m)))
    tv._sig = SIGNAL_FN(lambda u, code, m: signals.append((code, m)))
    tv.direct_error, tv.signal = tv._err, tv._sig
    return tv


def _push(rt, tv, frame, w, h, pts, interlaced=1, tff=1, fps=(30000, 1001), caption=None, disc=0):
    buf = rt.memory_take(tv.raw_video_pool, frame.nbytes)
    C.memmove(buf, frame.ctypes.data, frame.nbytes)
    m = C.cast(rt.memory_take(tv.msg_pool, C.sizeof(Msg)), C.POINTER(Msg))
    C.memset(m, 0, C.sizeof(Msg))
    m.contents.buffer, m.contents.buffer_size = buf, frame.nbytes
    m.contents.pts, m.contents.dts = pts, pts - 3003
    m.contents.width, m.contents.height = w, h
    m.contents.interlaced, m.contents.tff = interlaced, tff
    m.contents.fps_num, m.contents.fps_den = fps
    m.contents.aspect_num, m.contents.aspect_den = 16, 9
    m.contents.source_discontinuity = disc
    if caption:
        cb = C.CDLL(None).malloc
        cb.restype = C.c_void_p
        p = cb(len(caption))
        C.memmove(p, caption, len(caption))
        m.contents.caption_buffer, m.contents.caption_size, m.contents.caption_timestamp = p, len(caption), pts
    return m


def _drain(rt, tv, q, nbytes):
    out = []
    while True:
        m = rt.dataqueue_take_back(q)
        if not m:
            break
        c = m.contents
        data = np.ctypeslib.as_array(C.cast(c.buffer, C.POINTER(C.c_uint8)), (nbytes,)).copy()
        cap = C.string_at(c.caption_buffer, c.caption_size) if c.caption_buffer else None
        out.append(dict(data=data, pts=c.pts, dts=c.dts, w=c.width, h=c.height, idx=c.stream_index, il=c.interlaced,
                        tff=c.tff, fps=(c.fps_num, c.fps_den), caption=cap, cap_ptr=c.caption_buffer, size=c.buffer_size))
        rt.memory_return(tv.raw_video_pool, c.buffer)     # what the encoder thread does (transvideo.c:1456-1458)
        rt.memory_return(tv.msg_pool, m)
    return out


def test_thread_body_semantics(lib, og, oracle, rt, host):
    w, h = 192, 108
    outs = [(192, 108), (128, 72), (64, 36)]
    errors, signals = [], []
    tv = _make_tv(rt, outs, errors, signals)
    frames = interlaced_sequence(w, h, 6, seed=21)
    pool0 = rt.memory_unused(tv.raw_video_pool)
    ref = oracle.YadifStream(w, h)
    expected = []
    for i, f in enumerate(frames):
        cap = b"CC%02d" % i if i in (1, 3) else None
        m = _push(rt, tv, f, w, h, 90000 + 3003 * i, caption=cap)
        assert host.ottgpu_transvideo_process(C.byref(tv), m) == 0
        r = ref.push(f, 1, 1, tag=i)
        if r is not None:
            expected.append(r)
    assert not errors
    assert tv.frames_in == 6 and tv.frames_out == 5
    per_rung = [_drain(rt, tv, tv.encode_queue[k], oracle.i420_size(*outs[k])) for k in range(3)]
    for k, (dw, dh) in enumerate(outs):
        assert len(per_rung[k]) == 5
        for n, got in enumerate(per_rung[k]):
            deint, tag = expected[n]
            assert np.array_equal(got["data"], oracle.scale_i420(deint, w, h, dw, dh)), (k, n)
            assert (got["w"], got["h"], got["idx"], got["il"], got["tff"]) == (dw, dh, k, 0, 1)
            assert got["pts"] == 90000 + 3003 * tag and got["dts"] == got["pts"] - 3003     # sideband follows the frame
            assert got["fps"] == (30000, 1001) and got["size"] == oracle.i420_size(dw, dh)
            assert got["caption"] == (b"CC%02d" % tag if tag in (1, 3) else None)
    # last rendition owns the original caption buffer, the others got copies (transvideo.c:1649-1657)
    caps = [[g["cap_ptr"] for g in per_rung[k] if g["cap_ptr"]] for k in range(3)]
    assert len(set(caps[0]) | set(caps[1]) | set(caps[2])) == 6
    # every input buffer went back to the pool (transvideo.c:2410) and the outputs were returned by the "encoder"
    assert rt.memory_unused(tv.raw_video_pool) == pool0
    host.ottgpu_transvideo_close(C.byref(tv))


def test_thread_body_runs_as_pthread_with_discontinuity(lib, og, oracle, rt, host):
    import threading
    w, h = 192, 108
    outs = [(128, 72)]
    errors, signals = [], []
    tv = _make_tv(rt, outs, errors, signals)
    frames = [natural_i420(w, h, 7, shift=2.0 * i) for i in range(8)]
    t = threading.Thread(target=host.ottgpu_video_prepare_scale_thread, args=(C.addressof(tv),))
    t.start()
    for i, f in enumerate(frames):                      # 60000/1001 progressive-flagged => pass-through, one frame late
        m = _push(rt, tv, f, w, h, 1000 * i, interlaced=0, fps=(60000, 1001), disc=1 if i == 4 else 0)
        rt.dataqueue_put_front(tv.prepare_queue, m)
    deadline = time.time() + 30
    while tv.frames_in < 8 and time.time() < deadline:
        time.sleep(0.01)
    time.sleep(0.05)
    tv._running.value = 0
    t.join(10)
    assert not t.is_alive() and not errors
    got = _drain(rt, tv, tv.encode_queue[0], oracle.i420_size(128, 72))
    # discontinuity at frame 4 drops the window: frames 0,1,2 come out, then 4,5,6
    assert [g["pts"] for g in got] == [0, 1000, 2000, 4000, 5000, 6000]
    for g in got:
        assert np.array_equal(g["data"], oracle.scale_i420(frames[g["pts"] // 1000], w, h, 128, 72))


def test_av_sync_repeat_and_drop(lib, og, oracle, rt, host):
    # transvideo.c:2233-2382: frames are repeated / dropped when the frame count drifts from what dts implies
    w, h = 96, 64
    errors, signals = [], []
    tv = _make_tv(rt, [(64, 40)], errors, signals, first_ts=0)
    f = natural_i420(w, h, 7)
    # dts runs 4 frame-times per frame => the output falls behind => repeats
    for i in range(5):
        m = _push(rt, tv, f, w, h, 3003 + 4 * 3003 * i, interlaced=0, fps=(60000, 1001))
        assert host.ottgpu_transvideo_process(C.byref(tv), m) == 0
    got = _drain(rt, tv, tv.encode_queue[0], oracle.i420_size(64, 40))
    assert len(got) > 4 and any(g["pts"] == 0 for g in got) and signals and signals[0][0] == 0x11
    host.ottgpu_transvideo_close(C.byref(tv))
    # dts stands still => the output runs ahead => drops
    errors2, signals2 = [], []
    tv2 = _make_tv(rt, [(64, 40)], errors2, signals2, first_ts=0)
    for i in range(8):
        m = _push(rt, tv2, f, w, h, 3003, interlaced=0, fps=(60000, 1001))
        assert host.ottgpu_transvideo_process(C.byref(tv2), m) == 0
    got2 = _drain(rt, tv2, tv2.encode_queue[0], oracle.i420_size(64, 40))
    assert len(got2) < 7
    host.ottgpu_transvideo_close(C.byref(tv2))


def test_device_resident_inputs_from_the_converter(lib, og, oracle, rt, host):
    """Decode stage converts a 4:2:2 10-bit picture on the GPU and leaves the packed frame there
    (ottgpu_converter_convert, OTTGPU_MEM_DEVICE); the message is tagged OTTGPU_MSG_FLAG_DEVICE and the thread body feeds
    the frame to the channel without a host round trip, returning it to its pool when the yadif window lets go of it."""
    from conftest import decoded_planes
    w, h = 192, 108
    outs = [(192, 108), (96, 54)]
    errors, signals = [], []
    tv = _make_tv(rt, outs, errors, signals)
    in_pool = og.Pool(lib, 0, og.MEM_DEVICE, 5, lib.frame_bytes(og.FMT_I420, w, h))
    tv.input_pool = in_pool.handle
    cv = og.Converter(lib, og.PIX_YUV422P10, w, h)
    ref = oracle.YadifStream(w, h)
    expected = []
    for i in range(7):
        planes = decoded_planes("yuv422p10le", w, h, 400 + i, "noise", pad=4)
        frame = oracle.convert_to_i420(planes, oracle.PIX_YUV422P10, w, h)
        buf = in_pool.take()
        assert buf, "input pool ran dry: released frames are not coming back"
        cv.convert(planes, device_dst=buf)
        m = C.cast(rt.memory_take(tv.msg_pool, C.sizeof(Msg)), C.POINTER(Msg))
        C.memset(m, 0, C.sizeof(Msg))
        m.contents.buffer, m.contents.buffer_size, m.contents.flags = buf, frame.nbytes, 0x4f540001
        m.contents.pts, m.contents.dts = 9000 * i, 9000 * i - 3003
        m.contents.width, m.contents.height = w, h
        m.contents.interlaced, m.contents.tff = 1, 1
        m.contents.fps_num, m.contents.fps_den = 30000, 1001
        m.contents.source_discontinuity = 1 if i == 4 else 0
        assert host.ottgpu_transvideo_process(C.byref(tv), m) == 0
        if i == 4:
            ref.reset()
        r = ref.push(frame, 1, 1, tag=i)
        if r is not None:
            expected.append(r)
    assert not errors
    for k, (dw, dh) in enumerate(outs):
        got = _drain(rt, tv, tv.encode_queue[k], oracle.i420_size(dw, dh))
        assert [g["pts"] for g in got] == [9000 * tag for _, tag in expected]
        for g, (deint, tag) in zip(got, expected):
            assert np.array_equal(g["data"], oracle.scale_i420(deint, w, h, dw, dh)), (k, tag)
    assert 0 < tv.n_held <= 3                               # the window still references the newest frames ...
    host.ottgpu_transvideo_close(C.byref(tv))
    assert tv.n_held == 0 and in_pool.unused() == 5         # ... and closing hands every one of them back
    cv.close()
    in_pool.close()


def test_av_sync_repeat_with_device_surfaces(lib, og, oracle, rt, host):
    """NVENC mode (device_outputs=1): a repeated frame is a second device surface with the same pixels, duplicated on
    the device (transvideo.c:2271-2316 copies on the host)."""
    w, h = 96, 64
    dw, dh = 64, 40
    errors, signals = [], []
    tv = _make_tv(rt, [(dw, dh)], errors, signals, first_ts=0)
    tv.device_outputs, tv.out_fmt = 1, og.FMT_NV12
    rung = og.RungCfg(dw, dh, og.FMT_NV12, 0)
    nbytes = lib.rung_bytes(rung)
    pool = og.Pool(lib, 0, og.MEM_DEVICE, 16, nbytes)
    tv.surface_pool[0] = pool.handle
    f = natural_i420(w, h, 7)
    for i in range(4):                                   # dts runs 4 frame-times per frame => repeats
        m = _push(rt, tv, f, w, h, 3003 + 4 * 3003 * i, interlaced=0, fps=(60000, 1001))
        assert host.ottgpu_transvideo_process(C.byref(tv), m) == 0
    want = oracle.i420_to_nv12(oracle.scale_i420(f, w, h, dw, dh), dw, dh)
    seen, repeats = set(), 0
    while True:
        m = rt.dataqueue_take_back(tv.encode_queue[0])
        if not m:
            break
        c = m.contents
        assert c.flags & 0xffff00ff == 0x4f540001 and c.buffer not in seen
        seen.add(c.buffer)
        repeats += c.pts == 0
        assert np.array_equal(lib.from_device(c.buffer, nbytes), want)
        pool.give_back(c.buffer)
        rt.memory_return(tv.msg_pool, m)
    assert repeats >= 1 and len(seen) > 3 and not errors
    host.ottgpu_transvideo_close(C.byref(tv))
    assert pool.unused() == 16
    pool.close()


def test_pool_exhaustion_reports_like_the_reference(lib, og, oracle, rt, host):
    """transvideo.c:1596-1600, 2203-2207: when raw_video_pool has no buffer left the reference calls
    send_direct_error(SIGNAL_DIRECT_ERROR_RAWPOOL, "Out of Uncompressed Video Buffers ...") and exits; the thread body
    reports through the same hook (the hook decides about exit), returns every buffer it holds and fails the message."""
    w, h = 96, 64
    errors, signals = [], []
    tv = _make_tv(rt, [(64, 40), (48, 32)], errors, signals)
    rt.memory_unused.restype = C.c_int
    f = natural_i420(w, h, 7)
    m0 = _push(rt, tv, f, w, h, 1000, interlaced=0, fps=(60000, 1001))
    assert host.ottgpu_transvideo_process(C.byref(tv), m0) == 0          # primes the window
    m1 = _push(rt, tv, f, w, h, 2000, interlaced=0, fps=(60000, 1001))
    # drain the pool: the next memory_take for a rendition buffer returns NULL
    hoard = []
    while True:
        p = rt.memory_take(tv.raw_video_pool, 64)
        if not p:
            break
        hoard.append(p)
    msgs_before = rt.memory_unused(tv.msg_pool)
    assert host.ottgpu_transvideo_process(C.byref(tv), m1) == -1
    assert errors and b"Out of Uncompressed Video Buffers" in errors[0][1]
    assert rt.memory_unused(tv.msg_pool) == msgs_before + 1             # the input message went back to its pool
    assert not rt.dataqueue_take_back(tv.encode_queue[0])               # nothing was emitted for the failed message
    for p in hoard:
        rt.memory_return(tv.raw_video_pool, p)
    # the channel is still usable afterwards
    m2 = _push(rt, tv, f, w, h, 3000, interlaced=0, fps=(60000, 1001))
    assert host.ottgpu_transvideo_process(C.byref(tv), m2) == 0
    got = _drain(rt, tv, tv.encode_queue[0], oracle.i420_size(64, 40))
    assert len(got) == 1 and np.array_equal(got[0]["data"], oracle.scale_i420(f, w, h, 64, 40))
    _drain(rt, tv, tv.encode_queue[1], oracle.i420_size(48, 32))
    host.ottgpu_transvideo_close(C.byref(tv))


# ------------------------------------------------------------------------------------------------ N3: signal-loss filler
class MON(C.Structure):                       # ott-packager_b200/host/monitor_gpu.h
    _fields_ = [("monitor_queue", C.c_void_p), ("prepare_queue", C.c_void_p), ("raw_video_pool", C.c_void_p),
                ("msg_pool", C.c_void_p), ("device_pool", C.c_void_p), ("gpu", C.c_int), ("running", C.POINTER(C.c_int)),
                ("reinitialize_decoder", C.POINTER(C.c_int)), ("signal", SIGNAL_FN), ("direct_error", SIGNAL_FN),
                ("user", C.c_void_p), ("monitor_ready", C.c_int), ("monitor_width", C.c_int), ("monitor_height", C.c_int),
                ("monitor_num", C.c_int), ("monitor_den", C.c_int), ("monitor_tff", C.c_int), ("monitor_interlaced", C.c_int),
                ("monitor_aspect_num", C.c_int), ("monitor_aspect_den", C.c_int), ("monitor_anchor_dts", C.c_int64),
                ("monitor_anchor_pts", C.c_int64), ("total_dead_frames", C.c_int64),
                ("total_video_monitor_dead_time", C.c_int64), ("expected_dead_frames", C.c_double),
                ("dead_frames_triggered", C.c_int), ("saved_video_frame", C.c_void_p), ("saved_is_device", C.c_int),
                ("saved_bytes", C.c_size_t), ("frames_passed", C.c_int64), ("filler_frames", C.c_int64)]


def _make_mon(rt, tv, errors, signals):
    mon = MON()
    mon.monitor_queue = rt.dataqueue_create()
    mon.prepare_queue = tv.prepare_queue
    mon.raw_video_pool, mon.msg_pool, mon.gpu = tv.raw_video_pool, tv.msg_pool, 0
    mon._running = C.c_int(1)
    mon.running = C.pointer(mon._running)
    mon._reinit = C.c_int(0)
    mon.reinitialize_decoder = C.pointer(mon._reinit)
    mon._err = SIGNAL_FN(lambda u, code, m: errors.append((code, m)))
    mon._sig = SIGNAL_FN(lambda u, code, m: signals.append((code, m)))
    mon.direct_error, mon.signal = mon._err, mon._sig
    return mon


def test_monitor_filler_frames_follow_the_reference(lib, og, oracle, rt, host):
    """transvideo.c:1726-1812 / 1839-1884: after one second without input the saved frame is repeated at the stream's frame
    rate with pts/dts extrapolated from the last real frame; the next real frame carries source_discontinuity."""
    host.ottgpu_monitor_frame.argtypes = [C.POINTER(MON), C.POINTER(Msg)]
    host.ottgpu_monitor_dead_time.argtypes = [C.POINTER(MON)]
    host.ottgpu_monitor_close.argtypes = [C.POINTER(MON)]
    w, h = 96, 64
    errors, signals = [], []
    tv = _make_tv(rt, [(64, 40)], errors, signals)
    mon = _make_mon(rt, tv, errors, signals)
    frames = [natural_i420(w, h, 7, shift=3.0 * i) for i in range(3)]
    m = _push(rt, tv, frames[0], w, h, 900000, interlaced=0, fps=(30000, 1001))
    assert host.ottgpu_monitor_frame(C.byref(mon), m) == 0
    assert host.ottgpu_monitor_dead_time(C.byref(mon)) == 29          # int(29.97) + 0 drift correction
    assert host.ottgpu_monitor_dead_time(C.byref(mon)) == 29          # expected 29.97 vs 29 queued: drift 0.97 < 1, no correction yet
    m2 = _push(rt, tv, frames[1], w, h, 900000 + 3003 * 70, interlaced=0, fps=(30000, 1001))
    assert host.ottgpu_monitor_frame(C.byref(mon), m2) == 0
    m3 = _push(rt, tv, frames[2], w, h, 900000 + 3003 * 71, interlaced=0, fps=(30000, 1001))
    assert host.ottgpu_monitor_frame(C.byref(mon), m3) == 0
    assert mon._reinit.value == 1 and signals and signals[0][0] == 0x18 and b"Filler Video Frames" in signals[0][1]
    got = []
    while True:
        q = rt.dataqueue_take_back(tv.prepare_queue)
        if not q:
            break
        c = q.contents
        got.append((c.pts, c.dts, c.source_discontinuity, np.ctypeslib.as_array(C.cast(c.buffer, C.POINTER(C.c_uint8)), (frames[0].nbytes,)).copy()))
        rt.memory_return(tv.raw_video_pool, c.buffer)
        rt.memory_return(tv.msg_pool, q)
    assert len(got) == 1 + 58 + 2
    delta = 90000.0 / (30000 / 1001)
    for k in range(1, 59):                                            # filler frames: saved picture, extrapolated stamps
        assert got[k][0] == int(900000 + k * delta) and got[k][1] == int(900000 - 3003 + k * delta) and got[k][2] == 0
        assert np.array_equal(got[k][3], frames[0])
    assert got[59][2] == 1 and np.array_equal(got[59][3], frames[1])  # first real frame after the gap: discontinuity
    assert got[60][2] == 0 and np.array_equal(got[60][3], frames[2])
    assert mon.filler_frames == 58 and not errors
    host.ottgpu_monitor_close(C.byref(mon))


def test_monitor_keeps_a_device_stream_on_the_device(lib, og, oracle, rt, host):
    """A device-resident stream (converter output tagged OTTGPU_MSG_FLAG_DEVICE): the saved frame and the filler frames
    live in the stream's device pool, produced by device-to-device copies, and run through the channel like any frame."""
    host.ottgpu_monitor_frame.argtypes = [C.POINTER(MON), C.POINTER(Msg)]
    host.ottgpu_monitor_dead_time.argtypes = [C.POINTER(MON)]
    host.ottgpu_monitor_close.argtypes = [C.POINTER(MON)]
    w, h = 96, 64
    errors, signals = [], []
    tv = _make_tv(rt, [(64, 40)], errors, signals)
    nbytes = lib.frame_bytes(og.FMT_I420, w, h)
    pool = og.Pool(lib, 0, og.MEM_DEVICE, 80, nbytes)
    tv.input_pool = pool.handle
    mon = _make_mon(rt, tv, errors, signals)
    mon.device_pool = pool.handle
    f = natural_i420(w, h, 7)
    buf = pool.take()
    lib.to_device(buf, f)
    m = C.cast(rt.memory_take(tv.msg_pool, C.sizeof(Msg)), C.POINTER(Msg))
    C.memset(m, 0, C.sizeof(Msg))
    m.contents.buffer, m.contents.buffer_size, m.contents.flags = buf, nbytes, 0x4f540001
    m.contents.pts, m.contents.dts, m.contents.width, m.contents.height = 3003, 0, w, h
    m.contents.fps_num, m.contents.fps_den = 60000, 1001              # progressive-flagged: pass-through, one frame late
    assert host.ottgpu_monitor_frame(C.byref(mon), m) == 0
    assert mon.saved_is_device == 1
    assert host.ottgpu_monitor_dead_time(C.byref(mon)) == 59
    n = 0
    while True:                                                       # prepare/scale stage consumes real + filler frames
        q = rt.dataqueue_take_back(tv.prepare_queue)
        if not q:
            break
        assert q.contents.flags & 0xffff00ff == 0x4f540001
        assert host.ottgpu_transvideo_process(C.byref(tv), q) == 0
        n += 1
    assert n == 60 and not errors
    got = _drain(rt, tv, tv.encode_queue[0], oracle.i420_size(64, 40))
    want = oracle.scale_i420(f, w, h, 64, 40)
    assert len(got) == 59 and all(np.array_equal(g["data"], want) for g in got)
    host.ottgpu_transvideo_close(C.byref(tv))
    host.ottgpu_monitor_close(C.byref(mon))
    assert pool.unused() == 80                                        # every device frame (saved, filler, real) went back
    pool.close()


# ------------------------------------------------------------------------------------------------ N1: encoder hand-off
class ENC(C.Structure):                       # ott-packager_b200/host/encode_adapter.h
    _fields_ = [("raw_video_pool", C.c_void_p), ("msg_pool", C.c_void_p), ("surface_pool", C.c_void_p), ("gpu", C.c_int),
                ("width", C.c_int), ("height", C.c_int), ("fmt", C.c_int), ("pitch_align", C.c_int),
                ("staging_pool", C.c_void_p), ("staging", C.c_void_p), ("staging_bytes", C.c_size_t),
                ("frames", C.c_int64), ("d2h_frames", C.c_int64)]


class PLANES(C.Structure):
    _fields_ = [("plane", C.c_void_p * 3), ("stride", C.c_int * 3), ("planes", C.c_int)]


class NVSURF(C.Structure):
    _fields_ = [("device_ptr", C.c_void_p), ("chroma_offset", C.c_size_t), ("pitch", C.c_int), ("width", C.c_int),
                ("height", C.c_int), ("fmt", C.c_int), ("gpu", C.c_int)]


def _planes_to_i420(pl, w, h):
    out = []
    for k, (pw, ph) in enumerate(((w, h), ((w + 1) // 2, (h + 1) // 2), ((w + 1) // 2, (h + 1) // 2))):
        a = np.ctypeslib.as_array(C.cast(pl.plane[k], C.POINTER(C.c_uint8)), (ph * pl.stride[k],)).reshape(ph, pl.stride[k])[:, :pw]
        out.append(a.ravel())
    return np.concatenate(out)


def test_encoder_side_x264_x265_layout(lib, og, oracle, rt, host):
    """The rung arrives at the encoder thread in the contiguous w*h*3/2 buffer transvideo.c:1330-1338 (x264 memcpy #6) and
    960-967 (x265 plane pointers) read - from host pool buffers directly, from device surfaces through one pinned D2H."""
    host.ottgpu_encode_map_host.argtypes = [C.POINTER(ENC), C.POINTER(Msg), C.POINTER(PLANES)]
    host.ottgpu_encode_release.argtypes = [C.POINTER(ENC), C.POINTER(Msg)]
    host.ottgpu_encode_adapter_close.argtypes = [C.POINTER(ENC)]
    w, h = 192, 108
    outs = [(128, 72), (64, 36)]
    f = [natural_i420(w, h, 7, shift=2.0 * i) for i in range(3)]
    for device_outputs in (0, 1):
        errors, signals = [], []
        tv = _make_tv(rt, outs, errors, signals)
        pools = []
        if device_outputs:
            tv.device_outputs = 1
            for k, (dw, dh) in enumerate(outs):
                pools.append(og.Pool(lib, 0, og.MEM_DEVICE, 4, lib.rung_bytes(og.RungCfg(dw, dh, og.FMT_I420, 0))))
                tv.surface_pool[k] = pools[k].handle
        pool0 = rt.memory_unused(tv.raw_video_pool)
        for i in range(3):
            m = _push(rt, tv, f[i], w, h, 1000 * i, interlaced=0, fps=(60000, 1001))
            assert host.ottgpu_transvideo_process(C.byref(tv), m) == 0
        for k, (dw, dh) in enumerate(outs):
            ad = ENC()
            ad.raw_video_pool, ad.msg_pool = tv.raw_video_pool, tv.msg_pool
            ad.surface_pool = pools[k].handle if device_outputs else None
            ad.gpu, ad.width, ad.height, ad.fmt, ad.pitch_align = 0, dw, dh, og.FMT_I420, 0
            n = 0
            while True:
                m = rt.dataqueue_take_back(tv.encode_queue[k])
                if not m:
                    break
                pl = PLANES()
                assert host.ottgpu_encode_map_host(C.byref(ad), m, C.byref(pl)) == 0
                assert pl.planes == 3 and list(pl.stride) == [dw, (dw + 1) // 2, (dw + 1) // 2]      # x265 strides, 962-964
                if not device_outputs:                                                              # x265: pointers INTO the buffer
                    assert pl.plane[0] == m.contents.buffer and pl.plane[1] == m.contents.buffer + dw * dh
                assert np.array_equal(_planes_to_i420(pl, dw, dh), oracle.scale_i420(f[n], w, h, dw, dh))
                host.ottgpu_encode_release(C.byref(ad), m)
                n += 1
            assert n == 2 and ad.d2h_frames == (2 if device_outputs else 0)
            host.ottgpu_encode_adapter_close(C.byref(ad))
        assert not errors and rt.memory_unused(tv.raw_video_pool) == pool0
        host.ottgpu_transvideo_close(C.byref(tv))
        for p in pools:
            assert p.unused() == 4
            p.close()


def test_encoder_side_nvenc_surface_descriptor(lib, og, oracle, rt, host):
    """NVENC mode: the message carries a pitch-aligned NV12 device surface; the descriptor is what resource registration
    needs, and the surface holds interleave(I420 target-A result) - the bytes transvideo.c:532-557 builds on the CPU
    before av_hwframe_transfer_data (587)."""
    host.ottgpu_encode_nvenc_surface.argtypes = [C.POINTER(ENC), C.POINTER(Msg), C.POINTER(NVSURF)]
    host.ottgpu_encode_release.argtypes = [C.POINTER(ENC), C.POINTER(Msg)]
    host.ottgpu_nvenc_probe.argtypes = [C.c_char_p, C.c_size_t]
    w, h, dw, dh = 192, 108, 128, 72
    errors, signals = [], []
    tv = _make_tv(rt, [(dw, dh)], errors, signals)
    tv.device_outputs, tv.out_fmt = 1, og.FMT_NV12
    rung = og.RungCfg(dw, dh, og.FMT_NV12, 0)
    pool = og.Pool(lib, 0, og.MEM_DEVICE, 4, lib.rung_bytes(rung))
    tv.surface_pool[0] = pool.handle
    f = [natural_i420(w, h, 7, shift=2.0 * i) for i in range(2)]
    for i in range(2):
        assert host.ottgpu_transvideo_process(C.byref(tv), _push(rt, tv, f[i], w, h, 1000 * i, interlaced=0, fps=(60000, 1001))) == 0
    ad = ENC()
    ad.raw_video_pool, ad.msg_pool, ad.surface_pool = tv.raw_video_pool, tv.msg_pool, pool.handle
    ad.gpu, ad.width, ad.height, ad.fmt, ad.pitch_align = 0, dw, dh, og.FMT_NV12, 0
    m = rt.dataqueue_take_back(tv.encode_queue[0])
    s = NVSURF()
    assert host.ottgpu_encode_nvenc_surface(C.byref(ad), m, C.byref(s)) == 0
    assert (s.device_ptr, s.pitch, s.width, s.height, s.fmt, s.gpu, s.chroma_offset) == (m.contents.buffer, dw, dw, dh, og.FMT_NV12, 0, dw * dh)
    want = oracle.i420_to_nv12(oracle.scale_i420(f[0], w, h, dw, dh), dw, dh)
    assert np.array_equal(lib.from_device(s.device_ptr, lib.rung_bytes(rung)), want)
    host.ottgpu_encode_release(C.byref(ad), m)
    buf = C.create_string_buffer(256)
    assert host.ottgpu_nvenc_probe(buf, 256) in (0, 1) and b"libnvidia-encode" in buf.value
    host.ottgpu_transvideo_close(C.byref(tv))
    assert pool.unused() == 4 and not errors
    pool.close()


# ------------------------------------------------------------------------------------------------ N4: thumbnail JPEG thread
class Thumbnailer(C.Structure):               # ott-packager_b200/host/thumbnail_gpu.h
    _fields_ = [("thumbnail_queue", C.c_void_p), ("msg_pool", C.c_void_p), ("gpu", C.c_int), ("identity", C.c_int),
                ("directory", C.c_char_p), ("qscale", C.c_int), ("running", C.POINTER(C.c_int)), ("encoder", C.c_void_p),
                ("file_buffer", C.c_void_p), ("file_capacity", C.c_size_t), ("written", C.c_int64), ("failed", C.c_int64),
                ("last_bytes", C.c_size_t)]


class ScaleStruct(C.Structure):               # transvideo.c:124-127
    _fields_ = [("output_data", C.c_void_p * 4), ("output_stride", C.c_int * 4)]


def test_thumbnail_thread_writes_the_jpeg_on_the_gpu(lib, og, oracle, rt, host, tmp_path):
    """video_thumbnail_thread (transvideo.c:200-262) with save_frame_as_jpeg on the GPU: the scale_struct of the hand-off
    (one allocation, strides 176/88/88, transvideo.c:2176-2183) comes off the reference's own queue, the file appears as
    thumbnail<identity>.jpg, picture and struct are freed, the message goes back to the pool."""
    import threading
    libc = C.CDLL(None)
    libc.malloc.restype, libc.malloc.argtypes = C.c_void_p, [C.c_size_t]
    host.ottgpu_thumbnail_process.argtypes = [C.POINTER(Thumbnailer), C.POINTER(Msg)]
    host.ottgpu_video_thumbnail_thread.argtypes = [C.c_void_p]
    host.ottgpu_video_thumbnail_thread.restype = C.c_void_p
    host.ottgpu_thumbnailer_close.argtypes = [C.POINTER(Thumbnailer)]
    w, h = 176, 144
    th = Thumbnailer()
    th.thumbnail_queue = rt.dataqueue_create()
    th.msg_pool = rt.memory_create(16, C.sizeof(Msg))
    th.gpu, th.identity, th.directory, th.qscale = 0, 7, str(tmp_path).encode(), 0
    run = C.c_int(1)
    th.running = C.pointer(run)
    free0 = rt.memory_unused(th.msg_pool)

    def hand_off(frame):
        pic = libc.malloc(frame.nbytes)                              # av_image_alloc(176, 144, YUV420P, align 1): one allocation
        C.memmove(pic, frame.ctypes.data, frame.nbytes)
        ss = C.cast(libc.malloc(C.sizeof(ScaleStruct)), C.POINTER(ScaleStruct))
        ss.contents.output_data[0], ss.contents.output_data[1], ss.contents.output_data[2] = pic, pic + w * h, pic + w * h + 88 * 72
        ss.contents.output_data[3] = None
        ss.contents.output_stride[0], ss.contents.output_stride[1], ss.contents.output_stride[2], ss.contents.output_stride[3] = w, 88, 88, 0
        m = C.cast(rt.memory_take(th.msg_pool, C.sizeof(Msg)), C.POINTER(Msg))
        C.memset(m, 0, C.sizeof(Msg))
        m.contents.buffer = C.cast(ss, C.c_void_p)
        return m

    f1, f2 = natural_i420(w, h, 12), natural_i420(w, h, 13, shift=40.0)
    assert host.ottgpu_thumbnail_process(C.byref(th), hand_off(f1)) == 0
    path = tmp_path / "thumbnail7.jpg"
    assert path.read_bytes() == oracle.jpeg_encode(f1, w, h, 3)
    assert th.written == 1 and th.last_bytes == path.stat().st_size and rt.memory_unused(th.msg_pool) == free0
    # as the pthread body: polls the queue like the reference does
    t = threading.Thread(target=host.ottgpu_video_thumbnail_thread, args=(C.addressof(th),))
    t.start()
    rt.dataqueue_put_front(th.thumbnail_queue, hand_off(f2))
    for _ in range(300):
        if th.written == 2:
            break
        time.sleep(0.01)
    run.value = 0
    t.join(5)
    assert th.written == 2 and th.failed == 0
    assert path.read_bytes() == oracle.jpeg_encode(f2, w, h, 3)
    assert rt.memory_unused(th.msg_pool) == free0
    host.ottgpu_thumbnailer_close(C.byref(th))
