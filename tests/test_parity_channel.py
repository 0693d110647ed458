"""Channel / batch behaviour through the C ABI: yadif parity (vs oracle/ref_yadif.c, itself PARITY-UNPINNED, see its
header), frame scheduling of "yadif=0:-1:0|1" (transvideo.c:2036-2052, 2132-2141), sideband association across the
one-frame latency (transvideo.c:2103-2152), discontinuity reset (1959-1970), thumbnail cadence (2159-2186), the fused
multi-channel launch, and device-resident inputs/outputs (zero-copy surfaces)."""
import ctypes
import numpy as np
import pytest
from conftest import noise_i420, noise_p010, natural_i420, interlaced_sequence, split_i420


@pytest.mark.parametrize("size", [(64, 48), (130, 74), (854, 480), (36, 22), (320, 182), (328, 46)])
@pytest.mark.parametrize("tff", [1, 0])
def test_yadif_frame_vs_oracle(lib, og, oracle, size, tff):
    w, h = size
    fr = interlaced_sequence(w, h, 3, seed=w, tff=bool(tff))
    got = lib.yadif_frame(fr[0], fr[1], fr[2], og.FMT_I420, w, h, 1, tff)
    assert np.array_equal(got, oracle.yadif_frame(fr[0], fr[1], fr[2], w, h, 1, tff))
    n = [noise_i420(w, h, s) for s in (1, 2, 3)]
    got = lib.yadif_frame(n[0], n[1], n[2], og.FMT_I420, w, h, 1, tff)
    assert np.array_equal(got, oracle.yadif_frame(n[0], n[1], n[2], w, h, 1, tff))


@pytest.mark.parametrize("size", [(64, 48), (130, 74), (192, 108)])
def test_yadif_extreme_values(lib, og, oracle, size):
    """0 / 255 checker noise drives every |a-b| to 0 or 255 and every clamp to its limits (the packed 16-bit lanes of the
    fast path must never borrow or carry into their neighbours); flat frames exercise the all-equal paths."""
    w, h = size
    rng = np.random.default_rng(w)
    n = oracle.i420_size(w, h)
    for tff in (1, 0):
        fr = [(rng.integers(0, 2, n) * 255).astype(np.uint8) for _ in range(3)]
        got = lib.yadif_frame(fr[0], fr[1], fr[2], og.FMT_I420, w, h, 1, tff)
        assert np.array_equal(got, oracle.yadif_frame(fr[0], fr[1], fr[2], w, h, 1, tff))
        # two-level frames with long runs: large spatial-check terms next to zero temporal differences
        fr = [np.repeat((rng.integers(0, 2, (n + 15) // 16) * 255).astype(np.uint8), 16)[:n] for _ in range(3)]
        got = lib.yadif_frame(fr[0], fr[1], fr[1], og.FMT_I420, w, h, 1, tff)
        assert np.array_equal(got, oracle.yadif_frame(fr[0], fr[1], fr[1], w, h, 1, tff))
    for level in (0, 255, 128):
        flat = np.full(n, level, np.uint8)
        assert np.array_equal(lib.yadif_frame(flat, flat, flat, og.FMT_I420, w, h, 1, 1), flat)


@pytest.mark.gpu
@pytest.mark.parametrize("tff", [1, 0])
def test_yadif_full_size_1080i(real_lib, og, oracle, tff):
    w, h = 1920, 1080
    fr = interlaced_sequence(w, h, 3, seed=11, tff=bool(tff))
    got = real_lib.yadif_frame(fr[0], fr[1], fr[2], og.FMT_I420, w, h, 1, tff)
    assert np.array_equal(got, oracle.yadif_frame(fr[0], fr[1], fr[2], w, h, 1, tff))


def test_yadif_frame_16bit(lib, og, oracle):
    w, h = 130, 74
    fr = interlaced_sequence(w, h, 3, seed=9, bps=2)
    got = lib.yadif_frame(fr[0], fr[1], fr[2], og.FMT_I420P10, w, h, 1, 1)
    assert np.array_equal(got, oracle.yadif_frame(fr[0], fr[1], fr[2], w, h, 1, 1, bps=2))


def _oracle_ladder(oracle, frame, w, h, rungs, og):
    out = []
    for (dw, dh, fmt) in rungs:
        i420 = oracle.scale_i420(frame, w, h, dw, dh)
        out.append(oracle.i420_to_nv12(i420, dw, dh) if fmt == og.FMT_NV12 else i420)
    return out


def test_channel_stream_cfg1_shape(lib, og, oracle):
    """cfg1 in miniature: interlaced TFF source -> yadif(all) -> {full, 2/3, 1/3} I420 rungs, thumbnails every 3rd frame."""
    w, h = 192, 108
    rungs = [(192, 108, og.FMT_I420), (128, 72, og.FMT_I420), (64, 36, og.FMT_I420)]
    ch = og.Channel(lib, og.make_cfg(w, h, rungs, deint=og.DEINT_ALL, thumb=(44, 36, 3)))
    frames = interlaced_sequence(w, h, 8, seed=3)
    ref = oracle.YadifStream(w, h)
    produced = 0
    for i, f in enumerate(frames):
        r = ch.submit(f, interlaced=1, tff=1, pts=1000 + i, dts=900 + i, opaque=0x1000 + i)
        want = ref.push(f, 1, 1, tag=i)
        if want is None:
            assert r is None and i == 0
            continue
        produced += 1
        deint, tag = want
        assert r["opaque"] == 0x1000 + tag and r["pts"] == 1000 + tag and r["dts"] == 900 + tag
        for got, exp in zip(r["rungs"], _oracle_ladder(oracle, deint, w, h, rungs, og)):
            assert np.array_equal(got, exp)
        if produced % 3 == 0:
            assert np.array_equal(r["thumbnail"], oracle.scale_i420(deint, w, h, 44, 36))
        else:
            assert r["thumbnail"] is None
    ch.close()


def test_channel_reset_and_bff(lib, og, oracle):
    w, h = 96, 64
    rungs = [(64, 40, og.FMT_NV12)]                      # no full-size rung: yadif goes through the scratch frame
    ch = og.Channel(lib, og.make_cfg(w, h, rungs, deint=og.DEINT_ALL))
    frames = interlaced_sequence(w, h, 6, seed=4, tff=False)
    ref = oracle.YadifStream(w, h)
    for i, f in enumerate(frames):
        if i == 3:                                       # source_discontinuity (transvideo.c:1959-1970)
            ch.reset()
            ref.reset()
        r = ch.submit(f, interlaced=1, tff=0)
        want = ref.push(f, 1, 0)
        assert (r is None) == (want is None)
        if want is not None:
            assert np.array_equal(r["rungs"][0], _oracle_ladder(oracle, want[0], w, h, rungs, og)[0])
    ch.close()


def test_channel_flagged_only_passes_progressive_through(lib, og, oracle):
    """cfg2 in miniature: progressive-flagged 60000/1001 => "yadif=0:-1:1": frames pass through one frame late."""
    w, h = 192, 108
    rungs = [(192, 108, og.FMT_NV12), (128, 72, og.FMT_NV12), (96, 54, og.FMT_NV12), (64, 36, og.FMT_NV12), (42, 24, og.FMT_NV12)]
    ch = og.Channel(lib, og.make_cfg(w, h, rungs, deint=og.DEINT_FLAGGED))
    frames = [natural_i420(w, h, 7, shift=3.0 * i) for i in range(4)]
    assert ch.submit(frames[0], interlaced=0) is None
    for i in range(1, 4):
        r = ch.submit(frames[i], interlaced=0)
        for got, exp in zip(r["rungs"], _oracle_ladder(oracle, frames[i - 1], w, h, rungs, og)):
            assert np.array_equal(got, exp)
    # a flagged frame in the same stream is deinterlaced
    il = interlaced_sequence(w, h, 3, seed=8)
    ch.reset()
    ref = oracle.YadifStream(w, h, deint_flagged_only=True)
    for f, flag in ((frames[0], 0), (il[1], 1), (il[2], 1), (frames[1], 0)):
        r = ch.submit(f, interlaced=flag, tff=1)
        want = ref.push(f, flag, 1)
        assert (r is None) == (want is None)
        if want is not None:
            assert np.array_equal(r["rungs"][1], _oracle_ladder(oracle, want[0], w, h, rungs[1:2], og)[0])
    ch.close()


def test_bypass_is_plain_sws_scale(lib, og, oracle):
    w, h = 130, 74
    rungs = [(130, 74, og.FMT_I420), (66, 38, og.FMT_I420)]
    ch = og.Channel(lib, og.make_cfg(w, h, rungs, deint=og.DEINT_BYPASS, target=og.TARGET_B))
    for s in range(3):
        f = noise_i420(w, h, 50 + s)
        r = ch.submit(f, opaque=s + 1)
        assert r["opaque"] == s + 1
        assert np.array_equal(r["rungs"][0], f)
        assert np.array_equal(r["rungs"][1], oracle.scale_i420(f, w, h, 66, 38, oracle.BICUBIC, oracle.TARGET_B))
    ch.close()


def test_batch_mixed_channels_one_launch(lib, og, oracle):
    """Several channels with different geometry, depth and deinterlace policy advanced by ONE fused ladder launch."""
    specs = [
        dict(w=192, h=108, fmt=og.FMT_I420, rungs=[(192, 108, og.FMT_I420), (128, 72, og.FMT_I420)], deint=og.DEINT_ALL),
        dict(w=192, h=108, fmt=og.FMT_I420, rungs=[(128, 72, og.FMT_NV12), (64, 36, og.FMT_NV12)], deint=og.DEINT_FLAGGED),
        dict(w=256, h=144, fmt=og.FMT_P010, rungs=[(256, 144, og.FMT_P010), (128, 72, og.FMT_P010)], deint=og.DEINT_BYPASS),
        dict(w=130, h=74, fmt=og.FMT_I420, rungs=[(66, 38, og.FMT_I420)], deint=og.DEINT_ALL),
    ]
    chans = [og.Channel(lib, og.make_cfg(s["w"], s["h"], s["rungs"], src_fmt=s["fmt"], deint=s["deint"],
                                         filt=og.FILTER_LANCZOS if s["fmt"] == og.FMT_P010 else og.FILTER_BICUBIC))
             for s in specs]
    batch = og.Batch(lib, chans)
    seqs = [interlaced_sequence(192, 108, 4, seed=1), [natural_i420(192, 108, 7, shift=i) for i in range(4)],
            [noise_p010(256, 144, 60 + i) for i in range(4)], interlaced_sequence(130, 74, 4, seed=2)]
    refs = [oracle.YadifStream(192, 108), None, None, oracle.YadifStream(130, 74)]
    for t in range(4):
        frames = [seqs[k][t] for k in range(4)]
        if t == 2:
            frames[3] = None                              # a channel may sit a step out
        res = batch.submit(frames, interlaced=1, tff=1)
        assert batch.last_launches >= 1
        # channel 0: yadif all
        want = refs[0].push(frames[0], 1, 1)
        assert (res[0] is None) == (want is None)
        if want is not None:
            assert np.array_equal(res[0]["rungs"][0], want[0])
            assert np.array_equal(res[0]["rungs"][1], oracle.scale_i420(want[0], 192, 108, 128, 72))
        # channel 1: flagged-only, but flagged interlaced here => deinterlaced too; compare with stream semantics
        if t == 0:
            assert res[1] is None
        # channel 2: P010 bypass
        assert np.array_equal(res[2]["rungs"][0], frames[2])
        assert np.array_equal(res[2]["rungs"][1], oracle.scale_p010(frames[2], 256, 144, 128, 72, oracle.LANCZOS))
        # channel 3
        if frames[3] is None:
            assert res[3] is None
        else:
            want = refs[3].push(frames[3], 1, 1)
            assert (res[3] is None) == (want is None)
            if want is not None:
                assert np.array_equal(res[3]["rungs"][0], oracle.scale_i420(want[0], 130, 74, 66, 38))
    batch.close()
    for c in chans:
        c.close()


def test_device_resident_frames_and_pools(lib, og, oracle):
    """Zero-copy mode: source frames and rung surfaces live in device pools (mempool.h discipline); the library
    reports which inputs it has stopped referencing (memory_return point of transvideo.c:2410)."""
    w, h = 192, 108
    rungs = [(192, 108, og.FMT_NV12, 256), (128, 72, og.FMT_NV12, 256)]
    cfg = og.make_cfg(w, h, rungs, deint=og.DEINT_FLAGGED)
    ch = og.Channel(lib, cfg)
    src_pool = og.Pool(lib, 0, og.MEM_DEVICE, 4, lib.frame_bytes(og.FMT_I420, w, h))
    out_pools = [og.Pool(lib, 0, og.MEM_DEVICE, 2, ch.rung_bytes[r]) for r in range(2)]
    assert src_pool.unused() == 4
    frames = [noise_i420(w, h, 70 + i) for i in range(5)]
    held = []
    for i, f in enumerate(frames):
        d = src_pool.take()
        assert d
        held.append(d)
        lib.to_device(d, f)
        dsts = [p.take() for p in out_pools]
        r = ch.submit(d, interlaced=0, device_dst=dsts)
        if r is None:
            assert i == 0
        else:
            exp = frames[i - 1]
            for k, (dw, dh, fmt, al) in enumerate(rungs):
                got = lib.from_device(dsts[k], ch.rung_bytes[k])
                lay = lib.rung_layout(cfg.rung[k])
                ref = oracle.i420_to_nv12(oracle.scale_i420(exp, w, h, dw, dh), dw, dh)
                y = got[lay[0][0]: lay[0][0] + lay[0][1] * dh].reshape(dh, lay[0][1])[:, :dw]
                uv = got[lay[1][0]: lay[1][0] + lay[1][1] * (dh // 2)].reshape(dh // 2, lay[1][1])[:, :dw]
                assert np.array_equal(y.ravel(), ref[: dw * dh]) and np.array_equal(uv.ravel(), ref[dw * dh:])
            for p_ in r["released"]:
                assert p_ in held
                held.remove(p_)
                assert src_pool.give_back(p_) == 0
        for p, d_ in zip(out_pools, dsts):
            assert p.give_back(d_) == 0
    assert len(held) <= 3 and src_pool.unused() == 4 - len(held)
    assert src_pool.give_back(held[0]) == 0 and src_pool.give_back(held[0]) < 0      # double return is refused
    ch.close()
    src_pool.close()
    for p in out_pools:
        p.close()


def test_bad_arguments_are_refused(lib, og):
    with pytest.raises(og.OttGpuError):
        og.Channel(lib, og.make_cfg(192, 108, [(128, 72, og.FMT_P010)]))                     # depth mismatch
    with pytest.raises(og.OttGpuError):
        og.Channel(lib, og.make_cfg(192, 108, [(128, 72, og.FMT_P010)], src_fmt=og.FMT_P010, deint=og.DEINT_ALL))
    with pytest.raises(og.OttGpuError):
        og.Channel(lib, og.make_cfg(192, 108, [(128, 72, og.FMT_I420)] * 9))
    with pytest.raises(og.OttGpuError):
        og.Channel(lib, og.make_cfg(192, 108, [(0, 72, og.FMT_I420)]))


def test_eight_rungs_and_odd_geometry(lib, og, oracle):
    """MAX_TRANS_OUTPUTS = 8 renditions (include/fillet.h:57) incl. ragged sizes, one launch."""
    w, h = 322, 182
    rungs = [(322, 182, og.FMT_I420), (214, 120, og.FMT_I420), (160, 90, og.FMT_I420), (107, 61, og.FMT_I420),
             (86, 48, og.FMT_I420), (64, 36, og.FMT_I420), (50, 28, og.FMT_I420), (42, 24, og.FMT_I420)]
    ch = og.Channel(lib, og.make_cfg(w, h, rungs, deint=og.DEINT_BYPASS))
    f = noise_i420(w, h, 77)
    r = ch.submit(f)
    for got, (dw, dh, _) in zip(r["rungs"], rungs):
        assert np.array_equal(got, oracle.scale_i420(f, w, h, dw, dh)), (dw, dh)
    ch.close()


@pytest.mark.gpu
def test_cfg5_4k_channel_with_thumbnail_full_size(real_lib, og, oracle):
    """BASELINE config 5's 4K channel: 2160p 8-bit -> {2160p,1080p,720p,360p} + 176x144 thumbnail (the thumbnail's
    88/60-tap filters take the single-buffered launch)."""
    w, h = 3840, 2160
    rungs = [(3840, 2160, og.FMT_I420), (1920, 1080, og.FMT_I420), (1280, 720, og.FMT_I420), (640, 360, og.FMT_I420)]
    ch = og.Channel(real_lib, og.make_cfg(w, h, rungs, deint=og.DEINT_FLAGGED, thumb=(176, 144, 1)))
    f0, f1 = noise_i420(w, h, 5), natural_i420(w, h, 7)
    assert ch.submit(f0, interlaced=0) is None
    r = ch.submit(f1, interlaced=0)
    for got, (dw, dh, _) in zip(r["rungs"], rungs):
        assert np.array_equal(got, oracle.scale_i420(f0, w, h, dw, dh)), (dw, dh)
    assert np.array_equal(r["thumbnail"], oracle.scale_i420(f0, w, h, 176, 144))
    ch.close()


@pytest.mark.gpu
def test_cfg4_many_channels_one_batch_full_size(real_lib, og, oracle):
    """BASELINE config 4's per-GPU slice: 8 concurrent 1080i channels, yadif + 4 rungs, one launch per stage."""
    w, h = 1920, 1080
    rungs = [(1920, 1080, og.FMT_I420), (1280, 720, og.FMT_I420), (960, 540, og.FMT_I420), (640, 360, og.FMT_I420)]
    chans = [og.Channel(real_lib, og.make_cfg(w, h, rungs, deint=og.DEINT_ALL)) for _ in range(8)]
    batch = og.Batch(real_lib, chans)
    seq = interlaced_sequence(w, h, 3, seed=31)
    for t in range(3):
        res = batch.submit([np.roll(seq[t], 1000 * c) for c in range(8)], interlaced=1, tff=1)
    assert batch.last_launches == 2                          # one yadif launch + one ladder launch for all 8 channels
    for c in (0, 5):
        fr = [np.roll(seq[t], 1000 * c) for t in range(3)]
        deint = oracle.yadif_frame(fr[0], fr[1], fr[2], w, h, 1, 1)
        assert np.array_equal(res[c]["rungs"][0], deint)
        for k in (1, 3):
            assert np.array_equal(res[c]["rungs"][k], oracle.scale_i420(deint, w, h, rungs[k][0], rungs[k][1]))
    batch.close()
    for c in chans:
        c.close()


@pytest.mark.gpu
def test_1080p_thumbnail_slot_full_size(real_lib, og, oracle):
    """The thumbnail launch class at the reference's real geometry (1080p -> 176x144, transvideo.c:2155-2174): its tiles do
    not fit the double-buffered budget, so this runs the single-buffer TMA path of the kernel."""
    w, h = 1920, 1080
    ch = og.Channel(real_lib, og.make_cfg(w, h, [(1280, 720, og.FMT_NV12, 256)], deint=og.DEINT_BYPASS, thumb=(176, 144, 1)))
    for seed in (3, 4):
        f = natural_i420(w, h, 7, shift=seed) if seed == 3 else noise_i420(w, h, seed)
        r = ch.submit(f)
        assert np.array_equal(r["thumbnail"], oracle.scale_i420(f, w, h, 176, 144))
    ch.close()


def test_channels_driven_from_concurrent_threads(lib, og, oracle):
    """The reference runs one prepare/scale thread pair per channel (transvideo.c:3044-3099); a process hosting several
    channels drives each ottgpu_channel from its own thread.  ctypes drops the GIL during the calls, so the submits below
    really overlap; every channel must still produce exactly what the oracle says."""
    import threading
    specs = [(192, 108, [(128, 72, og.FMT_I420), (64, 36, og.FMT_NV12)], og.DEINT_ALL),
             (130, 74, [(130, 74, og.FMT_I420), (66, 38, og.FMT_I420)], og.DEINT_BYPASS),
             (96, 64, [(64, 40, og.FMT_NV12)], og.DEINT_ALL),
             (192, 108, [(128, 72, og.FMT_I420), (64, 36, og.FMT_NV12)], og.DEINT_FLAGGED)]   # same plan as channel 0
    n_frames = 5
    inputs, results, errors = [], [None] * len(specs), []
    for k, (w, h, rungs, deint) in enumerate(specs):
        inputs.append(interlaced_sequence(w, h, n_frames, seed=20 + k) if deint != og.DEINT_BYPASS
                      else [noise_i420(w, h, 70 + k + i) for i in range(n_frames)])
    start = threading.Barrier(len(specs))

    def drive(k):
        try:
            w, h, rungs, deint = specs[k]
            ch = og.Channel(lib, og.make_cfg(w, h, rungs, deint=deint))
            start.wait()
            out = []
            for i, f in enumerate(inputs[k]):
                r = ch.submit(f, interlaced=1, tff=1, opaque=1000 + 100 * k + i)
                out.append(None if r is None else (r["opaque"], [np.array(x) for x in r["rungs"]]))
            ch.close()
            results[k] = out
        except Exception as e:                                   # surfaced in the main thread
            errors.append((k, repr(e)))
            try:
                start.abort()
            except Exception:
                pass

    threads = [threading.Thread(target=drive, args=(k,)) for k in range(len(specs))]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors, errors
    for k, (w, h, rungs, deint) in enumerate(specs):
        ref = None if deint == og.DEINT_BYPASS else oracle.YadifStream(w, h, deint_flagged_only=deint == og.DEINT_FLAGGED)
        for i, f in enumerate(inputs[k]):
            want = (f, i) if ref is None else ref.push(f, 1, 1, tag=i)
            got = results[k][i]
            assert (got is None) == (want is None)
            if want is None:
                continue
            assert got[0] == 1000 + 100 * k + want[1]
            for a, b in zip(got[1], _oracle_ladder(oracle, want[0], w, h, rungs, og)):
                assert np.array_equal(a, b)


def test_thumbnail_cadence_runs_through_a_discontinuity(lib, og, oracle):
    """The reference's thumbnail_count is a local of the prepare thread that the source_discontinuity block leaves alone
    (transvideo.c:1925, 1959-1970, 2160-2185): the Nth PRODUCED frame gets the thumbnail, resets or not."""
    w, h = 96, 64
    ch = og.Channel(lib, og.make_cfg(w, h, [(64, 40, og.FMT_I420)], deint=og.DEINT_ALL, thumb=(44, 36, 4)))
    frames = interlaced_sequence(w, h, 12, seed=9)
    produced, with_thumb = 0, []
    for i, f in enumerate(frames):
        if i == 3 or i == 7:
            ch.reset()
        r = ch.submit(f, interlaced=1, tff=1)
        if r is None:
            continue
        produced += 1
        if r["thumbnail"] is not None:
            with_thumb.append(produced)
    assert produced == 12 - 3                            # one priming frame at the start and after each reset
    assert with_thumb == [4, 8]                          # every 4th produced frame, counted across the resets
    ch.close()


def test_failed_submit_changes_nothing(lib, og, oracle):
    """ottgpu_batch_submit is transactional: a call that fails (here: a rung without a destination) leaves the yadif
    window, the upload slots and the thumbnail counter as they were, and the stream continues bit-exactly."""
    import ctypes as C
    w, h = 96, 64
    rungs = [(96, 64, og.FMT_I420), (64, 40, og.FMT_I420)]
    ch = og.Channel(lib, og.make_cfg(w, h, rungs, deint=og.DEINT_ALL, thumb=(44, 36, 2)))
    frames = interlaced_sequence(w, h, 7, seed=11)
    ref = oracle.YadifStream(w, h)
    produced = 0
    for i, f in enumerate(frames):
        if i in (2, 4):                                  # a broken call in between: second rung has no buffer
            fin, fout = og.FrameIn(), og.FrameOut()
            ch.fill_in(fin, frames[(i + 3) % 7], 1, 1, 0, 0, 0x77)
            bufs = ch.alloc_out(fout)
            fout.dst[1] = None
            rc = lib.dll.ottgpu_channel_submit(ch.handle, C.byref(fin), C.byref(fout))
            assert rc == og.ERR_ARG and not fout.thumbnail and not fout.produced
            assert lib.dll.ottgpu_channel_pending_opaque(ch.handle) == (0x1000 + i - 1)
        r = ch.submit(f, interlaced=1, tff=1, opaque=0x1000 + i)
        want = ref.push(f, 1, 1, tag=i)
        assert (r is None) == (want is None)
        assert lib.dll.ottgpu_channel_pending_opaque(ch.handle) == 0x1000 + i
        if want is None:
            continue
        produced += 1
        deint, tag = want
        assert r["opaque"] == 0x1000 + tag
        for got, exp in zip(r["rungs"], _oracle_ladder(oracle, deint, w, h, rungs, og)):
            assert np.array_equal(got, exp)
        assert (r["thumbnail"] is not None) == (produced % 2 == 0)
    ch.reset()
    assert not lib.dll.ottgpu_channel_pending_opaque(ch.handle)
    ch.close()


def test_filter_table_refuses_more_taps_than_the_callers_buffer_holds(real_lib, og):
    import ctypes as C
    pos = (C.c_int32 * 144)()
    coef = (C.c_int16 * (144 * 256))()
    assert real_lib.dll.ottgpu_filter_table(16384, 144, og.FILTER_BICUBIC, og.TARGET_A, 0, pos, coef) == og.ERR_ARG
    assert real_lib.dll.ottgpu_filter_table(1080, 144, og.FILTER_BICUBIC, og.TARGET_A, 1, pos, coef) == 30


def test_zero_copy_inputs_from_a_large_pool_never_stall(lib, og, oracle):
    """The reference's raw pool hands out a fresh address on every memory_take (mempool.c:196-198, 512 buffers,
    fillet.h:84).  Device-resident input frames cycling through 72 distinct pool buffers go through ONE set of tensor
    maps (three-dimensional, the buffer index is a coordinate): no per-address encode, no device-wide synchronisation,
    bit-exact rungs; the frame path stays fast (time bound) instead of re-encoding / syncing per frame."""
    import time
    w, h = 192, 108
    rungs = [(128, 72, og.FMT_I420), (96, 54, og.FMT_NV12)]
    ch = og.Channel(lib, og.make_cfg(w, h, rungs, deint=og.DEINT_ALL))
    nb = 72
    pool = og.Pool(lib, 0, og.MEM_DEVICE, nb, lib.frame_bytes(og.FMT_I420, w, h))
    bufs = [pool.take() for _ in range(nb)]
    assert len(set(bufs)) == nb
    frames = interlaced_sequence(w, h, 6, seed=31)
    ref = oracle.YadifStream(w, h)
    t_first, t_rest = None, 0.0
    for i in range(nb):
        f = frames[i % 6]
        lib.to_device(bufs[i], f)
        t0 = time.perf_counter()
        r = ch.submit(bufs[i], interlaced=1, tff=1)
        dt = time.perf_counter() - t0
        if i < 4:
            t_first = dt if t_first is None else max(t_first, dt)
        else:
            t_rest += dt
        want = ref.push(f, 1, 1)
        assert (r is None) == (want is None)
        if want is not None:
            for got, exp in zip(r["rungs"], _oracle_ladder(oracle, want[0], w, h, rungs, og)):
                assert np.array_equal(got, exp)
    # steady state: no call may cost anywhere near a device-wide sync + tensor-map re-encode per frame
    assert t_rest / (nb - 4) < 0.05
    ch.close()
    pool.close()


def test_batch_uploads_travel_as_one_copy(lib, og, oracle):
    """Host frames that are consecutive buffers of one pinned pool reach the batch's upload arena in a single
    cudaMemcpyAsync per step (the per-GPU H2D of N channels is one DMA, not N)."""
    w, h, n = 96, 64, 5
    rungs = [(64, 40, og.FMT_I420)]
    chans = [og.Channel(lib, og.make_cfg(w, h, rungs, deint=og.DEINT_FLAGGED)) for _ in range(n)]
    batch = og.Batch(lib, chans)
    nbytes = lib.frame_bytes(og.FMT_I420, w, h)
    hpool = og.Pool(lib, 0, og.MEM_HOST, 2 * n, nbytes)
    host = [[hpool.take() for _ in range(n)] for _ in range(2)]          # one phase = n consecutive pool buffers
    frames = [[natural_i420(w, h, 7, shift=1.0 * (3 * s + c)) for c in range(n)] for s in range(4)]
    outs = [np.zeros(chans[0].rung_bytes[0], np.uint8) for _ in range(n)]
    import ctypes as C
    for s in range(4):
        for c in range(n):
            C.memmove(host[s % 2][c], frames[s][c].ctypes.data, nbytes)
            batch.fin[c].data, batch.fin[c].mem = host[s % 2][c], og.MEM_HOST
            batch.fin[c].interlaced, batch.fin[c].tff = 0, 1
            batch.fout[c].dst[0], batch.fout[c].dst_mem[0] = outs[c].ctypes.data, og.MEM_HOST
        batch.submit_raw(og.SUBMIT_SYNC)
        assert batch.last_uploads == 1
        if s:
            for c in range(n):
                assert batch.fout[c].produced == 1
                assert np.array_equal(outs[c], oracle.scale_i420(frames[s - 1][c], w, h, 64, 40))
    batch.close()
    for c in chans:
        c.close()
    hpool.close()


def test_batch_readbacks_travel_as_one_copy_per_rung(lib, og, oracle):
    """Host-destined rungs: the surfaces of one rung of all channels share one device allocation, so destinations that are
    consecutive buffers of one pinned pool come back with a single copy per rung (x264/x265 hand-off, transvideo.c:1610-1624);
    scattered destinations (every second pool buffer, plain numpy memory) still arrive bit-exact through pitched / single copies."""
    import ctypes as C
    w, h, n = 96, 64, 5
    rungs = [(96, 64, og.FMT_I420), (64, 40, og.FMT_I420)]
    chans = [og.Channel(lib, og.make_cfg(w, h, rungs, deint=og.DEINT_FLAGGED)) for _ in range(n)]
    batch = og.Batch(lib, chans)
    nbytes = lib.frame_bytes(og.FMT_I420, w, h)
    rb = chans[0].rung_bytes
    pools = [og.Pool(lib, 0, og.MEM_HOST, 2 * n, rb[r]) for r in range(2)]
    packed = [[pools[r].take() for _ in range(2 * n)] for r in range(2)]
    loose = [np.zeros(rb[1], np.uint8) for _ in range(n)]
    frames = [[natural_i420(w, h, 11, shift=1.0 * (5 * s + c)) for c in range(n)] for s in range(5)]
    for s in range(5):
        layout = s % 3                      # 0: consecutive buffers, 1: every second buffer (pitched copy), 2: rung 1 in numpy memory
        for c in range(n):
            batch.fin[c].data, batch.fin[c].mem = frames[s][c].ctypes.data, og.MEM_HOST
            batch.fin[c].interlaced, batch.fin[c].tff = 0, 1
            for r in range(2):
                dst = packed[r][c] if layout == 0 else packed[r][2 * c]
                if layout == 2 and r == 1:
                    dst = loose[c].ctypes.data
                batch.fout[c].dst[r], batch.fout[c].dst_mem[r] = dst, og.MEM_HOST
        batch.submit_raw(og.SUBMIT_SYNC)
        if s == 0:
            continue
        assert batch.last_readbacks == (2 if layout < 2 else 1 + n)
        for c in range(n):
            assert batch.fout[c].produced == 1
            want = [frames[s - 1][c], oracle.scale_i420(frames[s - 1][c], w, h, 64, 40)]
            for r in range(2):
                got = np.empty(rb[r], np.uint8)
                C.memmove(got.ctypes.data, batch.fout[c].dst[r], rb[r])
                assert np.array_equal(got, want[r]), (s, c, r)
    batch.close()
    for c in chans:
        c.close()
    for p in pools:
        p.close()


def test_batch_of_unlike_channels_with_host_destinations(lib, og, oracle):
    """Channels of one batch may differ in source size and ladder; the shared per-rung staging allocation is sized for the
    largest surface and every channel's host-destined rungs still arrive bit-exact."""
    specs = [(96, 64, [(96, 64), (64, 40)]), (128, 72, [(64, 36), (48, 28), (32, 18)]), (96, 64, [(48, 32)])]
    chans = [og.Channel(lib, og.make_cfg(w, h, [(dw, dh, og.FMT_I420) for dw, dh in rungs], deint=og.DEINT_FLAGGED)) for w, h, rungs in specs]
    batch = og.Batch(lib, chans)
    frames = [[natural_i420(w, h, 21 + s, shift=3.0 * s + c) for c, (w, h, _) in enumerate(specs)] for s in range(3)]
    for s in range(3):
        res = batch.submit(frames[s], interlaced=0)
        if s == 0:
            assert all(r is None for r in res)
            continue
        for c, (w, h, rungs) in enumerate(specs):
            for got, (dw, dh) in zip(res[c]["rungs"], rungs):
                want = frames[s - 1][c] if (dw, dh) == (w, h) else oracle.scale_i420(frames[s - 1][c], w, h, dw, dh)
                assert np.array_equal(got, want), (s, c, dw, dh)
    batch.close()
    for c in chans:
        c.close()
