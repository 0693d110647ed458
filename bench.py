#!/usr/bin/env python
"""bench.py -- ladder frames/s of the transvideo preprocessing hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--channels C] [--workload cfg4|cfg2|cfg1|cfg3]
                  [--repeats R] [--no-cpu] [--no-extra]

A *step* = one pass of the hot path over one batch: ONE frame of each of C concurrent channels per GPU (weak scaling:
C channels per GPU, channel-sharded, no collective on the data path -- SURVEY.md 8(e)).
The headline workload is cfg4 (BASELINE.json configs[3], the "1080i channels/box" half of the metric: yadif + 4 rungs,
64 channels per GPU); the other configs run in the same process with a shorter budget and are reported under `workloads`.
  value      : source frames/s fully processed to all rungs, inputs and outputs resident in HBM, CUDA-event timed; the
               timed region of exactly K steps is repeated R times (default 5), `value` is the median region.
  e2e        : the same through the C ABI with HOST (pinned) source frames: H2D of every frame inside the timed region.
               Where the rungs land follows the workload's consumer: x264/x265 (cfg1, cfg4, cfg3) = every rung copied
               back to pinned host memory in the contiguous layout transvideo.c:1330-1338 / 960-967 read; NVENC (cfg2) =
               pitch-aligned device surfaces, thumbnails to host.  Both variants are always reported (e2e_device_out,
               e2e_host_out).
  roofline   : the dominant kernel's algorithmic bytes / its CUDA-event launch time vs MEASURED_PEAKS.json hbm_gbs;
               roofline.kernels has ladder_kernel and yadif_kernel side by side.
  cpu_baseline / --impl reference: the reference's CPU path (real libswscale 8.0.1 via oracle/swscale_so.py when the
               bundled .so is present, else the C oracle port) on the box's host cores, free-running workers.
Nothing here reads /root/reference.  oracle/ is executed only by the cpu_baseline leg (which also checks the last step's GPU
output against the CPU result) and by --impl reference.
"""
import argparse
import gc
import hashlib
import ctypes as C
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

I420, NV12, P010 = 0, 1, 2
WORKLOADS = {
    # BASELINE.json configs[1]: 1080p60 progressive 8-bit -> 5-rung NV12 ladder, zero-copy surfaces for NVENC
    "cfg2": dict(desc="cfg2: 1080p60 progressive I420 8-bit -> 5-rung NV12 ladder 1080p/720p/540p/360p/416x234, "
                      "bicubic target A, yadif=0:-1:1 pass-through, 176x144 thumbnail every 150 frames",
                 w=1920, h=1080, rungs=[(1920, 1080), (1280, 720), (960, 540), (640, 360), (416, 234)], out_fmt=NV12,
                 deint=1, interlaced=0, pitch_align=256, consumer="nvenc (device surfaces)"),
    # BASELINE.json configs[0]: 1080i -> yadif + 3 rungs I420
    "cfg1": dict(desc="cfg1: 1080i29.97 TFF I420 8-bit -> yadif(all) + 3-rung I420 ladder 1080p/720p/360p, bicubic "
                      "target A, 176x144 thumbnail every 150 frames",
                 w=1920, h=1080, rungs=[(1920, 1080), (1280, 720), (640, 360)], out_fmt=I420, deint=0, interlaced=1,
                 pitch_align=0, consumer="x264 (packed host buffers)"),
    # BASELINE.json configs[3] per-GPU slice: 1080i -> yadif + 4 rungs
    "cfg4": dict(desc="cfg4: 1080i channels -> yadif(all) + 4-rung I420 ladder 1080p/720p/540p/360p, bicubic target A",
                 w=1920, h=1080, rungs=[(1920, 1080), (1280, 720), (960, 540), (640, 360)], out_fmt=I420, deint=0,
                 interlaced=1, pitch_align=0, consumer="x264 (packed host buffers)"),
    # BASELINE.json configs[2]: 2160p60 10-bit P010 -> 6-rung lanczos ladder (P010 rungs)
    "cfg3": dict(desc="cfg3: 2160p60 P010 10-bit progressive -> 6-rung P010 lanczos ladder 2160p/1440p/1080p/720p/540p/360p, "
                      "target A tables, yadif=0:-1:1 pass-through, 176x144 thumbnail every 150 frames",
                 w=3840, h=2160, rungs=[(3840, 2160), (2560, 1440), (1920, 1080), (1280, 720), (960, 540), (640, 360)],
                 out_fmt=P010, src_fmt=P010, bps=2, lanczos=1, deint=1, interlaced=0, pitch_align=0, consumer="x265 (packed host buffers)"),
}
THUMB = (176, 144, 150)


def frame_bytes(w, h):
    return w * h + 2 * ((w + 1) // 2) * ((h + 1) // 2)


def algorithmic_bytes(wl):
    """fused read (2.5*S with yadif, S without) + sum of rung bytes, per source frame."""
    bps = wl.get("bps", 1)
    s = frame_bytes(wl["w"], wl["h"]) * bps
    out = sum(frame_bytes(w, h) for w, h in wl["rungs"]) * bps
    return (2.5 * s if wl["deint"] == 0 else s) + out, s, out


def ladder_kernel_bytes(wl):
    """bytes the ladder kernel itself moves per frame: read S, write the rungs it produces (with yadif the full-size rung
    is written by the yadif kernel instead)."""
    bps = wl.get("bps", 1)
    s = frame_bytes(wl["w"], wl["h"])
    out = sum(frame_bytes(w, h) for w, h in wl["rungs"] if not (wl["deint"] == 0 and (w, h) == (wl["w"], wl["h"])))
    return (s + out) * bps


def synth_frames(w, h, n, interlaced, p010=False):
    """n distinct natural-like frames (sin/cos + checker + noise); woven fields when interlaced.
    p010: 10-bit noise MSB-aligned in 16 bits, Y plane then interleaved UV (seed 99)."""
    if p010:
        rng = np.random.default_rng(99)
        n_s = frame_bytes(w, h)
        return [((rng.integers(0, 1024, n_s, dtype=np.uint16)) << 6).astype(np.uint16).view(np.uint8) for _ in range(n)]
    rng = np.random.default_rng(7)
    cw, ch = (w + 1) // 2, (h + 1) // 2
    out = []
    if os.environ.get("OTT_BENCH_CONTENT") == "noise":       # experiments only (profiles/README.md, content dependence of the
        n_s = frame_bytes(w, h)                              # ladder's vertical pass): full-range uniform noise
        return [rng.integers(0, 256, n_s, dtype=np.uint8) for _ in range(n)]
    for i in range(n):
        def plane(pw, ph, k, shift):
            x = np.arange(pw, dtype=np.float32)[None, :] * k + shift
            y = np.arange(ph, dtype=np.float32)[:, None] * k
            v = 0.5 + 0.25 * np.sin(x / 37.0) + 0.2 * np.cos(y / 23.0)
            v = v + 0.05 * ((((x // 64).astype(np.int32) + (y // 64).astype(np.int32)) & 1) * 2 - 1)
            v = v + rng.normal(0, 0.01, (ph, pw)).astype(np.float32)
            return (np.clip(v, 0.0625, 0.92) * 255).astype(np.uint8)
        yp = plane(w, h, 1, 8.0 * i)
        if interlaced:
            yp2 = plane(w, h, 1, 8.0 * i + 4.0)
            yp[1::2] = yp2[1::2]
        out.append(np.concatenate([yp.ravel(), plane(cw, ch, 2, 8.0 * i).ravel(), plane(cw, ch, 2, 8.0 * i + 100).ravel()]))
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.ok = index, [], set(), False, None, False
        self.busy = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        while self.ok and not self.stop_flag:
            try:
                if self.busy:
                    self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                    try:
                        r = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                    except Exception:
                        r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
class CpuReference:
    """The reference's CPU path for one source frame: N x sws_scale (single-threaded contexts, flags=SWS_BICUBIC, exactly
    transvideo.c:1563-1581), NV12 repack when the workload asks for NV12 (transvideo.c:543-557), yadif via the C port
    (no libavfilter in the image), thumbnail every 150th frame (transvideo.c:2159-2174).  One instance per thread."""

    def __init__(self, wl, copies=False):
        from oracle import oracle as O, swscale_so as S
        self.O, self.wl = O, wl
        # copies=True adds the full-frame memcpys the reference makes around the arithmetic: #3 into the filter
        # frame (transvideo.c:1476-1502, 2130), #4 out of it (2217-2231), #5 per rendition into the pool buffer (1610-1624)
        self.copies = copies
        self.stage_in = np.empty(frame_bytes(wl["w"], wl["h"]) * (2 if wl.get("src_fmt") == P010 else 1), np.uint8)
        self.stage_out = np.empty_like(self.stage_in)
        self.kind = "reference" if S.available() else "port"
        self.w, self.h = wl["w"], wl["h"]
        self.count = 0
        self.prev = None
        self.p010 = wl.get("src_fmt") == P010
        if self.kind == "reference" and self.p010:
            S.load()
            self.S = S
            self.ctx = [S.Scaler(self.w, self.h, "p010le", dw, dh, "p010le", S.SWS_LANCZOS) for dw, dh in wl["rungs"]]
            self.thumb = S.Scaler(self.w, self.h, "p010le", THUMB[0], THUMB[1], "yuv420p10le", S.SWS_BICUBIC)
            self.out = [S._alloc_planes("p010le", dw, dh) for dw, dh in wl["rungs"]]
            self.tout = S._alloc_planes("yuv420p10le", THUMB[0], THUMB[1])
        elif self.kind == "reference":
            S.load()
            self.S = S
            self.ctx = [S.Scaler(self.w, self.h, "yuv420p", dw, dh, "yuv420p", S.SWS_BICUBIC) for dw, dh in wl["rungs"]]
            self.thumb = S.Scaler(self.w, self.h, "yuv420p", THUMB[0], THUMB[1], "yuv420p", S.SWS_BICUBIC)
            self.out = [S._alloc_planes("yuv420p", dw, dh) for dw, dh in wl["rungs"]]
            self.tout = S._alloc_planes("yuv420p", THUMB[0], THUMB[1])
        self.nv12 = [np.empty(frame_bytes(dw, dh), np.uint8) for dw, dh in wl["rungs"]]

    def planes(self, f):
        w, h = self.w, self.h
        cw, ch = (w + 1) // 2, (h + 1) // 2
        if self.p010:
            f16 = f.view(np.uint16)
            return [f16[: w * h].reshape(h, w), f16[w * h:].reshape(ch, 2 * cw)]
        return [f[: w * h].reshape(h, w), f[w * h: w * h + cw * ch].reshape(ch, cw), f[w * h + cw * ch:].reshape(ch, cw)]

    def frame(self, f, nxt):
        O, wl = self.O, self.wl
        if self.copies:
            np.copyto(self.stage_in, f.view(np.uint8))            # copy #3 (the filter graph's input frame)
        if wl["deint"] == 0:                                  # yadif on every frame
            f = O.yadif_frame(self.prev if self.prev is not None else f, f, nxt, self.w, self.h, 1, 1)
        self.prev = f
        self.count += 1
        if self.copies:
            np.copyto(self.stage_out, f.view(np.uint8))           # copy #4 (deinterlaced frame -> pool buffer)
        if self.kind == "reference":
            src = self.planes(f)
            for sc, out, (dw, dh), nv in zip(self.ctx, self.out, wl["rungs"], self.nv12):
                sc.scale_planes(src, out)
                if wl["out_fmt"] == NV12 or self.copies:
                    packed = np.concatenate([p.ravel() for p in out])     # copy #5: planes -> contiguous pool buffer
                if wl["out_fmt"] == NV12:
                    O.lib().ora_i420_to_nv12(packed.ctypes.data, dw, dh, nv.ctypes.data)
            if self.count % THUMB[2] == 0:
                self.thumb.scale_planes(src, self.tout)
        elif self.p010:
            for (dw, dh) in wl["rungs"]:
                O.scale_p010(f.view(np.uint16), self.w, self.h, dw, dh, O.LANCZOS)
        else:
            for (dw, dh), nv in zip(wl["rungs"], self.nv12):
                o = O.scale_i420(f, self.w, self.h, dw, dh)
                if wl["out_fmt"] == NV12:
                    O.lib().ora_i420_to_nv12(o.ctypes.data, dw, dh, nv.ctypes.data)
            if self.count % THUMB[2] == 0:
                O.scale_i420(f, self.w, self.h, THUMB[0], THUMB[1])


CPU_KIND_NOTE = ("kind=reference: the scaler is the real libswscale 8.0.1 in this image (the library the reference calls through "
                 "sws_getContext/sws_scale, x86 SIMD enabled), one single-threaded context per rendition per worker exactly as "
                 "transvideo.c:1563-1581; yadif and the NV12 repack run through the C port in oracle/ (no libavfilter in the image). "
                 "kind=port: everything through oracle/ (bundled libswscale missing)")



def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class CpuPool:
    """Free-running CPU workers: every thread owns its scaler contexts (like one fillet process per channel) and pulls
    the next frame index from a shared counter until the region's frame budget is used up - no per-round barrier, so a
    slow frame on one thread never idles the others (ctypes releases the GIL inside libswscale / the C oracle)."""

    def __init__(self, wl, frames, threads, copies=False):
        self.workers = [CpuReference(wl, copies) for _ in range(threads)]
        self.kind = self.workers[0].kind
        self.frames, self.threads = frames, threads

    def run(self, n_frames=None, seconds=None):
        """Process n_frames (or as many as fit in `seconds`) -> (frames_done, elapsed_s)."""
        lock = threading.Lock()
        state = {"next": 0, "done": 0}
        t_end = None if seconds is None else time.perf_counter() + seconds

        def loop(k):
            w, fr = self.workers[k], self.frames
            while True:
                with lock:
                    i = state["next"]
                    if (n_frames is not None and i >= n_frames) or (t_end is not None and time.perf_counter() >= t_end):
                        return
                    state["next"] = i + 1
                w.frame(fr[i % len(fr)], fr[(i + 1) % len(fr)])
                with lock:
                    state["done"] += 1

        ts = [threading.Thread(target=loop, args=(k,), daemon=True) for k in range(self.threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return state["done"], time.perf_counter() - t0


def static_config(wl, channels, world):
    """The part of `config` both arms print verbatim (the driver compares it)."""
    alg, s_in, s_out = algorithmic_bytes(wl)
    return {"workload": wl["desc"], "channels_per_gpu": channels, "frames_per_step": world * channels,
            "algorithmic_bytes_per_frame": alg,
            "l2": "inputs+outputs per step exceed L2 (%.0f MB in, %.0f MB out per GPU per step)" % (channels * s_in / 1e6, channels * s_out / 1e6),
            "consumer": wl["consumer"]}


def reference_main(args, wl, rank, world):
    if rank != 0:
        return
    threads = host_cores()
    frames = synth_frames(wl["w"], wl["h"], 4, wl["interlaced"], wl.get("src_fmt") == P010)
    per_step = world * args.channels                         # the same step as our arm: one frame of every channel
    pool = CpuPool(wl, frames, threads)
    if args.warmup > 0:
        pool.run(n_frames=args.warmup * per_step)
    done, elapsed = pool.run(n_frames=args.steps * per_step)
    fps = done / elapsed
    sample = "%d steps x %d frames (one frame of each of the %d channels per step), %d free-running host threads, %.1f s" % (
        args.steps, per_step, per_step, threads, elapsed)
    line = {"impl": "reference", "metric": "ladder_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * elapsed / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8" if wl.get("bps", 1) == 1 else "u16 (10 significant bits)",
            "data": "synthetic", "config": static_config(wl, args.channels, world),
            "run": {"host_threads": threads, "yadif_parity": "unpinned (C port of vf_yadif; no libavfilter in the image)" if wl["deint"] == 0 else "n/a"},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": pool.kind, "sample": sample, "note": CPU_KIND_NOTE},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm
class Ctx:
    pass


def median(v):
    v = sorted(v)
    return v[len(v) // 2] if len(v) % 2 else 0.5 * (v[len(v) // 2 - 1] + v[len(v) // 2])


def run_workload(cx, name, Cn, steps, warmup, repeats, do_e2e, do_parity):
    """One workload on this rank's GPU: device-resident region (value, kernel times), optional in-run parity check against
    the CPU path, optional end-to-end regions.  Every rank of a multi-GPU run executes the same sequence."""
    og, lib, torch, dist, stream, local_rank, world = cx.og, cx.lib, cx.torch, cx.dist, cx.stream, cx.local_rank, cx.world
    wl = WORKLOADS[name]
    w, h = wl["w"], wl["h"]
    fmt = wl["out_fmt"]
    rungs = [(dw, dh, fmt, wl["pitch_align"]) for dw, dh in wl["rungs"]]
    bps, src_fmt = wl.get("bps", 1), wl.get("src_fmt", I420)
    src_bytes = frame_bytes(w, h) * bps
    libc = C.CDLL(None)
    submit = lib.dll.ottgpu_batch_submit

    def make_channels():
        return [og.Channel(lib, og.make_cfg(w, h, rungs, src_fmt=src_fmt, filt=wl.get("lanczos", 0), deint=wl["deint"], gpu=local_rank, thumb=THUMB,
                                            stream=stream.cuda_stream, thumb_phase=(i * THUMB[2]) // Cn))
                for i in range(Cn)]

    base_frames = synth_frames(w, h, 4, wl["interlaced"], src_fmt == P010)
    RING = 4
    chans = make_channels()
    batch = og.Batch(lib, chans)
    rung_bytes = chans[0].rung_bytes
    src_pool = og.Pool(lib, local_rank, og.MEM_DEVICE, Cn * RING, src_bytes)
    dst_pools = [og.Pool(lib, local_rank, og.MEM_DEVICE, Cn * 2, rb) for rb in rung_bytes]
    src_dev = [[src_pool.take() for _ in range(RING)] for _ in range(Cn)]
    dst_dev = [[[p.take() for p in dst_pools] for _ in range(2)] for _ in range(Cn)]
    for c in range(Cn):
        for r in range(RING):
            lib.to_device(src_dev[c][r], np.roll(base_frames[(c + r) % 4], 64 * c), gpu=local_rank)

    def build_io(phase, host_src=None, host_dst=None):
        fin = (og.FrameIn * Cn)()
        fout = (og.FrameOut * Cn)()
        for c in range(Cn):
            if host_src is not None:
                fin[c].data, fin[c].mem = host_src[c][phase % 2], og.MEM_HOST
            else:
                fin[c].data, fin[c].mem = src_dev[c][phase % RING], og.MEM_DEVICE
            fin[c].interlaced, fin[c].tff = wl["interlaced"], 1
            for r in range(len(rungs)):
                if host_dst is not None:
                    fout[c].dst[r], fout[c].dst_mem[r] = host_dst[c][r], og.MEM_HOST
                else:
                    fout[c].dst[r], fout[c].dst_mem[r] = dst_dev[c][phase % 2][r], og.MEM_DEVICE
        return fin, fout

    def free_thumbs(fout):
        n = 0
        for c in range(Cn):
            if fout[c].thumbnail:
                libc.free(C.c_void_p(fout[c].thumbnail))
                fout[c].thumbnail = None
                n += 1
        return n

    def run_steps(b, ios, first, n, launches):
        thumbs = 0
        for k in range(first, first + n):
            fin, fout = ios[k % len(ios)]
            thumbs += free_thumbs(fout)                      # thumbnails of this set's previous use (step k-4: complete)
            lib.check(submit(b.handle, fin, fout, og.SUBMIT_ASYNC))
            launches[0] += b.last_launches
        return thumbs

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed_regions(b, ios, steps, warmup, repeats):
        """W warm-up steps, then `repeats` regions of exactly `steps` steps each, every region bracketed by a barrier +
        synchronize on both sides and timed with CUDA events on the batch's stream; max over ranks per region."""
        out = []
        with torch.cuda.stream(stream):
            k = 0
            run_steps(b, ios, k, warmup, [0])
            k += warmup
            b.wait()
            for _ in range(repeats):
                barrier()
                b.timing(True)
                cx.sampler.busy = True
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                gc.collect()
                gc.disable()                                 # a collector pause inside the loop would show up as GPU idle time
                launches = [0]
                e0.record(stream)
                h0 = time.perf_counter()
                thumbs = run_steps(b, ios, k, steps, launches)
                host_ms = (time.perf_counter() - h0) * 1000.0 / max(steps, 1)
                e1.record(stream)
                gc.enable()
                b.wait()
                barrier()
                cx.sampler.busy = False
                ms = e0.elapsed_time(e1)
                k += steps
                if dist is not None:
                    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                kms, kn = b.timing_read()
                yms, yn = b.timing_read_yadif()
                b.timing(False)
                out.append({"ms": ms, "launches": launches[0], "ladder_ms": kms / max(kn, 1), "ladder_n": kn,
                            "yadif_ms": yms / max(yn, 1) if yn else None, "thumbs": thumbs, "host_ms": host_ms, "last_step": k - 1})
            for fin, fout in ios:
                free_thumbs(fout)
        return out

    def pick(regs):
        """the median region (by time) and the spread"""
        order = sorted(range(len(regs)), key=lambda i: regs[i]["ms"])
        med = regs[order[len(order) // 2]]
        spread = {"n": len(regs), "ms_per_step": [r["ms"] / steps for r in regs],
                  "median": med["ms"] / steps, "min": regs[order[0]]["ms"] / steps, "max": regs[order[-1]]["ms"] / steps}
        return med, spread

    res = {"name": name, "channels": Cn, "steps": steps}
    ios = [build_io(p) for p in range(RING)]
    regs = timed_regions(batch, ios, steps, warmup, repeats)
    med, spread = pick(regs)
    res.update(ms_per_step=med["ms"] / steps, value=world * Cn * steps / (med["ms"] / 1000.0), repeats=spread,
               launches=med["launches"], host_loop_ms_per_step=med["host_ms"],
               ladder_ms=median([r["ladder_ms"] for r in regs]), ladder_launches_timed=sum(r["ladder_n"] for r in regs),
               yadif_ms=median([r["yadif_ms"] for r in regs]) if regs[0]["yadif_ms"] is not None else None)

    # in-run parity (first half of the cpu_baseline leg: rank 0, N=1, not with --no-cpu): the frames the LAST step of the
    # last region produced are compared with what the CPU path computes for the same inputs
    parity = None
    last_k = regs[-1]["last_step"]
    if do_parity and last_k >= 3 and src_fmt == P010:
        from oracle import oracle as O
        exp_src = base_frames[(((last_k - 1) % RING)) % 4].view(np.uint16)
        ok = True
        for r, (dw, dh) in enumerate(wl["rungs"]):
            if (dw, dh) not in ((3840, 2160), (640, 360)):
                continue                                     # the 4K lanczos oracle is slow: check the first and last rung
            got = lib.from_device(dst_dev[0][last_k % 2][r], rung_bytes[r], gpu=local_rank).view(np.uint16)
            ok = ok and np.array_equal(got, O.scale_p010(exp_src, w, h, dw, dh, O.LANCZOS))
        parity = "bit-exact vs oracle (channel 0, rungs 2160p + 360p)" if ok else "MISMATCH"
    elif do_parity and last_k >= 3:
        from oracle import oracle as O
        # step k used ios[k % RING]; its output is the frame submitted one step earlier (yadif latency).
        # First, middle and last channel of the launch (their tiles are spread over different CTAs and waves).
        checked = sorted({0, Cn // 2, Cn - 1})
        ok = True
        for c in checked:
            src_of = lambda d: np.roll(base_frames[(c + ((last_k + d) % RING)) % 4], 64 * c)
            exp_src = src_of(0) if wl["deint"] == 2 else src_of(-1)
            if wl["deint"] == 0:
                exp_src = O.yadif_frame(src_of(-2), src_of(-1), src_of(0), w, h, 1, 1)
            for r, (dw, dh) in enumerate(wl["rungs"]):
                got = lib.from_device(dst_dev[c][last_k % 2][r], rung_bytes[r], gpu=local_rank)
                ref = O.scale_i420(exp_src, w, h, dw, dh)
                if fmt == NV12:
                    ref = O.i420_to_nv12(ref, dw, dh)
                lay = lib.rung_layout(og.RungCfg(dw, dh, fmt, wl["pitch_align"]))
                ch_h = (dh + 1) // 2
                rows = [dh, ch_h] if fmt == NV12 else [dh, ch_h, ch_h]
                widths = [dw, 2 * ((dw + 1) // 2)] if fmt == NV12 else [dw, (dw + 1) // 2, (dw + 1) // 2]
                flat = np.concatenate([got[o: o + pch * n].reshape(n, pch)[:, :wd].ravel() for (o, pch), n, wd in zip(lay, rows, widths)])
                ok = ok and np.array_equal(flat, ref)
        parity = ("bit-exact vs CPU path (channels %s, %d rungs each%s)" % (",".join(str(c) for c in checked), len(rungs),
                  "; yadif vs the C port of vf_yadif - FFmpeg parity unpinned" if wl["deint"] == 0 else "")) if ok else "MISMATCH"
    res["parity"] = parity
    batch.close()
    for c in chans:
        c.close()

    # ---------------- end-to-end arms (host pinned source frames through the C ABI)
    if do_e2e:
        # OTT_WC=1: the upload ring in write-combined pinned memory (the CPU only writes it)
        host_pool = og.Pool(lib, local_rank, og.MEM_HOST_WC if os.environ.get("OTT_WC") == "1" else og.MEM_HOST, Cn * 2, src_bytes)
        host_src = [[None, None] for _ in range(Cn)]
        for r in range(2):                                   # one phase's frames of all channels are contiguous in pinned memory
            for c in range(Cn):
                host_src[c][r] = host_pool.take()
        for c in range(Cn):
            for r in range(2):
                f = np.roll(base_frames[(c + r) % 4], 64 * c)
                C.memmove(host_src[c][r], np.ascontiguousarray(f).ctypes.data, src_bytes)
        e2e_warm = max(warmup, 8)         # the channels allocate their upload slots / tensor maps lazily: settle first
        thumb_bytes = frame_bytes(THUMB[0], THUMB[1]) * bps
        chans = make_channels()
        batch = og.Batch(lib, chans)
        ios = [build_io(p, host_src=host_src) for p in range(RING)]
        med, spread = pick(timed_regions(batch, ios, steps, e2e_warm, max(3, repeats)))
        res["e2e_device_out"] = {"value": world * Cn * steps / (med["ms"] / 1000.0), "unit": "frames/s",
                                 "h2d_bytes_per_step": Cn * src_bytes + Cn * 424,
                                 "d2h_bytes_per_step": med["thumbs"] * thumb_bytes / float(steps), "ms_per_step": med["ms"] / steps,
                                 "repeats": spread, "outputs": "device surfaces (NVENC-style zero-copy hand-off); thumbnails to host"}
        batch.close()
        for c in chans:
            c.close()
        out_pools = [og.Pool(lib, local_rank, og.MEM_HOST, Cn, rb) for rb in rung_bytes]
        host_dst = [[p.take() for p in out_pools] for _ in range(Cn)]
        chans = make_channels()
        batch = og.Batch(lib, chans)
        ios = [build_io(p, host_src=host_src, host_dst=host_dst) for p in range(RING)]
        med, spread = pick(timed_regions(batch, ios, steps, e2e_warm, max(3, repeats)))
        res["e2e_host_out"] = {"value": world * Cn * steps / (med["ms"] / 1000.0), "unit": "frames/s",
                               "h2d_bytes_per_step": Cn * src_bytes + Cn * 424,
                               "d2h_bytes_per_step": Cn * sum(rung_bytes) + med["thumbs"] * thumb_bytes / float(steps),
                               "ms_per_step": med["ms"] / steps, "repeats": spread,
                               "outputs": "every rung copied to pinned host memory in the packed layout the x264/x265 threads read (transvideo.c:1330-1338, 960-967)"}
        batch.close()
        for c in chans:
            c.close()
        for p in out_pools:
            p.close()
        host_pool.close()
        # The box's own floor for one step's traffic: plain pinned copies of the same byte counts, every rank copying at the
        # same time (uploads alone for the device-out variant, uploads + read-backs at once for the host-out one).  Not part
        # of any timed region; it says how much of an end-to-end step is the PCIe / host-memory system of this box.
        def link_floor(h2d_bytes, d2h_bytes, reps=8):
            hin = torch.empty(int(h2d_bytes), dtype=torch.uint8).pin_memory()
            din = torch.empty(int(h2d_bytes), dtype=torch.uint8, device="cuda")
            hout = torch.empty(max(int(d2h_bytes), 1), dtype=torch.uint8).pin_memory()
            dout = torch.empty(max(int(d2h_bytes), 1), dtype=torch.uint8, device="cuda")
            s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

            def once():
                with torch.cuda.stream(s1):
                    din.copy_(hin, non_blocking=True)
                if d2h_bytes > 65536:
                    with torch.cuda.stream(s2):
                        hout.copy_(dout, non_blocking=True)
            once()
            once()
            barrier()
            best = None
            for _ in range(reps):                            # the floor is the best repetition: event-timed on the device
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                barrier()                                    # all ranks copy at the same time
                e0.record()
                s1.wait_event(e0)
                s2.wait_event(e0)
                once()
                torch.cuda.current_stream().wait_stream(s1)
                torch.cuda.current_stream().wait_stream(s2)
                e1.record()
                torch.cuda.synchronize()
                t = e0.elapsed_time(e1)
                best = t if best is None else min(best, t)
            ms = best
            if dist is not None:
                t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            del hin, din, hout, dout
            return ms
        for key in ("e2e_device_out", "e2e_host_out"):
            e = res[key]
            floor_ms = link_floor(e["h2d_bytes_per_step"], e["d2h_bytes_per_step"])
            e["link_floor_ms_per_step"] = floor_ms
            e["link_floor_share"] = floor_ms / e["ms_per_step"]
            e["link_floor_note"] = ("plain pinned cudaMemcpyAsync of the step's %d H2D%s bytes, all %d rank(s) at once: the PCIe / host-memory "
                                    "floor of THIS box at this rank count; share = floor / measured step" %
                                    (e["h2d_bytes_per_step"], (" + %d D2H" % e["d2h_bytes_per_step"]) if e["d2h_bytes_per_step"] > 65536 else "", world))
    for p in dst_pools:
        p.close()
    src_pool.close()
    res["base_frames"] = base_frames
    return res


def kernel_rooflines(wl, res, peak):
    """ladder_kernel / yadif_kernel: algorithmic bytes per launch (SURVEY 8(d) per-frame figures x frames per launch) over
    the CUDA-event launch time the library measured on its stream."""
    Cn = res["channels"]
    _, s_in, _ = algorithmic_bytes(wl)
    out = {}
    if res["ladder_ms"]:
        b = ladder_kernel_bytes(wl) * Cn
        a = b / (res["ladder_ms"] * 1e-3) / 1e9
        out["ladder_kernel"] = {"ms_per_launch": res["ladder_ms"], "bytes_per_launch": b, "achieved": a, "frac": a / peak}
    if res["yadif_ms"]:
        b = 3.5 * s_in * Cn                                  # yadif reads 2.5*S and writes S per frame
        a = b / (res["yadif_ms"] * 1e-3) / 1e9
        out["yadif_kernel"] = {"ms_per_launch": res["yadif_ms"], "bytes_per_launch": b, "achieved": a, "frac": a / peak}
    return out


def ours_main(args, wl, rank, world, local_rank):
    import torch
    og = importlib.import_module("ott-packager_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libottgpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if os.environ.get("OTT_PIN") == "1":                     # experiment: each rank on its own share of the host cores
        cores = sorted(os.sched_getaffinity(0))
        mine = cores[local_rank::max(world, 1)]
        if mine:
            os.sched_setaffinity(0, mine)
    dist = None
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at VERSION/INFO level
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL logs (version banner included) off stdout
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = og.Library()
    info = (lib.dll.ottgpu_build_info() or b"").decode()
    if "sm_100a" not in info:
        raise SystemExit("bench.py: %s is not the sm_100a product build (%s); refusing to measure it" % (lib.path, info))
    cx = Ctx()
    cx.og, cx.lib, cx.torch, cx.dist, cx.local_rank, cx.world = og, lib, torch, dist, local_rank, world
    cx.stream = torch.cuda.Stream()
    cx.sampler = ClockSampler(local_rank)
    cx.sampler.start()
    cpu_leg = rank == 0 and world == 1 and not args.no_cpu
    warm = max(args.warmup, 3)
    res = run_workload(cx, args.workload, args.channels, args.steps, warm, args.repeats, True, cpu_leg)
    extra = {}
    if not args.no_extra:
        for name in ("cfg2", "cfg1", "cfg3", "cfg4"):
            if name == args.workload:
                continue
            ch = min(args.channels, 16) if name == "cfg3" else args.channels
            r = run_workload(cx, name, ch, min(args.steps, 30), warm, 3, False, cpu_leg and name != "cfg3")
            extra[name] = r
    cx.sampler.stop_flag = True
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    kr = kernel_rooflines(wl, res, peak)
    dom = max(kr, key=lambda k: kr[k]["ms_per_launch"]) if kr else None
    # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture of this workload (profiles/), scaled
    # from the capture's frames per launch to this run's; null when no capture of that kernel/workload is committed
    traffic, traffic_note = None, None
    try:
        cap_file = os.path.join(ROOT, "profiles", "r2_%s_%s_traffic.json" % (args.workload, dom))
        if dom and os.path.exists(cap_file):
            cap = json.load(open(cap_file))
            traffic = cap["traffic_bytes_per_frame"] * res["channels"]
            traffic_note = "%s: %d-frame capture scaled per frame; writes still resident in L2 at kernel end are not counted by dram__bytes_write" % (
                os.path.basename(cap_file), cap["channels"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": kr[dom]["achieved"] if dom else None, "peak": peak, "unit": "GB/s",
                "frac": kr[dom]["frac"] if dom else None, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": "measured" if peaks else "fallback",
                "bytes_per_launch": kr[dom]["bytes_per_launch"] if dom else None, "ms_per_launch": kr[dom]["ms_per_launch"] if dom else None,
                "kernels": kr,
                "note": "per-kernel: algorithmic bytes / CUDA-event launch time on the library's stream; the step runs yadif_kernel then ladder_kernel back to back"}
    # ---------------- CPU baseline (rank 0, N=1 only), bounded sample
    cpu = None
    if cpu_leg:
        threads = host_cores()
        frames = res["base_frames"]
        pool = CpuPool(wl, frames, threads)
        pool.run(n_frames=2 * threads)
        done, el = pool.run(seconds=args.cpu_seconds)
        one = CpuPool(wl, frames, 1)
        one_done, one_el = one.run(seconds=min(3.0, args.cpu_seconds))
        cp = CpuPool(wl, frames, threads, copies=True)
        cp_done, cp_el = cp.run(seconds=min(4.0, args.cpu_seconds))
        cpu = {"value": done / el, "unit": "frames/s", "cores": threads, "kind": pool.kind, "note": CPU_KIND_NOTE,
               "sample": "%d frames of the same workload, %d free-running host threads, %.1f s" % (done, threads, el),
               "with_reference_copies": {"value": cp_done / cp_el, "unit": "frames/s",
                                         "note": "same, plus the reference's full-frame memcpys #3, #4 and #5 per rendition "
                                                 "(transvideo.c:1476-1502, 2217-2231, 1610-1624); `value` leaves them out"},
               "single_thread": {"value": one_done / one_el, "unit": "frames/s",
                                 "note": "one thread = how the reference runs one channel's scaler (transvideo.c:1504)"}}
    e2e_key = "e2e_host_out" if wl["consumer"].startswith("x26") else "e2e_device_out"
    sha = hashlib.sha256(open(lib.path, "rb").read()).hexdigest()[:16]
    line = {"metric": "ladder_frames_per_sec", "value": res["value"], "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8" if wl.get("bps", 1) == 1 else "u16 (10 significant bits)", "data": "synthetic",
            "config": static_config(wl, args.channels, world),
            "run": {"parity": res["parity"], "host_loop_ms_per_step": res["host_loop_ms_per_step"], "host_threads": host_cores(),
                    "realtime_channels_equiv": res["value"] / (60000.0 / 1001.0 if wl["interlaced"] == 0 else 30000.0 / 1001.0),
                    "library": {"path": os.path.relpath(lib.path, ROOT), "sha256_16": sha, "build": info},
                    "yadif_parity": "unpinned (GPU == C port of vf_yadif bit-exact; no libavfilter in the image to pin the port)" if wl["deint"] == 0 else "n/a"},
            "repeats": res["repeats"],
            "e2e": dict(res[e2e_key], variant=e2e_key), "e2e_device_out": res["e2e_device_out"], "e2e_host_out": res["e2e_host_out"],
            "gpu_launches": res["launches"], "roofline": roofline, "clocks": cx.sampler.result()}
    if extra:
        wls = {}
        for name, r in extra.items():
            k2 = kernel_rooflines(WORKLOADS[name], r, peak)
            wls[name] = {"workload": WORKLOADS[name]["desc"], "channels_per_gpu": r["channels"], "steps": r["steps"],
                         "value": r["value"], "unit": "frames/s", "ms_per_step": r["ms_per_step"], "repeats": r["repeats"],
                         "kernels": k2, "parity": r["parity"], "gpu_launches": r["launches"]}
        line["workloads"] = wls
    if cpu is not None:
        line["cpu_baseline"] = cpu
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


_RESULT_FD = None


def claim_stdout():
    """The driver reads ONE JSON line from stdout.  Libraries print there too (NCCL's version banner, for one), so the
    real stdout is set aside for emit() and everything else written to fd 1 from here on goes to stderr."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    data = (json.dumps(line) + "\n").encode()
    os.write(_RESULT_FD if _RESULT_FD is not None else 1, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=64, help="concurrent channels per GPU (one frame each per step)")
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    ap.add_argument("--repeats", type=int, default=5, help="timed regions of exactly --steps steps; value = the median region")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the `workloads` block (the other BASELINE configs)")
    args = ap.parse_args()
    claim_stdout()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_main(args, wl, rank, world)
    else:
        ours_main(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
