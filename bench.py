#!/usr/bin/env python
"""bench.py -- ladder frames/s of the transvideo preprocessing hot path on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--channels C] [--workload cfg2|cfg1|cfg4]

A *step* = one pass of the hot path over one batch: ONE frame of each of C concurrent channels per GPU (weak scaling:
C channels per GPU, channel-sharded, no collective on the data path -- SURVEY.md 8(e)).
  value      : source frames/s fully processed to all rungs, inputs and outputs resident in HBM, CUDA-event timed.
  e2e        : the same through the C ABI with HOST (pinned) source frames: H2D of every frame inside the timed region,
               rungs written to device surfaces (the workload's NVENC zero-copy hand-off), thumbnails read back to host.
  e2e_host_out: additionally every rung copied back to pinned host memory (the x264/x265 consumer layout).
  roofline   : ladder_kernel algorithmic bytes / its CUDA-event launch time vs MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference: the reference's CPU path (real libswscale 8.0.1 via oracle/swscale_so.py when the
               bundled .so is present, else the C oracle port) on the box's host cores.
Nothing here reads /root/reference.  oracle/ is executed only by the cpu_baseline leg (which also checks the last step's GPU
output against the CPU result) and by --impl reference.
"""
import argparse
import gc
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
                 deint=1, interlaced=0, pitch_align=256),
    # BASELINE.json configs[0]: 1080i -> yadif + 3 rungs I420
    "cfg1": dict(desc="cfg1: 1080i29.97 TFF I420 8-bit -> yadif(all) + 3-rung I420 ladder 1080p/720p/360p, bicubic "
                      "target A, 176x144 thumbnail every 150 frames",
                 w=1920, h=1080, rungs=[(1920, 1080), (1280, 720), (640, 360)], out_fmt=I420, deint=0, interlaced=1,
                 pitch_align=0),
    # BASELINE.json configs[3] per-GPU slice: 1080i -> yadif + 4 rungs
    "cfg4": dict(desc="cfg4: 1080i channels -> yadif(all) + 4-rung I420 ladder 1080p/720p/540p/360p, bicubic target A",
                 w=1920, h=1080, rungs=[(1920, 1080), (1280, 720), (960, 540), (640, 360)], out_fmt=I420, deint=0,
                 interlaced=1, pitch_align=0),
    # BASELINE.json configs[2]: 2160p60 10-bit P010 -> 6-rung lanczos ladder (P010 rungs)
    "cfg3": dict(desc="cfg3: 2160p60 P010 10-bit progressive -> 6-rung P010 lanczos ladder 2160p/1440p/1080p/720p/540p/360p, "
                      "target A tables, yadif=0:-1:1 pass-through, 176x144 thumbnail every 150 frames",
                 w=3840, h=2160, rungs=[(3840, 2160), (2560, 1440), (1920, 1080), (1280, 720), (960, 540), (640, 360)],
                 out_fmt=P010, src_fmt=P010, bps=2, lanczos=1, deint=1, interlaced=0, pitch_align=0),
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


def run_cpu_rounds(wl, frames, threads, rounds=None, seconds=None, copies=False):
    """Each round = every worker thread processes ONE source frame (ctypes releases the GIL inside libswscale / the C
    oracle).  Returns (frames_done, elapsed_s, kind, per-round seconds)."""
    workers = [CpuReference(wl, copies) for _ in range(threads)]
    kind = workers[0].kind
    barrier = threading.Barrier(threads + 1)
    state = {"go": True}

    def loop(k):
        i = k
        while True:
            barrier.wait()
            if not state["go"]:
                return
            workers[k].frame(frames[i % len(frames)], frames[(i + 1) % len(frames)])
            i += 1
            barrier.wait()

    ts = [threading.Thread(target=loop, args=(k,), daemon=True) for k in range(threads)]
    for t in ts:
        t.start()
    per_round, done, t_start = [], 0, time.perf_counter()
    while True:
        t0 = time.perf_counter()
        barrier.wait()
        barrier.wait()
        per_round.append(time.perf_counter() - t0)
        done += threads
        if rounds is not None and len(per_round) >= rounds:
            break
        if seconds is not None and time.perf_counter() - t_start >= seconds:
            break
    state["go"] = False
    barrier.wait()
    return done, sum(per_round), kind, per_round


def reference_main(args, wl, rank, world):
    if rank != 0:
        return
    threads = host_cores()
    frames = synth_frames(wl["w"], wl["h"], 4, wl["interlaced"], wl.get("src_fmt") == P010)
    total_rounds = args.warmup + args.steps
    _, _, kind, per_round = run_cpu_rounds(wl, frames, threads, rounds=total_rounds)
    timed = per_round[args.warmup:]
    elapsed = sum(timed)
    fps = threads * len(timed) / elapsed
    sample = "%d rounds x %d threads x 1 frame (one frame per host thread per step)" % (len(timed), threads)
    line = {"impl": "reference", "metric": "ladder_frames_per_sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * elapsed / len(timed),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": wl["desc"], "frames_per_step": threads},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": threads, "kind": kind, "sample": sample, "note": CPU_KIND_NOTE},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------ our arm
def ours_main(args, wl, rank, world, local_rank):
    import torch
    og = importlib.import_module("ott-packager_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; libottgpu has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        # keep stdout to the single JSON line: NCCL prints its version banner there at VERSION/INFO level
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL logs (version banner included) off stdout
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = og.Library()
    stream = torch.cuda.Stream()
    w, h, Cn = wl["w"], wl["h"], args.channels
    fmt = wl["out_fmt"]
    rungs = [(dw, dh, fmt, wl["pitch_align"]) for dw, dh in wl["rungs"]]
    bps, src_fmt = wl.get("bps", 1), wl.get("src_fmt", I420)
    src_bytes = frame_bytes(w, h) * bps

    def make_channels():
        return [og.Channel(lib, og.make_cfg(w, h, rungs, src_fmt=src_fmt, filt=wl.get("lanczos", 0), deint=wl["deint"], gpu=local_rank, thumb=THUMB,
                                            stream=stream.cuda_stream, thumb_phase=(i * THUMB[2]) // Cn))
                for i in range(Cn)]

    sampler = ClockSampler(local_rank)
    sampler.start()
    base_frames = synth_frames(w, h, 4, wl["interlaced"], src_fmt == P010)
    RING = 4
    # ---------------- device-resident arm
    chans = make_channels()
    batch = og.Batch(lib, chans)
    rung_bytes = chans[0].rung_bytes
    src_pool = og.Pool(lib, local_rank, og.MEM_DEVICE, Cn * RING, src_bytes)
    dst_pools = [og.Pool(lib, local_rank, og.MEM_DEVICE, Cn * 2, rb) for rb in rung_bytes]
    src_dev = [[src_pool.take() for _ in range(RING)] for _ in range(Cn)]
    dst_dev = [[[p.take() for p in dst_pools] for _ in range(2)] for _ in range(Cn)]
    for c in range(Cn):
        for r in range(RING):
            f = np.roll(base_frames[(c + r) % 4], 64 * c)
            lib.to_device(src_dev[c][r], f, gpu=local_rank)

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

    libc = C.CDLL(None)
    submit = lib.dll.ottgpu_batch_submit

    step_events = [] if os.environ.get("OTT_STEP_EVENTS") else None       # debugging aid: one CUDA event per step

    def run_steps(b, ios, n, launches):
        thumbs = 0
        for k in range(n):
            fin, fout = ios[k % len(ios)]
            thumbs += free_thumbs(fout)                      # thumbnails of this set's previous use (step k-4: complete)
            lib.check(submit(b.handle, fin, fout, og.SUBMIT_ASYNC))
            launches[0] += b.last_launches
            if step_events is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record(stream)
                step_events.append(ev)
        return thumbs

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    yadif_ms = [0.0, 0]
    host_ms = [0.0]

    def timed_region(b, ios, steps, warmup):
        launches = [0]
        with torch.cuda.stream(stream):
            run_steps(b, ios, warmup, [0])
            b.wait()
            barrier()
            b.timing(True)
            sampler.busy = True
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            gc.collect()
            gc.disable()                                     # a collector pause inside the loop would show up as GPU idle time
            e0.record(stream)
            h0 = time.perf_counter()
            thumbs = run_steps(b, ios, steps, launches)
            host_ms[0] = (time.perf_counter() - h0) * 1000.0 / max(steps, 1)      # wall time of the submission loop per step (includes waiting for a free descriptor set)
            e1.record(stream)
            gc.enable()
            b.wait()
            torch.cuda.synchronize()
            sampler.busy = False
            ms = e0.elapsed_time(e1)
            if step_events:
                tail = step_events[-steps:]
                d = [tail[i].elapsed_time(tail[i + 1]) for i in range(len(tail) - 1)]
                worst = sorted(range(len(d)), key=lambda i: -d[i])[:6]
                sys.stderr.write("step deltas ms: median %.4f, worst %s\n" % (sorted(d)[len(d) // 2], [(i, round(d[i], 3)) for i in worst]))
                del step_events[:]
            for fin, fout in ios:
                thumbs += free_thumbs(fout)
        if dist is not None:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        kms, kn = b.timing_read()
        yadif_ms[0], yadif_ms[1] = b.timing_read_yadif()
        b.timing(False)
        return ms, launches[0], kms, kn, thumbs

    ios = [build_io(p) for p in range(RING)]
    ms, launches, kms, kn, _ = timed_region(batch, ios, args.steps, max(args.warmup, 3))
    yadif_stat = list(yadif_ms)
    host_submit_ms = host_ms[0]
    frames_per_step = Cn
    value = world * frames_per_step * args.steps / (ms / 1000.0)

    # First half of the cpu_baseline leg (rank 0, N=1, not with --no-cpu): the frames the last timed step produced are
    # compared with what the CPU path computes for the same inputs.  oracle/ is only ever executed by this leg and by
    # --impl reference, as the checker / the baseline - never on the measured path.
    parity = None
    cpu_leg = rank == 0 and world == 1 and not args.no_cpu
    if cpu_leg and args.steps >= 3 and src_fmt == P010:
        from oracle import oracle as O
        last_phase = (args.steps - 1) % RING
        exp_src = base_frames[((args.steps - 2) % RING) % 4].view(np.uint16)
        ok = True
        for r, (dw, dh) in enumerate(wl["rungs"]):
            if (dw, dh) not in ((3840, 2160), (640, 360)):
                continue                                     # the 4K lanczos oracle is slow: check the first and last rung
            got = lib.from_device(dst_dev[0][last_phase % 2][r], rung_bytes[r], gpu=local_rank).view(np.uint16)
            ok = ok and np.array_equal(got, O.scale_p010(exp_src, w, h, dw, dh, O.LANCZOS))
        parity = "bit-exact vs oracle (channel 0, rungs 2160p + 360p)" if ok else "MISMATCH"
    elif cpu_leg and args.steps >= 3:
        from oracle import oracle as O
        # timed step k used ios[k % RING]; the last step's output is the frame submitted one step earlier (yadif latency).
        # First, middle and last channel of the launch (their tiles are spread over different CTAs and waves).
        last_phase = (args.steps - 1) % RING
        checked = sorted({0, Cn // 2, Cn - 1})
        ok = True
        for c in checked:
            src_of = lambda k: np.roll(base_frames[(c + ((args.steps - 1 + k) % RING)) % 4], 64 * c)
            exp_src = src_of(0) if wl["deint"] == 2 else src_of(-1)
            if wl["deint"] == 0:
                exp_src = O.yadif_frame(src_of(-2), src_of(-1), src_of(0), w, h, 1, 1)
            for r, (dw, dh) in enumerate(wl["rungs"]):
                got = lib.from_device(dst_dev[c][last_phase % 2][r], rung_bytes[r], gpu=local_rank)
                ref = O.scale_i420(exp_src, w, h, dw, dh)
                if fmt == NV12:
                    ref = O.i420_to_nv12(ref, dw, dh)
                lay = lib.rung_layout(og.RungCfg(dw, dh, fmt, wl["pitch_align"]))
                ch_h = (dh + 1) // 2
                rows = [dh, ch_h] if fmt == NV12 else [dh, ch_h, ch_h]
                widths = [dw, 2 * ((dw + 1) // 2)] if fmt == NV12 else [dw, (dw + 1) // 2, (dw + 1) // 2]
                flat = np.concatenate([got[o: o + pch * n].reshape(n, pch)[:, :wd].ravel() for (o, pch), n, wd in zip(lay, rows, widths)])
                ok = ok and np.array_equal(flat, ref)
        parity = "bit-exact vs oracle (channels %s, %d rungs each)" % (",".join(str(c) for c in checked), len(rungs)) if ok else "MISMATCH"

    # ---------------- end-to-end arms (host pinned source frames through the C ABI)
    batch.close()
    for c in chans:
        c.close()
    host_pool = og.Pool(lib, local_rank, og.MEM_HOST, Cn * 2, src_bytes)
    host_src = [[host_pool.take() for _ in range(2)] for _ in range(Cn)]
    for c in range(Cn):
        for r in range(2):
            f = np.roll(base_frames[(c + r) % 4], 64 * c)
            C.memmove(host_src[c][r], np.ascontiguousarray(f).ctypes.data, src_bytes)
    e2e_steps = max(4, min(args.steps, 400))          # same K as the device-resident region (a short region is hostage to one hiccup)
    chans = make_channels()
    batch = og.Batch(lib, chans)
    ios = [build_io(p, host_src=host_src) for p in range(RING)]
    e2e_warm = max(args.warmup, 8)        # the channels allocate their upload slots / tensor maps lazily: settle first
    ems, elaunch, _, _, thumbs = timed_region(batch, ios, e2e_steps, e2e_warm)
    e2e_value = world * Cn * e2e_steps / (ems / 1000.0)
    thumb_bytes = frame_bytes(THUMB[0], THUMB[1]) * bps
    e2e = {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": Cn * src_bytes + Cn * 424,
           "d2h_bytes_per_step": thumbs * thumb_bytes / float(e2e_steps), "ms_per_step": ems / e2e_steps,
           "outputs": "device surfaces (NVENC-style zero-copy hand-off); thumbnails to host"}
    batch.close()
    for c in chans:
        c.close()
    # every rung back to pinned host memory (CPU-encoder consumers)
    out_pools = [og.Pool(lib, local_rank, og.MEM_HOST, Cn, rb) for rb in rung_bytes]
    host_dst = [[p.take() for p in out_pools] for _ in range(Cn)]
    chans = make_channels()
    batch = og.Batch(lib, chans)
    ios = [build_io(p, host_src=host_src, host_dst=host_dst) for p in range(RING)]
    hms, _, _, _, thumbs2 = timed_region(batch, ios, e2e_steps, e2e_warm)
    e2e_host = {"value": world * Cn * e2e_steps / (hms / 1000.0), "unit": "frames/s",
                "h2d_bytes_per_step": Cn * src_bytes + Cn * 424,
                "d2h_bytes_per_step": Cn * sum(rung_bytes) + thumbs2 * thumb_bytes / float(e2e_steps),
                "ms_per_step": hms / e2e_steps, "outputs": "all rungs copied to pinned host memory"}
    batch.close()
    for c in chans:
        c.close()
    sampler.stop_flag = True

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---------------- roofline of the dominant kernel
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg, s_in, s_out = algorithmic_bytes(wl)
    k_bytes = ladder_kernel_bytes(wl) * frames_per_step
    k_ms = kms / max(kn, 1)
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    # DRAM bytes of the same kernel from the committed `ncu --set full` capture (profiles/), scaled from the capture's
    # frames per launch to this run's; null for workloads that were not captured
    traffic, traffic_note = None, None
    try:
        import glob
        caps = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_ladder_traffic.json")))
        if caps and args.workload == "cfg2":
            cap = json.load(open(caps[-1]))
            traffic = cap["traffic_bytes_per_frame"] * frames_per_step
            traffic_note = "%s: %d-frame capture scaled per frame; rung writes still resident in L2 at kernel end are not counted by dram__bytes_write" % (
                os.path.basename(caps[-1]), cap["channels"])
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "ladder_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
                "peak_source": "measured" if peaks else "fallback",
                "bytes_per_launch": k_bytes, "ms_per_launch": k_ms, "launches_timed": kn}
    if yadif_stat[1]:
        y_ms = yadif_stat[0] / yadif_stat[1]
        y_bytes = 3.5 * s_in * frames_per_step              # yadif reads 2.5*S and writes S per frame
        roofline["yadif_kernel"] = {"ms_per_launch": y_ms, "bytes_per_launch": y_bytes,
                                    "achieved": y_bytes / (y_ms * 1e-3) / 1e9, "frac": y_bytes / (y_ms * 1e-3) / 1e9 / peak}
    # ---------------- CPU baseline (rank 0, N=1 only), bounded sample
    cpu = None
    if cpu_leg:
        threads = host_cores()
        done, el, kind, _ = run_cpu_rounds(wl, base_frames, threads, seconds=args.cpu_seconds)
        one_done, one_el, _, _ = run_cpu_rounds(wl, base_frames, 1, seconds=min(3.0, args.cpu_seconds))
        cp_done, cp_el, _, _ = run_cpu_rounds(wl, base_frames, threads, seconds=min(4.0, args.cpu_seconds), copies=True)
        cpu = {"value": done / el, "unit": "frames/s", "cores": threads, "kind": kind, "note": CPU_KIND_NOTE,
               "sample": "%d frames of the same workload (%d rounds x %d host threads, one frame per thread per round, ~%.0f s)"
                         % (done, done // threads, threads, el),
               "with_reference_copies": {"value": cp_done / cp_el, "unit": "frames/s",
                                         "note": "same, plus the reference's full-frame memcpys #3, #4 and #5 per rendition "
                                                 "(transvideo.c:1476-1502, 2217-2231, 1610-1624); `value` leaves them out"},
               "single_thread": {"value": one_done / one_el, "unit": "frames/s",
                                 "note": "one thread = how the reference runs one channel's scaler (transvideo.c:1504)"}}
    line = {"metric": "ladder_frames_per_sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8" if bps == 1 else "u16 (10 significant bits)", "data": "synthetic",
            "config": {"workload": wl["desc"], "channels_per_gpu": Cn, "frames_per_step": world * Cn,
                       "algorithmic_bytes_per_frame": alg, "l2": "inputs+outputs per step exceed L2 (%.0f MB in, %.0f MB out per GPU per step)"
                       % (Cn * s_in / 1e6, Cn * s_out / 1e6), "parity": parity,
                       "host_loop_ms_per_step": host_submit_ms,
                       "realtime_channels_equiv": value / (60000.0 / 1001.0 if wl["interlaced"] == 0 else 30000.0 / 1001.0)},
            "e2e": e2e, "e2e_host_out": e2e_host, "gpu_launches": launches, "roofline": roofline,
            "clocks": sampler.result()}
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
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--channels", type=int, default=64, help="concurrent channels per GPU (one frame each per step)")
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true")
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
