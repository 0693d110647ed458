#!/usr/bin/env python
"""Decode-side converter on one B200: device-resident pictures -> packed I420 on the device, timed
with CUDA events on the converter's stream; plus the end-to-end figure with pageable/pinned host pictures in and the
packed frame left on the device (what the prepare stage consumes).  Prints one JSON line per pixel format.

  python scripts/bench_convert.py [--size 1920x1080] [--steps 200]"""
import argparse, importlib, json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", default="1920x1080")
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--batch", type=int, default=32, help="pictures per launch in the batched figure")
    ap.add_argument("--pictures", type=int, default=24, help="distinct device pictures cycled through (defeats L2 reuse)")
    args = ap.parse_args()
    w, h = [int(v) for v in args.size.split("x")]
    import torch
    from conftest import decoded_planes, DECODED_FORMATS
    og = importlib.import_module("ott-packager_b200")
    lib = og.Library()
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    peak = float(peaks.get("hbm_gbs", 6650.0))
    stream = torch.cuda.Stream()
    out_bytes = lib.frame_bytes(og.FMT_I420, w, h)
    for fmt, (dt, mx, is422) in DECODED_FORMATS.items():
        pix = {"yuv420p": og.PIX_YUV420P, "yuv422p": og.PIX_YUV422P, "yuv420p10le": og.PIX_YUV420P10, "yuv422p10le": og.PIX_YUV422P10}[fmt]
        planes = decoded_planes(fmt, w, h, 1)
        in_bytes = sum(p.nbytes for p in planes)
        cv = og.Converter(lib, pix, w, h, stream=stream.cuda_stream)
        n = args.pictures
        dev = [[torch.from_numpy(p.view(np.uint8).reshape(p.shape[0], -1)).cuda() for p in planes] for _ in range(n)]
        outs = [torch.empty(out_bytes, dtype=torch.uint8, device="cuda") for _ in range(n)]
        pics = [cv.picture([t.data_ptr() for t in d], og.MEM_DEVICE, [t.stride(0) for t in d]) for d in dev]

        def run(count):
            for i in range(count):
                lib.check(lib.dll.ottgpu_converter_convert(cv.handle, pics[i % n], outs[i % n].data_ptr(), og.MEM_DEVICE, og.SUBMIT_ASYNC))

        with torch.cuda.stream(stream):
            run(10)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            run(args.steps)
            e1.record(stream)
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        # end to end: pinned host picture in, packed frame left on the device
        pinned = [torch.from_numpy(np.ascontiguousarray(p).view(np.uint8).reshape(p.shape[0], -1)).pin_memory() for p in planes]
        hpic = cv.picture([t.data_ptr() for t in pinned], og.MEM_DEVICE, [t.stride(0) for t in pinned])
        hpic.mem = og.MEM_HOST
        with torch.cuda.stream(stream):
            for i in range(6):
                lib.check(lib.dll.ottgpu_converter_convert(cv.handle, hpic, outs[i % n].data_ptr(), og.MEM_DEVICE, og.SUBMIT_ASYNC))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            k = max(args.steps // 2, 10)
            for i in range(k):
                lib.check(lib.dll.ottgpu_converter_convert(cv.handle, hpic, outs[i % n].data_ptr(), og.MEM_DEVICE, og.SUBMIT_ASYNC))
            torch.cuda.synchronize()
            e2e_ms = (time.perf_counter() - t0) * 1000.0 / k
        cv.close()
        # batched: one launch converts the pictures of `--batch` converters (a multi-channel process converting one
        # picture per channel per step)
        nb = args.batch
        cvs = [og.Converter(lib, pix, w, h, stream=stream.cuda_stream) for _ in range(nb)]
        import ctypes as C
        handles = (C.c_void_p * nb)(*[c.handle for c in cvs])
        bpics = (og.Picture * nb)()
        for i in range(nb):
            src = pics[i % n]
            bpics[i].mem = src.mem
            for k in range(3):
                bpics[i].plane[k], bpics[i].stride[k] = src.plane[k], src.stride[k]
        bouts = [torch.empty(out_bytes, dtype=torch.uint8, device="cuda") for _ in range(nb)]
        bdst = (C.c_void_p * nb)(*[t.data_ptr() for t in bouts])
        with torch.cuda.stream(stream):
            for _ in range(5):
                lib.check(lib.dll.ottgpu_converter_convert_batch(handles, bpics, bdst, nb, og.MEM_DEVICE, og.SUBMIT_ASYNC))
            torch.cuda.synchronize()
            e0.record(stream)
            reps = max(args.steps // 4, 10)
            for _ in range(reps):
                lib.check(lib.dll.ottgpu_converter_convert_batch(handles, bpics, bdst, nb, og.MEM_DEVICE, og.SUBMIT_ASYNC))
            e1.record(stream)
            torch.cuda.synchronize()
        bms = e0.elapsed_time(e1) / reps / nb
        for c in cvs:
            c.close()
        gbs = (in_bytes + out_bytes) / (ms * 1e-3) / 1e9
        bgbs = (in_bytes + out_bytes) / (bms * 1e-3) / 1e9
        print(json.dumps({"metric": "decode_convert_frames_per_sec", "pix_fmt": fmt, "size": "%dx%d" % (w, h), "value": 1000.0 / ms,
                          "unit": "frames/s", "us_per_frame": ms * 1000.0, "algorithmic_bytes": in_bytes + out_bytes,
                          "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak},
                          "batched": {"pictures_per_launch": nb, "us_per_frame": bms * 1000.0, "value": 1000.0 / bms, "unit": "frames/s",
                                      "roofline": {"bound": "hbm", "achieved": bgbs, "peak": peak, "unit": "GB/s", "frac": bgbs / peak}},
                          "e2e": {"value": 1000.0 / e2e_ms, "unit": "frames/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": 0}}))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
