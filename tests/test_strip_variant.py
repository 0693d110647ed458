"""The strip-mode work distribution of the 8-bit ladder (device_types.h OTTGPU_STRIP=1: one column group per warp for both
passes, no CTA barrier) is a build option that was measured and not made the default (profiles/README.md).  It must stay
bit-exact: the kernel source is compiled for the CPU with the option on (tests/emu, test-only) and checked against the oracle."""
import os, subprocess
import numpy as np
import pytest
from conftest import ROOT, noise_i420, natural_i420


@pytest.fixture(scope="module")
def strip_lib(og):
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu"), "strip"])
    return og.Library(os.path.join(ROOT, "tests", "emu", "libottgpu_emu_strip.so"))


@pytest.mark.parametrize("w,h,rungs", [
    (640, 360, [(640, 360), (426, 240), (320, 180), (208, 118)]),      # same-size copy, odd ratios, a partial column group
    (854, 480, [(640, 360), (176, 144)]),                              # width not a multiple of 8
    (1920, 1080, [(1280, 720), (416, 234)]),                           # the cfg2 shapes: 12-group tiles, two-lane-per-row tiles
])
@pytest.mark.parametrize("fmt", ["I420", "NV12"])
def test_strip_build_matches_oracle(strip_lib, og, oracle, w, h, rungs, fmt):
    f = getattr(og, "FMT_" + fmt)
    cfg = og.make_cfg(w, h, [(rw, rh, f) for rw, rh in rungs], deint=og.DEINT_FLAGGED)
    ch = og.Channel(strip_lib, cfg)
    frames = [noise_i420(w, h, 3), natural_i420(w, h, 5)]
    outs = [ch.submit(fr, interlaced=0) for fr in frames] + [ch.submit(frames[0], interlaced=0)]
    got = [o for o in outs if o is not None]
    assert len(got) == 2
    for fr, o in zip(frames, got):
        for (rw, rh), plane in zip(rungs, o["rungs"]):
            want = fr if (rw, rh) == (w, h) else oracle.scale_i420(fr, w, h, rw, rh)
            if fmt == "NV12":
                want = oracle.i420_to_nv12(want, rw, rh)
            assert np.array_equal(plane, want), (rw, rh)
    ch.close()
