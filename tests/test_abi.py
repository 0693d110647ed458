"""CPU tier: libottgpu.so builds, loads, exports every symbol include/ottgpu.h declares, and its host-side arithmetic
(polyphase tables = what sws_getContext builds, transvideo.c:1563-1565; buffer geometry, transvideo.c:1559) is right.
No compute call is made here."""
import os, re
import numpy as np
import pytest
from conftest import ROOT


def _declared():
    text = open(os.path.join(ROOT, "include", "ottgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ottgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_and_library_agree(real_lib, og):
    names = _declared()
    assert names == og.EXPORTS, (set(names) ^ set(og.EXPORTS))
    for n in names:
        assert hasattr(real_lib.dll, n), n
    assert real_lib.dll.ottgpu_abi_version() == 1


def test_no_oracle_or_fallback_in_product_sources():
    pkg = os.path.join(ROOT, "ott-packager_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cpp", ".cu", ".cuh", ".h", ".c", "Makefile")):
                src = open(os.path.join(dirpath, f)).read()
                assert "liboracle" not in src and "import oracle" not in src and "from oracle" not in src, f
                assert "libottgpu_emu" not in src, f


@pytest.mark.parametrize("s,d", [(1920, 1280), (1920, 960), (1920, 640), (1920, 416), (1920, 176), (1080, 720), (1080, 234),
                                 (1080, 144), (3840, 2560), (3840, 640), (2160, 360), (720, 1080), (854, 427), (1920, 854),
                                 (1080, 1080), (1080, 1076), (640, 1920), (96, 30), (46, 22)])
def test_polyphase_tables_equal_oracle(real_lib, oracle, og, s, d):
    for filt, kern in ((og.FILTER_BICUBIC, oracle.BICUBIC), (og.FILTER_LANCZOS, oracle.LANCZOS)):
        for target in (og.TARGET_A, og.TARGET_B):
            for vertical in (0, 1):
                taps, pos, coef = real_lib.filter_table(s, d, filt, target, vertical)
                align = 1 if target == og.TARGET_B else (2 if vertical else 4)
                rt, rp, rc = oracle.make_filter(s, d, kern, align, (1 << 12) if vertical else (1 << 14), target == og.TARGET_B)
                assert taps == rt and np.array_equal(pos, rp) and np.array_equal(coef, rc), (s, d, filt, target, vertical)


def test_geometry(real_lib, og):
    assert real_lib.frame_bytes(og.FMT_I420, 1920, 1080) == 3110400
    assert real_lib.frame_bytes(og.FMT_P010, 3840, 2160) == 24883200
    assert real_lib.rung_bytes(og.RungCfg(416, 234, og.FMT_NV12, 0)) == 146016
    assert real_lib.rung_bytes(og.RungCfg(176, 144, og.FMT_I420, 0)) == 38016
    lay = real_lib.rung_layout(og.RungCfg(854, 480, og.FMT_I420, 0))
    assert lay == [(0, 854), (854 * 480, 427), (854 * 480 + 427 * 240, 427)]
    lay = real_lib.rung_layout(og.RungCfg(1280, 720, og.FMT_NV12, 256))
    assert lay == [(0, 1280), (1280 * 720, 1280)]
    lay = real_lib.rung_layout(og.RungCfg(416, 234, og.FMT_NV12, 256))
    assert lay == [(0, 512), (512 * 234, 512)]
    lay = real_lib.rung_layout(og.RungCfg(640, 360, og.FMT_P010, 0))
    assert lay == [(0, 1280), (1280 * 360, 1280)]


def test_compute_fails_loudly_without_gpu(real_lib, og):
    from conftest import _have_gpu
    if _have_gpu():
        pytest.skip("a GPU is present")
    src = np.zeros(real_lib.frame_bytes(og.FMT_I420, 64, 48), np.uint8)
    with pytest.raises(og.OttGpuError):
        real_lib.scale_frame(src, og.FMT_I420, 64, 48, 32, 24)
    with pytest.raises(og.OttGpuError):
        og.Channel(real_lib, og.make_cfg(64, 48, [(32, 24, og.FMT_I420)]))
