"""Shared fixtures / generators.  GPU tests are marked @pytest.mark.gpu (run on the B200 box only)."""
import os, sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); skipped by -m 'not gpu'")


# ---------------------------------------------------------------- synthetic frames
def chroma_dims(w, h):
    return (w + 1) // 2, (h + 1) // 2


def noise_i420(w, h, seed=1234):
    """uniform noise, planes generated in order Y,U,V -> packed I420 (strides w, ceil(w/2))"""
    cw, ch = chroma_dims(w, h)
    rng = np.random.default_rng(seed)
    y = rng.integers(0, 256, (h, w), dtype=np.uint8)
    u = rng.integers(0, 256, (ch, cw), dtype=np.uint8)
    v = rng.integers(0, 256, (ch, cw), dtype=np.uint8)
    return np.concatenate([y.ravel(), u.ravel(), v.ravel()])


def noise_p010(w, h, seed=99):
    """10-bit noise MSB-aligned in 16 bits: Y plane then interleaved UV plane"""
    cw, ch = chroma_dims(w, h)
    rng = np.random.default_rng(seed)
    y = (rng.integers(0, 1024, (h, w), dtype=np.uint16) << 6).astype(np.uint16)
    uv = (rng.integers(0, 1024, (ch, cw * 2), dtype=np.uint16) << 6).astype(np.uint16)
    return np.concatenate([y.ravel(), uv.ravel()])


DECODED_FORMATS = {"yuv420p": (np.uint8, 256, False), "yuv422p": (np.uint8, 256, True),
                   "yuv420p10le": (np.uint16, 1024, False), "yuv422p10le": (np.uint16, 1024, True)}


def decoded_planes(fmt, w, h, seed=1, kind="noise", pad=0):
    """Three planes as a decoder hands them over (transvideo.c:2788-2795): Y, U, V with `pad` extra samples per row
    (linesize > width).  kind: noise | extreme (0 / max checker noise) | ramp."""
    dt, mx, is422 = DECODED_FORMATS[fmt]
    cw, ch = (w + 1) // 2, (h if is422 else (h + 1) // 2)
    rng = np.random.default_rng(seed)
    out = []
    for pw, ph in ((w, h), (cw, ch), (cw, ch)):
        if kind == "noise":
            a = rng.integers(0, mx, (ph, pw + pad))
        elif kind == "extreme":
            a = rng.integers(0, 2, (ph, pw + pad)) * (mx - 1)
        else:
            a = (np.arange(pw + pad)[None, :] * 3 + np.arange(ph)[:, None] * 5) % mx
        out.append(np.ascontiguousarray(a.astype(dt))[:, :pw])
    return out


def natural_i420(w, h, seed=7, shift=0.0):
    cw, ch = chroma_dims(w, h)
    rng = np.random.default_rng(seed)

    def plane(pw, ph, k):
        x = np.arange(pw)[None, :] * k + shift
        y = np.arange(ph)[:, None] * k
        v = 0.5 + 0.25 * np.sin(x / 37.0) + 0.2 * np.cos(y / 23.0)
        v = v + 0.05 * ((((x // 64).astype(int) + (y // 64).astype(int)) & 1) * 2 - 1)
        v = v + rng.normal(0, 0.01, (ph, pw))
        return np.clip(v, 0.0625, 0.92)

    y = (plane(w, h, 1) * 255).astype(np.uint8)
    u = (plane(cw, ch, 2) * 255).astype(np.uint8)
    v = (plane(cw, ch, 2)[::-1] * 255).astype(np.uint8)
    return np.concatenate([y.ravel(), u.ravel(), np.ascontiguousarray(v).ravel()])


def interlaced_sequence(w, h, n, seed=5, tff=True, speed=4, bps=1):
    """n woven frames of a bright bar moving `speed` px per FIELD over a noisy textured background."""
    cw, ch = chroma_dims(w, h)
    rng = np.random.default_rng(seed)
    mx = 255 if bps == 1 else 1023
    dt = np.uint8 if bps == 1 else np.uint16
    bg = rng.integers(mx // 8, mx // 2, (h, w))
    bgc = rng.integers(mx // 4, 3 * mx // 4, (2, ch, cw))
    frames = []
    for f in range(n):
        fields = []
        for k in range(2):
            t = 2 * f + k
            img = bg + rng.integers(0, 6, (h, w))
            x0 = (t * speed) % max(w - 16, 1)
            img[:, x0:x0 + 12] = mx - 20
            fields.append(img)
        first, second = fields
        y = np.empty((h, w), dt)
        top, bot = (first, second) if tff else (second, first)
        y[0::2] = top[0::2]
        y[1::2] = bot[1::2]
        c = []
        for p in range(2):
            cp = np.empty((ch, cw), dt)
            a = bgc[p] + (2 * f) % 7
            b = bgc[p] + (2 * f + 1) % 7
            ta, tb = (a, b) if tff else (b, a)
            cp[0::2] = ta[0::2]
            cp[1::2] = tb[1::2]
            c.append(cp)
        frames.append(np.concatenate([y.ravel(), c[0].ravel(), c[1].ravel()]))
    return frames


def split_i420(buf, w, h):
    cw, ch = chroma_dims(w, h)
    y = buf[: w * h].reshape(h, w)
    u = buf[w * h: w * h + cw * ch].reshape(ch, cw)
    v = buf[w * h + cw * ch: w * h + 2 * cw * ch].reshape(ch, cw)
    return y, u, v


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O


@pytest.fixture(scope="session")
def sws():
    from oracle import swscale_so as S
    if not S.available():
        pytest.skip("bundled libswscale (opencv_python_headless.libs) not present")
    S.load()
    return S


# ---------------------------------------------------------------- the library under test
import importlib
import subprocess


def og_module():
    return importlib.import_module("ott-packager_b200")


def _have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


@pytest.fixture(scope="session")
def og():
    return og_module()


@pytest.fixture(scope="session")
def real_lib(og):
    """libottgpu.so itself (host-arithmetic entry points work without a GPU; compute calls need the B200)."""
    if not os.path.exists(og.DEFAULT_LIB):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "ott-packager_b200", "csrc")])
    return og.Library()


@pytest.fixture(scope="session")
def emu_lib(og):
    """TEST-ONLY CPU emulation of the kernel source (tests/emu) behind the same C ABI."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu")])
    return og.Library(os.path.join(ROOT, "tests", "emu", "libottgpu_emu.so"))


@pytest.fixture(scope="session", params=[pytest.param("emu"), pytest.param("gpu", marks=pytest.mark.gpu)])
def lib(request, og):
    """Parity tests run twice: against the CPU emulation of the kernel source (CPU tier: catches logic errors before
    GPU time is spent) and, marked gpu, against the real sm_100a kernels through libottgpu.so."""
    if request.param == "emu":
        return request.getfixturevalue("emu_lib")
    if not _have_gpu():
        pytest.fail("gpu tier requested but no CUDA device is visible (there is no CPU fallback)")
    return request.getfixturevalue("real_lib")
