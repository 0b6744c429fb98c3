"""The 70-iteration impulse accumulation (nans_projekat_b200/csrc/solver_accum.cuh, device code) compiled for the
host: the fast (interval) and pipelined (min/max) forms must give the literal loop's bits for every increment
triple -- random magnitudes, the friction-saturation border, zeros, infinities, denormals, NaN dispatch."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
BUILD = os.path.join(HERE, "_build")
f32p = C.POINTER(C.c_float)
KF = 1.4142135623730951 * float(np.float32(0.1))


@pytest.fixture(scope="module")
def accum():
    if not os.path.exists("/usr/local/cuda/include/cuda_runtime.h"):
        pytest.skip("CUDA headers not found")
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "libaccum_host.so")
    src = os.path.join(HERE, "solver_accum_host_shim.cpp")
    hdr = os.path.join(HERE, "..", "nans_projekat_b200", "csrc", "solver_accum.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-shared",
                               "-I/usr/local/cuda/include", "-D__device__=",
                               "-D__forceinline__=inline __attribute__((always_inline))", "-o", out, src])
    lib = C.CDLL(out)
    lib.accum_fast.restype = C.c_long
    return lib


def _compare(lib, ln, lt1, lt2, what, max_fallback_frac=None):
    a = [np.ascontiguousarray(x, np.float32) for x in (ln, lt1, lt2)]
    n = len(a[0])
    ref, fast, pipe = (np.zeros((n, 3), np.float32) for _ in range(3))
    ptr = [x.ctypes.data_as(f32p) for x in a]
    lib.accum_literal(C.c_long(n), *ptr, ref.ctypes.data_as(f32p))
    fb = lib.accum_fast(C.c_long(n), *ptr, fast.ctypes.data_as(f32p))
    lib.accum_pipelined(C.c_long(n), *ptr, pipe.ctypes.data_as(f32p))
    for name, got in (("fast", fast), ("pipelined", pipe)):
        bad = (got.view(np.uint32) != ref.view(np.uint32)) & ~(np.isnan(got) & np.isnan(ref))
        assert not bad.any(), f"{what}/{name}: {int(bad.any(1).sum())} of {n} triples differ, first {a[0][bad.any(1)][0]!r}"
    if max_fallback_frac is not None:
        assert fb <= max_fallback_frac * n, f"{what}: {fb} fallbacks in {n}"


def _logu(rng, n, lo=-30, hi=20):
    return ((2.0 ** rng.uniform(lo, hi, n)) * rng.choice([-1, 1], n)).astype(np.float32)


def test_random_magnitudes(accum):
    rng = np.random.default_rng(0)
    n = 400_000
    _compare(accum, _logu(rng, n), _logu(rng, n), _logu(rng, n), "log-uniform", 1e-3)
    _compare(accum, rng.normal(0.5, 1, n), rng.normal(0, 0.3, n), rng.normal(0, 0.3, n), "contact-like", 1e-3)
    ln = np.abs(_logu(rng, n, -8, 8))
    ratio = 2.0 ** rng.uniform(-7, 1, n) * rng.choice([-1, 1], n)          # |LT| / LN across the saturation ratio 0.1414
    _compare(accum, ln, ln * ratio, ln * ratio[::-1], "ratio sweep", 1e-3)


def test_saturation_border(accum):
    rng = np.random.default_rng(1)
    n = 400_000
    ln = np.abs(_logu(rng, n, -10, 10)).astype(np.float32)
    lt = (KF * ln.astype(np.float64)).astype(np.float32)
    lt = (lt.view(np.int32) + rng.integers(-6, 7, n).astype(np.int32)).view(np.float32)
    lt *= rng.choice([-1, 1], n).astype(np.float32)
    _compare(accum, ln, lt, lt[::-1].copy(), "border")                       # many fallbacks, all must agree


def test_special_values(accum):
    rng = np.random.default_rng(2)
    sp = np.array([0.0, -0.0, np.inf, -np.inf, 1e-45, -1e-45, 1e-38, 3e38, -3e38, 1.0, -1.0, 1e-20, 0.1, 7.0], np.float32)
    g = np.array(np.meshgrid(sp, sp, sp)).reshape(3, -1)
    _compare(accum, g[0], g[1], g[2], "special grid")
    n = 300_000
    mix = lambda: np.where(rng.random(n) < 0.2, rng.choice(sp, n), _logu(rng, n)).astype(np.float32)
    _compare(accum, mix(), mix(), mix(), "special mix")
    _compare(accum, rng.integers(1, 64, n) / 8.0, rng.integers(-64, 64, n) / 64.0, rng.integers(-64, 64, n) / 1024.0, "dyadic")
    x = _logu(rng, 50_000); x[::7] = np.nan
    y = _logu(rng, 50_000); y[::11] = np.nan
    _compare(accum, x, y, _logu(rng, 50_000), "NaN increments take the literal loop")


def test_regime_thresholds(accum):
    """accumulate_fast decides free / saturated by one compare against LN: increments within a few ulps of both
    thresholds (where the chains run closest to the clamp), across the magnitude guards 1e-30 / 1e30, and the
    band in between (pipelined fallback) must all give the literal loop's bits."""
    rng = np.random.default_rng(7)
    n = 1_000_000
    for lo, hi in ((-99, -95), (-40, 40), (95, 101)):       # around 1e-30, the bulk, around 1e30
        ln = (2.0 ** rng.uniform(lo, hi, n)).astype(np.float32)
        for c in (0.1414071, 0.1417043, 0.14142136, 0.1413, 0.1420):
            lt = (np.float32(c) * ln).astype(np.float32)
            lt = (lt.view(np.int32) + rng.integers(-4, 5, n).astype(np.int32)).view(np.float32)
            lt = np.where(np.isfinite(lt), lt, np.float32(1.0)) * rng.choice([-1, 1], n).astype(np.float32)
            _compare(accum, ln, lt, lt[::-1].copy(), f"threshold {c} 2^[{lo},{hi}]")
    ln = np.abs(_logu(rng, n, -20, 20))
    _compare(accum, ln, ln * np.float32(0.05), -ln * np.float32(3.0), "free + saturated", 1e-6)
