"""GPU: the parallel exact scan reproduces the left-to-right f64 loop bit for bit on adversarial inputs, and agrees
with the engine's own single-chain kernels."""
import numpy as np
import pytest

import montecarlolocalisation_b200 as m

pytestmark = pytest.mark.gpu


def seq_cumsum(w):
    return np.add.accumulate(w.astype(np.float64))        # numpy accumulates strictly left to right


def cases(rng):
    for n in (1, 2, 7, 255, 2047, 2048, 2049, 4097, 100_003, 1_000_000, 3_000_001):
        yield "sensor-like", (40.0 * rng.random(n)).astype(np.float32)
        yield "wide-range", np.ldexp(rng.random(n), rng.integers(-55, 6, n)).astype(np.float32)
        yield "ties", np.ldexp(rng.integers(0, 64, n).astype(np.float64), -rng.integers(0, 40, n)).astype(np.float32)
        z = (12.0 * rng.random(n)).astype(np.float32)
        z[rng.random(n) < 0.75] = 0
        yield "mostly-zero", z
        lead = rng.random(n).astype(np.float32)
        lead[: n // 3] = 0
        yield "leading-zeros", lead
        yield "normalised", (rng.random(n) / n).astype(np.float32)
        dust = np.ldexp(rng.random(n), -30).astype(np.float32)
        dust[min(7, n - 1)] = 1e6
        yield "giant-then-dust", dust
    yield "all-zero", np.zeros(5000, np.float32)
    yield "nan", np.array([1.0, np.nan, 2.0] * 1000, np.float32)
    yield "inf", np.array([1.0, np.inf, 2.0] * 1000, np.float32)


def test_exact_scan_matches_sequential_loop():
    pf = m.ParticleFilter()
    rng = np.random.default_rng(7)
    fallbacks = {}
    for name, w in cases(rng):
        want = seq_cumsum(w)
        cdf, total, fb = pf.exactScan(w)
        assert np.array_equal(cdf.view(np.uint64), want.view(np.uint64)) or (np.isnan(want).any() and np.array_equal(cdf, want, equal_nan=True)), \
            "%s n=%d first mismatch at %d" % (name, len(w), int(np.argmax(cdf != want)))
        assert total == want[-1] or (np.isnan(total) and np.isnan(want[-1]))
        fallbacks[name] = fallbacks.get(name, 0) + fb
    # realistic weight distributions never need the sequential fallback
    for name in ("sensor-like", "mostly-zero", "leading-zeros", "normalised"):
        assert fallbacks[name] == 0, fallbacks


def test_parallel_path_agrees_with_single_chain_kernels():
    pf = m.ParticleFilter()
    rng = np.random.default_rng(8)
    w = (rng.random(500_000) * np.ldexp(1.0, rng.integers(-20, 4, 500_000))).astype(np.float32)
    a, ta, fba = pf.exactScan(w)
    pf.forceSequential(True)
    b, tb, fbb = pf.exactScan(w)
    pf.forceSequential(False)
    assert fba == 0 and fbb == 1
    assert np.array_equal(a, b) and ta == tb
