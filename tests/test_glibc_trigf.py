"""csrc/glibc_trigf.cuh restates glibc 2.39's sinf/cosf (what the reference binary calls at MC:644-645 and MC:747-748) so the
engine's float trig can be bit-identical to the compiled reference. CPU: the host instantiation against this machine's
libm over a 1-in-61 sample of all float bit patterns (the exhaustive run, stride 1, takes 35 s on 8 cores and was green for
the FMA build on this image: DESIGN.md). GPU: the device instantiation the kernels call, against the host libm."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# the only arguments (and their negatives) at which glibc's FMA and SSE2 builds of sinf / cosf differ (exhaustive search)
DIFFER = [0x4255b0a9, 0x42a35c07, 0x42a35d44, 0x42a97360, 0x42cf5854, 0x42e87a55, 0x418a3adb, 0x418a3adc, 0x418a3add, 0x418a3ade,
          0x41bc76d9, 0x4202eb4b, 0x42687a55, 0x4280ce28, 0x42870e40, 0x42c55faa, 0x42d8d23e]


def test_host_restatement_equals_libm_on_a_sample_of_all_floats(tmp_path):
    exe = str(tmp_path / "glibc_trigf_check")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-mfma", "-pthread", "-o", exe,
                    os.path.join(ROOT, "tests", "native", "glibc_trigf_check.cpp"), "-lm"], check=True)
    r = subprocess.run([exe, "61"], capture_output=True, text=True)
    assert r.returncode == 0 and "reproduced bit for bit" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_device_trig_equals_host_libm():
    """2^24 random bit patterns, every exponent's edges, the arguments where the two glibc builds differ, NaN and Inf."""
    import montecarlolocalisation_b200 as m
    from oracle.pyoracle import libm_trigf
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 2**32, 1 << 24, dtype=np.uint64).astype(np.uint32)
    edges = np.array([(e << 23) + d for e in range(256) for d in (0, 1, 0x7fffff, 0x400000)], np.uint32)
    special = np.array(DIFFER + [b | 0x80000000 for b in DIFFER] + [0x7f800000, 0xff800000, 0x7fc00000, 0x7f800001, 0xffc12345, 0, 0x80000000],
                       np.uint32)
    typical = rng.uniform(-7.0, 7.0, 1 << 22).astype(np.float32).view(np.uint32)          # where particle headings live
    x = np.concatenate([bits, edges, edges | np.uint32(0x80000000), special, typical]).view(np.float32)
    pf = m.ParticleFilter()
    s, c, kind = pf.trigf(x)
    hs, hc = libm_trigf(x)
    assert kind in (0, 1)                                                # MCL_TRIG_LIBM resolved to one of glibc's builds
    assert np.array_equal(s.view(np.uint32), hs.view(np.uint32))
    assert np.array_equal(c.view(np.uint32), hc.view(np.uint32))
    # the portable mode: (float)sin((double)x), within one ulp of libm everywhere and different from it somewhere
    pc = m.ParticleFilter(trig_mode=m.TRIG_CORRECTLY_ROUNDED)
    s2, c2, kind2 = pc.trigf(typical.view(np.float32))
    ref = np.sin(typical.view(np.float32).astype(np.float64)).astype(np.float32)
    assert kind2 == 2 and (s2 != ref).mean() < 1e-6
    assert 0 < (s2 != hs[-len(typical):]).mean() < 0.02
