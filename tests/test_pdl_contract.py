"""Programmatic dependent launch contract (csrc/engine_internal.hpp, csrc/mcl_device.cuh), checked on the sources and on
the built library, no GPU needed: a kernel launched with LAUNCH_PDL may be made resident before its predecessor has
finished, so it must execute pdl_enter() (griddepcontrol.wait) before it touches anything a kernel of the tick writes.
A kernel that is launched that way without the wait would race silently - this test is the guard."""
import os
import re
import subprocess

import pytest

from montecarlolocalisation_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "montecarlolocalisation_b200", "csrc")

# kernels that read only tables no tick kernel writes (map, ray directions, beams, field) before the wait: the names of the
# only device buffers their pre-wait code may mention
LATE_WAIT = {"k_ref_update_v2": ("P.radii", "P.lut", "P.beams", "P.occ", "P.occ_pad", "P.gauss", "ref_gauss(P"),
             "k_ns_update": ("F.lf", "F.codes", "beams[")}


def _read(name):
    return open(os.path.join(CSRC, name)).read()


def _strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def pdl_launched_kernels():
    names = set()
    for f in ("mcl_engine.cu", "mcl_engine_ns.cu", "mcl_engine_next.cu"):
        src = _strip_comments(_read(f))
        for m_ in re.finditer(r"\bLAUNCH_PDL\(\s*\w+\s*,\s*\(?\s*((?:\w+::)*\w+)", src):
            k = m_.group(1).split("::")[-1]
            names.add("k_ns_update" if k == "kernel" else k)      # ns_launch_update launches k_ns_update<KIND, PACK> through a lambda
    return sorted(names)


def kernel_body(name):
    for f in ("kernels_ref.cuh", "kernels_ns.cuh", "kernels_next.cuh", "exact_scan.cuh", "exact_scan_fused.cuh"):
        src = _strip_comments(_read(f))
        m_ = re.search(r"__global__\s+void\s+(?:__launch_bounds__\([^{;]*?\)\s+)?" + name + r"\s*\(", src)
        if not m_:
            continue
        i, depth = m_.end(), 1
        while depth:
            depth += {"(": 1, ")": -1}.get(src[i], 0)
            i += 1
        j = src.index("{", i)
        k, depth = j + 1, 1
        while depth:
            depth += {"{": 1, "}": -1}.get(src[k], 0)
            k += 1
        return src[j + 1:k - 1]
    raise AssertionError("no definition of " + name)


def test_every_pdl_launched_kernel_waits_first():
    names = pdl_launched_kernels()
    assert len(names) >= 20 and "k_ref_resample" in names and "k_ns_update" in names
    for name in names:
        body = kernel_body(name)
        assert "pdl_enter();" in body, name + " is launched with LAUNCH_PDL but never calls pdl_enter()"
        before = body.split("pdl_enter();")[0]
        if name in LATE_WAIT:
            # everything dereferenced before the wait must be one of the static tables
            for tok in re.findall(r"\b(?:part|w_dense|ll_out|max_bits)\s*\[", before):
                raise AssertionError("%s touches %s before pdl_enter()" % (name, tok))
            assert any(t in before for t in LATE_WAIT[name])
        else:
            assert before.strip() == "", name + ": pdl_enter() must be the first statement, found before it: " + before.strip()[:80]


def test_pdl_enter_is_wait_then_launch_dependents():
    dev = _strip_comments(_read("mcl_device.cuh"))
    body = dev[dev.index("void pdl_enter()"):]
    body = body[:body.index("}")]
    assert body.index("griddepcontrol.wait") < body.index("griddepcontrol.launch_dependents")


def test_built_library_carries_the_instructions():
    """SASS of libmcl_b200.so: one ACQBULK (griddepcontrol.wait) and one PREEXIT (launch_dependents) per instantiation of a
    programmatically launched kernel."""
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    try:
        sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    except (OSError, subprocess.TimeoutExpired):
        pytest.skip("cuobjdump not available")
    if "Function :" not in sass:
        pytest.skip("cuobjdump gave no SASS")
    per_fn = {}
    for chunk in sass.split("Function :")[1:]:
        fn = chunk.split("\n", 1)[0].strip()
        per_fn[fn] = (chunk.count("ACQBULK"), chunk.count("PREEXIT"))
    for name in pdl_launched_kernels():
        hits = [v for k, v in per_fn.items() if re.search(r"\d+" + name + r"(?:I|E|P|v)", k)]
        assert hits, "no SASS for " + name
        for acq, pre in hits:
            assert acq >= 1 and pre >= 1, (name, acq, pre)
