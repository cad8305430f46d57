"""CPU model check of the parallel bit-exact accumulation (montecarlolocalisation_b200/csrc/exact_scan_core.cuh):
compiles tests/native/exact_scan_model.cpp with g++ and runs it on adversarial weight vectors."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_exact_scan_model(tmp_path):
    exe = str(tmp_path / "xs_model")
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-o", exe,
                    os.path.join(ROOT, "tests", "native", "exact_scan_model.cpp")], check=True)
    r = subprocess.run([exe, "10"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mismatching-cases 0" in r.stdout
