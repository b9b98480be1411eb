"""gmix_b200/csrc/dmath.cuh compiled for the host against this box's glibc libm (expf, logf, tanhf,
expm1f), strided sweep of the 2^32 bit patterns (the full sweep is `dmath_check 1`, ~30 s on 8 cores;
the GPU-side full sweep is gmx_selftest_math, run by tests/test_gpu_parity.py with a stride and in
full by scripts/gpu_run1.sh)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))


def test_device_math_models_match_host_libm(tmp_path):
    exe = str(tmp_path / "dmath_check")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-pthread", os.path.join(HERE, "native", "dmath_check.cpp"), "-o", exe], check=True)
    r = subprocess.run([exe, "61"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    assert "expf mismatches 0" in r.stdout and "logf mismatches 0" in r.stdout and "tanhf mismatches 0" in r.stdout
