"""CPU-side checks of the C++ host layer: the runner builds, mirrors the reference CLI's argument
handling, and fails loudly (no CPU fallback) when no GPU is present. The host coder is checked against
the oracle's coder through a tiny C++ harness."""
import os
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
RUNNER = os.path.join(ROOT, "gmix_b200", "lib", "gmixb200")


@pytest.fixture(scope="module", autouse=True)
def _built():
    from gmix_b200.build import build_library
    build_library()
    assert os.path.exists(RUNNER)


def test_runner_help_and_argument_errors():
    r = subprocess.run([RUNNER], capture_output=True, text=True)
    assert r.returncode != 0 and "Compress" in r.stdout
    r = subprocess.run([RUNNER, "-g", "ckpt", "prompt", "out", "10", "1.0"], capture_output=True, text=True)
    assert r.returncode != 0 and "Can not open: prompt" in r.stdout          # reference wording, runner-utils.cpp:164
    r = subprocess.run([RUNNER, "-g", "ckpt", "prompt", "out"], capture_output=True, text=True)
    assert r.returncode != 0 and "Wrong number of arguments" in r.stdout     # runner.cpp:41
    r = subprocess.run([RUNNER, "-t", "a", "b", "c", "d"], capture_output=True, text=True)
    assert r.returncode != 0 and "Wrong number of arguments" in r.stdout
    r = subprocess.run([RUNNER, "-c", "/nonexistent/in", "/tmp/out"], capture_output=True, text=True)
    assert r.returncode != 0 and "Error opening" in r.stdout


def test_runner_has_no_cpu_fallback(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    src = tmp_path / "in"
    src.write_bytes(b"hello")
    r = subprocess.run([RUNNER, "-c", str(src), str(tmp_path / "out")], capture_output=True, text=True)
    assert r.returncode != 0
    assert "no CUDA device" in r.stdout or "CUDA" in r.stdout
    assert not (tmp_path / "out").exists()


def test_host_coder_roundtrip(tmp_path):
    """gmix_b200/host/coder.h: Encoder/Decoder are inverse for an arbitrary probability sequence, and the
    encoder's bytes equal the oracle's coder bytes for the same (bit, p) sequence."""
    harness = tmp_path / "coder_check.cpp"
    harness.write_text(r'''
#include <stdio.h>
#include <vector>
#include "%s/gmix_b200/host/coder.h"
int main() {
  std::vector<int> bits; std::vector<float> ps;
  unsigned long long x = 88172645463325252ull;
  for (int i = 0; i < 20000; ++i) {
    x ^= x << 13; x ^= x >> 7; x ^= x << 17;
    float p = 0.0001f + 0.9998f * (float)((x >> 11) & 0xffff) / 65535.0f;
    ps.push_back(p); bits.push_back(((x >> 40) & 0xffff) < (unsigned)(p * 65536.0f));
  }
  std::vector<uint8_t> out;
  gmixb::Encoder e(&out);
  for (size_t i = 0; i < bits.size(); ++i) e.Encode(bits[i], ps[i]);
  e.Flush();
  gmixb::Decoder d(out.data(), out.size());
  for (size_t i = 0; i < bits.size(); ++i) if (d.Decode(ps[i]) != bits[i]) { printf("mismatch at %%zu\n", i); return 1; }
  printf("ok %%zu\n", out.size());
  return 0;
}
''' % ROOT)
    exe = tmp_path / "coder_check"
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-o", str(exe), str(harness)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout
