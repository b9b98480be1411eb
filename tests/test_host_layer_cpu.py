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


def test_cxx_partitioner_matches_the_python_one(tmp_path):
    """gmix_b200/host/shard.h (what `gmixb200 -C` shards its streams over the GPUs with) against gmix_b200/shard.py (what the
    torch.distributed harness uses), on random length lists incl. empty streams, all-empty lists and more ranks than streams."""
    import random
    from gmix_b200 import shard
    harness = tmp_path / "shard_check.cpp"
    harness.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "%s/gmix_b200/host/shard.h"
int main(int argc, char** argv) {   // world n len0 len1 ... -> "lo hi" per rank
  const int world = atoi(argv[1]);
  std::vector<uint64_t> len;
  for (int i = 3; i < argc; ++i) len.push_back(strtoull(argv[i], nullptr, 10));
  for (auto& r : gmixb::ShardRanges(len, world)) printf("%%u %%u\n", r.first, r.second);
  const unsigned char abc[3] = {'a', 'b', 'c'};
  printf("%%llu\n", (unsigned long long)gmixb::Fnv1a64(abc, 3));
  return 0;
}
''' % ROOT)
    exe = tmp_path / "shard_check"
    subprocess.run(["g++", "-std=c++17", "-O2", "-o", str(exe), str(harness)], check=True)
    rng = random.Random(5)
    cases = [([], 4), ([0, 0, 0], 2), ([5], 8), ([1000000] * 100, 8), ([1000000] * 100, 3)]
    for _ in range(40):
        n = rng.randint(1, 60)
        cases.append(([rng.choice([0, 1, 7, 4096, 65536, rng.randint(0, 10 ** 6)]) for _ in range(n)], rng.randint(1, 9)))
    for lengths, world in cases:
        r = subprocess.run([str(exe), str(world), str(len(lengths))] + [str(x) for x in lengths], capture_output=True, text=True, check=True)
        lines = r.stdout.split("\n")
        got = [tuple(int(x) for x in l.split()) for l in lines[:world]]
        assert got == [tuple(x) for x in shard.shard_ranges(lengths, world)], (lengths, world)
        assert int(lines[world]) == shard.fnv1a64(b"abc")
