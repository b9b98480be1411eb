// Sweeps gmix_b200/csrc/dmath.cuh (compiled for the host) against the host glibc libm:
// gm_expf and gm_tanhf on all 2^32 bit patterns, gm_logf on all positive normal floats
// (its documented domain). Prints mismatch counts; exit code 0 iff all are zero (NaN payloads
// compare by NaN-ness). Usage: dmath_check [stride]  (stride>1 subsamples for quick CI runs)
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include <thread>
#include <vector>
#include "../../gmix_b200/csrc/dmath.cuh"

static bool same(float a, float b) {
  if (a != a && b != b) return true;
  return gmx::f2u(a) == gmx::f2u(b);
}

int main(int argc, char** argv) {
  uint64_t stride = argc > 1 ? strtoull(argv[1], 0, 10) : 1;
  unsigned nt = std::thread::hardware_concurrency();
  if (!nt) nt = 4;
  std::atomic<uint64_t> bad_exp{0}, bad_log{0}, bad_tanh{0}, bad_expm1{0};
  std::atomic<uint32_t> first_exp{0}, first_log{0}, first_tanh{0};
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t] {
    uint64_t be = 0, bl = 0, bt = 0, bm = 0;
    for (uint64_t u = t * stride; u < (1ull << 32); u += nt * stride) {
      float x = gmx::u2f((uint32_t)u);
      if (!same(gmx::gm_expf(x), expf(x))) { if (!be++) first_exp = (uint32_t)u; }
      if (!same(gmx::gm_tanhf(x), tanhf(x))) { if (!bt++) first_tanh = (uint32_t)u; }
      if (!same(gmx::gm_expm1f(x), expm1f(x))) bm++;
      if (u >= 0x00800000u && u < 0x7f800000u && !same(gmx::gm_logf(x), logf(x))) { if (!bl++) first_log = (uint32_t)u; }
    }
    bad_exp += be; bad_log += bl; bad_tanh += bt; bad_expm1 += bm;
  });
  for (auto& x : th) x.join();
  printf("expf mismatches %llu (first 0x%08x)\nlogf mismatches %llu (first 0x%08x)\ntanhf mismatches %llu (first 0x%08x)\nexpm1f mismatches %llu\n",
         (unsigned long long)bad_exp, first_exp.load(), (unsigned long long)bad_log, first_log.load(),
         (unsigned long long)bad_tanh, first_tanh.load(), (unsigned long long)bad_expm1);
  return (bad_exp || bad_log || bad_tanh) ? 1 : 0;
}
