// Small stand-alone driver of the C ABI for profiler captures (ncu --replay-mode application re-runs the
// whole process once per pass, so it must start fast: no Python, no torch).
//   ncu_case <n_streams> <stream_bytes> [input_file]
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../include/gmix_b200.h"

int main(int argc, char** argv) {
  const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 8;
  const uint64_t len = argc > 2 ? (uint64_t)atoll(argv[2]) : 1024;
  const char* path = argc > 3 ? argv[3] : "tests/data/english.dic";
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); return 2; }
  std::vector<uint8_t> file;
  uint8_t buf[65536];
  size_t got;
  while ((got = fread(buf, 1, sizeof(buf), f)) > 0) file.insert(file.end(), buf, buf + got);
  fclose(f);
  std::vector<uint8_t> in((size_t)n * len);
  for (uint32_t i = 0; i < n; ++i)
    for (uint64_t k = 0; k < len; ++k) in[i * len + k] = file[((uint64_t)i * 3001 + k) % file.size()];
  const uint64_t cap = gmx_compress_bound(len);
  std::vector<uint64_t> in_off(n + 1), out_off(n + 1), out_len(n);
  for (uint32_t i = 0; i <= n; ++i) { in_off[i] = i * len; out_off[i] = i * cap; }
  std::vector<uint8_t> out((size_t)n * cap);
  std::vector<uint32_t> status(n);
  gmx_ctx* c = nullptr;
  if (gmx_create(0, &c)) { fprintf(stderr, "gmx_create: %s\n", gmx_global_error()); return 1; }
  if (gmx_configure(c, len, n)) { fprintf(stderr, "gmx_configure: %s\n", gmx_last_error(c)); return 1; }
  if (gmx_compress_batch(c, in.data(), in_off.data(), n, out.data(), out_off.data(), out_len.data(), status.data())) {
    fprintf(stderr, "gmx_compress_batch: %s\n", gmx_last_error(c));
    return 1;
  }
  uint64_t total = 0;
  for (uint32_t i = 0; i < n; ++i) total += out_len[i];
  printf("%u x %llu B -> %llu B, kernel %.2f ms, resident %u, arena %llu MiB\n", n, (unsigned long long)len,
         (unsigned long long)total, gmx_last_kernel_ms(c), gmx_resident_streams(c), (unsigned long long)(gmx_arena_bytes(c) >> 20));
  gmx_destroy(c);
  return 0;
}
