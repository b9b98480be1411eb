// gmixb200 — command-line runner over libgmix_b200.so, mirroring the reference CLI (reference
// src/runner/runner.cpp:14-104, src/runner/runner-utils.cpp:88-156):
//
//   gmixb200 -c input output        compress, one stream; bytes identical to `gmix -c` (strict build)
//   gmixb200 -d input output        decompress a stream written by either program
//   gmixb200 -C bytes input output  split input into chunks of `bytes`, compress them as independent
//                                   streams in one GPU batch (container: "GMXB", u32 count, u64 sizes[], streams)
//   gmixb200 -D input output        inverse of -C
//   gmixb200 -p input output        compress one stream through the Predictor facade + host coder
//                                   (Predict/Perceive/Learn per bit; slow, for checking the drop-in interface)
//
// Not implemented on the GPU path yet: the optional checkpoint argument, -g and -t.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <fstream>
#include <iterator>
#include <string>
#include <vector>

#include "../../include/gmix_b200.h"
#include "coder.h"
#include "predictor.h"

namespace {

int Help() {
  printf("gmixb200 (B200 build of the gmix per-bit path)\n"
         "Compress:    gmixb200 -c input output\n"
         "Decompress:  gmixb200 -d input output\n"
         "Chunked:     gmixb200 -C chunk_bytes input output   /   gmixb200 -D input output\n"
         "Via facade:  gmixb200 -p input output\n"
         "Checkpoints, -g (generate) and -t (train) are not available on the GPU path yet.\n");
  return -1;
}

bool ReadFile(const std::string& path, std::vector<uint8_t>* data) {
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) return false;
  data->assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
  return true;
}
bool WriteFile(const std::string& path, const uint8_t* p, size_t n) {
  std::ofstream f(path, std::ios::binary);
  if (!f.is_open()) return false;
  f.write((const char*)p, (std::streamsize)n);
  return f.good();
}

struct Batch {
  std::vector<uint8_t> out;
  std::vector<uint64_t> out_off, out_len;
};

// n streams in[in_off[i]..in_off[i+1]) -> Batch, through the C ABI.
bool RunBatch(gmx_ctx* ctx, bool compress, const std::vector<uint8_t>& in, const std::vector<uint64_t>& in_off, Batch* b) {
  const uint32_t n = (uint32_t)in_off.size() - 1;
  b->out_off.assign(n + 1, 0);
  for (uint32_t i = 0; i < n; ++i) {
    uint64_t cap;
    if (compress) cap = gmx_compress_bound(in_off[i + 1] - in_off[i]);
    else {
      cap = 0;
      for (uint64_t k = 0; k < 5 && in_off[i] + k < in_off[i + 1]; ++k) cap = (cap << 8) + in[in_off[i] + k];
      cap += 8;
    }
    b->out_off[i + 1] = b->out_off[i] + cap;
  }
  b->out.assign(b->out_off[n] + 1, 0);
  b->out_len.assign(n, 0);
  std::vector<uint32_t> status(n, 0);
  static const uint8_t kNone = 0;
  const uint8_t* src = in.empty() ? &kNone : in.data();
  const int rc = compress ? gmx_compress_batch(ctx, src, in_off.data(), n, b->out.data(), b->out_off.data(), b->out_len.data(), status.data())
                          : gmx_decompress_batch(ctx, src, in_off.data(), n, b->out.data(), b->out_off.data(), b->out_len.data(), status.data());
  if (rc != 0) { printf("%s\n", gmx_last_error(ctx)); return false; }
  return true;
}

bool CompressViaPredictor(const std::vector<uint8_t>& in, std::vector<uint8_t>* out) {
  gmixb::Gpu gpu(0);
  gmixb::Predictor p(gpu, in.size() + 1);
  const uint64_t n = in.size();
  for (int i = 4; i >= 0; --i) out->push_back((uint8_t)(n >> (8 * i)));   // WriteHeader runner-utils.cpp:22-27
  p.EnableAnalysis((int)(8 * n / 1000));                                    // runner-utils.cpp:47
  gmixb::Encoder e(out);
  for (uint64_t pos = 0; pos < n; ++pos) {
    for (int j = 7; j >= 0; --j) {                                          // runner-utils.cpp:50-58
      const int bit = (in[pos] >> j) & 1;
      e.Encode(bit, p.Predict());
      p.Perceive(bit);
      p.Learn();
    }
  }
  e.Flush();
  return true;
}

}  // namespace

int main(int argc, char* argv[]) {
  if (argc < 4 || strlen(argv[1]) != 2 || argv[1][0] != '-') return Help();
  const char mode = argv[1][1];
  if (mode == 'g' || mode == 't') { printf("-%c is not available on the GPU path yet.\n", mode); return Help(); }
  if ((mode == 'c' || mode == 'd') && argc == 5) { printf("Checkpoints are not available on the GPU path yet.\n"); return Help(); }
  const bool chunked_c = mode == 'C';
  if ((chunked_c && argc != 5) || (!chunked_c && argc != 4) || !strchr("cdCDp", mode)) return Help();
  const std::string input_path = argv[chunked_c ? 3 : 2], output_path = argv[chunked_c ? 4 : 3];
  std::vector<uint8_t> in;
  if (!ReadFile(input_path, &in)) { printf("Error opening: %s\n", input_path.c_str()); return Help(); }
  const clock_t start = clock();
  std::vector<uint8_t> result;
  try {
    if (mode == 'p') {
      if (!CompressViaPredictor(in, &result)) return -1;
    } else {
      gmixb::Gpu gpu(0);
      std::vector<uint64_t> in_off{0};
      std::vector<uint8_t> payload;
      const std::vector<uint8_t>* src = &in;
      if (mode == 'c' || mode == 'd') {
        in_off.push_back(in.size());
      } else if (mode == 'C') {
        const uint64_t chunk = strtoull(argv[2], nullptr, 10);
        if (chunk == 0) return Help();
        for (uint64_t o = 0; o < in.size(); o += chunk) in_off.push_back(o + chunk < in.size() ? o + chunk : in.size());
        if (in.empty()) in_off.push_back(0);
      } else {  // 'D': parse the container
        if (in.size() < 8 || memcmp(in.data(), "GMXB", 4) != 0) { printf("Not a gmixb200 -C container.\n"); return -1; }
        uint32_t count;
        memcpy(&count, &in[4], 4);
        if (in.size() < 8 + 8ull * count) { printf("Truncated container.\n"); return -1; }
        uint64_t off = 0;
        for (uint32_t i = 0; i < count; ++i) { uint64_t sz; memcpy(&sz, &in[8 + 8ull * i], 8); off += sz; in_off.push_back(off); }
        payload.assign(in.begin() + 8 + 8ull * count, in.end());
        if (payload.size() < off) { printf("Truncated container.\n"); return -1; }
        src = &payload;
      }
      Batch b;
      if (!RunBatch(gpu.ctx(), mode == 'c' || mode == 'C', *src, in_off, &b)) return -1;
      const uint32_t n = (uint32_t)in_off.size() - 1;
      if (mode == 'C') {
        result.assign({'G', 'M', 'X', 'B'});
        result.resize(8 + 8ull * n);
        memcpy(&result[4], &n, 4);
        for (uint32_t i = 0; i < n; ++i) memcpy(&result[8 + 8ull * i], &b.out_len[i], 8);
      }
      for (uint32_t i = 0; i < n; ++i) result.insert(result.end(), b.out.begin() + b.out_off[i], b.out.begin() + b.out_off[i] + b.out_len[i]);
    }
  } catch (const std::exception& e) {
    printf("%s\n", e.what());
    return -1;
  }
  if (!WriteFile(output_path, result.data(), result.size())) { printf("Error opening: %s\n", output_path.c_str()); return Help(); }
  printf("%zu bytes -> %zu bytes in %1.2f s.\n", in.size(), result.size(), ((double)clock() - start) / CLOCKS_PER_SEC);
  if (mode == 'c' || mode == 'C' || mode == 'p')
    printf("cross entropy: %1.3f\n", in.empty() ? 0.0 : 8.0 * result.size() / in.size());
  return 0;
}
