// TEST INFRASTRUCTURE ONLY — CLI around the CPU restatement (gmix_oracle.h), same sub-commands and
// trace format as oracle/ref_driver.cpp so the two can be compared with cmp(1).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "gmix_oracle.h"

static std::vector<uint8_t> ReadAll(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  std::vector<uint8_t> v;
  uint8_t buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) v.insert(v.end(), buf, buf + n);
  fclose(f);
  return v;
}
static void WriteAll(const char* path, const uint8_t* d, size_t n) {
  FILE* f = fopen(path, "wb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  fwrite(d, 1, n, f);
  fclose(f);
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: gmix_oracle compress|decompress|trace <in> <out> [level]\n"); return 2; }
  std::string mode = argv[1];
  std::vector<uint8_t> in = ReadAll(argv[2]);
  if (mode == "compress") {
    std::vector<uint8_t> out(in.size() * 2 + 64);
    uint64_t n = 0;
    if (gmo_compress(in.data(), in.size(), out.data(), out.size(), &n)) return 1;
    WriteAll(argv[3], out.data(), n);
    return 0;
  }
  if (mode == "decompress") {
    uint64_t len = 0;
    for (int i = 0; i < 5 && i < (int)in.size(); ++i) len = (len << 8) + in[i];
    std::vector<uint8_t> out(len + 1);
    uint64_t n = 0;
    if (gmo_decompress(in.data(), in.size(), out.data(), out.size(), &n)) return 1;
    WriteAll(argv[3], out.data(), n);
    return 0;
  }
  if (mode == "trace") {
    int level = argc > 4 ? atoi(argv[4]) : 1;
    FILE* tr = fopen(argv[3], "wb");
    gmo_predictor* p = gmo_new();
    gmo_set_analysis(p, (8 * in.size() / 1000) > 0);
    for (size_t pos = 0; pos < in.size(); ++pos) {
      char c = (char)in[pos];
      for (int j = 7; j >= 0; --j) {
        int bit = (c >> j) & 1;
        float prob = gmo_predict(p);
        uint32_t p16 = 1 + 65534 * prob;
        fwrite(&prob, 4, 1, tr);
        fwrite(&p16, 4, 1, tr);
        if (level >= 2) {
          float preds[90], l0[24], l1[8], fin; uint32_t mask[3];
          gmo_peek(p, preds, mask, l0, l1, &fin);
          fwrite(preds, 4, 90, tr); fwrite(mask, 4, 3, tr); fwrite(l0, 4, 24, tr); fwrite(l1, 4, 8, tr); fwrite(&fin, 4, 1, tr);
        }
        if (level >= 3 && j == 7) {
          float ppm[256], lstm[256];
          gmo_peek_bytes(p, ppm, lstm);
          fwrite(ppm, 4, 256, tr); fwrite(lstm, 4, 256, tr);
        }
        gmo_perceive(p, bit);
        gmo_learn(p);
      }
    }
    gmo_free(p);
    fclose(tr);
    return 0;
  }
  return 2;
}
