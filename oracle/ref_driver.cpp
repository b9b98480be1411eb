// TEST INFRASTRUCTURE ONLY (see oracle/Makefile). Our own driver, linked against the UNMODIFIED
// reference objects compiled from /root/reference/src, used to (a) produce golden vectors and
// per-bit traces that pin oracle/gmix_oracle.cpp, (b) serve as the "reference" CPU baseline.
//
// It drives the reference exactly the way runner_utils::Compress does
// (reference src/runner/runner-utils.cpp:43-67): EnableAnalysis(8*n/1000), then per bit
// Predict -> Encode -> Perceive -> Learn, MSB first, then Flush; 5-byte big-endian header first.
//
//   ref_driver compress   <in> <out>
//   ref_driver decompress <in> <out>
//   ref_driver trace      <in> <out.trace> <level>   level 1: {f32 prob,u32 p16} per bit
//                                                    level 2: + predictions[90], active mask, mixer outs
//                                                    level 3: + per byte ppm_predictions[256], lstm probs[256]
//   ref_driver train      <in> <ckpt_prefix>         Predict/Perceive/Learn over <in>, WriteCheckpoint
//   ref_driver tables     <out>                      dumps Nonstationary/RunMap tables (512+512 bytes)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <memory>
#include <string>
#include <unistd.h>
#include <unordered_map>
#include <valarray>
#include <vector>
#include <array>
#include <numeric>
#include <algorithm>
#include <math.h>

#define private public
#define protected public
#include "predictor.h"
#include "coder/decoder.h"
#include "coder/encoder.h"
#include "models/lstm-model.h"
#include "runner/runner-utils.h"
#undef private
#undef protected

namespace fs = std::filesystem;

static std::string Abs(const char* p) { return fs::absolute(p).string(); }

// The reference writes analysis/*.tsv relative to the cwd; keep that out of the repo.
static void EnterScratch() {
  std::string d = "/tmp/gmix_ref_scratch_" + std::to_string(getpid());
  fs::create_directories(d);
  if (chdir(d.c_str()) != 0) { perror("chdir"); exit(2); }
}

static std::vector<unsigned char> ReadAll(const std::string& path) {
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
  return std::vector<unsigned char>((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
}

static int Trace(const std::string& in, const std::string& out, int level) {
  std::vector<unsigned char> data = ReadAll(in);
  std::ofstream tr(out, std::ios::binary);
  std::ofstream sink("/dev/null", std::ios::binary);
  Predictor p;
  Encoder e(&sink);
  p.EnableAnalysis(8 * data.size() / 1000);
  LstmModel* lstm = nullptr;
  for (auto& m : p.models_) if (auto* l = dynamic_cast<LstmModel*>(m.get())) lstm = l;
  auto& stm = p.short_term_memory_;
  for (size_t pos = 0; pos < data.size(); ++pos) {
    char c = data[pos];
    for (int j = 7; j >= 0; --j) {
      int bit = (c >> j) & 1;
      float prob = p.Predict();
      unsigned int p16 = 1 + 65534 * prob;  // Encoder::Discretize, encoder.cpp:8
      tr.write((char*)&prob, 4);
      tr.write((char*)&p16, 4);
      if (level >= 2) {
        unsigned int mask[3] = {0, 0, 0};
        for (int i : stm.active_models) mask[i >> 5] |= 1u << (i & 31);
        tr.write((char*)&stm.predictions[0], 4 * 90);
        tr.write((char*)mask, 12);
        tr.write((char*)&stm.mixer_layer0_outputs[0], 4 * 24);
        tr.write((char*)&stm.mixer_layer1_outputs[0], 4 * 8);
        tr.write((char*)&stm.final_mixer_output, 4);
      }
      if (level >= 3 && j == 7) {
        tr.write((char*)&stm.ppm_predictions[0], 4 * 256);
        tr.write((char*)&lstm->probs_[0], 4 * 256);
      }
      e.Encode(bit, prob);
      p.Perceive(bit);
      p.Learn();
    }
  }
  return 0;
}

static int Train(const std::string& in, const std::string& ckpt) {
  std::vector<unsigned char> data = ReadAll(in);
  Predictor p;
  for (size_t pos = 0; pos < data.size(); ++pos) {
    char c = data[pos];
    for (int j = 7; j >= 0; --j) {
      p.Predict();
      p.Perceive((c >> j) & 1);
      p.Learn();
    }
  }
  p.WriteCheckpoint(ckpt);
  return 0;
}

static int Tables(const std::string& out) {
  Nonstationary ns;
  RunMap rm;
  std::ofstream f(out, std::ios::binary);
  for (int s = 0; s < 256; ++s)
    for (int b = 0; b < 2; ++b) f.put((char)ns.Next(s, b));
  for (int s = 0; s < 256; ++s)
    for (int b = 0; b < 2; ++b) f.put((char)rm.Next(s, b));
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: see header of oracle/ref_driver.cpp\n"); return 2; }
  std::string mode = argv[1];
  srand(0xDEADBEEF);  // runner.cpp:37
  if (mode == "tables") { return Tables(Abs(argv[2])); }
  if (argc < 4) return 2;
  std::string a = Abs(argv[2]), b = Abs(argv[3]);
  EnterScratch();
  unsigned long long ib = 0, ob = 0;
  int rc = 0;
  if (mode == "compress") rc = runner_utils::RunCompression("", a, b, &ib, &ob) ? 0 : 1;
  else if (mode == "decompress") rc = runner_utils::RunDecompression("", a, b, &ib, &ob) ? 0 : 1;
  else if (mode == "trace") rc = Trace(a, b, argc > 4 ? atoi(argv[4]) : 1);
  else if (mode == "train") rc = Train(a, b);
  else rc = 2;
  fs::remove_all(fs::current_path());
  return rc;
}
