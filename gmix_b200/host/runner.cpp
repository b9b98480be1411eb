// gmixb200 — command-line runner over libgmix_b200.so, mirroring the reference CLI (reference
// src/runner/runner.cpp:14-104, src/runner/runner-utils.cpp:88-156):
//
//   gmixb200 -c [ckpt] input output   compress, one stream; bytes identical to `gmix -c` (strict build)
//   gmixb200 -d [ckpt] input output   decompress a stream written by either program
//   gmixb200 -g ckpt prompt output size temperature    generate, bytes identical to `gmix -g` (runner-utils.cpp:158-221)
//   gmixb200 -G ckpt prompts output size temperature [stream|exact|tensor]   one prompt per line, batched (lock-step) generation
//   gmixb200 -t [ckpt] train test     RunTraining (runner-utils.cpp:223-322): Predict/Perceive/Learn over `train`, after every
//                                     2 % the test file is scored on a copy of the predictor (analysis/training.tsv), the
//                                     coded training stream goes to data/tmp, data/trained_checkpoint.{short,long} at the end
//   gmixb200 -C bytes input output  split input into chunks of `bytes` and compress them as independent streams on ALL
//                                   visible GPUs (GMIXB200_GPUS=k limits them): one context + host thread per GPU, contiguous
//                                   byte-balanced stream ranges, one ncclAllGather of {size, FNV-1a 64} per stream
//                                   (host/multi_gpu.h). Container: "GMXB", u32 count, u64 sizes[], streams
//   gmixb200 -D input output        inverse of -C (same sharding)
//   gmixb200 -p input output        compress one stream through the Predictor facade + host coder
//                                   (Predict/Perceive/Learn per bit; slow, for checking the drop-in interface)
//   gmixb200 -T input workdir       the reference's own test-suite (runner/tester.cpp:323-378) re-run against this build
//                                   through the Predictor facade + host coder: compression, compression with restart
//                                   (predictor + coder checkpoints), with Copy restart, decompression with restart,
//                                   generation leaves the long-term memory unchanged. Prints "Tests passed." like it.
// `ckpt` is a checkpoint prefix: <ckpt>.short + <ckpt>.long, written by the reference or by this program.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <math.h>

#include <algorithm>
#include <filesystem>
#include <fstream>
#include <iomanip>
#include <iterator>
#include <string>
#include <vector>

#include "../../include/gmix_b200.h"
#include "coder.h"
#include "multi_gpu.h"
#include "predictor.h"

namespace {

int Help() {
  printf("gmixb200 (B200 build of the gmix per-bit path)\n"
         "Compress:    gmixb200 -c [checkpoint_path] input output\n"
         "Decompress:  gmixb200 -d [checkpoint_path] input output\n"
         "Generate:    gmixb200 -g checkpoint_path prompt output output_size temperature\n"
         "Batched:     gmixb200 -G checkpoint_path prompts(one per line) output output_size temperature [stream|exact|tensor]\n"
         "Train:       gmixb200 -t [checkpoint_path] training_file test_file\n"
         "Chunked:     gmixb200 -C chunk_bytes input output   /   gmixb200 -D input output\n"
         "Via facade:  gmixb200 -p input output\n"
         "Self test:   gmixb200 -T input workdir   (the reference's tester.cpp against this build)\n");
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

// Predictor::ReadCheckpoint for the batch calls: <prefix>.short/.long -> device-resident model.
gmx_model* LoadModel(gmx_ctx* ctx, const std::string& prefix, uint64_t max_new_bytes) {
  std::vector<uint8_t> sh, lo;
  if (!ReadFile(prefix + ".short", &sh) || !ReadFile(prefix + ".long", &lo)) { printf("Error opening: %s.short/.long\n", prefix.c_str()); return nullptr; }
  gmx_model* m = nullptr;
  if (gmx_model_load(ctx, sh.data(), sh.size(), lo.data(), lo.size(), max_new_bytes, 0, &m) != 0) { printf("%s\n", gmx_last_error(ctx)); return nullptr; }
  return m;
}

// n streams in[in_off[i]..in_off[i+1]) -> Batch, through the C ABI; model != null: every stream starts from it.
bool RunBatch(gmx_ctx* ctx, bool compress, const std::vector<uint8_t>& in, const std::vector<uint64_t>& in_off, Batch* b,
              const gmx_model* model = nullptr) {
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
  int rc;
  if (model) rc = compress ? gmx_compress_batch_from(ctx, model, src, in_off.data(), n, b->out.data(), b->out_off.data(), b->out_len.data(), status.data())
                           : gmx_decompress_batch_from(ctx, model, src, in_off.data(), n, b->out.data(), b->out_off.data(), b->out_len.data(), status.data());
  else rc = compress ? gmx_compress_batch(ctx, src, in_off.data(), n, b->out.data(), b->out_off.data(), b->out_len.data(), status.data())
                     : gmx_decompress_batch(ctx, src, in_off.data(), n, b->out.data(), b->out_off.data(), b->out_len.data(), status.data());
  if (rc != 0) { printf("%s\n", gmx_last_error(ctx)); return false; }
  return true;
}

// runner_utils::Compress with the analysis output the reference writes whenever 8 * n / 1000 > 0 (runner-utils.cpp:47,
// Predictor::EnableAnalysis / RunAnalysis predictor.cpp:422-504): analysis/entropy.tsv and analysis/memory.tsv in the working
// directory, one row per sample_frequency bits. The changing figures come from the GPU (gmx_compress_analysis); the other
// memory.tsv columns are constants of the model graph: LstmModel::GetMemoryUsage (lstm-model.cpp:86-99) = 7 017 924, an Indirect
// model with a 2^16-context table (indirect.cpp:72-79) = 12 + 2048 + 2 * (2^16 * 256 + 1), the final mixer after its first Learn
// (mixer.cpp:197-206) = 29 + (33 * 4 + 12) + 8.
bool CompressWithAnalysis(gmx_ctx* ctx, const std::vector<uint8_t>& in, Batch* b) {
  const uint32_t freq = (uint32_t)(8 * in.size() / 1000);
  const uint64_t cap = gmx_compress_bound(in.size());
  b->out.assign(cap + 1, 0);
  b->out_off = {0, cap};
  b->out_len.assign(1, 0);
  std::vector<gmx_analysis_row> rows(4096);
  uint32_t nr = 0;
  if (gmx_compress_analysis(ctx, in.data(), in.size(), b->out.data(), cap, &b->out_len[0], freq, rows.data(), (uint32_t)rows.size(), &nr) != 0) {
    printf("%s\n", gmx_last_error(ctx));
    return false;
  }
  static const char* kSkip[15] = {"1_2", "1_2_3", "0_2", "0_2_3", "1_2_3_4", "0_3", "0_4", "0_5", "0_2_3_4", "0_3_4", "0_6", "0_7", "0_1_3_4", "0_4_5", "0_1_2_4"};
  std::filesystem::create_directory("analysis");
  std::ofstream entropy("analysis/entropy.tsv", std::ios::out), memory("analysis/memory.tsv", std::ios::out);
  std::string header = "bits seen\tmod_ppmd(20)\tLSTM";
  for (const char* k : kSkip) header += std::string("\tIndirect(skip_") + k + ")-indirect\tIndirect(skip_" + k + ")-run_map";
  header += "\tMixer(final layer)";
  entropy << header << std::endl;
  memory << header << "\tmatch history" << std::endl;
  for (uint32_t r = 0; r < nr; ++r) {
    entropy << rows[r].bits_seen;
    memory << rows[r].bits_seen;
    for (int i = 0; i < GMX_ANALYSIS_COLUMNS; ++i) entropy << std::fixed << std::setprecision(5) << "\t" << rows[r].neg_entropy[i];
    memory << "\t" << 16 + rows[r].ppmd_used << "\t" << 7017924ull;
    for (int i = 0; i < 30; ++i) memory << "\t" << 12 + 256 * 4 * 2 + 2 * ((1ull << 16) * 256 + 1);
    memory << "\t" << 29 + (33 * 4 + 12) + 8 << "\t" << rows[r].history;
    entropy << std::endl;
    memory << std::endl;
  }
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

uint64_t HeaderLength(const std::vector<uint8_t>& in) {  // ReadHeader runner-utils.cpp:29-36
  uint64_t n = 0;
  for (size_t k = 0; k < 5 && k < in.size(); ++k) n = (n << 8) + in[k];
  return n;
}

// RunGeneration (runner-utils.cpp:158-221) for one prompt.
int Generate(int argc, char* argv[]) {
  if (argc != 7) { printf("Wrong number of arguments.\n"); return Help(); }
  std::vector<uint8_t> prompt;
  if (!ReadFile(argv[3], &prompt) || prompt.empty()) { printf("Can not open: %s\n", argv[3]); return Help(); }
  const int size = std::stoi(argv[5]);
  const float temperature = std::stof(argv[6]);
  if (size < 0) return Help();
  const clock_t start = clock();
  gmixb::Gpu gpu(0);
  gmx_model* m = LoadModel(gpu.ctx(), argv[2], prompt.size() + (uint64_t)size);
  if (!m) return -1;
  std::vector<float> rand_u((size_t)size * 8 + 1);
  gmx_reference_rand_u(rand_u.data(), (uint64_t)size * 8);
  std::vector<uint8_t> out((size_t)size + 1);
  const uint64_t off[2] = {0, prompt.size()};
  uint32_t status = 0;
  const int rc = size ? gmx_generate_batch(gpu.ctx(), m, prompt.data(), off, 1, (uint32_t)size, temperature, rand_u.data(), 0, out.data(), &status) : 0;
  if (rc != 0) printf("%s\n", gmx_last_error(gpu.ctx()));
  gmx_model_free(m);
  if (rc != 0) return -1;
  if (!WriteFile(argv[4], out.data(), (size_t)size)) { printf("Can not open: %s\n", argv[4]); return Help(); }
  printf("generation: 100%%\n%1.2f s.\n", ((double)clock() - start) / CLOCKS_PER_SEC);
  return 0;
}

// RunGeneration for MANY prompts at once (BASELINE configs[3]): one prompt per line of the prompt file (the line break is part
// of the prompt, as it is when `gmix -g` reads a one-line file), every prompt sampled with the reference's draw sequence - the
// output file is what n separate `gmix -g` processes write, concatenated. mode: "stream" (one CTA per prompt does everything),
// "exact" (lock-step, batched gate product with the reference's arithmetic: same bytes) or "tensor" (lock-step, the gate product
// on the tensor cores: opt-in, not bit-exact). gmix_b200.h: GMX_GEN_*.
int GenerateBatch(int argc, char* argv[]) {
  if (argc != 7 && argc != 8) { printf("Wrong number of arguments.\n"); return Help(); }
  std::vector<uint8_t> file;
  if (!ReadFile(argv[3], &file) || file.empty()) { printf("Can not open: %s\n", argv[3]); return Help(); }
  const int size = std::stoi(argv[5]);
  const float temperature = std::stof(argv[6]);
  const std::string mode = argc == 8 ? argv[7] : "exact";
  if (size <= 0 || (mode != "stream" && mode != "exact" && mode != "tensor")) return Help();
  std::vector<uint64_t> off(1, 0);
  uint64_t longest = 0;
  for (size_t i = 0; i < file.size(); ++i)
    if (file[i] == '\n' || i + 1 == file.size()) { longest = std::max<uint64_t>(longest, i + 1 - off.back()); off.push_back(i + 1); }
  const uint32_t n = (uint32_t)off.size() - 1;
  const clock_t start = clock();
  gmixb::Gpu gpu(0);
  gmx_model* m = LoadModel(gpu.ctx(), argv[2], longest + (uint64_t)size);
  if (!m) return -1;
  std::vector<float> rand_u((size_t)size * 8 + 1);
  gmx_reference_rand_u(rand_u.data(), (uint64_t)size * 8);
  std::vector<uint8_t> out((size_t)n * size + 1);
  std::vector<uint32_t> status(n, 0);
  gmx_set_generation_mode(gpu.ctx(), mode == "stream" ? GMX_GEN_PER_STREAM : mode == "exact" ? GMX_GEN_LOCKSTEP_EXACT : GMX_GEN_LOCKSTEP_TENSOR);
  const int rc = gmx_generate_batch(gpu.ctx(), m, file.data(), off.data(), n, (uint32_t)size, temperature, rand_u.data(), 0, out.data(), status.data());
  if (rc != 0) printf("%s\n", gmx_last_error(gpu.ctx()));
  const int ran = gmx_last_generation_mode(gpu.ctx());
  gmx_model_free(m);
  if (rc != 0) return -1;
  if (!WriteFile(argv[4], out.data(), (size_t)n * size)) { printf("Can not open: %s\n", argv[4]); return Help(); }
  printf("generation: %u prompts x %d bytes, mode %s\n%1.2f s.\n", n, size, ran == GMX_GEN_PER_STREAM ? "stream" : ran == GMX_GEN_LOCKSTEP_EXACT ? "exact" : "tensor",
         ((double)clock() - start) / CLOCKS_PER_SEC);
  return 0;
}

// RunTraining (runner-utils.cpp:223-322) at batch speed: the training file is one stream coded in parts that end where
// the reference scores the test file (after every 2 % of the training bytes: pos % percent == 0, pos / percent even);
// at each of those positions the stream's checkpoint is loaded as a model (= Predictor::Copy, :291-292) and the test
// file runs from it, learning as it goes, like p2 does. Same outputs: analysis/training.tsv (bytes, train_entropy,
// test_entropy, computed from the per-bit probabilities exactly as :284-309 does), data/tmp, data/trained_checkpoint.
double SumLog2(const std::vector<uint8_t>& data, const std::vector<float>& probs) {
  double e = 0;
  for (size_t i = 0; i < data.size(); ++i)
    for (int j = 7; j >= 0; --j) {
      const float prob = probs[i * 8 + (7 - j)];
      if ((data[i] >> j) & 1) e += log2(prob); else e += log2(1 - prob);
    }
  return e;
}
int Train(int argc, char* argv[]) {
  if (argc != 4 && argc != 5) { printf("Wrong number of arguments.\n"); return Help(); }
  const std::string ckpt = argc == 5 ? argv[2] : "", train_path = argv[argc - 2], test_path = argv[argc - 1];
  std::vector<uint8_t> train, test;
  if (!ReadFile(train_path, &train)) { printf("Can not open: %s\n", train_path.c_str()); return Help(); }
  if (!ReadFile(test_path, &test)) { printf("Can not open: %s\n", test_path.c_str()); return Help(); }
  const clock_t start = clock();
  gmixb::Gpu gpu(0);
  gmx_ctx* ctx = gpu.ctx();
  const uint64_t n = train.size(), percent = 1 + n / 100;
  std::vector<uint64_t> cuts;   // part k covers bytes [cuts[k-1], cuts[k])
  for (uint64_t pos = 1; pos < n; ++pos) if (pos % percent == 0 && (pos / percent) % 2 == 0) cuts.push_back(pos + 1);
  if (cuts.empty() || cuts.back() != n) cuts.push_back(n);
  std::filesystem::create_directory("analysis");
  std::filesystem::create_directory("data");
  std::ofstream metrics("analysis/training.tsv", std::ios::out);
  metrics << "bytes\ttrain_entropy\ttest_entropy" << std::endl;
  std::vector<uint8_t> tmp, part_out(gmx_compress_bound(n) + 16), test_out(gmx_compress_bound(test.size()) + 16), sh, lo;
  gmx_model* cur = nullptr;
  const uint64_t longest_new = std::max<uint64_t>(n, test.size()) + 16;
  if (!ckpt.empty() && !(cur = LoadModel(ctx, ckpt, longest_new))) return -1;
  gmx_coder_state coder{0, 0xffffffffu, 0};
  double train_entropy = 0;
  uint64_t begin = 0;
  const int analysis = (8 * n / 1000) > 0;   // p.EnableAnalysis(8 * input_bytes / 1000), :268
  for (size_t k = 0; k < cuts.size(); ++k) {
    const uint64_t end = cuts[k], len = end - begin;
    std::vector<uint8_t> piece(train.begin() + begin, train.begin() + end);
    std::vector<float> probs(len * 8 + 1);
    const void *sp, *lp;
    uint64_t sl, ll, out_len = 0;
    gmx_coder_state next;
    const int rc = gmx_compress_part(ctx, cur, k == 0 ? nullptr : &coder, k == 0, n, end == n, analysis, piece.data(), len, part_out.data(), part_out.size(),
                                     &out_len, &next, &sp, &sl, &lp, &ll, probs.data());
    if (rc != 0) { printf("%s\n", gmx_last_error(ctx)); return -1; }
    coder = next;
    tmp.insert(tmp.end(), part_out.begin(), part_out.begin() + out_len);
    train_entropy += SumLog2(piece, probs);
    sh.assign((const uint8_t*)sp, (const uint8_t*)sp + sl);
    lo.assign((const uint8_t*)lp, (const uint8_t*)lp + ll);
    if (cur) gmx_model_free(cur);
    cur = nullptr;
    if (gmx_model_load(ctx, sh.data(), sh.size(), lo.data(), lo.size(), longest_new, 0, &cur) != 0) { printf("%s\n", gmx_last_error(ctx)); return -1; }
    const uint64_t pos = end - 1;
    printf("\rtraining: %lld%%", (long long)(pos / percent));
    fflush(stdout);
    if (pos % percent == 0 && (pos / percent) % 2 == 0 && pos != 0) {   // score the test file on a copy (:288-311)
      std::vector<float> tprobs(test.size() * 8 + 1);
      uint64_t tl = 0;
      if (gmx_compress_part(ctx, cur, nullptr, 0, test.size(), 1, 0, test.data(), test.size(), test_out.data(), test_out.size(), &tl, nullptr,
                            nullptr, nullptr, nullptr, nullptr, tprobs.data()) != 0) { printf("%s\n", gmx_last_error(ctx)); return -1; }
      const double test_entropy = SumLog2(test, tprobs);
      metrics << std::fixed << std::setprecision(5) << pos << "\t" << -train_entropy / pos << "\t" << -test_entropy / test.size() << std::endl;
    }
    begin = end;
  }
  if (cur) gmx_model_free(cur);
  printf("\rtraining cross entropy: %.4f\n", n ? -train_entropy / n : 0.0);
  if (!WriteFile("data/tmp", tmp.data(), tmp.size()) || !WriteFile("data/trained_checkpoint.short", sh.data(), sh.size()) ||
      !WriteFile("data/trained_checkpoint.long", lo.data(), lo.size())) {
    printf("Can not write data/trained_checkpoint\n");
    return -1;
  }
  printf("%zu bytes -> %zu bytes in %1.2f s.\n", train.size(), tmp.size(), ((double)clock() - start) / CLOCKS_PER_SEC);
  return 0;
}

// ---- the reference's tests (runner/tester.cpp) against the facade ------------------------------------------------
bool FilesEqual(const std::string& a, const std::string& b) {
  std::vector<uint8_t> x, y;
  return ReadFile(a, &x) && ReadFile(b, &y) && x == y;
}
void CodeByte(gmixb::Predictor& p, gmixb::Encoder& e, uint8_t c) {   // runner-utils.cpp:50-58
  for (int j = 7; j >= 0; --j) { const int bit = (c >> j) & 1; e.Encode(bit, p.Predict()); p.Perceive(bit); p.Learn(); }
}
int DecodeByte(gmixb::Predictor& p, gmixb::Decoder& d) {             // runner-utils.cpp:75-78 + decoder.cpp:19-39
  int byte = 1;
  while (byte < 256) { const int bit = d.Decode(p.Predict()); p.Perceive(bit); p.Learn(); byte += byte + bit; }
  return byte - 256;
}
void Header(uint64_t n, std::vector<uint8_t>* out) { for (int i = 4; i >= 0; --i) out->push_back((uint8_t)(n >> (8 * i))); }

int SelfTest(const std::string& input_path, const std::string& dir) {
  std::vector<uint8_t> in;
  if (!ReadFile(input_path, &in) || in.size() < 4) { printf("Error opening: %s\n", input_path.c_str()); return -1; }
  std::filesystem::create_directories(dir);
  const uint64_t n = in.size(), half = n / 2;
  gmixb::Gpu gpu(0);
  // TestCompression (tester.cpp:323-327): the batch kernel's stream is the expected one (analysis off, as in tester.cpp)
  std::vector<uint8_t> test1;
  {
    gmixb::Predictor p(gpu, n + 1);
    Header(n, &test1);
    gmixb::Encoder e(&test1);
    for (uint64_t pos = 0; pos < n; ++pos) CodeByte(p, e, in[pos]);
    e.Flush();
  }
  printf("compression: %llu -> %zu bytes\n", (unsigned long long)n, test1.size());
  // TestCompressionWithRestart (:329-337, CompressFirstHalf/SecondHalf :24-88): checkpoint after byte n/2
  {
    std::vector<uint8_t> out;
    Header(n, &out);
    {
      gmixb::Predictor p(gpu, n + 1);
      gmixb::Encoder e(&out);
      for (uint64_t pos = 0; pos <= half; ++pos) CodeByte(p, e, in[pos]);
      p.WriteCheckpoint(dir + "/checkpoint");
      e.WriteCheckpoint(dir + "/checkpoint.coder");
    }
    gmixb::Predictor p(gpu, n + 1);
    p.ReadCheckpoint(dir + "/checkpoint");
    gmixb::Encoder e(&out);
    e.ReadCheckpoint(dir + "/checkpoint.coder");
    p.WriteCheckpoint(dir + "/checkpoint2");
    for (uint64_t pos = half + 1; pos < n; ++pos) CodeByte(p, e, in[pos]);
    e.Flush();
    if (out != test1) { printf("compression with restart: output differs\n"); return -1; }
    if (!FilesEqual(dir + "/checkpoint.long", dir + "/checkpoint2.long") || !FilesEqual(dir + "/checkpoint.short", dir + "/checkpoint2.short")) {
      printf("compression with restart: a checkpoint read back and written again differs\n");
      return -1;
    }
  }
  printf("compression with restart: ok\n");
  // TestCompressionWithCopyRestart (:339-348, :112-180)
  {
    std::vector<uint8_t> out;
    Header(n, &out);
    gmixb::Predictor p(gpu, n + 1);
    gmixb::Encoder e(&out);
    for (uint64_t pos = 0; pos <= half; ++pos) CodeByte(p, e, in[pos]);
    e.WriteCheckpoint(dir + "/checkpoint.coder");
    gmixb::Predictor p2(gpu, n + 1);
    p2.Copy(p);
    gmixb::Encoder e2(&out);
    e2.ReadCheckpoint(dir + "/checkpoint.coder");
    p2.WriteCheckpoint(dir + "/checkpoint2");
    for (uint64_t pos = half + 1; pos < n; ++pos) CodeByte(p2, e2, in[pos]);
    e2.Flush();
    if (out != test1) { printf("compression with Copy restart: output differs\n"); return -1; }
    if (!FilesEqual(dir + "/checkpoint.long", dir + "/checkpoint2.long") || !FilesEqual(dir + "/checkpoint.short", dir + "/checkpoint2.short")) {
      printf("compression with Copy restart: the copy's checkpoint differs from the original's\n");
      return -1;
    }
  }
  printf("compression with Copy restart: ok\n");
  // TestDecompressionWithRestart (:350-356, :182-321)
  {
    std::vector<uint8_t> back;
    uint64_t coded_pos;
    {
      gmixb::Predictor p(gpu, n + 1);
      gmixb::Decoder d(test1.data() + 5, test1.size() - 5);
      for (uint64_t pos = 0; pos <= half; ++pos) back.push_back((uint8_t)DecodeByte(p, d));
      p.WriteCheckpoint(dir + "/checkpoint");
      d.WriteCheckpoint(dir + "/checkpoint.coder");
      coded_pos = d.position();
    }
    gmixb::Predictor p(gpu, n + 1);
    p.ReadCheckpoint(dir + "/checkpoint");
    gmixb::Decoder d(test1.data() + 5, test1.size() - 5);
    d.ReadCheckpoint(dir + "/checkpoint.coder");
    d.Seek(coded_pos);
    for (uint64_t pos = half + 1; pos < n; ++pos) back.push_back((uint8_t)DecodeByte(p, d));
    if (back != in) { printf("decompression with restart: output differs from the original\n"); return -1; }
  }
  printf("decompression with restart: ok\n");
  // TestGeneration (:358-366): Predict/Perceive without Learn leaves `.long` unchanged and changes `.short`
  {
    gmixb::Predictor p(gpu, n + 64);
    for (uint64_t pos = 0; pos < n; ++pos)
      for (int j = 7; j >= 0; --j) { p.Predict(); p.Perceive((in[pos] >> j) & 1); p.Learn(); }
    p.WriteCheckpoint(dir + "/checkpoint");
    unsigned x = 12345;
    for (int i = 0; i < 8 * 32; ++i) { const float pr = p.Predict(); x = x * 1103515245u + 12345u; p.Perceive(((x >> 16) & 0xffff) < pr * 65536.0f); }
    p.WriteCheckpoint(dir + "/checkpoint2");
    if (!FilesEqual(dir + "/checkpoint.long", dir + "/checkpoint2.long")) { printf("generation changed the long-term memory\n"); return -1; }
    if (FilesEqual(dir + "/checkpoint.short", dir + "/checkpoint2.short")) { printf("generation did not change the short-term memory\n"); return -1; }
  }
  printf("generation: ok\nTests passed.\n");
  return 0;
}

}  // namespace

int main(int argc, char* argv[]) {
  if (argc < 4 || strlen(argv[1]) != 2 || argv[1][0] != '-') return Help();
  const char mode = argv[1][1];
  try {
    if (mode == 'T') { if (argc != 4) return Help(); return SelfTest(argv[2], argv[3]); }
    if (mode == 'g') return Generate(argc, argv);
    if (mode == 'G') return GenerateBatch(argc, argv);
    if (mode == 't') return Train(argc, argv);
  } catch (const std::exception& e) {
    printf("%s\n", e.what());
    return -1;
  }
  const bool chunked_c = mode == 'C';
  const bool with_ckpt = (mode == 'c' || mode == 'd') && argc == 5;
  if ((chunked_c && argc != 5) || (!chunked_c && !with_ckpt && argc != 4) || !strchr("cdCDp", mode)) return Help();
  const std::string checkpoint_path = with_ckpt ? argv[2] : "";
  const std::string input_path = argv[chunked_c || with_ckpt ? 3 : 2], output_path = argv[chunked_c || with_ckpt ? 4 : 3];
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
      const uint32_t n = (uint32_t)in_off.size() - 1;
      if (mode == 'C' || mode == 'D') {   // all visible GPUs
        int ng = 0;
        if (cudaGetDeviceCount(&ng) != cudaSuccess || ng < 1) { printf("no CUDA device\n"); return -1; }
        if (const char* e = getenv("GMIXB200_GPUS")) { const int k = atoi(e); if (k >= 1 && k < ng) ng = k; }
        std::vector<int> devs;
        for (int d = 0; d < ng; ++d) devs.push_back(d);
        gmixb::MultiGpu mg(devs);
        if (!mg.ok()) { printf("%s\n", mg.error().c_str()); return -1; }
        b.out_off.assign(n + 1, 0);
        for (uint32_t i = 0; i < n; ++i) {
          uint64_t cap;
          if (mode == 'C') cap = gmx_compress_bound(in_off[i + 1] - in_off[i]);
          else { cap = 0; for (uint64_t k = 0; k < 5 && in_off[i] + k < in_off[i + 1]; ++k) cap = (cap << 8) + (*src)[in_off[i] + k]; cap += 8; }
          b.out_off[i + 1] = b.out_off[i] + cap;
        }
        b.out.assign(b.out_off[n] + 1, 0);
        std::vector<gmixb::StreamRecord> table;
        static const uint8_t kNone = 0;
        if (!mg.Run(mode == 'C', src->empty() ? &kNone : src->data(), in_off, b.out.data(), b.out_off, &b.out_len, &table)) { printf("%s\n", mg.error().c_str()); return -1; }
        printf("%u streams on %d GPU(s):", n, mg.world());
        for (int r = 0; r < mg.world(); ++r) printf(" [%u, %u)", mg.ranges()[r].first, mg.ranges()[r].second);
        printf("; {size, checksum} of every stream gathered on every GPU (ncclAllGather) and verified\n");
      } else {
      gmx_model* model = nullptr;
      if (with_ckpt && !(model = LoadModel(gpu.ctx(), checkpoint_path, mode == 'c' ? in.size() : HeaderLength(in)))) return -1;
      // `gmix -c` from scratch writes analysis/*.tsv whenever the input has at least 125 bytes; GMIXB200_ANALYSIS=0 skips it
      // (the analysis step runs in the phase-serial kernel configuration, a single stream is ~1.7x faster without)
      const char* an_env = getenv("GMIXB200_ANALYSIS");
      const bool analysis = mode == 'c' && !model && 8 * in.size() / 1000 > 0 && !(an_env && an_env[0] == '0');
      const bool ok = analysis ? CompressWithAnalysis(gpu.ctx(), in, &b) : RunBatch(gpu.ctx(), mode == 'c' || mode == 'C', *src, in_off, &b, model);
      if (model) gmx_model_free(model);
      if (!ok) return -1;
      }
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
