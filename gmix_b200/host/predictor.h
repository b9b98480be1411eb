// Host-side mirror of the reference's `class Predictor` (reference src/predictor.h:13-38) on top of
// the C ABI (include/gmix_b200.h). Same call contract: Predict() -> Perceive(bit) -> Learn(); Learn() is
// optional once training is over (generation). Every call is a kernel launch on the GPU that owns the
// stream; bulk work should use gmx_compress_batch / gmx_decompress_batch instead.
#ifndef GMIX_B200_HOST_PREDICTOR_H_
#define GMIX_B200_HOST_PREDICTOR_H_
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/gmix_b200.h"

namespace gmixb {

class Gpu {  // one gmx_ctx per GPU
 public:
  explicit Gpu(int device = 0) {
    if (gmx_create(device, &ctx_) != 0) throw std::runtime_error(std::string("gmx_create: ") + gmx_global_error());
  }
  ~Gpu() { gmx_destroy(ctx_); }
  Gpu(const Gpu&) = delete;
  Gpu& operator=(const Gpu&) = delete;
  gmx_ctx* ctx() const { return ctx_; }

 private:
  gmx_ctx* ctx_ = nullptr;
};

class Predictor {
 public:
  // max_stream_len: upper bound on the bytes this predictor will see (sizes its device arena).
  Predictor(Gpu& gpu, unsigned long long max_stream_len) : gpu_(gpu) {
    if (gmx_pred_new(gpu.ctx(), max_stream_len, &p_) != 0) throw std::runtime_error(gmx_last_error(gpu.ctx()));
  }
  ~Predictor() { gmx_pred_free(p_); }
  Predictor(const Predictor&) = delete;
  Predictor& operator=(const Predictor&) = delete;

  // Probability that the next bit is 1 (reference predictor.cpp:360-376).
  float Predict() {
    float p = 0.5f;
    Check(gmx_pred_predict(p_, &p));
    return p;
  }
  void Perceive(int bit) { Check(gmx_pred_perceive(p_, bit)); }   // predictor.cpp:378-381
  void Learn() { Check(gmx_pred_learn(p_)); }                       // predictor.cpp:383-387
  void Copy(const Predictor& p) { Check(gmx_pred_copy(p_, p.p_)); } // predictor.cpp:42-48 (same Gpu, same max_stream_len)
  // Only the path-visible effect of EnableAnalysis (predictions zeroed each Predict, predictor.cpp:362-365).
  void EnableAnalysis(int sample_frequency) { Check(gmx_pred_enable_analysis(p_, sample_frequency > 0)); }
  // `path`.short + `path`.long in the reference's format (predictor.cpp:389-420); like the reference, a file that
  // cannot be opened is silently ignored. Call at a byte boundary (after Learn() of a byte's last bit).
  void WriteCheckpoint(const std::string& path) {
    const void *sp, *lp;
    uint64_t sl, ll;
    Check(gmx_pred_write_checkpoint(p_, &sp, &sl, &lp, &ll));
    std::ofstream fs(path + ".short", std::ios::out | std::ios::binary);
    if (!fs.is_open()) return;
    std::ofstream fl(path + ".long", std::ios::out | std::ios::binary);
    if (!fl.is_open()) return;
    fs.write((const char*)sp, (std::streamsize)sl);
    fl.write((const char*)lp, (std::streamsize)ll);
  }
  void ReadCheckpoint(const std::string& path) {
    std::ifstream fs(path + ".short", std::ios::in | std::ios::binary);
    if (!fs.is_open()) return;
    std::ifstream fl(path + ".long", std::ios::in | std::ios::binary);
    if (!fl.is_open()) return;
    const std::vector<char> s((std::istreambuf_iterator<char>(fs)), std::istreambuf_iterator<char>());
    const std::vector<char> l((std::istreambuf_iterator<char>(fl)), std::istreambuf_iterator<char>());
    Check(gmx_pred_read_checkpoint(p_, s.data(), s.size(), l.data(), l.size()));
  }

 private:
  void Check(int rc) { if (rc != 0) throw std::runtime_error(gmx_last_error(gpu_.ctx())); }
  Gpu& gpu_;
  gmx_pred* p_ = nullptr;
};

}  // namespace gmixb
#endif
