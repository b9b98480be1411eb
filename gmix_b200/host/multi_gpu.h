// Multi-GPU host of the stream batch calls (BASELINE.json configs[4]): one gmx_ctx and one host
// thread per GPU of the box, every GPU compresses (or decompresses) its own contiguous, byte-balanced range of the
// streams with no communication on the data path; afterwards ONE collective gathers {compressed size, FNV-1a 64} per
// stream on every GPU (ncclAllGather over NVLink, padded to the largest range). The gathered table is what a
// distributed writer needs to lay the container out without moving stream bytes between GPUs; here it is also checked
// against what the host saw.
// Uses the device-pointer entry points of the C ABI + the CUDA runtime + NCCL; no kernels of its own.
#ifndef GMIX_B200_HOST_MULTI_GPU_H_
#define GMIX_B200_HOST_MULTI_GPU_H_
#include <cuda_runtime_api.h>
#include <nccl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gmix_b200.h"
#include "shard.h"

namespace gmixb {

struct StreamRecord { uint64_t len, fnv1a64; };

// The host threads of the GPUs meet here. cudaMalloc / cudaFree synchronise devices implicitly and can deadlock against a
// collective kernel that is spinning for its peers, so every allocation happens before the first rank launches the
// gather and every release after the last rank has finished it.
class HostBarrier {
 public:
  explicit HostBarrier(int n) : n_(n) {}
  void Wait() {
    std::unique_lock<std::mutex> lk(m_);
    const unsigned gen = gen_;
    if (++count_ == n_) { count_ = 0; ++gen_; cv_.notify_all(); }
    else cv_.wait(lk, [&] { return gen_ != gen; });
  }
 private:
  std::mutex m_; std::condition_variable cv_; int n_, count_ = 0; unsigned gen_ = 0;
};

class MultiGpu {
 public:
  // devices: CUDA device ordinals, one rank each. Throws nothing: ok() tells, error() explains.
  explicit MultiGpu(const std::vector<int>& devices) : dev_(devices), ctx_(devices.size(), nullptr), comm_(devices.size(), nullptr), buf_(devices.size()) {
    for (size_t r = 0; r < dev_.size(); ++r)
      if (gmx_create(dev_[r], &ctx_[r]) != 0) { err_ = std::string("gmx_create: ") + gmx_global_error(); return; }
    if (getenv("GMIXB200_VERBOSE")) fprintf(stderr, "[gmixb200] %zu contexts created, ncclCommInitAll ...\n", dev_.size());
    const ncclResult_t rc = ncclCommInitAll(comm_.data(), (int)dev_.size(), dev_.data());
    if (rc != ncclSuccess) { err_ = std::string("ncclCommInitAll: ") + ncclGetErrorString(rc); return; }
    ok_ = true;
  }
  ~MultiGpu() {
    for (ncclComm_t c : comm_) if (c) ncclCommDestroy(c);
    for (gmx_ctx* c : ctx_) if (c) gmx_destroy(c);
  }
  MultiGpu(const MultiGpu&) = delete;
  MultiGpu& operator=(const MultiGpu&) = delete;
  bool ok() const { return ok_; }
  const std::string& error() const { return err_; }
  int world() const { return (int)dev_.size(); }
  const std::vector<std::pair<uint32_t, uint32_t>>& ranges() const { return ranges_; }

  // n streams in[in_off[i] .. in_off[i+1]) -> out[out_off[i] ..] (capacities out_off[i+1] - out_off[i]), out_len[i];
  // table[i] = {size, checksum} of stream i as gathered over NCCL (identical on every GPU, checked).
  bool Run(bool compress, const uint8_t* in, const std::vector<uint64_t>& in_off, uint8_t* out, const std::vector<uint64_t>& out_off,
           std::vector<uint64_t>* out_len, std::vector<StreamRecord>* table) {
    const uint32_t n = (uint32_t)in_off.size() - 1;
    std::vector<uint64_t> work(n);   // bytes a stream costs: its uncompressed length
    for (uint32_t i = 0; i < n; ++i) work[i] = compress ? in_off[i + 1] - in_off[i] : out_off[i + 1] - out_off[i];
    ranges_ = ShardRanges(work, world());
    uint32_t width = 1;
    for (auto& r : ranges_) width = r.second - r.first > width ? r.second - r.first : width;
    out_len->assign(n, 0);
    table->assign(n, StreamRecord{0, 0});
    std::vector<std::string> errs((size_t)world());
    std::vector<std::vector<uint64_t>> gathered((size_t)world());
    std::vector<std::thread> th;
    HostBarrier bar(world());
    std::atomic<int> failed{0};
    for (int r = 0; r < world(); ++r)
      th.emplace_back([&, r] { errs[r] = Rank(r, compress, in, in_off, out, out_off, out_len->data(), width, &gathered[r], &bar, &failed); });
    for (auto& t : th) t.join();
    for (int r = 0; r < world(); ++r) if (!errs[r].empty()) { err_ = "GPU " + std::to_string(dev_[r]) + ": " + errs[r]; return false; }
    // every GPU holds the whole table: [rank][2][width]
    for (int r = 0; r < world(); ++r) if (gathered[r] != gathered[0]) { err_ = "gathered tables differ between GPUs"; return false; }
    for (int r = 0; r < world(); ++r)
      for (uint32_t i = ranges_[r].first; i < ranges_[r].second; ++i) {
        const uint64_t* slot = gathered[0].data() + (size_t)r * 2 * width;
        (*table)[i] = StreamRecord{slot[i - ranges_[r].first], slot[width + i - ranges_[r].first]};
        if ((*table)[i].len != (*out_len)[i] || (*table)[i].fnv1a64 != Fnv1a64(out + out_off[i], (*out_len)[i])) {
          err_ = "stream " + std::to_string(i) + ": gathered {size, checksum} does not match the bytes the host received";
          return false;
        }
      }
    return true;
  }

 private:
#define GMIXB_CUDA(expr) do { cudaError_t e_ = (expr); if (e_ != cudaSuccess) return std::string(#expr) + ": " + cudaGetErrorString(e_); } while (0)
#define GMIXB_NCCL(expr) do { ncclResult_t e_ = (expr); if (e_ != ncclSuccess) return std::string(#expr) + ": " + ncclGetErrorString(e_); } while (0)
  // Everything GPU r does, on its own host thread. Returns an error text or "".
  std::string Rank(int r, bool compress, const uint8_t* in, const std::vector<uint64_t>& in_off, uint8_t* out, const std::vector<uint64_t>& out_off,
                   uint64_t* out_len, uint32_t width, std::vector<uint64_t>* gathered, HostBarrier* bar, std::atomic<int>* failed) {
    // compute (allocates), then - all ranks together - the gather; a rank that fails still walks through the barriers and
    // nobody enters the collective without it
    const bool verbose = getenv("GMIXB200_VERBOSE") != nullptr;
    std::string err = Compute(r, compress, in, in_off, out_off, width);
    if (verbose) fprintf(stderr, "[gmixb200] GPU %d: streams [%u, %u) done%s%s\n", dev_[r], ranges_[r].first, ranges_[r].second, err.empty() ? "" : ": ", err.c_str());
    if (!err.empty()) failed->store(1);
    bar->Wait();
    if (!failed->load()) err = Gather(r, out, out_off, out_len, width, gathered);
    else if (err.empty()) err = "another GPU failed";
    if (verbose) fprintf(stderr, "[gmixb200] GPU %d: gather done%s%s\n", dev_[r], err.empty() ? "" : ": ", err.c_str());
    bar->Wait();
    Release(r);
    return err;
  }

  struct RankBuffers {
    cudaStream_t st = nullptr;
    uint8_t *d_in = nullptr, *d_out = nullptr;
    uint64_t *d_io = nullptr, *d_oo = nullptr, *d_send = nullptr, *d_recv = nullptr;
    uint32_t* d_status = nullptr;
  };

  std::string Compute(int r, bool compress, const uint8_t* in, const std::vector<uint64_t>& in_off, const std::vector<uint64_t>& out_off, uint32_t width) {
    RankBuffers& B = buf_[r];
    cudaStream_t& st = B.st;
    uint8_t*& d_in = B.d_in; uint8_t*& d_out = B.d_out;
    uint64_t*& d_io = B.d_io; uint64_t*& d_oo = B.d_oo; uint64_t*& d_send = B.d_send; uint64_t*& d_recv = B.d_recv;
    uint32_t*& d_status = B.d_status;
    const uint32_t lo = ranges_[r].first, hi = ranges_[r].second, m = hi - lo;
    GMIXB_CUDA(cudaSetDevice(dev_[r]));
    GMIXB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    gmx_set_cuda_stream(ctx_[r], st);
    const uint64_t in_bytes = in_off[hi] - in_off[lo], out_bytes = out_off[hi] - out_off[lo];
    GMIXB_CUDA(cudaMalloc(&d_in, in_bytes + 16));
    GMIXB_CUDA(cudaMalloc(&d_out, out_bytes + 16));
    GMIXB_CUDA(cudaMalloc(&d_io, (m + 1) * 8));
    GMIXB_CUDA(cudaMalloc(&d_oo, (m + 1) * 8));
    GMIXB_CUDA(cudaMalloc(&d_send, 2ull * width * 8));             // {len[width], sum[width]}
    GMIXB_CUDA(cudaMalloc(&d_recv, 2ull * width * 8 * world()));
    GMIXB_CUDA(cudaMalloc(&d_status, (m + 1) * 4));
    GMIXB_CUDA(cudaMemsetAsync(d_send, 0, 2ull * width * 8, st));
    std::vector<uint64_t> io(m + 1), oo(m + 1);
    uint64_t max_len = 0;
    for (uint32_t i = 0; i <= m; ++i) { io[i] = in_off[lo + i] - in_off[lo]; oo[i] = out_off[lo + i] - out_off[lo]; }
    for (uint32_t i = 0; i < m; ++i) {
      const uint64_t len = compress ? io[i + 1] - io[i] : oo[i + 1] - oo[i];
      max_len = len > max_len ? len : max_len;
    }
    if (m) {
      GMIXB_CUDA(cudaMemcpyAsync(d_in, in + in_off[lo], in_bytes, cudaMemcpyHostToDevice, st));
      GMIXB_CUDA(cudaMemcpyAsync(d_io, io.data(), (m + 1) * 8, cudaMemcpyHostToDevice, st));
      GMIXB_CUDA(cudaMemcpyAsync(d_oo, oo.data(), (m + 1) * 8, cudaMemcpyHostToDevice, st));
      const int rc = compress ? gmx_compress_batch_device(ctx_[r], d_in, d_io, m, d_out, d_oo, d_send, d_status, max_len)
                              : gmx_decompress_batch_device(ctx_[r], d_in, d_io, m, d_out, d_oo, d_send, d_status, max_len);
      if (rc != 0) return gmx_last_error(ctx_[r]);
      if (gmx_checksum_device(ctx_[r], d_out, d_oo, d_send, m, d_send + width) != 0) return gmx_last_error(ctx_[r]);
    }
    GMIXB_CUDA(cudaStreamSynchronize(st));
    return "";
  }

  std::string Gather(int r, uint8_t* out, const std::vector<uint64_t>& out_off, uint64_t* out_len, uint32_t width, std::vector<uint64_t>* gathered) {
    RankBuffers& B = buf_[r];
    cudaStream_t st = B.st;
    uint8_t* d_out = B.d_out; uint64_t* d_send = B.d_send; uint64_t* d_recv = B.d_recv; uint32_t* d_status = B.d_status;
    const uint32_t lo = ranges_[r].first, hi = ranges_[r].second, m = hi - lo;
    const uint64_t out_bytes = out_off[hi] - out_off[lo];
    GMIXB_CUDA(cudaSetDevice(dev_[r]));
    // the one collective: sizes + checksums of every stream to every GPU
    GMIXB_NCCL(ncclAllGather(d_send, d_recv, 2ull * width, ncclUint64, comm_[r], st));
    gathered->resize(2ull * width * world());
    GMIXB_CUDA(cudaMemcpyAsync(gathered->data(), d_recv, gathered->size() * 8, cudaMemcpyDeviceToHost, st));
    std::vector<uint32_t> status(m + 1, 0);
    if (m) {
      GMIXB_CUDA(cudaMemcpyAsync(out + out_off[lo], d_out, out_bytes, cudaMemcpyDeviceToHost, st));
      GMIXB_CUDA(cudaMemcpyAsync(out_len + lo, d_send, m * 8ull, cudaMemcpyDeviceToHost, st));
      GMIXB_CUDA(cudaMemcpyAsync(status.data(), d_status, m * 4ull, cudaMemcpyDeviceToHost, st));
    }
    GMIXB_CUDA(cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < m; ++i) if (status[i]) return "stream " + std::to_string(lo + i) + " failed with status " + std::to_string(status[i]);
    return "";
  }

  void Release(int r) {
    RankBuffers& B = buf_[r];
    cudaSetDevice(dev_[r]);
    gmx_set_cuda_stream(ctx_[r], nullptr);
    for (void* p : {(void*)B.d_in, (void*)B.d_out, (void*)B.d_io, (void*)B.d_oo, (void*)B.d_send, (void*)B.d_recv, (void*)B.d_status}) if (p) cudaFree(p);
    if (B.st) cudaStreamDestroy(B.st);
    B = RankBuffers();
  }
#undef GMIXB_CUDA
#undef GMIXB_NCCL

  std::vector<int> dev_;
  std::vector<gmx_ctx*> ctx_;
  std::vector<ncclComm_t> comm_;
  std::vector<RankBuffers> buf_;
  std::vector<std::pair<uint32_t, uint32_t>> ranges_;
  bool ok_ = false;
  std::string err_;
};

}  // namespace gmixb
#endif
