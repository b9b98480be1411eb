// libgmix_b200.so — host side of the C ABI declared in include/gmix_b200.h.
// Owns device memory (stream arenas, libm tables), launches the sm_100a stream kernels
// (stream_kernel.cuh) and the device-math self test. No CPU fallback exists.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gmix_b200.h"
#include "checkpoint.h"
#include "kernels.h"
#include "layout.h"

namespace {

std::string g_global_error;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
};

}  // namespace

struct gmx_ctx {
  int device = 0;
  int sm_count = 0;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string error;
  // arenas: `layout`/`d_arenas` is the normal class every stream is first run in; `roomy` is the
  // worst-case-sized class, allocated on demand for the streams that overflowed a normal arena
  uint64_t cfg_max_len = 0;
  uint32_t cfg_max_resident = 0;
  const gmx_model* cfg_model = nullptr;   // arenas currently use this model's layout (streams start from its checkpoint)
  uint32_t cfg_factor = 1;                // arenas per resident CTA slot (ConfigureForModel)
  uint64_t cfg_ov_learn = 0, cfg_ov_new = 0;   // != 0: ... in overlay mode, sized for this many learned / new bytes per stream
  std::vector<uint8_t> ck_short, ck_long; // last checkpoint written through this ctx (gmx_train_checkpoint, gmx_pred_write_checkpoint)
  uint32_t n_arenas = 0;
  gmx::ArenaLayout layout;
  uint8_t* d_arenas = nullptr;
  gmx::ArenaLayout* d_layout = nullptr;
  uint32_t n_roomy = 0;
  gmx::ArenaLayout roomy_layout;
  uint8_t* d_roomy = nullptr;
  gmx::ArenaLayout* d_roomy_layout = nullptr;
  uint64_t retried_streams = 0;
  float* d_lstm_init = nullptr;
  float* d_adam = nullptr;
  float* d_decay = nullptr;
  uint32_t decay_len = 0;
  uint32_t* d_queue = nullptr;
  std::vector<float> h_decay;
  // staging buffers of the host-pointer entry points
  DevBuf b_in, b_out, b_in_off, b_out_off, b_out_len, b_status, b_trace, b_ptrace, b_prof, b_ids, b_usage, b_rand, b_final, b_coder;
  // lock-step batched generation (gate_gemm.cuh): parked stream states, operand planes / byte in front / pre-activations of the
  // batched gate product, the model's tiled weight planes
  DevBuf b_park, b_gx, b_gsym, b_gg, b_wt, b_an;
  enum { kGroups = 16 };
  cudaStream_t gen_stream[kGroups] = {};   // one per group of lock-step streams
  cudaEvent_t gen_done[kGroups] = {}, gen_fork = nullptr;
  int gen_mode = 0;        // GMX_GEN_* requested by gmx_set_generation_mode
  int last_gen_mode = 0;   // what the last generation call actually ran
  bool profile = false;
  uint32_t prof_streams = 0;
  uint32_t usage_streams = 0;
  uint32_t last_grid = 0;
  int kcfg = gmx::kThroughputConfig;   // kernel configuration (kernels.h) the arenas are currently sized for
  int kcfg_user = -1;      // -1: chosen per call from the batch shape (AutoConfig), else pinned by gmx_set_kernel_config
  uint64_t launches = 0;
  double last_ms = 0;
};

// A loaded checkpoint (Predictor::ReadCheckpoint predictor.cpp:406-420): one parked stream on the device
// (arena image + StreamSmem image) that batch calls clone into every stream arena. Read-only after load.
struct gmx_model {
  gmx_ctx* ctx = nullptr;
  gmx::ArenaLayout layout;
  gmx::Preload pre;
  uint64_t max_new_bytes = 0;
  uint32_t l_epoch = 0;      // LSTM epoch slot of the parked stream: bytes learned since the model's last BPTT pass (lstm.cpp:57-79)
  uint8_t* d_arena = nullptr;
  uint32_t* d_state = nullptr;
  gmx::ArenaLayout* d_layout = nullptr;   // the model's layout on the device (overlay-mode streams read the model's arena through it)
};

// Per-call options of RunDevice beyond the stream slices.
struct RunOpts {
  const gmx_model* model = nullptr;
  int analysis = -1;
  uint32_t* d_final_state = nullptr;
  uint32_t gen_bytes = 0; float temperature = 1.0f; const float* d_rand_u = nullptr; uint64_t rand_stride = 0;
  uint64_t overlay_learn = 0;   // generation: longest prompt (the bytes a stream learns)
  uint64_t* d_bit_trace = nullptr; float* d_pred_trace = nullptr;
  // one stream coded in parts (gmx_compress_part / gmx_decompress_part)
  uint32_t part = 0, part_header = 0, part_last = 0; uint64_t part_total = 0;
  const uint32_t* d_coder_in = nullptr; uint32_t* d_coder_out = nullptr;
  // analysis output of a single stream (gmx_compress_analysis)
  double* d_an_entropy = nullptr; gmx::AnalysisRow* d_an_rows = nullptr; uint32_t an_freq = 0, an_max_rows = 0;
};

// One stream stepped bit by bit (the Predictor facade). Owns a worst-case-sized arena.
struct gmx_pred {
  gmx_ctx* ctx = nullptr;
  gmx::ArenaLayout layout;
  uint8_t* d_arena = nullptr;
  gmx::ArenaLayout* d_layout = nullptr;
  uint32_t* d_state = nullptr;
  float* d_prob = nullptr;       // {prob, status}
  int pending_bit = -1;
  int analysis = 0;
  uint64_t bits = 0, max_bits = 0;
};

namespace {

int Fail(gmx_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (c) c->error = buf; else g_global_error = buf;
  return code;
}

#define GMX_CUDA(c, expr)                                                                          \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return Fail((c), GMX_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
  } while (0)

int Reserve(gmx_ctx* c, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap && b.p) return 0;
  if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
  size_t want = bytes < 256 ? 256 : bytes;
  GMX_CUDA(c, cudaMalloc(&b.p, want));
  b.cap = want;
  return 0;
}

// decay table: one entry per bit step (mixer.cpp:111), shared by all streams of the ctx
int EnsureDecay(gmx_ctx* c, uint64_t max_stream_len) {
  const uint64_t need = max_stream_len * 8 + 16;
  if (need <= c->decay_len) return 0;
  gmx::FillDecayTable(c->h_decay, need);
  if (c->d_decay) cudaFree(c->d_decay);
  c->d_decay = nullptr;
  GMX_CUDA(c, cudaMalloc(&c->d_decay, c->h_decay.size() * 4));
  GMX_CUDA(c, cudaMemcpy(c->d_decay, c->h_decay.data(), c->h_decay.size() * 4, cudaMemcpyHostToDevice));
  c->decay_len = (uint32_t)c->h_decay.size();
  return 0;
}

void FreeArenas(gmx_ctx* c) {
  if (c->d_arenas) cudaFree(c->d_arenas);
  if (c->d_roomy) cudaFree(c->d_roomy);
  c->d_arenas = nullptr;
  c->d_roomy = nullptr;
  c->n_arenas = 0;
  c->n_roomy = 0;
  c->cfg_max_len = 0;
  c->cfg_model = nullptr;
  c->cfg_ov_learn = c->cfg_ov_new = 0;
  c->cfg_factor = 1;
}

bool Retryable(uint32_t st) {
  return st == gmx::GMX_ERR_PPMD_ARENA || st == gmx::GMX_ERR_MIXER_POOL || st == gmx::GMX_ERR_SPARSE_FULL;
}

int Launch(gmx_ctx* c, int mode, const gmx::StreamParams& P, uint32_t grid) {
  GMX_CUDA(c, cudaMemsetAsync(c->d_queue, 0, sizeof(uint32_t), c->stream));
  GMX_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  GMX_CUDA(c, mode == gmx::MODE_COMPRESS ? (P.prof ? gmx::LaunchCompressProf(c->kcfg, P, grid, c->stream) : gmx::LaunchCompress(c->kcfg, P, grid, c->stream))
              : mode == gmx::MODE_DECOMPRESS ? gmx::LaunchDecompress(c->kcfg, P, grid, c->stream) : gmx::LaunchGenerate(c->kcfg, P, grid, c->stream));
  GMX_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  GMX_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms = 0;
  GMX_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->last_ms += ms;
  c->launches += 1;
  return 0;
}

// Streams whose normal arena overflowed (incompressible data touches far more table slots and gate
// contexts than text) are re-run from scratch in worst-case-sized arenas.
int RetryInRoomyArenas(gmx_ctx* c, int mode, gmx::StreamParams P, uint32_t n, uint32_t* d_status) {
  std::vector<uint32_t> st(n), ids;
  GMX_CUDA(c, cudaMemcpyAsync(st.data(), d_status, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  GMX_CUDA(c, cudaStreamSynchronize(c->stream));
  for (uint32_t i = 0; i < n; ++i) if (Retryable(st[i])) ids.push_back(i);
  if (ids.empty()) return 0;
  if (!c->d_roomy) {
    if (c->cfg_model) return Fail(c, GMX_E_STREAM, "%zu streams overflowed the arena sized by gmx_model_load (reload the model with a larger max_new_bytes)", ids.size());
    c->roomy_layout = gmx::MakeLayout(c->cfg_max_len, true);
    size_t free_b = 0, total_b = 0;
    GMX_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
    const uint64_t usable = free_b > (1ull << 30) ? free_b - (1ull << 30) : 0;
    uint64_t want = ids.size() < c->n_arenas ? ids.size() : c->n_arenas;
    if (usable / c->roomy_layout.total < want) want = usable / c->roomy_layout.total;
    if (want == 0) return Fail(c, GMX_E_NOMEM, "%zu streams overflowed their arena and no worst-case arena (%llu MiB) fits",
                               ids.size(), (unsigned long long)(c->roomy_layout.total >> 20));
    GMX_CUDA(c, cudaMalloc(&c->d_roomy, want * c->roomy_layout.total));
    if (!c->d_roomy_layout) GMX_CUDA(c, cudaMalloc(&c->d_roomy_layout, sizeof(gmx::ArenaLayout)));
    GMX_CUDA(c, cudaMemcpy(c->d_roomy_layout, &c->roomy_layout, sizeof(gmx::ArenaLayout), cudaMemcpyHostToDevice));
    c->n_roomy = (uint32_t)want;
  }
  int rc = Reserve(c, c->b_ids, ids.size() * 4);
  if (rc) return rc;
  GMX_CUDA(c, cudaMemcpyAsync(c->b_ids.p, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, c->stream));
  P.ids = (const uint32_t*)c->b_ids.p;
  P.n_streams = (uint32_t)ids.size();
  P.arenas = c->d_roomy; P.arena_stride = c->roomy_layout.total; P.layout = c->d_roomy_layout;
  if (ids[0] != 0) { P.bit_trace = nullptr; P.pred_trace = nullptr; }  // traces belong to stream 0: re-traced only when it is re-run
  c->retried_streams += ids.size();
  return Launch(c, mode, P, P.n_streams < c->n_roomy ? P.n_streams : c->n_roomy);
}

// Arenas sized by a loaded model's layout (every stream is a clone of the model's parked stream), or - ov_new != 0 - overlay
// arenas on top of the model's arena (MakeOverlayLayout: the model's tables are shared, a stream keeps its changes only).
// factor: arenas per resident CTA slot (lock-step generation keeps more streams in flight than CTAs fit the SMs).
int ConfigureForModel(gmx_ctx* c, const gmx_model* m, uint64_t ov_learn = 0, uint64_t ov_new = 0, uint32_t factor = 1) {
  if (c->cfg_model == m && c->n_arenas && c->cfg_factor == factor &&
      ((ov_new == 0 && c->cfg_ov_new == 0) || (ov_new != 0 && c->cfg_ov_learn >= ov_learn && c->cfg_ov_new >= ov_new))) return 0;
  const uint32_t max_resident = c->cfg_max_resident;
  FreeArenas(c);
  c->cfg_factor = factor;
  c->layout = ov_new ? gmx::MakeOverlayLayout(m->layout, m->pre, ov_learn, ov_new) : m->layout;
  c->cfg_ov_learn = ov_learn; c->cfg_ov_new = ov_new;
  int rc = EnsureDecay(c, m->pre.steps / 8 + m->max_new_bytes + 2);
  if (rc) return rc;
  int per_sm = 0;
  GMX_CUDA(c, gmx::OccupancyCompress(c->kcfg, &per_sm));
  if (per_sm > gmx::KernelConfig(c->kcfg).minb) per_sm = gmx::KernelConfig(c->kcfg).minb;   // a configuration is tuned for exactly this residency
  if (per_sm < 1) per_sm = 1;
  uint64_t want = (uint64_t)per_sm * c->sm_count;
  if (max_resident && max_resident < want) want = max_resident;
  want *= factor;
  size_t free_b = 0, total_b = 0;
  GMX_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
  const uint64_t usable = free_b > (2ull << 30) ? free_b - (2ull << 30) : 0;
  const uint64_t fit = usable / c->layout.total;
  if (fit == 0) return Fail(c, GMX_E_NOMEM, "one stream arena of this model needs %llu MiB but only %llu MiB are free",
                            (unsigned long long)(c->layout.total >> 20), (unsigned long long)(free_b >> 20));
  if (fit < want) want = fit;
  GMX_CUDA(c, cudaMalloc(&c->d_arenas, want * c->layout.total));
  GMX_CUDA(c, cudaMemcpy(c->d_layout, &c->layout, sizeof(c->layout), cudaMemcpyHostToDevice));
  c->n_arenas = (uint32_t)want;
  c->cfg_model = m;
  return 0;
}

int RunDevice(gmx_ctx* c, int mode, const uint8_t* d_in, const uint64_t* d_in_off, uint32_t n, uint8_t* d_out,
              const uint64_t* d_out_off, uint64_t* d_out_len, uint32_t* d_status, uint64_t max_len, const RunOpts& o) {
  if (!c) return GMX_E_ARG;
  if (n == 0) return 0;
  if (!d_in || !d_in_off || !d_out || !d_out_len || !d_status || (mode != gmx::MODE_GENERATE && !d_out_off)) return Fail(c, GMX_E_ARG, "null pointer argument");
  GMX_CUDA(c, cudaSetDevice(c->device));
  {
    // Few streams (at most one per SM): the latency configuration (one CTA per SM, pipelined roles, gate weights resident in
    // shared memory) runs each stream ~1.8x faster; a full wave of streams is served best by the phase-serial configuration
    // at 8 CTAs per SM (measurements: profiles/r02_*). A pinned configuration (gmx_set_kernel_config) is left alone.
    const int want = c->kcfg_user >= 0 ? c->kcfg_user : (n <= (uint32_t)c->sm_count ? gmx::kLatencyConfig : gmx::kThroughputConfig);
    if (want != c->kcfg) { c->kcfg = want; FreeArenas(c); }
  }
  if (o.model) {
    if (o.model->ctx != c) return Fail(c, GMX_E_ARG, "model belongs to another context");
    if (max_len > o.model->max_new_bytes) return Fail(c, GMX_E_ARG, "stream of %llu bytes exceeds the max_new_bytes (%llu) the model was loaded with",
                                                       (unsigned long long)max_len, (unsigned long long)o.model->max_new_bytes);
    // generation never writes a checkpoint of its streams: they run as overlays of the shared model
    const bool overlay = mode == gmx::MODE_GENERATE && !o.d_final_state;
    int rc = overlay ? ConfigureForModel(c, o.model, o.overlay_learn ? o.overlay_learn : max_len, max_len) : ConfigureForModel(c, o.model);
    if (rc) return rc;
  } else if (max_len > c->cfg_max_len || c->n_arenas == 0 || c->cfg_model) {
    int rc = gmx_configure(c, max_len > c->cfg_max_len || c->cfg_model ? max_len : c->cfg_max_len, c->cfg_max_resident);
    if (rc) return rc;
  }
  gmx::StreamParams P;
  memset(&P, 0, sizeof(P));
  P.in = d_in; P.in_off = d_in_off; P.out = d_out; P.out_off = d_out_off; P.out_len = d_out_len; P.status = d_status;
  P.n_streams = n; P.queue = c->d_queue;
  P.arenas = c->d_arenas; P.arena_stride = c->layout.total; P.layout = c->d_layout;
  P.lstm_init = c->d_lstm_init; P.decay = c->d_decay; P.decay_len = c->decay_len; P.adam = c->d_adam;
  P.bit_trace = o.d_bit_trace; P.pred_trace = o.d_pred_trace;
  P.analysis = o.analysis; P.final_state = o.d_final_state;
  P.gen_bytes = o.gen_bytes; P.temperature = o.temperature; P.rand_u = o.d_rand_u; P.rand_stride = o.rand_stride;
  P.part = o.part; P.part_header = o.part_header; P.part_last = o.part_last; P.part_total = o.part_total; P.coder_in = o.d_coder_in; P.coder_out = o.d_coder_out;
  P.an_entropy = o.d_an_entropy; P.an_rows = o.d_an_rows; P.an_freq = o.an_freq; P.an_max_rows = o.an_max_rows;
  if (o.model) { P.tmpl_arena = o.model->d_arena; P.tmpl_state = o.model->d_state; P.tmpl_layout = c->layout.ov ? o.model->d_layout : nullptr; }
  const uint32_t grid = n < c->n_arenas ? n : c->n_arenas;
  {
    int rc = Reserve(c, c->b_usage, (size_t)n * 32);
    if (rc) return rc;
    P.usage = (uint32_t*)c->b_usage.p;
    c->usage_streams = n;
  }
  if (c->profile && mode == gmx::MODE_COMPRESS) {
    int rc = Reserve(c, c->b_prof, (size_t)n * gmx::GMX_PROF_SLOTS * 8);
    if (rc) return rc;
    GMX_CUDA(c, cudaMemsetAsync(c->b_prof.p, 0, (size_t)n * gmx::GMX_PROF_SLOTS * 8, c->stream));
    P.prof = (unsigned long long*)c->b_prof.p;
    c->prof_streams = n;
  }
  c->last_ms = 0;
  c->last_grid = grid;
  int rc = Launch(c, mode, P, grid);
  if (rc) return rc;
  if (o.d_final_state) return 0;   // the parked state belongs to the arena the stream ran in: no re-run elsewhere
  return RetryInRoomyArenas(c, mode, P, n, d_status);
}

int RunHost(gmx_ctx* c, int mode, const uint8_t* in, const uint64_t* in_off, uint32_t n, uint8_t* out,
            const uint64_t* out_off, uint64_t* out_len, uint32_t* status, uint64_t* h_bit_trace, float* h_pred_trace,
            RunOpts o = RunOpts()) {
  if (!c) return GMX_E_ARG;
  if (n == 0) return 0;
  if (!in || !in_off || !out || !out_off || !out_len || !status) return Fail(c, GMX_E_ARG, "null pointer argument");
  GMX_CUDA(c, cudaSetDevice(c->device));
  const uint64_t in_total = in_off[n] - in_off[0], out_total = out_off[n] - out_off[0];
  uint64_t max_len = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (in_off[i + 1] < in_off[i] || out_off[i + 1] < out_off[i]) return Fail(c, GMX_E_ARG, "offsets must be non-decreasing");
    uint64_t len;
    if (mode == gmx::MODE_COMPRESS) len = in_off[i + 1] - in_off[i];
    else {  // uncompressed length from the 5-byte big-endian header (runner-utils.cpp:29-36)
      len = 0;
      const uint64_t avail = in_off[i + 1] - in_off[i];
      for (uint64_t k = 0; k < 5 && k < avail; ++k) len = (len << 8) + in[in_off[i] + k];
      const uint64_t cap = out_off[i + 1] - out_off[i];
      if (len > cap) len = cap;  // the kernel reports GMX_S_OUTPUT_CAP for this stream
    }
    if (len > max_len) max_len = len;
  }
  // rebase offsets to 0 for the device copies
  std::vector<uint64_t> io(n + 1), oo(n + 1);
  for (uint32_t i = 0; i <= n; ++i) { io[i] = in_off[i] - in_off[0]; oo[i] = out_off[i] - out_off[0]; }
  int rc;
  if ((rc = Reserve(c, c->b_in, in_total + 16))) return rc;
  if ((rc = Reserve(c, c->b_out, out_total + 16))) return rc;
  if ((rc = Reserve(c, c->b_in_off, (n + 1) * 8))) return rc;
  if ((rc = Reserve(c, c->b_out_off, (n + 1) * 8))) return rc;
  if ((rc = Reserve(c, c->b_out_len, (size_t)n * 8))) return rc;
  if ((rc = Reserve(c, c->b_status, (size_t)n * 4))) return rc;
  uint64_t* d_bt = nullptr;
  float* d_pt = nullptr;
  const uint64_t nbits0 = (in_off[1] - in_off[0]) * 8;
  if (h_bit_trace) { if ((rc = Reserve(c, c->b_trace, nbits0 * 8 + 8))) return rc; d_bt = (uint64_t*)c->b_trace.p; }
  if (h_pred_trace) { if ((rc = Reserve(c, c->b_ptrace, nbits0 * 126 * 4 + 8))) return rc; d_pt = (float*)c->b_ptrace.p; }
  GMX_CUDA(c, cudaMemcpyAsync(c->b_in.p, in + in_off[0], in_total, cudaMemcpyHostToDevice, c->stream));
  GMX_CUDA(c, cudaMemcpyAsync(c->b_in_off.p, io.data(), (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  GMX_CUDA(c, cudaMemcpyAsync(c->b_out_off.p, oo.data(), (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  GMX_CUDA(c, cudaMemsetAsync(c->b_out_len.p, 0, (size_t)n * 8, c->stream));
  GMX_CUDA(c, cudaMemsetAsync(c->b_status.p, 0xff, (size_t)n * 4, c->stream));
  o.d_bit_trace = d_bt; o.d_pred_trace = d_pt;
  rc = RunDevice(c, mode, (const uint8_t*)c->b_in.p, (const uint64_t*)c->b_in_off.p, n, (uint8_t*)c->b_out.p,
                 (const uint64_t*)c->b_out_off.p, (uint64_t*)c->b_out_len.p, (uint32_t*)c->b_status.p, max_len, o);
  if (rc == GMX_E_STREAM) {   // per-stream codes are part of the contract of GMX_E_STREAM (gmix_b200.h)
    const std::string keep = c->error;
    cudaMemcpyAsync(out_len, c->b_out_len.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream);
    cudaMemcpyAsync(status, c->b_status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream);
    cudaStreamSynchronize(c->stream);
    c->error = keep;
  }
  if (rc) return rc;
  GMX_CUDA(c, cudaMemcpyAsync(out + out_off[0], c->b_out.p, out_total, cudaMemcpyDeviceToHost, c->stream));
  GMX_CUDA(c, cudaMemcpyAsync(out_len, c->b_out_len.p, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  GMX_CUDA(c, cudaMemcpyAsync(status, c->b_status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  if (h_bit_trace) GMX_CUDA(c, cudaMemcpyAsync(h_bit_trace, d_bt, nbits0 * 8, cudaMemcpyDeviceToHost, c->stream));
  if (h_pred_trace) GMX_CUDA(c, cudaMemcpyAsync(h_pred_trace, d_pt, nbits0 * 126 * 4, cudaMemcpyDeviceToHost, c->stream));
  GMX_CUDA(c, cudaStreamSynchronize(c->stream));
  for (uint32_t i = 0; i < n; ++i)
    if (status[i] != 0) return Fail(c, GMX_E_STREAM, "stream %u failed with status %u", i, status[i]);
  return 0;
}

// ---- device math self test ----------------------------------------------------------------------
__global__ void MathSweepKernel(uint32_t first, uint32_t stride, uint32_t count, float* e, float* l, float* t) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float x = gmx::u2f(first + i * stride);
  e[i] = gmx::gm_expf(x);
  t[i] = gmx::gm_tanhf(x);
  const uint32_t u = first + i * stride;
  l[i] = (u >= 0x00800000u && u < 0x7f800000u) ? gmx::gm_logf(x) : 0.0f;
}

// FNV-1a 64 of every stream's output slice; one thread per stream (slices are tens of KB).
__global__ void ChecksumKernel(const uint8_t* data, const uint64_t* off, const uint64_t* len, uint32_t n, uint64_t* sum) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* p = data + off[i];
  uint64_t h = 0xcbf29ce484222325ull;
  for (uint64_t k = 0, e = len[i]; k < e; ++k) h = (h ^ p[k]) * 0x100000001b3ull;
  sum[i] = h;
}

bool SameFloat(float a, float b) {
  if (a != a && b != b) return true;
  uint32_t x, y;
  memcpy(&x, &a, 4); memcpy(&y, &b, 4);
  return x == y;
}

}  // namespace

extern "C" {

const char* gmx_version(void) { return "gmix_b200 0.1 (sm_100a)"; }
const char* gmx_global_error(void) { return g_global_error.c_str(); }

int gmx_create(int device, gmx_ctx** out) {
  if (!out) return GMX_E_ARG;
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return Fail(nullptr, GMX_E_NODEVICE, "no CUDA device available (%s); gmix_b200 has no CPU fallback",
                e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
  if (device < 0 || device >= count) return Fail(nullptr, GMX_E_ARG, "device %d out of range (%d devices)", device, count);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return Fail(nullptr, GMX_E_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10) return Fail(nullptr, GMX_E_NODEVICE, "device %d is sm_%d%d; this library contains sm_100a code only", device, prop.major, prop.minor);
  gmx_ctx* c = new gmx_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreate(&c->ev0) != cudaSuccess || cudaEventCreate(&c->ev1) != cudaSuccess ||
      cudaMalloc(&c->d_queue, 256) != cudaSuccess || cudaMalloc(&c->d_layout, sizeof(gmx::ArenaLayout)) != cudaSuccess) {
    Fail(nullptr, GMX_E_CUDA, "CUDA initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete c;
    return GMX_E_CUDA;
  }
  c->stream = c->own_stream;
  if (const char* e = getenv("GMIX_B200_KERNEL_CONFIG")) { const int k = atoi(e); if (k >= 0 && k < gmx::kNumKernelConfigs) { c->kcfg = k; c->kcfg_user = k; } }
  std::vector<float> linit, adam;
  gmx::FillLstmInit(linit);
  gmx::FillAdamTable(adam);
  if (cudaMalloc(&c->d_lstm_init, linit.size() * 4) != cudaSuccess || cudaMalloc(&c->d_adam, adam.size() * 4) != cudaSuccess ||
      cudaMemcpy(c->d_lstm_init, linit.data(), linit.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(c->d_adam, adam.data(), adam.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    Fail(nullptr, GMX_E_CUDA, "table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    gmx_destroy(c);
    return GMX_E_CUDA;
  }
  *out = c;
  return 0;
}

void gmx_destroy(gmx_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  FreeArenas(c);
  for (DevBuf* b : {&c->b_in, &c->b_out, &c->b_in_off, &c->b_out_off, &c->b_out_len, &c->b_status, &c->b_trace, &c->b_ptrace, &c->b_prof, &c->b_ids, &c->b_usage, &c->b_rand, &c->b_final, &c->b_coder,
                     &c->b_park, &c->b_gx, &c->b_gsym, &c->b_gg, &c->b_wt, &c->b_an})
    if (b->p) cudaFree(b->p);
  if (c->d_layout) cudaFree(c->d_layout);
  if (c->d_roomy_layout) cudaFree(c->d_roomy_layout);
  if (c->d_lstm_init) cudaFree(c->d_lstm_init);
  if (c->d_adam) cudaFree(c->d_adam);
  if (c->d_decay) cudaFree(c->d_decay);
  if (c->d_queue) cudaFree(c->d_queue);
  for (int g = 0; g < gmx_ctx::kGroups; ++g) {
    if (c->gen_stream[g]) cudaStreamDestroy(c->gen_stream[g]);
    if (c->gen_done[g]) cudaEventDestroy(c->gen_done[g]);
  }
  if (c->gen_fork) cudaEventDestroy(c->gen_fork);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->own_stream) cudaStreamDestroy(c->own_stream);
  delete c;
}

const char* gmx_last_error(const gmx_ctx* c) { return c ? c->error.c_str() : g_global_error.c_str(); }

int gmx_set_kernel_config(gmx_ctx* c, int cfg) {
  if (!c) return GMX_E_ARG;
  if (cfg < -1 || cfg >= gmx::kNumKernelConfigs) return Fail(c, GMX_E_ARG, "kernel configuration %d out of range (-1 = automatic, 0 .. %d)", cfg, gmx::kNumKernelConfigs - 1);
  c->kcfg_user = cfg;
  if (cfg >= 0 && cfg != c->kcfg) { c->kcfg = cfg; FreeArenas(c); }   // residency (arena count) depends on the configuration
  return 0;
}
int gmx_get_kernel_config(const gmx_ctx* c) { return c ? c->kcfg : -1; }
int gmx_kernel_config_count(void) { return gmx::kNumKernelConfigs; }
int gmx_kernel_config_info(int cfg, int* bit_warps, int* lstm_warps, int* ctas_per_sm, int* serial, int* resident_weights) {
  if (cfg < 0 || cfg >= gmx::kNumKernelConfigs) return GMX_E_ARG;
  const gmx::KernelConfigInfo k = gmx::KernelConfig(cfg);
  if (bit_warps) *bit_warps = k.wb;
  if (lstm_warps) *lstm_warps = k.wl;
  if (ctas_per_sm) *ctas_per_sm = k.minb;
  if (serial) *serial = k.serial;
  if (resident_weights) *resident_weights = k.ws;
  return 0;
}

int gmx_set_cuda_stream(gmx_ctx* c, void* s) {
  if (!c) return GMX_E_ARG;
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  return 0;
}

int gmx_configure(gmx_ctx* c, uint64_t max_stream_len, uint32_t max_resident) {
  if (!c) return GMX_E_ARG;
  GMX_CUDA(c, cudaSetDevice(c->device));
  if (max_stream_len * 8 + 16 >= (1ull << 32)) return Fail(c, GMX_E_ARG, "streams of 512 MiB or more are not supported (32-bit bit-step counters)");
  FreeArenas(c);
  c->layout = gmx::MakeLayout(max_stream_len);
  {
    int rc = EnsureDecay(c, max_stream_len);
    if (rc) return rc;
  }
  int per_sm = 0;
  GMX_CUDA(c, gmx::OccupancyCompress(c->kcfg, &per_sm));
  if (per_sm > gmx::KernelConfig(c->kcfg).minb) per_sm = gmx::KernelConfig(c->kcfg).minb;   // a configuration is tuned for exactly this residency
  if (per_sm < 1) per_sm = 1;
  uint64_t want = (uint64_t)per_sm * c->sm_count;
  if (max_resident && max_resident < want) want = max_resident;
  size_t free_b = 0, total_b = 0;
  GMX_CUDA(c, cudaMemGetInfo(&free_b, &total_b));
  const uint64_t usable = free_b > (2ull << 30) ? free_b - (2ull << 30) : 0;  // leave 2 GiB for I/O buffers
  const uint64_t fit = usable / c->layout.total;
  if (fit == 0) return Fail(c, GMX_E_NOMEM, "one stream arena needs %llu MiB but only %llu MiB are free",
                            (unsigned long long)(c->layout.total >> 20), (unsigned long long)(free_b >> 20));
  if (fit < want) want = fit;
  GMX_CUDA(c, cudaMalloc(&c->d_arenas, want * c->layout.total));
  GMX_CUDA(c, cudaMemcpy(c->d_layout, &c->layout, sizeof(c->layout), cudaMemcpyHostToDevice));
  c->n_arenas = (uint32_t)want;
  c->cfg_max_len = max_stream_len;
  c->cfg_max_resident = max_resident;
  return 0;
}

uint64_t gmx_compress_bound(uint64_t n) { return n + n / 16 + 64; }

int gmx_compress_batch(gmx_ctx* c, const uint8_t* in, const uint64_t* in_off, uint32_t n, uint8_t* out,
                       const uint64_t* out_off, uint64_t* out_len, uint32_t* status) {
  return RunHost(c, gmx::MODE_COMPRESS, in, in_off, n, out, out_off, out_len, status, nullptr, nullptr);
}
int gmx_decompress_batch(gmx_ctx* c, const uint8_t* in, const uint64_t* in_off, uint32_t n, uint8_t* out,
                         const uint64_t* out_off, uint64_t* out_len, uint32_t* status) {
  return RunHost(c, gmx::MODE_DECOMPRESS, in, in_off, n, out, out_off, out_len, status, nullptr, nullptr);
}
int gmx_compress_batch_device(gmx_ctx* c, const uint8_t* d_in, const uint64_t* d_in_off, uint32_t n, uint8_t* d_out,
                              const uint64_t* d_out_off, uint64_t* d_out_len, uint32_t* d_status, uint64_t max_len) {
  return RunDevice(c, gmx::MODE_COMPRESS, d_in, d_in_off, n, d_out, d_out_off, d_out_len, d_status, max_len, RunOpts());
}
int gmx_decompress_batch_device(gmx_ctx* c, const uint8_t* d_in, const uint64_t* d_in_off, uint32_t n, uint8_t* d_out,
                                const uint64_t* d_out_off, uint64_t* d_out_len, uint32_t* d_status, uint64_t max_len) {
  return RunDevice(c, gmx::MODE_DECOMPRESS, d_in, d_in_off, n, d_out, d_out_off, d_out_len, d_status, max_len, RunOpts());
}

// ---- checkpoints and generation ---------------------------------------------------------------------
int gmx_model_load(gmx_ctx* c, const void* short_blob, uint64_t short_len, const void* long_blob, uint64_t long_len,
                   uint64_t max_new_bytes, int roomy, gmx_model** out) {
  if (!c || !out || !short_blob || !long_blob) return GMX_E_ARG;
  *out = nullptr;
  GMX_CUDA(c, cudaSetDevice(c->device));
  if (max_new_bytes * 8 + 16 >= (1ull << 32)) return Fail(c, GMX_E_ARG, "max_new_bytes out of range");
  gmx::ckpt::Image im;
  std::string err;
  if (!gmx::ckpt::Parse(short_blob, short_len, long_blob, long_len, &im, &err)) return Fail(c, GMX_E_ARG, "checkpoint: %s", err.c_str());
  gmx_model* m = new gmx_model();
  m->ctx = c;
  m->pre = gmx::ckpt::Count(im);
  if (m->pre.steps + max_new_bytes * 8 + 16 >= (1ull << 32)) {   // Mixer::steps_ and the decay table index are 32-bit on the device
    delete m;
    return Fail(c, GMX_E_ARG, "checkpoint trained on %llu bytes + %llu new bytes exceeds the 512 MiB (2^32 bit steps) limit",
                (unsigned long long)(im.mixer[0].steps / 8), (unsigned long long)max_new_bytes);
  }
  m->max_new_bytes = max_new_bytes;
  std::vector<uint8_t> arena;
  std::vector<uint32_t> state(sizeof(gmx::StreamSmem) / 4 + 4);
  bool ok = false;
  for (int attempt = roomy ? 1 : 0; attempt < 2 && !ok; ++attempt) {   // typical sizing first, worst case if the checkpoint does not fit it
    m->layout = gmx::MakeLayout(max_new_bytes, attempt == 1, &m->pre);
    arena.assign(m->layout.total, 0);
    ok = gmx::ckpt::ToArena(im, m->layout, arena.data(), (gmx::StreamSmem*)state.data(), &err);
  }
  if (!ok) { delete m; return Fail(c, GMX_E_ARG, "checkpoint: %s", err.c_str()); }
  m->l_epoch = ((const gmx::StreamSmem*)state.data())->l_epoch;
  if (cudaMalloc(&m->d_arena, m->layout.total) != cudaSuccess || cudaMalloc(&m->d_state, sizeof(gmx::StreamSmem)) != cudaSuccess ||
      cudaMalloc(&m->d_layout, sizeof(gmx::ArenaLayout)) != cudaSuccess ||
      cudaMemcpy(m->d_layout, &m->layout, sizeof(gmx::ArenaLayout), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(m->d_arena, arena.data(), m->layout.total, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(m->d_state, state.data(), sizeof(gmx::StreamSmem), cudaMemcpyHostToDevice) != cudaSuccess) {
    const char* what = cudaGetErrorString(cudaGetLastError());
    gmx_model_free(m);
    return Fail(c, GMX_E_NOMEM, "cannot place a %llu MiB model on the device: %s", (unsigned long long)(m->layout.total >> 20), what);
  }
  *out = m;
  return 0;
}

void gmx_model_free(gmx_model* m) {
  if (!m) return;
  cudaSetDevice(m->ctx->device);
  if (m->ctx->cfg_model == m) FreeArenas(m->ctx);
  if (m->d_arena) cudaFree(m->d_arena);
  if (m->d_state) cudaFree(m->d_state);
  if (m->d_layout) cudaFree(m->d_layout);
  delete m;
}

uint64_t gmx_model_arena_bytes(const gmx_model* m) { return m ? m->layout.total : 0; }
uint64_t gmx_model_trained_bytes(const gmx_model* m) { return m ? m->pre.steps / 8 : 0; }

int gmx_compress_batch_from(gmx_ctx* c, const gmx_model* model, const uint8_t* in, const uint64_t* in_off, uint32_t n, uint8_t* out,
                            const uint64_t* out_off, uint64_t* out_len, uint32_t* status) {
  RunOpts o; o.model = model;
  return RunHost(c, gmx::MODE_COMPRESS, in, in_off, n, out, out_off, out_len, status, nullptr, nullptr, o);
}
int gmx_decompress_batch_from(gmx_ctx* c, const gmx_model* model, const uint8_t* in, const uint64_t* in_off, uint32_t n, uint8_t* out,
                              const uint64_t* out_off, uint64_t* out_len, uint32_t* status) {
  RunOpts o; o.model = model;
  return RunHost(c, gmx::MODE_DECOMPRESS, in, in_off, n, out, out_off, out_len, status, nullptr, nullptr, o);
}

namespace {
// Lock-step batched generation (stream_kernel.cuh: StreamParams::lockstep, GenStepKernel; gate_gemm.cuh). Waves of at most n_arenas
// streams: one launch consumes the prompts (with learning, per stream), then every sampled byte is one batched gate product
// (exact SIMT or tcgen05) + one GenStepKernel launch over all streams of the wave.
constexpr int kGenGroups = gmx_ctx::kGroups;
constexpr uint32_t kGenOversubscribe = 2;
int RunLockstepGenerate(gmx_ctx* c, const gmx_model* model, const uint8_t* d_prompts, const uint64_t* d_prompt_off, uint32_t n, uint32_t out_bytes,
                        float temperature, const float* d_rand_u, uint64_t rand_stride, uint8_t* d_out, uint64_t* d_out_len, uint32_t* d_status,
                        uint64_t max_prompt_len, bool tensor) {
  GMX_CUDA(c, cudaSetDevice(c->device));
  if (model->ctx != c) return Fail(c, GMX_E_ARG, "model belongs to another context");
  const uint64_t max_len = max_prompt_len + out_bytes;
  if (max_len > model->max_new_bytes) return Fail(c, GMX_E_ARG, "stream of %llu bytes exceeds the max_new_bytes (%llu) the model was loaded with",
                                                  (unsigned long long)max_len, (unsigned long long)model->max_new_bytes);
  if (c->kcfg != gmx::kThroughputConfig) { c->kcfg = gmx::kThroughputConfig; FreeArenas(c); }   // GenStepKernel is that configuration's CTA
  // Streams in flight = kGenOversubscribe x the CTAs that fit the SMs, in GROUPS of two streams per SM that advance independently,
  // each in its own CUDA stream: a byte step is a barrier over the streams of ONE group only, and while a group waits for its
  // slowest stream, runs its (small) gate product or sits between two launches, the step kernels of the other groups keep every
  // CTA slot of the SMs busy.
  uint32_t oversub = kGenOversubscribe, group_streams = 2u * (uint32_t)c->sm_count;   // measured: profiles/r02_lockstep_generation.md
  if (const char* e = getenv("GMX_GEN_OVERSUB")) oversub = (uint32_t)std::max(1, atoi(e));        // development: A/B of the two knobs
  if (const char* e = getenv("GMX_GEN_GROUP")) group_streams = (uint32_t)std::max(1, atoi(e));
  int rc = ConfigureForModel(c, model, max_prompt_len ? max_prompt_len : max_len, max_len, oversub);
  if (rc) return rc;
  const uint32_t cap = c->n_arenas;
  const uint32_t ng = std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)kGenGroups, (cap + group_streams - 1) / group_streams));
  const uint32_t gs = (cap + ng - 1) / ng;                                   // arena slots per group
  const size_t state_bytes = sizeof(gmx::StreamSmem);
  const size_t gx_floats = gmx::GateXFloats(gs), gsym_words = (size_t)gs + 128, gg_floats = ((size_t)gs + gmx::GG_M) * gmx::GG_N;
  if ((rc = Reserve(c, c->b_park, (size_t)cap * state_bytes))) return rc;
  if ((rc = Reserve(c, c->b_gx, ng * gx_floats * 4))) return rc;
  if ((rc = Reserve(c, c->b_gsym, ng * gsym_words * 4))) return rc;
  if ((rc = Reserve(c, c->b_gg, ng * gg_floats * 4))) return rc;
  if ((rc = Reserve(c, c->b_wt, (size_t)gmx::GateWtFloats() * 4))) return rc;
  if ((rc = Reserve(c, c->b_usage, (size_t)n * 32))) return rc;
  c->usage_streams = n;
  for (uint32_t g = 0; g < ng; ++g) {
    if (!c->gen_stream[g]) GMX_CUDA(c, cudaStreamCreateWithFlags(&c->gen_stream[g], cudaStreamNonBlocking));
    if (!c->gen_done[g]) GMX_CUDA(c, cudaEventCreateWithFlags(&c->gen_done[g], cudaEventDisableTiming));
  }
  if (!c->gen_fork) GMX_CUDA(c, cudaEventCreateWithFlags(&c->gen_fork, cudaEventDisableTiming));
  GMX_CUDA(c, cudaEventRecord(c->ev0, c->stream));
  GMX_CUDA(c, cudaMemsetAsync(c->b_gx.p, 0, ng * gx_floats * 4, c->stream));    // rows of a partial last tile stay finite
  GMX_CUDA(c, cudaMemsetAsync(c->b_gsym.p, 0, ng * gsym_words * 4, c->stream));
  const float* d_w = (const float*)(model->d_arena + model->layout.l_w);
  if (tensor) GMX_CUDA(c, gmx::LaunchGateWeightPrep(d_w, (float*)c->b_wt.p, c->stream));
  GMX_CUDA(c, cudaEventRecord(c->gen_fork, c->stream));
  gmx::StreamParams P;
  memset(&P, 0, sizeof(P));
  P.in = d_prompts; P.in_off = d_prompt_off; P.out = d_out; P.out_len = d_out_len; P.status = d_status;
  P.n_streams = n; P.queue = c->d_queue;
  P.arena_stride = c->layout.total; P.layout = c->d_layout;
  P.lstm_init = c->d_lstm_init; P.decay = c->d_decay; P.decay_len = c->decay_len; P.adam = c->d_adam;
  P.analysis = -1; P.usage = (uint32_t*)c->b_usage.p;
  P.gen_bytes = out_bytes; P.temperature = temperature; P.rand_u = d_rand_u; P.rand_stride = rand_stride;
  P.tmpl_arena = model->d_arena; P.tmpl_state = model->d_state; P.tmpl_layout = c->layout.ov ? model->d_layout : nullptr;
  P.lockstep = 1;
  c->last_ms = 0;
  c->last_grid = n < cap ? n : cap;
  gmx::GenStepParams Q[kGenGroups];
  for (uint32_t g = 0; g < ng; ++g) {   // block b of a group's launches = arena slot g * gs + b
    gmx::StreamParams& G = Q[g].P;
    G = P;
    G.arenas = c->d_arenas + (uint64_t)g * gs * c->layout.total;
    G.park = (uint32_t*)c->b_park.p + (size_t)g * gs * (state_bytes / 4);
    G.final_state = G.park;
    G.gate_x = (float*)c->b_gx.p + g * gx_floats; G.gate_sym = (uint32_t*)c->b_gsym.p + g * gsym_words; G.gate_g = (const float*)c->b_gg.p + g * gg_floats;
    GMX_CUDA(c, cudaStreamWaitEvent(c->gen_stream[g], c->gen_fork, 0));
  }
  for (uint32_t base = 0; base < n;) {   // rounds: every group takes the next range of streams that fits its slots
    uint32_t active = 0;
    for (uint32_t g = 0; g < ng && base < n; ++g) {
      const uint32_t room = std::min<uint32_t>(gs, cap - g * gs);
      Q[g].n_slots = std::min<uint32_t>(room, n - base);
      Q[g].P.stream_base = base;
      base += Q[g].n_slots;
      active = g + 1;
      GMX_CUDA(c, gmx::LaunchGenerate(c->kcfg, Q[g].P, Q[g].n_slots, c->gen_stream[g]));
      c->launches += 1;
    }
    for (uint32_t i = 0; i < out_bytes; ++i)
      for (uint32_t g = 0; g < active; ++g) {
        const gmx::StreamParams& G = Q[g].P;
        if (tensor) GMX_CUDA(c, gmx::LaunchGateTc(G.gate_x, (const float*)c->b_wt.p, d_w, G.gate_sym, (float*)G.gate_g, Q[g].n_slots, c->gen_stream[g]));
        else GMX_CUDA(c, gmx::LaunchGateExact(d_w, G.gate_x, G.gate_sym, (float*)G.gate_g, Q[g].n_slots, (unsigned)c->sm_count, c->gen_stream[g]));
        Q[g].byte_index = i; Q[g].last = i + 1 == out_bytes;
        GMX_CUDA(c, gmx::LaunchGenStep(Q[g], Q[g].n_slots, c->gen_stream[g]));
        c->launches += 2;
      }
  }
  for (uint32_t g = 0; g < ng; ++g) {
    GMX_CUDA(c, cudaEventRecord(c->gen_done[g], c->gen_stream[g]));
    GMX_CUDA(c, cudaStreamWaitEvent(c->stream, c->gen_done[g], 0));
  }
  GMX_CUDA(c, cudaEventRecord(c->ev1, c->stream));
  GMX_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms = 0;
  GMX_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
  c->last_ms = ms;
  return 0;
}
}  // namespace

int gmx_set_generation_mode(gmx_ctx* c, int mode) {
  if (!c) return GMX_E_ARG;
  if (mode < GMX_GEN_PER_STREAM || mode > GMX_GEN_LOCKSTEP_TENSOR) return Fail(c, GMX_E_ARG, "generation mode %d out of range", mode);
  c->gen_mode = mode;
  return 0;
}
int gmx_last_generation_mode(const gmx_ctx* c) { return c ? c->last_gen_mode : -1; }

int gmx_generate_batch_device(gmx_ctx* c, const gmx_model* model, const uint8_t* d_prompts, const uint64_t* d_prompt_off, uint32_t n,
                              uint32_t out_bytes, float temperature, const float* d_rand_u, uint64_t rand_stride, uint8_t* d_out,
                              uint64_t* d_out_len, uint32_t* d_status, uint64_t max_prompt_len) {
  if (!c || !model) return GMX_E_ARG;
  if (!d_rand_u) return Fail(c, GMX_E_ARG, "null pointer argument");
  const float temp = temperature < (float)0.001 ? (float)0.001 : temperature;   // runner-utils.cpp:170
  // The batched gate product needs the gate matrix to be the model's in every stream: no stream may reach a BPTT pass while it
  // learns its prompt (all but the prompt's last byte are learned; a pass runs when the epoch slot wraps at L_HORIZON).
  const bool shared_gates = max_prompt_len >= 1 && model->l_epoch + (max_prompt_len - 1) < (uint64_t)gmx::L_HORIZON;
  if (c->gen_mode != GMX_GEN_PER_STREAM && shared_gates && n && out_bytes) {
    c->last_gen_mode = c->gen_mode;
    return RunLockstepGenerate(c, model, d_prompts, d_prompt_off, n, out_bytes, temp, d_rand_u, rand_stride, d_out, d_out_len, d_status, max_prompt_len,
                               c->gen_mode == GMX_GEN_LOCKSTEP_TENSOR);
  }
  c->last_gen_mode = GMX_GEN_PER_STREAM;
  RunOpts o; o.model = model; o.gen_bytes = out_bytes; o.d_rand_u = d_rand_u; o.rand_stride = rand_stride; o.overlay_learn = max_prompt_len;
  o.temperature = temp;
  return RunDevice(c, gmx::MODE_GENERATE, d_prompts, d_prompt_off, n, d_out, nullptr, d_out_len, d_status, max_prompt_len + out_bytes, o);
}

int gmx_generate_batch(gmx_ctx* c, const gmx_model* model, const uint8_t* prompts, const uint64_t* prompt_off, uint32_t n,
                       uint32_t out_bytes, float temperature, const float* rand_u, uint64_t rand_stride, uint8_t* out, uint32_t* status) {
  if (!c || !model) return GMX_E_ARG;
  if (n == 0) return 0;
  if (!prompts || !prompt_off || !rand_u || !out || !status) return Fail(c, GMX_E_ARG, "null pointer argument");
  GMX_CUDA(c, cudaSetDevice(c->device));
  uint64_t max_len = 0;
  for (uint32_t i = 0; i < n; ++i) {
    if (prompt_off[i + 1] <= prompt_off[i]) return Fail(c, GMX_E_ARG, "prompt %u is empty (the reference needs at least one byte)", i);
    max_len = std::max(max_len, prompt_off[i + 1] - prompt_off[i]);
  }
  std::vector<uint64_t> po(n + 1);
  for (uint32_t i = 0; i <= n; ++i) po[i] = prompt_off[i] - prompt_off[0];
  const uint64_t in_total = po[n], out_total = (uint64_t)n * out_bytes;
  const uint64_t n_rand = (rand_stride ? (uint64_t)(n - 1) * rand_stride : 0) + (uint64_t)out_bytes * 8;
  int rc;
  if ((rc = Reserve(c, c->b_in, in_total + 16))) return rc;
  if ((rc = Reserve(c, c->b_out, out_total + 16))) return rc;
  if ((rc = Reserve(c, c->b_in_off, (n + 1) * 8))) return rc;
  if ((rc = Reserve(c, c->b_out_len, (size_t)n * 8))) return rc;
  if ((rc = Reserve(c, c->b_status, (size_t)n * 4))) return rc;
  if ((rc = Reserve(c, c->b_rand, n_rand * 4 + 16))) return rc;
  GMX_CUDA(c, cudaMemcpyAsync(c->b_in.p, prompts + prompt_off[0], in_total, cudaMemcpyHostToDevice, c->stream));
  GMX_CUDA(c, cudaMemcpyAsync(c->b_in_off.p, po.data(), (n + 1) * 8, cudaMemcpyHostToDevice, c->stream));
  GMX_CUDA(c, cudaMemcpyAsync(c->b_rand.p, rand_u, n_rand * 4, cudaMemcpyHostToDevice, c->stream));
  GMX_CUDA(c, cudaMemsetAsync(c->b_status.p, 0xff, (size_t)n * 4, c->stream));
  rc = gmx_generate_batch_device(c, model, (const uint8_t*)c->b_in.p, (const uint64_t*)c->b_in_off.p, n, out_bytes, temperature,
                                 (const float*)c->b_rand.p, rand_stride, (uint8_t*)c->b_out.p, (uint64_t*)c->b_out_len.p,
                                 (uint32_t*)c->b_status.p, max_len);
  if (rc) return rc;
  GMX_CUDA(c, cudaMemcpyAsync(out, c->b_out.p, out_total, cudaMemcpyDeviceToHost, c->stream));
  GMX_CUDA(c, cudaMemcpyAsync(status, c->b_status.p, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
  GMX_CUDA(c, cudaStreamSynchronize(c->stream));
  for (uint32_t i = 0; i < n; ++i)
    if (status[i] != 0) return Fail(c, GMX_E_STREAM, "stream %u failed with status %u", i, status[i]);
  return 0;
}

void gmx_reference_rand_u(float* out, uint64_t n) {
  // what `gmix -g` draws: Predictor() seeds srand(0xDEADBEEF) (predictor.cpp:18), the LSTM initialisation takes
  // 3*50*563 draws (lstm-layer.cpp:176-195), sampling continues that sequence (runner-utils.cpp:18-20,203)
  std::vector<float> skip;
  gmx::FillLstmInit(skip);
  for (uint64_t i = 0; i < n; ++i) out[i] = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
}

namespace {
// arena + parked state of stream `slot` of the ctx arenas -> reference checkpoint bytes in c->ck_short / c->ck_long
int SerializeParked(gmx_ctx* c, const gmx::ArenaLayout& L, const uint8_t* d_arena, const uint32_t* d_state) {
  std::vector<uint8_t> arena(L.total);
  std::vector<uint32_t> state(sizeof(gmx::StreamSmem) / 4 + 4);
  GMX_CUDA(c, cudaMemcpy(arena.data(), d_arena, L.total, cudaMemcpyDeviceToHost));
  GMX_CUDA(c, cudaMemcpy(state.data(), d_state, sizeof(gmx::StreamSmem), cudaMemcpyDeviceToHost));
  gmx::ckpt::Image im;
  std::string err;
  if (!gmx::ckpt::FromArena(L, arena.data(), *(const gmx::StreamSmem*)state.data(), &im, &err)) return Fail(c, GMX_E_ARG, "checkpoint: %s", err.c_str());
  gmx::ckpt::Serialize(im, &c->ck_short, &c->ck_long);
  return 0;
}
}  // namespace

int gmx_train_checkpoint(gmx_ctx* c, const gmx_model* from, const uint8_t* data, uint64_t n, const void** short_blob,
                         uint64_t* short_len, const void** long_blob, uint64_t* long_len) {
  if (!c || !short_blob || !short_len || !long_blob || !long_len || (!data && n)) return GMX_E_ARG;
  GMX_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = Reserve(c, c->b_final, sizeof(gmx::StreamSmem)))) return rc;
  std::vector<uint8_t> out(gmx_compress_bound(n));
  const uint64_t io[2] = {0, n}, oo[2] = {0, out.size()};
  uint64_t out_len = 0;
  uint32_t status = 0;
  const uint8_t dummy = 0;
  RunOpts o; o.model = from; o.analysis = 0; o.d_final_state = (uint32_t*)c->b_final.p;
  const uint32_t keep_resident = c->cfg_max_resident;
  rc = RunHost(c, gmx::MODE_COMPRESS, n ? data : &dummy, io, 1, out.data(), oo, &out_len, &status, nullptr, nullptr, o);
  (void)keep_resident;
  bool roomy = false;
  if (rc == GMX_E_STREAM && Retryable(status) && !from) {
    // the text-sized arena overflowed (the parked state must come from the arena the stream ran in, so the batch
    // calls' retry in a second arena class does not apply): run again in one worst-case-sized arena
    FreeArenas(c);
    c->layout = gmx::MakeLayout(n, true);
    if (cudaMalloc(&c->d_arenas, c->layout.total) != cudaSuccess)
      return Fail(c, GMX_E_NOMEM, "cannot allocate a worst-case arena of %llu MiB: %s", (unsigned long long)(c->layout.total >> 20),
                  cudaGetErrorString(cudaGetLastError()));
    GMX_CUDA(c, cudaMemcpy(c->d_layout, &c->layout, sizeof(c->layout), cudaMemcpyHostToDevice));
    c->n_arenas = 1;
    c->cfg_max_len = n;
    roomy = true;
    c->retried_streams += 1;
    rc = RunHost(c, gmx::MODE_COMPRESS, n ? data : &dummy, io, 1, out.data(), oo, &out_len, &status, nullptr, nullptr, o);
  }
  if (rc) { if (roomy) FreeArenas(c); return rc; }
  rc = SerializeParked(c, c->layout, c->d_arenas, (const uint32_t*)c->b_final.p);   // one stream: it ran in arena 0
  if (roomy) FreeArenas(c);   // the next batch call sizes its arenas afresh
  if (rc) return rc;
  *short_blob = c->ck_short.data(); *short_len = c->ck_short.size();
  *long_blob = c->ck_long.data(); *long_len = c->ck_long.size();
  return 0;
}

namespace {
// One stream, one part, parked afterwards and serialised as the reference's checkpoint files. `from` == null: the part
// starts a stream from scratch.
int RunPart(gmx_ctx* c, int mode, const gmx_model* from, const gmx_coder_state* coder_in, RunOpts o, const uint8_t* in, uint64_t n_in, uint8_t* out,
            uint64_t cap, uint64_t* out_len, gmx_coder_state* coder_out, uint64_t* in_consumed, const void** short_blob, uint64_t* short_len,
            const void** long_blob, uint64_t* long_len, float* probs = nullptr) {
  GMX_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = Reserve(c, c->b_final, sizeof(gmx::StreamSmem)))) return rc;
  if ((rc = Reserve(c, c->b_coder, 64))) return rc;
  uint32_t h[8] = {0, 0xffffffffu, 0, 0, 0, 0, 0, 0};
  if (coder_in) { h[0] = coder_in->x1; h[1] = coder_in->x2; h[2] = coder_in->x; }
  GMX_CUDA(c, cudaMemcpy(c->b_coder.p, h, 32, cudaMemcpyHostToDevice));
  o.model = from; o.d_final_state = (uint32_t*)c->b_final.p; o.part = 1;
  o.d_coder_in = coder_in ? (const uint32_t*)c->b_coder.p : nullptr; o.d_coder_out = (uint32_t*)c->b_coder.p + 4;
  const uint64_t io[2] = {0, n_in}, oo[2] = {0, cap};
  uint32_t status = 0;
  const uint8_t dummy = 0;
  std::vector<uint64_t> bt(probs ? n_in * 8 + 1 : 0);
  uint64_t* trace = probs ? bt.data() : nullptr;
  rc = RunHost(c, mode, n_in ? in : &dummy, io, 1, out, oo, out_len, &status, trace, nullptr, o);
  bool roomy = false;
  if (rc == GMX_E_STREAM && Retryable(status) && !from) {   // as gmx_train_checkpoint: once more in one worst-case-sized arena
    const uint64_t len = mode == gmx::MODE_COMPRESS ? n_in : cap;
    FreeArenas(c);
    c->layout = gmx::MakeLayout(len, true);
    if (cudaMalloc(&c->d_arenas, c->layout.total) != cudaSuccess)
      return Fail(c, GMX_E_NOMEM, "cannot allocate a worst-case arena of %llu MiB: %s", (unsigned long long)(c->layout.total >> 20),
                  cudaGetErrorString(cudaGetLastError()));
    GMX_CUDA(c, cudaMemcpy(c->d_layout, &c->layout, sizeof(c->layout), cudaMemcpyHostToDevice));
    c->n_arenas = 1;
    c->cfg_max_len = len;
    roomy = true;
    c->retried_streams += 1;
    rc = RunHost(c, mode, n_in ? in : &dummy, io, 1, out, oo, out_len, &status, trace, nullptr, o);
  }
  if (rc) { if (roomy) FreeArenas(c); return rc; }
  for (uint64_t i = 0; probs && i < n_in * 8; ++i) { const uint32_t lo = (uint32_t)bt[i]; memcpy(&probs[i], &lo, 4); }
  GMX_CUDA(c, cudaMemcpy(h, (uint32_t*)c->b_coder.p + 4, 16, cudaMemcpyDeviceToHost));
  if (coder_out) { coder_out->x1 = h[0]; coder_out->x2 = h[1]; coder_out->x = h[2]; }
  if (in_consumed) *in_consumed = h[3];
  if (short_blob) {
    rc = SerializeParked(c, c->layout, c->d_arenas, (const uint32_t*)c->b_final.p);   // one stream: it ran in arena 0
    if (roomy) FreeArenas(c);
    if (rc) return rc;
    *short_blob = c->ck_short.data(); *short_len = c->ck_short.size();
    *long_blob = c->ck_long.data(); *long_len = c->ck_long.size();
  } else if (roomy) {
    FreeArenas(c);
  }
  return 0;
}
}  // namespace

int gmx_compress_part(gmx_ctx* c, const gmx_model* from, const gmx_coder_state* coder_in, int write_header, uint64_t total_len, int last, int analysis,
                      const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_len, gmx_coder_state* coder_out,
                      const void** short_blob, uint64_t* short_len, const void** long_blob, uint64_t* long_len, float* probs) {
  if (!c || !out || !out_len || (!in && n) || (short_blob && (!short_len || !long_blob || !long_len))) return GMX_E_ARG;
  RunOpts o;
  o.analysis = analysis; o.part_header = write_header != 0; o.part_last = last != 0; o.part_total = total_len;
  return RunPart(c, gmx::MODE_COMPRESS, from, coder_in, o, in, n, out, cap, out_len, coder_out, nullptr, short_blob, short_len, long_blob, long_len, probs);
}

int gmx_decompress_part(gmx_ctx* c, const gmx_model* from, const gmx_coder_state* coder_in, int analysis, const uint8_t* in, uint64_t n_in,
                        uint64_t out_bytes, uint8_t* out, uint64_t* in_consumed, gmx_coder_state* coder_out,
                        const void** short_blob, uint64_t* short_len, const void** long_blob, uint64_t* long_len) {
  if (!c || !out || (!in && n_in) || (short_blob && (!short_len || !long_blob || !long_len))) return GMX_E_ARG;
  RunOpts o;
  o.analysis = analysis; o.part_header = coder_in == nullptr; o.part_total = out_bytes;
  uint64_t out_len = 0;
  return RunPart(c, gmx::MODE_DECOMPRESS, from, coder_in, o, in, n_in, out, out_bytes + 8, &out_len, coder_out, in_consumed, short_blob, short_len, long_blob, long_len);
}

int gmx_checksum_device(gmx_ctx* c, const uint8_t* d_data, const uint64_t* d_off, const uint64_t* d_len, uint32_t n, uint64_t* d_sum) {
  if (!c) return GMX_E_ARG;
  if (n == 0) return 0;
  if (!d_data || !d_off || !d_len || !d_sum) return Fail(c, GMX_E_ARG, "null pointer argument");
  GMX_CUDA(c, cudaSetDevice(c->device));
  ChecksumKernel<<<(n + 63) / 64, 64, 0, c->stream>>>(d_data, d_off, d_len, n, d_sum);
  GMX_CUDA(c, cudaGetLastError());
  c->launches += 1;
  return 0;
}

int gmx_compress_analysis(gmx_ctx* c, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_len, uint32_t sample_frequency,
                          gmx_analysis_row* rows, uint32_t max_rows, uint32_t* n_rows) {
  static_assert(sizeof(gmx_analysis_row) == sizeof(gmx::AnalysisRow) && GMX_ANALYSIS_COLUMNS == gmx::AN_COLS, "gmix_b200.h mirrors stream_kernel.cuh");
  if (!c || !in || !out || !out_len || !rows || !n_rows || sample_frequency == 0) return GMX_E_ARG;
  GMX_CUDA(c, cudaSetDevice(c->device));
  *n_rows = 0;
  const uint64_t want = n ? (8 * n - 1) / sample_frequency : 0;   // samples at bits_seen = f, 2f, ... <= 8n - 1
  const uint32_t nr = (uint32_t)std::min<uint64_t>(want, max_rows);
  int rc = Reserve(c, c->b_an, gmx::AN_COLS * 8 + (size_t)(nr + 1) * sizeof(gmx::AnalysisRow));
  if (rc) return rc;
  double init[gmx::AN_COLS];
  for (double& e : init) e = -1;   // predictor.cpp:37-38
  GMX_CUDA(c, cudaMemcpyAsync(c->b_an.p, init, sizeof(init), cudaMemcpyHostToDevice, c->stream));
  GMX_CUDA(c, cudaMemsetAsync((char*)c->b_an.p + sizeof(init), 0, (size_t)(nr + 1) * sizeof(gmx::AnalysisRow), c->stream));
  RunOpts o;
  o.analysis = 1;
  o.d_an_entropy = (double*)c->b_an.p; o.d_an_rows = (gmx::AnalysisRow*)((char*)c->b_an.p + sizeof(init));
  o.an_freq = sample_frequency; o.an_max_rows = nr;
  const uint64_t in_off[2] = {0, n}, out_off[2] = {0, cap};
  uint32_t status = 0;
  const int keep = c->kcfg_user;
  c->kcfg_user = 0;   // the analysis step lives in the phase-serial order (SerialCompress)
  rc = RunHost(c, gmx::MODE_COMPRESS, in, in_off, 1, out, out_off, out_len, &status, nullptr, nullptr, o);
  c->kcfg_user = keep;
  if (rc) return rc;
  GMX_CUDA(c, cudaMemcpy(rows, o.d_an_rows, (size_t)nr * sizeof(gmx::AnalysisRow), cudaMemcpyDeviceToHost));
  *n_rows = nr;
  return 0;
}

int gmx_compress_trace(gmx_ctx* c, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_len,
                       float* probs, uint32_t* p16, float* blackboard) {
  if (!c || !probs || !p16) return GMX_E_ARG;
  std::vector<uint64_t> bt(n * 8 + 1);
  const uint64_t io[2] = {0, n}, oo[2] = {0, cap};
  uint32_t status = 0;
  int rc = RunHost(c, gmx::MODE_COMPRESS, in, io, 1, out, oo, out_len, &status, bt.data(), blackboard);
  if (rc) return rc;
  for (uint64_t i = 0; i < n * 8; ++i) {
    const uint32_t lo = (uint32_t)bt[i];
    memcpy(&probs[i], &lo, 4);
    p16[i] = (uint32_t)(bt[i] >> 32);
  }
  return 0;
}

int gmx_set_profile(gmx_ctx* c, int on) {
  if (!c) return GMX_E_ARG;
  c->profile = on != 0;
  return 0;
}
int gmx_get_profile(gmx_ctx* c, uint64_t* out, uint32_t max_streams) {
  if (!c || !out) return GMX_E_ARG;
  const uint32_t n = c->prof_streams < max_streams ? c->prof_streams : max_streams;
  if (n == 0 || !c->b_prof.p) return 0;
  GMX_CUDA(c, cudaMemcpy(out, c->b_prof.p, (size_t)n * gmx::GMX_PROF_SLOTS * 8, cudaMemcpyDeviceToHost));
  return (int)n;
}

int gmx_get_usage(gmx_ctx* c, uint32_t* out, uint32_t max_streams) {
  if (!c || !out) return GMX_E_ARG;
  const uint32_t n = c->usage_streams < max_streams ? c->usage_streams : max_streams;
  if (n == 0 || !c->b_usage.p) return 0;
  GMX_CUDA(c, cudaMemcpy(out, c->b_usage.p, (size_t)n * 32, cudaMemcpyDeviceToHost));
  return (int)n;
}

// ---- Predictor facade (reference src/predictor.h:20-38) -------------------------------------------
static int PredStep(gmx_pred* p, int op, float* prob) {
  gmx_ctx* c = p->ctx;
  GMX_CUDA(c, cudaSetDevice(c->device));
  gmx::StepParams Q;
  memset(&Q, 0, sizeof(Q));
  Q.P.arenas = p->d_arena; Q.P.arena_stride = p->layout.total; Q.P.layout = p->d_layout;
  Q.P.lstm_init = c->d_lstm_init; Q.P.decay = c->d_decay; Q.P.decay_len = c->decay_len; Q.P.adam = c->d_adam;
  Q.state = p->d_state; Q.op = op; Q.has_bit = p->pending_bit >= 0; Q.bit = p->pending_bit > 0; Q.analysis = p->analysis;
  Q.prob_out = p->d_prob; Q.status_out = (uint32_t*)(p->d_prob + 1);
  GMX_CUDA(c, gmx::LaunchStep(Q, c->stream));
  c->launches += 1;
  p->pending_bit = -1;
  float host[2];
  GMX_CUDA(c, cudaMemcpyAsync(host, p->d_prob, 8, cudaMemcpyDeviceToHost, c->stream));
  GMX_CUDA(c, cudaStreamSynchronize(c->stream));
  uint32_t st;
  memcpy(&st, &host[1], 4);
  if (st != 0) return Fail(c, GMX_E_STREAM, "predictor stream failed with status %u", st);
  if (prob) *prob = host[0];
  return 0;
}

int gmx_pred_new(gmx_ctx* c, uint64_t max_stream_len, gmx_pred** out) {
  if (!c || !out) return GMX_E_ARG;
  *out = nullptr;
  GMX_CUDA(c, cudaSetDevice(c->device));
  if (max_stream_len == 0 || max_stream_len * 8 + 16 >= (1ull << 32)) return Fail(c, GMX_E_ARG, "max_stream_len out of range (1 .. 512 MiB)");
  int rc = EnsureDecay(c, max_stream_len);
  if (rc) return rc;
  gmx_pred* p = new gmx_pred();
  p->ctx = c;
  p->layout = gmx::MakeLayout(max_stream_len, true);
  p->max_bits = max_stream_len * 8;
  if (cudaMalloc(&p->d_arena, p->layout.total) != cudaSuccess || cudaMalloc(&p->d_layout, sizeof(gmx::ArenaLayout)) != cudaSuccess ||
      cudaMalloc(&p->d_state, gmx::StepStateBytes()) != cudaSuccess || cudaMalloc(&p->d_prob, 8) != cudaSuccess ||
      cudaMemcpy(p->d_layout, &p->layout, sizeof(gmx::ArenaLayout), cudaMemcpyHostToDevice) != cudaSuccess) {
    gmx_pred_free(p);
    return Fail(c, GMX_E_NOMEM, "cannot allocate a %llu MiB predictor arena: %s", (unsigned long long)(p->layout.total >> 20),
                cudaGetErrorString(cudaGetLastError()));
  }
  rc = PredStep(p, gmx::STEP_INIT, nullptr);
  if (rc) { gmx_pred_free(p); return rc; }
  *out = p;
  return 0;
}
void gmx_pred_free(gmx_pred* p) {
  if (!p) return;
  cudaSetDevice(p->ctx->device);
  if (p->d_arena) cudaFree(p->d_arena);
  if (p->d_layout) cudaFree(p->d_layout);
  if (p->d_state) cudaFree(p->d_state);
  if (p->d_prob) cudaFree(p->d_prob);
  delete p;
}
int gmx_pred_enable_analysis(gmx_pred* p, int on) {
  if (!p) return GMX_E_ARG;
  p->analysis = on != 0;
  return 0;
}
int gmx_pred_predict(gmx_pred* p, float* prob) {
  if (!p || !prob) return GMX_E_ARG;
  if (p->bits >= p->max_bits) return Fail(p->ctx, GMX_E_ARG, "predictor was created for %llu bits", (unsigned long long)p->max_bits);
  return PredStep(p, gmx::STEP_PREDICT, prob);
}
int gmx_pred_perceive(gmx_pred* p, int bit) {
  if (!p) return GMX_E_ARG;
  p->pending_bit = bit != 0;
  p->bits++;
  return 0;
}
int gmx_pred_learn(gmx_pred* p) {
  if (!p) return GMX_E_ARG;
  return PredStep(p, gmx::STEP_LEARN, nullptr);
}

// Predictor::Copy (predictor.cpp:42-48): dst becomes a deep copy of src (device-to-device: arena + parked state).
int gmx_pred_copy(gmx_pred* dst, const gmx_pred* src) {
  if (!dst || !src) return GMX_E_ARG;
  gmx_ctx* c = dst->ctx;
  if (src->ctx != c) return Fail(c, GMX_E_ARG, "predictors belong to different contexts");
  if (memcmp(&dst->layout, &src->layout, sizeof(gmx::ArenaLayout)) != 0) return Fail(c, GMX_E_ARG, "predictors were created with different max_stream_len");
  GMX_CUDA(c, cudaSetDevice(c->device));
  GMX_CUDA(c, cudaMemcpyAsync(dst->d_arena, src->d_arena, src->layout.total, cudaMemcpyDeviceToDevice, c->stream));
  GMX_CUDA(c, cudaMemcpyAsync(dst->d_state, src->d_state, gmx::StepStateBytes(), cudaMemcpyDeviceToDevice, c->stream));
  GMX_CUDA(c, cudaStreamSynchronize(c->stream));
  dst->pending_bit = src->pending_bit; dst->analysis = src->analysis; dst->bits = src->bits;
  return 0;
}

int gmx_pred_write_checkpoint(gmx_pred* p, const void** short_blob, uint64_t* short_len, const void** long_blob, uint64_t* long_len) {
  if (!p || !short_blob || !short_len || !long_blob || !long_len) return GMX_E_ARG;
  gmx_ctx* c = p->ctx;
  GMX_CUDA(c, cudaSetDevice(c->device));
  if (p->pending_bit >= 0) {   // Perceive travels with the next launch: apply it to the parked state first
    GMX_CUDA(c, cudaMemcpy((uint8_t*)p->d_state + offsetof(gmx::StreamSmem, new_bit), &p->pending_bit, 4, cudaMemcpyHostToDevice));
  }
  int rc = SerializeParked(c, p->layout, p->d_arena, p->d_state);
  if (rc) return rc;
  *short_blob = c->ck_short.data(); *short_len = c->ck_short.size();
  *long_blob = c->ck_long.data(); *long_len = c->ck_long.size();
  return 0;
}

int gmx_pred_read_checkpoint(gmx_pred* p, const void* short_blob, uint64_t short_len, const void* long_blob, uint64_t long_len) {
  if (!p || !short_blob || !long_blob) return GMX_E_ARG;
  gmx_ctx* c = p->ctx;
  GMX_CUDA(c, cudaSetDevice(c->device));
  gmx::ckpt::Image im;
  std::string err;
  if (!gmx::ckpt::Parse(short_blob, short_len, long_blob, long_len, &im, &err)) return Fail(c, GMX_E_ARG, "checkpoint: %s", err.c_str());
  // the predictor's arena was sized (worst case) for max_stream_len bytes from scratch; the checkpoint must fit it
  std::vector<uint8_t> arena(p->layout.total);
  std::vector<uint32_t> state(sizeof(gmx::StreamSmem) / 4 + 4);
  if (!gmx::ckpt::ToArena(im, p->layout, arena.data(), (gmx::StreamSmem*)state.data(), &err))
    return Fail(c, GMX_E_ARG, "checkpoint does not fit this predictor (create it with a larger max_stream_len): %s", err.c_str());
  const uint64_t trained_bits = im.mixer[0].steps;
  if (trained_bits + p->max_bits + 16 >= (1ull << 32)) return Fail(c, GMX_E_ARG, "checkpoint + predictor length exceed the 2^32 bit-step limit");
  int rc = EnsureDecay(c, trained_bits / 8 + p->max_bits / 8 + 2);
  if (rc) return rc;
  ((gmx::StreamSmem*)state.data())->analysis = p->analysis;
  GMX_CUDA(c, cudaMemcpy(p->d_arena, arena.data(), p->layout.total, cudaMemcpyHostToDevice));
  GMX_CUDA(c, cudaMemcpy(p->d_state, state.data(), sizeof(gmx::StreamSmem), cudaMemcpyHostToDevice));
  p->pending_bit = -1;
  p->bits = 0;
  return 0;
}

uint32_t gmx_resident_streams(const gmx_ctx* c) { return c ? c->last_grid : 0; }
uint32_t gmx_arena_count(const gmx_ctx* c) { return c ? c->n_arenas : 0; }
uint64_t gmx_arena_bytes(const gmx_ctx* c) { return c ? c->layout.total : 0; }
uint64_t gmx_retried_streams(const gmx_ctx* c) { return c ? c->retried_streams : 0; }
uint64_t gmx_kernel_launches(const gmx_ctx* c) { return c ? c->launches : 0; }
double gmx_last_kernel_ms(const gmx_ctx* c) { return c ? c->last_ms : 0; }
int gmx_device_sm_count(const gmx_ctx* c) { return c ? c->sm_count : 0; }

int gmx_selftest_gate(gmx_ctx* c, uint32_t n_slots, uint32_t seed, uint64_t* exact_mismatches, double err[3]) {
  if (!c || !exact_mismatches || !err || n_slots == 0 || n_slots > 65536) return GMX_E_ARG;
  GMX_CUDA(c, cudaSetDevice(c->device));
  uint64_t rs = seed * 2654435761ull + 88172645463325252ull;
  auto rnd = [&]() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (float)((rs >> 40) & 0xffffff) / 16777216.0f; };   // [0, 1)
  std::vector<float> W(gmx::L_WSIZE), X(gmx::GateXFloats(n_slots), 0.0f), xv((size_t)n_slots * gmx::GG_K, 0.0f);
  std::vector<uint32_t> sym(n_slots);
  for (auto& w : W) w = (rnd() - 0.5f) * 0.6f;
  for (uint32_t s = 0; s < n_slots; ++s) {
    float sum = 0;
    for (int k = 0; k < 256; ++k) { const float r = rnd(); xv[(size_t)s * gmx::GG_K + k] = r * r * r * r; sum += r * r * r * r; }
    for (int k = 0; k < 256; ++k) xv[(size_t)s * gmx::GG_K + k] /= sum;
    for (int k = 256; k < 306; ++k) xv[(size_t)s * gmx::GG_K + k] = rnd() * 2.0f - 1.0f;
    xv[(size_t)s * gmx::GG_K + 306] = 1.0f;
    sym[s] = (uint32_t)(rnd() * 256.0f) & 0xffu;
    for (int k = 0; k < gmx::GG_K; ++k) {
      const float v = xv[(size_t)s * gmx::GG_K + k], hi = gmx::Tf32Hi(v);
      X[gmx::GateXIndex(s, k, 0)] = hi; X[gmx::GateXIndex(s, k, 1)] = v - hi;
    }
  }
  float *dW = nullptr, *dX = nullptr, *dWt = nullptr, *dG = nullptr, *dG2 = nullptr; uint32_t* dS = nullptr;
  const size_t g_bytes = ((size_t)n_slots + gmx::GG_M) * gmx::GG_N * 4;
  GMX_CUDA(c, cudaMalloc(&dW, W.size() * 4)); GMX_CUDA(c, cudaMalloc(&dX, X.size() * 4)); GMX_CUDA(c, cudaMalloc(&dWt, (size_t)gmx::GateWtFloats() * 4));
  GMX_CUDA(c, cudaMalloc(&dG, g_bytes)); GMX_CUDA(c, cudaMalloc(&dG2, g_bytes)); GMX_CUDA(c, cudaMalloc(&dS, (size_t)n_slots * 4 + 512));
  GMX_CUDA(c, cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice));
  GMX_CUDA(c, cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
  GMX_CUDA(c, cudaMemset(dS, 0, (size_t)n_slots * 4 + 512));
  GMX_CUDA(c, cudaMemcpy(dS, sym.data(), (size_t)n_slots * 4, cudaMemcpyHostToDevice));
  GMX_CUDA(c, cudaMemset(dG, 0, g_bytes)); GMX_CUDA(c, cudaMemset(dG2, 0, g_bytes));
  GMX_CUDA(c, gmx::LaunchGateWeightPrep(dW, dWt, c->stream));
  GMX_CUDA(c, gmx::LaunchGateExact(dW, dX, dS, dG, n_slots, (unsigned)c->sm_count, c->stream));
  GMX_CUDA(c, gmx::LaunchGateTc(dX, dWt, dW, dS, dG2, n_slots, c->stream));
  c->launches += 3;
  GMX_CUDA(c, cudaStreamSynchronize(c->stream));
  std::vector<float> G((size_t)n_slots * gmx::GG_N), G2((size_t)n_slots * gmx::GG_N);
  GMX_CUDA(c, cudaMemcpy(G.data(), dG, G.size() * 4, cudaMemcpyDeviceToHost));
  GMX_CUDA(c, cudaMemcpy(G2.data(), dG2, G2.size() * 4, cudaMemcpyDeviceToHost));
  cudaFree(dW); cudaFree(dX); cudaFree(dWt); cudaFree(dG); cudaFree(dG2); cudaFree(dS);
  uint64_t bad = 0;
  double e_tc = 0, e_seq = 0, mag = 0;
  for (uint32_t s = 0; s < n_slots; ++s)
    for (int r = 0; r < gmx::GG_ROWS; ++r) {
      const int g = r / gmx::L_CELLS, i = r - g * gmx::L_CELLS;
      volatile float f = W[gmx::LstmW(g, (int)sym[s], i)];
      double d = (double)f;
      for (int k = 0; k < gmx::GG_KUSED; ++k) {
        const float x = xv[(size_t)s * gmx::GG_K + k], w = W[gmx::LstmW(g, gmx::L_NOUT + k, i)];
        volatile float pr = x * w;
        f = f + pr;
        d += (double)x * (double)w;
      }
      const float fe = f;
      if (!SameFloat(fe, G[(size_t)s * gmx::GG_N + r])) ++bad;
      e_tc = std::max(e_tc, fabs((double)G2[(size_t)s * gmx::GG_N + r] - d));
      e_seq = std::max(e_seq, fabs((double)fe - d));
      mag = std::max(mag, fabs(d));
    }
  *exact_mismatches = bad;
  err[0] = e_tc; err[1] = e_seq; err[2] = mag;
  return 0;
}

int gmx_selftest_math(gmx_ctx* c, uint32_t stride, uint64_t mismatches[3], uint32_t first_bad[3]) {
  if (!c || !mismatches || !first_bad || stride == 0) return GMX_E_ARG;
  GMX_CUDA(c, cudaSetDevice(c->device));
  const uint32_t chunk = 1u << 24;
  float *d[3] = {nullptr, nullptr, nullptr};
  for (int k = 0; k < 3; ++k) GMX_CUDA(c, cudaMalloc(&d[k], (size_t)chunk * 4));
  std::vector<float> h[3];
  for (int k = 0; k < 3; ++k) h[k].resize(chunk);
  std::atomic<uint64_t> bad[3];
  std::atomic<uint32_t> first[3];
  for (int k = 0; k < 3; ++k) { bad[k] = 0; first[k] = 0xffffffffu; }
  const uint64_t total = ((1ull << 32) + stride - 1) / stride;
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 4;
  if (nt > 64) nt = 64;
  for (uint64_t done = 0; done < total; done += chunk) {
    const uint32_t count = (uint32_t)((total - done) < chunk ? (total - done) : chunk);
    const uint32_t first_u = (uint32_t)(done * stride);
    MathSweepKernel<<<(count + 255) / 256, 256, 0, c->stream>>>(first_u, stride, count, d[0], d[1], d[2]);
    GMX_CUDA(c, cudaGetLastError());
    c->launches += 1;
    for (int k = 0; k < 3; ++k) GMX_CUDA(c, cudaMemcpyAsync(h[k].data(), d[k], (size_t)count * 4, cudaMemcpyDeviceToHost, c->stream));
    GMX_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t] {
      for (uint32_t i = t; i < count; i += nt) {
        const uint32_t u = first_u + i * stride;
        float x;
        memcpy(&x, &u, 4);
        if (!SameFloat(h[0][i], expf(x))) { bad[0]++; uint32_t f = first[0]; while (u < f && !first[0].compare_exchange_weak(f, u)) {} }
        if (u >= 0x00800000u && u < 0x7f800000u && !SameFloat(h[1][i], logf(x))) { bad[1]++; uint32_t f = first[1]; while (u < f && !first[1].compare_exchange_weak(f, u)) {} }
        if (!SameFloat(h[2][i], tanhf(x))) { bad[2]++; uint32_t f = first[2]; while (u < f && !first[2].compare_exchange_weak(f, u)) {} }
      }
    });
    for (auto& x : th) x.join();
  }
  for (int k = 0; k < 3; ++k) { cudaFree(d[k]); mismatches[k] = bad[k]; first_bad[k] = first[k]; }
  return 0;
}

}  // extern "C"
