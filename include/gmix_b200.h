/* gmix_b200 — C ABI of the B200-native gmix per-bit path (libgmix_b200.so).
 *
 * This is the drop-in boundary: plain C, plain pointers and sizes, no CUDA or torch types. The
 * reference has no FFI of its own (it is one C++ program); what it exposes for this path is the
 * Predictor class and the runner entry points, so each entry point below cites the reference
 * interface it replaces (paths relative to the reference's src/):
 *
 *   gmx_compress_batch / _device    runner_utils::RunCompression + Compress   runner/runner-utils.cpp:43-67,88-121
 *   gmx_decompress_batch / _device  runner_utils::RunDecompression + Decompress runner/runner-utils.cpp:69-86,123-156
 *                                   (coder: coder/encoder.cpp:8-34, coder/decoder.cpp:3-39;
 *                                    per bit: Predictor::Predict/Perceive/Learn predictor.cpp:360-387)
 *   gmx_compress_trace              same loop, additionally exporting what Predictor::Predict returns per bit
 *   gmx_model_load, gmx_*_from      Predictor::ReadCheckpoint predictor.cpp:406-420 (+ `gmix -c/-d <ckpt>` runner-utils.cpp:110-116)
 *   gmx_generate_batch / _device    runner_utils::RunGeneration runner/runner-utils.cpp:158-221
 *   gmx_train_checkpoint,           Predictor::WriteCheckpoint predictor.cpp:389-404
 *   gmx_pred_write/read_checkpoint
 *
 * A per-bit host<->device call is a non-starter (2.1e9 bit steps in the 4096 x 64 KiB config), so
 * the ABI is stream-batch granular: n independent streams, each compressed from scratch exactly
 * like one `gmix -c` process would, one CTA per stream. Stream framing is the reference's:
 * 5-byte big-endian length, then the arithmetic-coder bytes.
 *
 * Conventions: return 0 on success, negative gmx error code otherwise (never throws); the caller
 * owns every buffer passed in; the library owns its device memory; a gmx_ctx is bound to one GPU
 * and must be used from one host thread at a time (one ctx per GPU for multi-GPU). There is NO CPU
 * fallback: without a usable CUDA device gmx_create fails.
 */
#ifndef GMIX_B200_H_
#define GMIX_B200_H_
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct gmx_ctx gmx_ctx;

enum {
  GMX_E_OK = 0,
  GMX_E_CUDA = -1,        /* CUDA runtime error, see gmx_last_error */
  GMX_E_ARG = -2,         /* invalid argument */
  GMX_E_NOMEM = -3,       /* not enough device memory for even one stream arena */
  GMX_E_STREAM = -4,      /* at least one stream failed; per-stream codes in status[] */
  GMX_E_NODEVICE = -5     /* no CUDA device / wrong architecture */
};

/* Per-stream status codes written to status[] (0 = ok). */
enum {
  GMX_S_OK = 0, GMX_S_PPMD_ARENA = 1, GMX_S_MIXER_POOL = 2, GMX_S_OUTPUT_CAP = 3,
  GMX_S_MATCH_RANGE = 4, GMX_S_HISTORY_CAP = 5, GMX_S_BAD_HEADER = 6, GMX_S_SPARSE_FULL = 7, GMX_S_INTERNAL = 8
};

const char* gmx_version(void);
/* Library-level error text for failures that happen before a ctx exists. */
const char* gmx_global_error(void);

int gmx_create(int device, gmx_ctx** out);
void gmx_destroy(gmx_ctx* ctx);
const char* gmx_last_error(const gmx_ctx* ctx);

/* Kernel configuration = the role split of the stream CTA (warps of bit role / LSTM role + one PPMd warp, resident CTAs
 * per SM). Every configuration produces the same bytes; they differ in throughput per workload shape. Default -1: chosen
 * per call (at most one stream per SM -> the latency configuration: pipelined roles, LSTM gate weights resident in shared
 * memory, one CTA per SM; more streams -> the throughput configuration: 8 CTAs per SM). A value >= 0 (or the environment
 * variable GMIX_B200_KERNEL_CONFIG) pins one. Changing it frees the arenas (the next call re-sizes them). */
int gmx_set_kernel_config(gmx_ctx* ctx, int cfg);
int gmx_get_kernel_config(const gmx_ctx* ctx);
int gmx_kernel_config_count(void);
int gmx_kernel_config_info(int cfg, int* bit_warps, int* lstm_warps, int* ctas_per_sm, int* serial, int* resident_weights);

/* Launch kernels on an existing CUDA stream (cudaStream_t passed as void*), e.g. torch's current
 * stream, so the caller can bracket calls with its own CUDA events. NULL = the ctx's own stream. */
int gmx_set_cuda_stream(gmx_ctx* ctx, void* cuda_stream);

/* Size the per-CTA stream arenas: max_stream_len = longest uncompressed stream in bytes,
 * max_resident = upper bound on concurrently resident streams (0 = as many as fit, at most one
 * wave of co-resident CTAs). Called implicitly by the batch calls when needed.
 * Arenas are sized for what text-like data touches; a stream that needs more (incompressible data)
 * is transparently re-run from scratch in a worst-case-sized arena by the same batch call, so
 * statuses 1, 2 and 7 only surface when even that does not fit the GPU.
 * Limit: a stream's bit-step counter (trained bits of the checkpoint it starts from + 8 * its own bytes) is
 * 32-bit on the device, so trained bytes + stream bytes must stay below 512 MiB; larger requests are
 * rejected with GMX_E_ARG here, in gmx_model_load and in gmx_pred_new. */
int gmx_configure(gmx_ctx* ctx, uint64_t max_stream_len, uint32_t max_resident);

/* Worst-case compressed size of an n-byte stream (header + coder bytes). */
uint64_t gmx_compress_bound(uint64_t n);

/* Compress n_streams independent streams. Stream i is in[in_off[i] .. in_off[i+1]); its output goes
 * to out[out_off[i] ..] with capacity out_off[i+1]-out_off[i]; out_len[i] receives its length and
 * status[i] its per-stream code. Host-pointer variant: copies in, runs, copies out. */
int gmx_compress_batch(gmx_ctx* ctx, const uint8_t* in, const uint64_t* in_off, uint32_t n_streams,
                       uint8_t* out, const uint64_t* out_off, uint64_t* out_len, uint32_t* status);
int gmx_decompress_batch(gmx_ctx* ctx, const uint8_t* in, const uint64_t* in_off, uint32_t n_streams,
                         uint8_t* out, const uint64_t* out_off, uint64_t* out_len, uint32_t* status);
/* Device-pointer variants: every pointer is a device pointer on the ctx's GPU; max_stream_len is the
 * longest uncompressed stream (for arena sizing). Enqueued on the ctx stream and synchronised
 * before returning. */
int gmx_compress_batch_device(gmx_ctx* ctx, const uint8_t* d_in, const uint64_t* d_in_off, uint32_t n_streams,
                              uint8_t* d_out, const uint64_t* d_out_off, uint64_t* d_out_len, uint32_t* d_status,
                              uint64_t max_stream_len);
int gmx_decompress_batch_device(gmx_ctx* ctx, const uint8_t* d_in, const uint64_t* d_in_off, uint32_t n_streams,
                                uint8_t* d_out, const uint64_t* d_out_off, uint64_t* d_out_len, uint32_t* d_status,
                                uint64_t max_stream_len);

/* ---- Checkpoints (reference src/memory: `<path>.short` + `<path>.long`) and generation --------------------
 * gmx_model_load replaces Predictor::ReadCheckpoint (predictor.cpp:406-420): it parses the two files exactly as the
 * reference writes them (every Model::WriteToDisk in construction order + ShortTermMemory::WriteToDisk
 * memory/short-term-memory.cpp:3-59; LongTermMemory::WriteToDisk memory/long-term-memory.cpp:6-108) and parks the
 * resulting stream on the device. The *_from calls then start every stream as a clone of it: `gmix -c <ckpt> in out`,
 * `gmix -d <ckpt> in out` (runner-utils.cpp:88-156). max_new_bytes bounds the bytes any stream adds on top (stream
 * length, or prompt + generated bytes); roomy != 0 sizes the tables for worst-case instead of text-like data.
 * Only byte-boundary checkpoints exist in practice (the reference writes them after whole files) and only those load. */
typedef struct gmx_model gmx_model;
int gmx_model_load(gmx_ctx* ctx, const void* short_blob, uint64_t short_len, const void* long_blob, uint64_t long_len,
                   uint64_t max_new_bytes, int roomy, gmx_model** out);
void gmx_model_free(gmx_model* model);
uint64_t gmx_model_arena_bytes(const gmx_model* model);     /* bytes of one stream arena cloned from this model */
uint64_t gmx_model_trained_bytes(const gmx_model* model);   /* bytes the checkpoint had learned (Mixer::steps_ / 8) */
int gmx_compress_batch_from(gmx_ctx* ctx, const gmx_model* model, const uint8_t* in, const uint64_t* in_off, uint32_t n_streams,
                            uint8_t* out, const uint64_t* out_off, uint64_t* out_len, uint32_t* status);
int gmx_decompress_batch_from(gmx_ctx* ctx, const gmx_model* model, const uint8_t* in, const uint64_t* in_off, uint32_t n_streams,
                              uint8_t* out, const uint64_t* out_off, uint64_t* out_len, uint32_t* status);

/* runner_utils::RunGeneration (runner-utils.cpp:158-221) for n prompts at once, one CTA per prompt: prompt i =
 * prompts[prompt_off[i] .. prompt_off[i+1]) is consumed WITH learning except its last byte, then out_bytes bytes are
 * sampled without Learn into out[i * out_bytes ..]. rand_u holds the rand()/RAND_MAX draws, one per generated bit;
 * stream i reads rand_u[i * rand_stride + k] (rand_stride 0: all streams share one sequence, which is what n separate
 * `gmix -g` processes do). temperature is clamped to >= 0.001 as the reference does. */
int gmx_generate_batch(gmx_ctx* ctx, const gmx_model* model, const uint8_t* prompts, const uint64_t* prompt_off, uint32_t n,
                       uint32_t out_bytes, float temperature, const float* rand_u, uint64_t rand_stride, uint8_t* out,
                       uint32_t* status);
int gmx_generate_batch_device(gmx_ctx* ctx, const gmx_model* model, const uint8_t* d_prompts, const uint64_t* d_prompt_off, uint32_t n,
                              uint32_t out_bytes, float temperature, const float* d_rand_u, uint64_t rand_stride, uint8_t* d_out,
                              uint64_t* d_out_len, uint32_t* d_status, uint64_t max_prompt_len);
/* Predictor::EnableAnalysis(sample_frequency) + RunAnalysis (predictor.cpp:422-504) for ONE compressed stream: the stream is
 * compressed as gmx_compress_batch does (analysis on: inactive predictions read as 0.5, predictor.cpp:362-365) and every
 * sample_frequency bits one row is recorded - what the reference appends to analysis/entropy.tsv and analysis/memory.tsv.
 * Columns: mod_ppmd(20), LSTM, the 15 skip-context Indirect models' "-indirect" and "-run_map" predictions, Mixer(final layer)
 * (the models the reference constructs with enable_analysis = true). neg_entropy = the value the reference prints
 * (-entropy, exponential average with alpha 1e-5 from a start value of -1); ppmd_used = ppmd_Model::GetUsedMemory(); history =
 * LongTermMemory::history.size(). The other memory.tsv columns are constants of the model graph (gmixb200 -c writes both files). */
enum { GMX_ANALYSIS_COLUMNS = 33 };
typedef struct gmx_analysis_row { uint64_t bits_seen; double neg_entropy[GMX_ANALYSIS_COLUMNS]; uint64_t ppmd_used, history; } gmx_analysis_row;
int gmx_compress_analysis(gmx_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_len,
                          uint32_t sample_frequency, gmx_analysis_row* rows, uint32_t max_rows, uint32_t* n_rows);

/* How gmx_generate_batch[_device] runs the sampling phase.
 *   GMX_GEN_PER_STREAM      (default) one persistent CTA per prompt does everything (RunGeneration as written, per stream).
 *   GMX_GEN_LOCKSTEP_EXACT  all streams advance one sampled byte per launch; the LSTM gate products of a byte step
 *                           (LstmLayer::ForwardPass lstm-layer.cpp:198-204) are ONE batched kernel over all streams with the
 *                           reference's arithmetic - the same bytes as GMX_GEN_PER_STREAM and as `gmix -g`.
 *   GMX_GEN_LOCKSTEP_TENSOR the batched product runs on the tensor cores (tcgen05.mma kind::tf32, 3xTF32 split, TMEM
 *                           accumulators, TMA bulk operand loads): fp32-accurate but not the reference's summation order, so
 *                           sampled bytes can differ from `gmix -g` (opt-in; divergence reported by bench.py / the tests).
 * The lock-step modes need the gate matrix to be shared by all streams, i.e. no stream may reach an LSTM BPTT pass while it
 * learns its prompt: (bytes the model learned since its last pass) + longest prompt - 1 < 100 (lstm.cpp:57-79). Otherwise
 * the call runs GMX_GEN_PER_STREAM; gmx_last_generation_mode tells which mode the last call used. */
enum { GMX_GEN_PER_STREAM = 0, GMX_GEN_LOCKSTEP_EXACT = 1, GMX_GEN_LOCKSTEP_TENSOR = 2 };
int gmx_set_generation_mode(gmx_ctx* ctx, int mode);
int gmx_last_generation_mode(const gmx_ctx* ctx);
/* The draws one `gmix -g` process makes for sampling: srand(0xDEADBEEF) (predictor.cpp:18), 84450 draws consumed by
 * the LSTM initialisation (lstm-layer.cpp:176-195), then n x rand()/RAND_MAX. Uses and reseeds the host libc rand(). */
void gmx_reference_rand_u(float* out, uint64_t n);

/* Predict/Perceive/Learn over data[0..n) (analysis off, as RunTraining and a plain Predictor loop do), starting from
 * `from` or from scratch, then Predictor::WriteCheckpoint (predictor.cpp:389-404): *short_blob / *long_blob point at
 * library-owned buffers valid until the next call on ctx. The `.long` bytes equal the reference's; `.short` differs
 * only in fields the reference rebuilds before reading (gmix_b200/csrc/checkpoint.h). */
int gmx_train_checkpoint(gmx_ctx* ctx, const gmx_model* from, const uint8_t* data, uint64_t n, const void** short_blob,
                         uint64_t* short_len, const void** long_blob, uint64_t* long_len);

/* FNV-1a 64 checksum of every stream slice d_data[d_off[i] .. d_off[i]+d_len[i]) into d_sum[i] (device
 * pointers; enqueued on the ctx stream, not synchronised). Used to gather {size, checksum} per stream
 * across GPUs without moving the payload. */
int gmx_checksum_device(gmx_ctx* ctx, const uint8_t* d_data, const uint64_t* d_off, const uint64_t* d_len,
                        uint32_t n_streams, uint64_t* d_sum);

/* Single-stream compress that also returns, per input bit, what Predictor::Predict returned
 * (probs[8n]) and the coder's 16-bit probability (p16[8n]); blackboard[8n*126] (optional, may be
 * NULL) receives the 90 stretched predictions, 3 active-mask words and 24+8+1 mixer outputs. */
int gmx_compress_trace(gmx_ctx* ctx, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_len,
                       float* probs, uint32_t* p16, float* blackboard);

/* ---- One stream coded in parts: Encoder/Decoder::WriteCheckpoint + ReadCheckpoint (coder/encoder.cpp:36-51,
 * coder/decoder.cpp:41-57) together with Predictor::Write/ReadCheckpoint, i.e. what the reference's restart tests do
 * (runner/tester.cpp:24-110, 182-321) at batch speed. A part is a whole number of bytes. `from` = the model the part
 * starts from (NULL = a fresh Predictor; load it with max_new_bytes >= the part's bytes + 8); coder_in = the coder state
 * the part starts with (NULL = a fresh Encoder / a Decoder that reads the 5-byte header and its first 4 bytes from `in`).
 * After the part: coder_out = what Encoder/Decoder::WriteCheckpoint would write, and - when short_blob != NULL - the
 * predictor checkpoint (valid until the next checkpoint-producing call on the ctx). analysis: -1 = as the runner would
 * for a stream of total_len bytes, 0 / 1 = forced (the reference's tester never enables it).
 * gmx_compress_part: write_header != 0 puts the 5-byte header of a total_len-byte stream first; last != 0 ends with
 * Encoder::Flush. gmx_decompress_part: produces exactly out_bytes bytes; in_consumed = coded bytes the decoder has
 * taken from `in` (the next part's `in` starts there). */
typedef struct { uint32_t x1, x2, x; } gmx_coder_state;
int gmx_compress_part(gmx_ctx* ctx, const gmx_model* from, const gmx_coder_state* coder_in, int write_header, uint64_t total_len, int last,
                      int analysis, const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_len, gmx_coder_state* coder_out,
                      const void** short_blob, uint64_t* short_len, const void** long_blob, uint64_t* long_len,
                      float* probs /* optional: 8 * n values, what Predictor::Predict returned for every bit of the part */);
int gmx_decompress_part(gmx_ctx* ctx, const gmx_model* from, const gmx_coder_state* coder_in, int analysis, const uint8_t* in, uint64_t n_in,
                        uint64_t out_bytes, uint8_t* out, uint64_t* in_consumed, gmx_coder_state* coder_out,
                        const void** short_blob, uint64_t* short_len, const void** long_blob, uint64_t* long_len);

/* ---- Predictor facade: one stream stepped bit by bit -------------------------------------------
 * Mirrors `class Predictor` (reference src/predictor.h:20-38): Predict() -> Perceive(bit) -> Learn(), Learn
 * optional (generation). Every Predict/Learn is one kernel launch plus a device->host read, so this is
 * the compatibility / debugging path, not the fast one (use the batch calls). The stream arena is
 * worst-case sized for max_stream_len bytes. gmx_pred_enable_analysis mirrors the one path-visible
 * effect of Predictor::EnableAnalysis (predictor.cpp:362-365, predictions zeroed each Predict), which
 * runner_utils::Compress turns on for inputs of 125 bytes or more (runner-utils.cpp:47). */
typedef struct gmx_pred gmx_pred;
int gmx_pred_new(gmx_ctx* ctx, uint64_t max_stream_len, gmx_pred** out);
void gmx_pred_free(gmx_pred* pred);
int gmx_pred_enable_analysis(gmx_pred* pred, int on);
int gmx_pred_predict(gmx_pred* pred, float* prob);
int gmx_pred_perceive(gmx_pred* pred, int bit);
int gmx_pred_learn(gmx_pred* pred);
/* Predictor::Copy (predictor.cpp:42-48): dst becomes a deep copy of src (same ctx, same max_stream_len). */
int gmx_pred_copy(gmx_pred* dst, const gmx_pred* src);

/* Predictor::WriteCheckpoint / ReadCheckpoint (predictor.cpp:389-420) of the stepped stream, at a byte boundary. */
int gmx_pred_write_checkpoint(gmx_pred* pred, const void** short_blob, uint64_t* short_len, const void** long_blob, uint64_t* long_len);
int gmx_pred_read_checkpoint(gmx_pred* pred, const void* short_blob, uint64_t short_len, const void* long_blob, uint64_t long_len);

/* Introspection for benchmarks. */
uint32_t gmx_resident_streams(const gmx_ctx* ctx);   /* CTAs (= arenas) the last launch used */
uint32_t gmx_arena_count(const gmx_ctx* ctx);        /* stream arenas currently allocated (= most streams resident at once) */
uint64_t gmx_arena_bytes(const gmx_ctx* ctx);        /* bytes of one stream arena */
uint64_t gmx_retried_streams(const gmx_ctx* ctx);    /* streams re-run in a worst-case arena so far (see gmx_configure) */
uint64_t gmx_kernel_launches(const gmx_ctx* ctx);    /* kernels launched by this ctx so far */
double gmx_last_kernel_ms(const gmx_ctx* ctx);       /* device time of the last stream kernel (CUDA events) */
int gmx_device_sm_count(const gmx_ctx* ctx);

/* Phase profiler: when on, every stream accumulates 24 cycle counters (slots documented in
 * gmix_b200/csrc/stream_kernel.cuh); gmx_get_profile copies up to max_streams x 32 counters of the
 * last launch and returns the number of streams copied. */
int gmx_set_profile(gmx_ctx* ctx, int on);
int gmx_get_profile(gmx_ctx* ctx, uint64_t* out, uint32_t max_streams);

/* Per-stream counters of the last batch call: 8 words per stream {entries in the shared sparse table,
 * mixer weight sets, PPMd unit bytes, Match history bytes, SM id, start and end time in us (device
 * globaltimer, low 32 bits), 0}; returns the number of streams copied. */
int gmx_get_usage(gmx_ctx* ctx, uint32_t* out, uint32_t max_streams);

/* Exhaustive check of the device expf/logf/tanhf against the host libm: inputs are the bit
 * patterns 0, stride, 2*stride, ... < 2^32 (logf: positive normals only). mismatches[3] and
 * first_bad[3] are indexed expf, logf, tanhf. */
int gmx_selftest_math(gmx_ctx* ctx, uint32_t stride, uint64_t mismatches[3], uint32_t first_bad[3]);
/* Self test of the batched gate product kernels on seeded random operands (n_slots streams): exact_mismatches = entries of
 * the exact kernel that differ bitwise from a host loop in the reference's order; err[0] = largest |tensor-core result - fp64
 * sum|, err[1] = largest |sequential fp32 - fp64 sum| (the reference arithmetic's own rounding, for scale), err[2] = largest
 * |result|. */
int gmx_selftest_gate(gmx_ctx* ctx, uint32_t n_slots, uint32_t seed, uint64_t* exact_mismatches, double err[3]);

#ifdef __cplusplus
}
#endif
#endif /* GMIX_B200_H_ */
