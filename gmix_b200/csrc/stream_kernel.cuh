// One CTA per stream, three concurrent ROLES per CTA: the complete gmix bit step (reference src/predictor.cpp:360-387
// and every Model::Predict/Learn it calls) plus the binary arithmetic coder (src/coder/*.cpp), for compress, decompress
// and generation. Persistent CTAs pull stream ids from an atomic queue and reuse their arena (global-memory workspace).
//
// Roles (warp-specialised, software-pipelined inside the stream):
//   PPMd role  (1 warp)    byte b: ModPPMD update with byte b-1, symbol distribution of byte b     (mod_ppmd.cpp)
//   LSTM role  (WL warps)  byte b: forward pass on that distribution, then Lstm::Perceive(byte b)    (lstm*.cpp)
//   bit role   (WB warps)  byte b: contexts, 8 x { lookups, mixer network, coder, Learn }            (everything else)
// In compress every byte is known in advance and both byte models are functions of the input prefix only, so the PPMd
// and LSTM roles run AHEAD of the bit role (up to PKT_RING bytes) and hand it one small packet per byte: the eight
// interval-node predictions of each byte model along the known byte's path + the LSTM's argmax context. Decompress and
// generation learn each byte only when its last bit is decided: there the same roles run in LOCKSTEP (the byte models
// start on byte b the moment it is known, overlapping the bit role's Learn of its last bit).
// Roles synchronise internally with named barriers (bar.sync id, count) and with each other through monotonic
// counters in shared memory (release/acquire fences, nanosleep back-off); nothing in a role ever waits on a CTA-wide
// barrier while a stream is running.
//
// Exactness rules: every fp32 operation goes through gmx::f_* (single IEEE
// rounding, never contracted), every dot product is accumulated by ONE thread in the reference's index order, libm calls
// go through gmx::gm_* (dmath.cuh). Parallelism inside a stream is only across independent quantities (the 41 Indirect
// models, the 24 layer-0 mixers, the 150 LSTM gate rows, the 2697 mixer weights, ...) and across the three roles.
//
// The file is plain CUDA C++ that also compiles for the host under tests/emu/cuda_emu.h (a fiber based SIMT emulator
// used by the CPU test-suite to debug the kernel logic without a GPU). The shipped library only contains the nvcc build.
#ifndef GMIX_B200_STREAM_KERNEL_CUH_
#define GMIX_B200_STREAM_KERNEL_CUH_
#include <stddef.h>
#include <stdint.h>

#include "dmath.cuh"
#include "ppmd.cuh"
#include "spec.cuh"

namespace gmx {

enum : uint32_t {
  GMX_OK = 0,
  GMX_ERR_PPMD_ARENA = 1,    // backed part of the PPMd heap exhausted
  GMX_ERR_MIXER_POOL = 2,    // mixer weight-set pool exhausted
  GMX_ERR_OUTPUT_CAP = 3,    // output slice too small
  GMX_ERR_MATCH_RANGE = 4,   // Match pointer outside history (reference would throw, match.cpp:55)
  GMX_ERR_HISTORY_CAP = 5,
  GMX_ERR_BAD_HEADER = 6,
  GMX_ERR_SPARSE_FULL = 7,   // shared sparse table over its load limit (host retries with a roomier arena)
  GMX_ERR_INTERNAL = 8,      // a hand-over between kernels of the lock-step generation happened off a byte boundary
};

// Shared-memory staging of the 33 selected weight sets: layer-0 sets (<= 113 weights) at stride 132 words,
// layer-1/final sets (<= 33 weights) at stride 36 words. Both strides are 16-byte multiples (float4
// loads) and = 4 mod 32 banks, which makes lane-per-neuron float4 reads conflict free.
enum : int { WSTRIDE0 = 132, WSTRIDE1 = 36, WTOTAL = 24 * WSTRIDE0 + 9 * WSTRIDE1 };
#define GMX_PRAGMA_(x) _Pragma(#x)
#define GMX_UNROLL(n) GMX_PRAGMA_(unroll n)
#ifndef GMX_BPTT_UNROLL
#define GMX_BPTT_UNROLL 4
#endif
// Packets the byte-model roles may run ahead of the bit role in compress. The LSTM role stops for a whole BPTT pass
// (~25 byte times of the bit role) every 100 bytes; the ring lets the bit role keep going through it.
#ifndef GMX_PKT_RING
#define GMX_PKT_RING 32
#endif
enum : int { PKT_RING = GMX_PKT_RING };

// Cycle-counter slots of the PROF kernel variant. Bit role: 0 bookkeeping, 1 wait for the byte packet, 2 byte-boundary
// contexts, 3 lookups, 4 gate selection, 5 weight-set swap, 6 layer-0 dot products, 7 layer-0 chain, 8 layer 1, 9 final
// neuron, 10 rest of predict (barrier), 11 learn scalars + tables + coder, 12 weight update, 13 trace, 14 stream init.
// LSTM role: 16 wait, 17 forward gate products, 18 rest of forward, 19 output-layer step, 20 BPTT epochs, 21 BPTT
// gradients + Adam. PPMd role: 24 wait, 25 model update + distribution, 26 normalise + path nodes.
enum : int { GMX_PROF_SLOTS = 32 };
#if defined(__CUDA_ARCH__)
#define GMX_CLOCK() clock64()
#else
#define GMX_CLOCK() 0ll
#endif

// Byte offsets (from the arena base) of every per-stream table. Filled by the host (layout.h).
struct ArenaLayout {
  // Big tables (Indirect 2^15/2^16, Match 2^21/2^24, IndirectHash 2^24) can live in ONE shared sparse
  // open-addressing map instead of dense arrays: sid = table id inside the map (1..63), 0 = dense.
  uint64_t sparse;            // u64 entries {key:31 = sid<<25 | index, value:32}; 0 = empty
  uint32_t sparse_mask;       // capacity - 1 (power of two), 0 = no sparse map
  uint32_t sparse_limit;      // maximum number of entries (load limit)
  uint8_t ind_sid[NIND], match_sid[NMATCH], ih_sid[NIH];
  uint64_t ind_tab[NIND];     // u16 {ns | rm<<8} per slot, (2^log2*256+1) slots (indirect.cpp:15-19)
  uint32_t ind_size[NIND];
  uint64_t ind_pred;          // float [NIND][2][256]
  uint64_t match_tab[NMATCH]; // u32 history pointers (5-byte pointers in the reference, match.cpp:48-49)
  uint64_t match_pred;        // float [NMATCH][256]
  uint64_t match_cnt;         // int   [NMATCH][256]
  uint64_t history; uint64_t history_cap;
  uint64_t ih_tab[NIH];       // u32
  uint64_t mix_dir[NMIX];     // u32 pool id per gate context (0 = no weight set yet)
  uint64_t mix_pool; uint32_t mix_pool_sets; uint32_t mix_set_stride;  // floats per set record
  // LSTM
  uint64_t l_w, l_m, l_v;     // float [3][L_ROWQ][L_CELLS][4], element (gate, column, cell) at LstmW() (spec.cuh)
  uint64_t l_gb;              // float [8][3][L_CELLS]: gamma, beta, gamma_m, gamma_v, beta_m, beta_v, gamma_u, beta_u
  uint64_t l_wout;            // float [L_HORIZON][L_HID][L_NOUT]
  uint64_t l_lin;             // float [L_HORIZON][L_NIN + 1]
  uint64_t l_out;             // float [L_HORIZON][L_NOUT]
  uint64_t l_gstate, l_norm;  // float [3][L_HORIZON][L_CELLS]
  uint64_t l_ivar;            // float [3][L_HORIZON]
  uint64_t l_tanh, l_ig, l_last;  // float [L_HORIZON][L_CELLS]
  uint64_t l_errh;            // float [3][L_HORIZON][L_CELLS]
  uint64_t l_wt;              // float [3][L_CELLS][L_CELLS] transposed recurrent weights (BPTT scratch)
  // PPMd
  uint64_t p_state; uint64_t p_heap; uint32_t p_mask; uint32_t p_text_cap; uint32_t p_units_cap;  // see ppmd.cuh
  uint32_t p_seg_lo, p_lo_cap, p_hi_cap;   // p_mask == 0: segmented backing (ppmd.cuh), overlay arenas only
  uint64_t total;             // arena bytes
  // Overlay mode (ov != 0; streams that start from a loaded model WITHOUT cloning its tables: batched generation). The
  // big tables then stay in the model's arena, read-only and shared by every stream; this arena holds what the stream
  // changes: every table (all of them have a sid here, mixer directories use sid OV_SID_MIX + m) is written to the
  // shared sparse map of THIS arena and looked up there first, weight sets the stream creates or changes live in a
  // local pool (ids >= base_sets), history bytes it appends behind base_hist. Small / dense-updated state (logit maps,
  // Match predictions, LSTM, PPMd) is a private copy made at stream start.
  uint32_t ov, base_sets, base_hist, ov_pad;
};
// Overlay mode exists in the generation kernels only (kernel_generate.cu defines GMX_OVERLAY 1 before this header; the
// CPU emulator too): everywhere else the checks below are compile-time false, so that the compress / decompress kernels
// carry neither the branches nor the code behind them (their per-bit path is instruction-cache and register bound).
#ifndef GMX_OVERLAY
#define GMX_OVERLAY 0
#endif
#define GMX_IS_OV(L) (GMX_OVERLAY && (L).ov)
enum : uint32_t { OV_SID_IND = 1, OV_SID_MATCH = OV_SID_IND + NIND, OV_SID_IH = OV_SID_MATCH + NMATCH, OV_SID_MIX = OV_SID_IH + NIH };
static_assert(OV_SID_MIX + NMIX <= 128, "overlay table ids must fit the 7 bits above the 25-bit index of a sparse key");

// One sample of analysis/entropy.tsv + analysis/memory.tsv (predictor.cpp:476-503): the columns are mod_ppmd(20), LSTM, the 15
// skip-context Indirect models' two predictions each, the final mixer (the models constructed with enable_analysis = true,
// predictor.cpp:28-29,124,353); the memory figures that change over a stream are PPMd's GetUsedMemory and the match history.
enum : int { AN_COLS = 33 };
struct AnalysisRow { uint64_t bits_seen; double neg_entropy[AN_COLS]; uint64_t ppmd_used, history; };
struct StreamParams {
  const uint8_t* in; const uint64_t* in_off;     // n_streams + 1 offsets
  uint8_t* out; const uint64_t* out_off;         // n_streams + 1 offsets (capacity slices)
  uint64_t* out_len; uint32_t* status;           // per stream
  uint32_t n_streams; uint32_t* queue;           // atomic stream counter
  const uint32_t* ids;                           // optional: queue position -> stream id (retry launches), or null
  uint8_t* arenas; uint64_t arena_stride;
  const ArenaLayout* layout;
  const float* lstm_init;    // [L_WSIZE] initial gate weights in the arena layout (host glibc rand(), lstm-layer.cpp:176-195)
  const float* decay;        // decay[s] = (float)(0.9 / pow(1e-7*s + 0.8, 0.8)) (mixer.cpp:111), host libm
  uint32_t decay_len;
  const float* adam;         // [L_UPDATE_LIMIT + 1][4]: alpha, 1-b1^t, 1-b2^t (lstm-layer.cpp:16-33), host libm
  uint64_t* bit_trace;       // optional: {f32 prob, u32 p16} per bit of stream 0 (debug/parity), or null
  float* pred_trace;         // optional: 90 predictions + 3 mask words + 33 mixer outs per bit of stream 0
  uint32_t* usage;           // optional: 8 words per stream {sparse entries, mixer sets, PPMd unit bytes, history bytes,
                             // SM id, start us, end us (globaltimer, low 32 bits), 0}, or null
  unsigned long long* prof;  // optional: GMX_PROF_SLOTS cycle counters per stream (phase breakdown), or null
  // Start every stream from a parked stream (a loaded checkpoint: Predictor::ReadCheckpoint predictor.cpp:406-420)
  // instead of from scratch: arena image of layout->total bytes + StreamSmem image. null = from scratch.
  const uint8_t* tmpl_arena; const uint32_t* tmpl_state;
  const ArenaLayout* tmpl_layout;   // layout of tmpl_arena when it differs from `layout` (overlay mode, ArenaLayout::ov), else null
  uint32_t* final_state;     // optional: the stream's StreamSmem is parked here at its end (n_streams x sizeof(StreamSmem)), or null
  int32_t analysis;          // -1: what the reference runner does for this mode; 0/1: forced (Predictor::EnableAnalysis)
  // generation (runner_utils::RunGeneration runner-utils.cpp:158-221): `in` holds the prompts, out[sid * gen_bytes ..] the samples
  uint32_t gen_bytes; float temperature;
  const float* rand_u; uint64_t rand_stride;   // rand()/RAND_MAX draws, one per generated bit; stream sid reads rand_u[sid * rand_stride + k]
  // One stream coded in PARTS (Encoder/Decoder::WriteCheckpoint + ReadCheckpoint, encoder.cpp:36-51, decoder.cpp:41-57):
  // single-stream launches only. part != 0 switches the framing off: no header is written / read unless part_header,
  // the coder starts from coder_in {x1, x2, x} when given, ends without Encoder::Flush unless part_last, and leaves its
  // state in coder_out {x1, x2, x, coded bytes consumed by the decoder}.
  uint32_t part, part_header, part_last;
  uint64_t part_total;         // compress: length the header announces; decompress: bytes this part produces
  const uint32_t* coder_in; uint32_t* coder_out;
  // Analysis output (Predictor::EnableAnalysis / RunAnalysis predictor.cpp:422-504) of a single compress stream run in the
  // phase-serial order: cross-entropy accumulators of the 33 models the reference enables analysis for (an_entropy, start
  // value -1) and one AnalysisRow per an_freq bits. null = off (only the reference's side effect on the inactive predictions,
  // StreamSmem::analysis, is mirrored).
  double* an_entropy; struct AnalysisRow* an_rows; uint32_t an_freq, an_max_rows;
  // LOCK-STEP batched generation (gate_gemm.cuh; host.cu RunLockstepGenerate). While no stream of a batch has reached a BPTT pass,
  // the 3 x 50 x 563 LSTM gate matrix is the loaded model's for ALL streams (lstm.cpp:57-79 changes it once per 100 learned
  // bytes), so the gate products of one byte step of all streams are ONE dense contraction [streams x 307] . [307 x 150].
  // lockstep != 0: MODE_GENERATE runs stream stream_base + blockIdx.x in arena blockIdx.x, consumes the prompt, stops in front
  // of the first sampled byte's gate product (input vector in gate_x, byte in front in gate_sym) and parks its state in `park`;
  // GenStepKernel then advances every stream by one sampled byte per launch, reading the gate pre-activations the batched
  // kernel left in gate_g.
  uint32_t lockstep, stream_base;
  float* gate_x;               // GateXIndex(slot, k, plane): tf32-truncated value / exact residual planes, tiled for the MMA
  uint32_t* gate_sym;          // [slots] byte in front of the boundary (the one-hot column of the gate matrices)
  const float* gate_g;         // [slots][GG_N] gate pre-activations of this byte step
  uint32_t* park;              // [slots][sizeof(StreamSmem) / 4]
};

// ---- layouts of the batched gate product (gate_gemm.cuh) --------------------------------------------------------------
// M = 128 streams per tile, N = 160 (150 gate rows, padded), K = 320 (256 PPMd probabilities, 50 hidden, bias, padded) in chunks
// of 32. Both operands are K-major and stored in global memory as the shared-memory image tcgen05.mma reads without swizzle:
// 8-row x 16-byte core matrices (128 contiguous bytes), row groups adjacent (SBO = 128 B), the eight 4-float k-slices of a chunk
// behind each other (LBO = rows / 8 * 128 B). One (tile, chunk, plane) block is therefore one contiguous TMA bulk copy.
enum : int { GG_M = 128, GG_N = 160, GG_K = 320, GG_KC = 32, GG_NCHUNK = GG_K / GG_KC, GG_ROWS = 3 * L_CELLS, GG_KUSED = L_NIN };
enum : int { GG_A_FLOATS = GG_M * GG_KC, GG_B_FLOATS = GG_N * GG_KC };
GMX_HD size_t GateXIndex(uint32_t slot, int k, int plane) {
  const uint32_t tile = slot / GG_M, r = slot % GG_M;
  const int chunk = k / GG_KC, kk = k % GG_KC;
  return (((size_t)tile * GG_NCHUNK + chunk) * 2 + plane) * GG_A_FLOATS + (size_t)(((kk >> 2) * (GG_M / 8) + (r >> 3)) * 32 + (r & 7) * 4 + (kk & 3));
}
GMX_HD size_t GateWIndex(int row, int k, int plane) {
  const int chunk = k / GG_KC, kk = k % GG_KC;
  return ((size_t)chunk * 2 + plane) * GG_B_FLOATS + (size_t)(((kk >> 2) * (GG_N / 8) + (row >> 3)) * 32 + (row & 7) * 4 + (kk & 3));
}
GMX_HD size_t GateXFloats(uint32_t slots) { return (size_t)((slots + GG_M - 1) / GG_M) * GG_NCHUNK * 2 * GG_A_FLOATS; }
// tf32 keeps the top 19 bits of an fp32 word; the residual v - hi is exact in fp32
GMX_HD float Tf32Hi(float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xffffe000u; float r; memcpy(&r, &u, 4); return r; }

// ---- device constant tables --------------------------------------------------------------------
#if defined(__CUDACC__)
#define GMX_CONST_TABLE static __device__ const
#else
#define GMX_CONST_TABLE static const
#endif
GMX_CONST_TABLE IndirectSpec kInd[NIND] = {GMX_INDIRECT_SPECS};
GMX_CONST_TABLE SkipSpec kSkip[20] = {GMX_SKIP_SPECS};
GMX_CONST_TABLE IntervalSpec kInterval[9] = {GMX_INTERVAL_SPECS};
GMX_CONST_TABLE IHSpec kIH[NIH] = {GMX_IH_SPECS};
GMX_CONST_TABLE MatchSpec kMatch[NMATCH] = {GMX_MATCH_SPECS};
GMX_CONST_TABLE MixerSpec kMixer[NMIX] = {GMX_MIXER_SPECS};
GMX_CONST_TABLE uint8_t kNonstationary[512] = {
#include "nonstationary.inc"
};

// Read-only per launch: the arena layout and the model-graph tables, staged in shared memory because
// every lane indexes them with its own model number on the per-bit path.
struct StreamTables {
  ArenaLayout L;
  IndirectSpec ind[NIND]; SkipSpec skip[20]; IntervalSpec interval[9]; IHSpec ih[NIH]; MatchSpec match[NMATCH];
  MixerSpec mixer[NMIX];
  uint8_t nonstationary[512];
};

// One packet per byte boundary from the byte-model roles to the bit role.
struct BytePacket {
  float node[2][8];      // AHEAD only: Logit of the interval-node probability of bit j, [0] PPMd, [1] LSTM (IntervalNode)
  uint32_t flags;        // AHEAD only: 2 bits per node (bit 0: denom != 0, bit 1: p != 0.5), PPMd nodes in bits 0..15
  uint32_t lstm_ctx;     // lstm_prediction_context of this byte (lstm-model.cpp:26-33)
};

// ---- per-stream state staged in shared memory ---------------------------------------------------
// Everything behind T is the stream's short-term state: it is what StepKernel parks between launches and, with the
// arena, what checkpoint.h turns into the reference's `.short` file.
struct StreamSmem {
  StreamTables T;
  // blackboard (ShortTermMemory) -- bit role
  alignas(16) float preds[NPRED + 2];
  alignas(4) uint8_t act[NPRED + NL0 + 2];   // prediction i is active; entries 90.. (layer-0 outputs) are always 1
  uint32_t ctx[C_COUNT + 2];
  alignas(16) float l0_out[NL0]; float l1_out[NL1], final_out, prob;   // l0_out | l1_out contiguous (final mixer input)
  uint8_t ring[32];            // last bytes (the reference keeps 1000, short-term-memory.h:23; only 10 are ever read)
  uint32_t ring_pos;
  int32_t new_bit, recent_bits, bb, first_prediction, analysis;
  uint32_t error;                      // first error of the stream (any role, SetError); sticky
  uint32_t steps;                      // Mixer::steps_ (identical for all 33 mixers)
  // mixers -- bit role
  alignas(16) float w[WTOTAL];
  alignas(16) float xe[NPRED + NL0 + 2];    // layer-0 input vector: predictions (inactive ones zeroed) | layer-0 outputs
  uint32_t set_steps[NMIX], max_steps[NMIX], set_idx[NMIX], set_pool[NMIX];
  uint32_t swap_old[NMIX], swap_new[NMIX], swap_oldidx[NMIX], nswap;   // queued set swaps of this bit
  uint8_t swap_m[NMIX + 3], shrink[NMIX + 3];
  uint8_t set_dirty[NMIX + 3];               // the staged set has learned since it was staged (its pool record is stale)
  float upd[NMIX];
  uint32_t pool_next;
  // indirect -- bit role
  uint32_t ind_base[NIND], ind_slot[NIND];   // ind_slot: dense slot, or position in the sparse map
  uint16_t ind_state[NIND + 1];
  uint8_t ind_found[NIND + 3];               // sparse tables: entry exists at ind_slot
  float ind_pa[NIND], ind_pb[NIND];          // the two logit-map entries Learn will update, as read by Predict
  uint32_t sparse_used;
  uint32_t t_start_us;
  // match -- bit role
  uint32_t m_cur[NMATCH]; uint8_t m_byte[NMATCH], m_bitpos[NMATCH], m_len[NMATCH];
  uint32_t hist_len;
  // indirect hash -- bit role
  uint64_t ih_outer[NIH]; uint32_t ih_hash[NIH];
  // byte models. ppm: PPMd role -> LSTM role (and the bit role's interval nodes in lockstep modes); PrepareByte leaves
  // the integer pseudo-probabilities in the same words (sqp) and the PPMd role normalises them in place.
  union alignas(16) { uint32_t sqp[256]; float ppm[256]; };
  alignas(16) float lprob[256];            // LSTM role: softmax output of the last forward pass
  alignas(16) float l_err256[256];         // LSTM role scratch (output pre-activations, BPTT error vector, symbol lists)
  alignas(16) float l_hidden[L_HID + 1]; float l_state[L_CELLS], l_state_err[L_CELLS], l_stored_err[L_CELLS], l_hidden_err[L_CELLS];
  float l_gate[3][L_CELLS], l_gerr[3][L_CELLS];
  float l_red[40];               // 0..2 ivar, 3 sum, 4..19 / 20..35 per-warp partial results, 36..38 BPTT sums
  uint32_t p_masked[8];          // PPMd: bit sym = CharMask[sym] == EscCount (ppmd.cuh)
  uint8_t l_hist[L_HORIZON], l_symin[L_HORIZON];
  uint32_t l_epoch, l_update_steps, l_old_input, l_fused;
  // coder -- bit role
  uint32_t x1, x2, x;
  uint64_t out_pos, out_cap, in_pos, in_len;
  // ---- role pipeline (reset at every stream start; not part of a checkpoint) ----
  BytePacket pkt[PKT_RING];
  uint32_t n_ppm;          // byte boundaries whose PPMd distribution has been published in ppm
  uint32_t n_ppm_used;     // byte boundaries whose distribution the LSTM role no longer reads
  uint32_t n_pkt;          // byte boundaries whose packet (and lprob) is published
  uint32_t n_done;         // bytes the bit role has finished (frees packet slots)
  uint32_t bit_stop, lstm_stop;   // role-uniform copies of `error != 0`, refreshed at role-defined points
  uint32_t wphase;         // phase of the gate-weight mbarrier (latency configurations)
  uint32_t byte0;          // byte in front of the stream (0 from scratch: the reference's phantom first byte)
};

// Cycle counters of the PROF kernel variant (separate from StreamSmem so that the product kernel does not pay for them).
struct ProfSmem { unsigned long long acc[GMX_PROF_SLOTS]; };
template <bool PROF> struct ProfHolder { ProfSmem m; GMX_DEV ProfSmem* get() { return &m; } };
template <> struct ProfHolder<false> { GMX_DEV ProfSmem* get() { return nullptr; } };
// Lap timer of one role's elected thread.
template <bool PROF>
struct Lap {
  ProfSmem* p; long long t; bool on;
  GMX_DEV void start(ProfSmem* pp, bool elected) { p = pp; on = PROF && elected; t = on ? GMX_CLOCK() : 0; }
  GMX_DEV void mark(int slot) {
    if (PROF && on) { const long long n = GMX_CLOCK(); p->acc[slot] += (unsigned long long)(n - t); t = n; }
  }
};

struct SparseMap { unsigned long long* tab; uint32_t mask; };

struct Arena {
  uint8_t* base;
  const ArenaLayout* L;
  const uint8_t* base0 = nullptr;       // overlay mode: the model's arena and its layout (global memory, read-only)
  const ArenaLayout* BL = nullptr;
  template <typename T> GMX_DEV T* at(uint64_t off) const { return (T*)(base + off); }
  GMX_DEV SparseMap map() const { return SparseMap{(unsigned long long*)(base + L->sparse), L->sparse_mask}; }
};

// ---- small helpers -----------------------------------------------------------------------------
GMX_DEV inline uint32_t Rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
GMX_DEV inline uint32_t MurmurMix(uint32_t h, uint32_t k) {
  k *= 0xcc9e2d51u; k = Rotl32(k, 15); k *= 0x1b873593u;
  h ^= k; h = Rotl32(h, 13); return h * 5 + 0xe6546b64u;
}
GMX_DEV inline uint32_t MurmurFinal(uint32_t h, uint32_t len) {
  h ^= len; h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16; return h;
}
// MurmurHash3_x86_32 of a little-endian u64 / u32, seed 0xDEADBEEF (murmur-hash.cpp:94-146).
GMX_DEV inline uint32_t Murmur64(uint64_t v) {
  uint32_t h = 0xDEADBEEFu;
  h = MurmurMix(h, (uint32_t)v); h = MurmurMix(h, (uint32_t)(v >> 32));
  return MurmurFinal(h, 8);
}
GMX_DEV inline uint32_t Murmur32(uint32_t v) { return MurmurFinal(MurmurMix(0xDEADBEEFu, v), 4); }

// L2 eviction-priority hints (createpolicy): the gate weights of the resident streams alone are almost twice the L2, so
// they are marked evict-first; what the per-bit path re-reads (logit maps, sparse-map lines) can be marked evict-last.
GMX_DEV inline uint64_t PolicyEvictFirst() {
#if defined(__CUDA_ARCH__)
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
#else
  return 0;
#endif
}
GMX_DEV inline uint64_t PolicyEvictLast() {
#if defined(__CUDA_ARCH__)
  uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
#else
  return 0;
#endif
}
// ---- shared sparse map (exact: keyed by table id + full index, linear probing, never deletes) ----
GMX_DEV inline uint32_t SparseKey(uint32_t sid, uint32_t index) { return (sid << 25) | index; }
GMX_DEV inline uint32_t SparseHash(uint32_t k) {
  k ^= k >> 16; k *= 0x85ebca6bu; k ^= k >> 13; k *= 0xc2b2ae35u; k ^= k >> 16; return k;
}
// Returns the position of `key`, or of the first empty slot of its probe sequence; *entry = slot content
// (0 when absent).
GMX_DEV inline uint32_t SparseFind(const SparseMap& M, uint32_t key, unsigned long long* entry) {
  uint32_t pos = SparseHash(key) & M.mask;
  for (;;) {
    const unsigned long long e = M.tab[pos];
    if (e == 0ull || (uint32_t)(e >> 32) == key) { *entry = e; return pos; }
    pos = (pos + 1) & M.mask;
  }
}
// Store `value` for `key`. pos/found come from a SparseFind of the same key; other threads may have
// inserted OTHER keys since (never this one: every table is owned by one thread per phase).
GMX_DEV inline void SparsePut(const SparseMap& M, uint32_t* used, uint32_t limit, uint32_t* error, uint32_t key,
                              uint32_t pos, bool found, uint32_t value) {
  const unsigned long long e = ((unsigned long long)key << 32) | value;
  if (found) { M.tab[pos] = e; return; }
  if (atomicAdd(used, 1u) >= limit) { atomicCAS(error, 0u, (uint32_t)GMX_ERR_SPARSE_FULL); return; }
  for (;;) {
    const unsigned long long old = atomicCAS(&M.tab[pos], 0ull, e);
    if (old == 0ull) return;
    pos = (pos + 1) & M.mask;
  }
}
GMX_DEV inline uint32_t SparseGet(const SparseMap& M, uint32_t key) {  // value, 0 when absent
  unsigned long long e;
  SparseFind(M, key, &e);
  return (uint32_t)e;
}
GMX_DEV inline void SparseSet(const SparseMap& M, uint32_t* used, uint32_t limit, uint32_t* error, uint32_t key, uint32_t value) {
  unsigned long long e;
  const uint32_t pos = SparseFind(M, key, &e);
  SparsePut(M, used, limit, error, key, pos, e != 0ull, value);
}

GMX_DEV inline uint32_t RecentByte(const StreamSmem& s, int ago) {  // short-term-memory.cpp:215-219
  return s.ring[(s.ring_pos - (uint32_t)ago) & 31u];
}
GMX_DEV inline int WOff(int m) { return m < NL0 ? m * WSTRIDE0 : NL0 * WSTRIDE0 + (m - NL0) * WSTRIDE1; }
GMX_DEV inline int MixerNW(int m) { return m < NL0 ? NPRED + m : m < NL0 + NL1 ? NL0 + (m - NL0) + 1 : NL0 + NL1 + 1; }


GMX_DEV inline void SetError(StreamSmem& s, uint32_t code) { atomicCAS(&s.error, 0u, code); }

// ---- overlay mode: what the model's arena (base) holds for a table entry this stream has not written -------------
GMX_DEV inline uint32_t BaseInd(const Arena& A, int k, uint32_t slot) {   // Indirect state {ns | rm << 8}; never written = {255, 0}
  const ArenaLayout& B = *A.BL;
  if (B.ind_sid[k]) {
    unsigned long long e;
    SparseFind(SparseMap{(unsigned long long*)(A.base0 + B.sparse), B.sparse_mask}, SparseKey(B.ind_sid[k], slot), &e);
    return e ? (uint32_t)e & 0xffffu : 0x00ffu;
  }
  return ((const uint16_t*)(A.base0 + B.ind_tab[k]))[slot];
}
GMX_DEV inline uint32_t BaseMatch(const Arena& A, int k, uint32_t idx) {
  const ArenaLayout& B = *A.BL;
  if (B.match_sid[k]) return SparseGet(SparseMap{(unsigned long long*)(A.base0 + B.sparse), B.sparse_mask}, SparseKey(B.match_sid[k], idx));
  return ((const uint32_t*)(A.base0 + B.match_tab[k]))[idx];
}
GMX_DEV inline uint32_t BaseIH(const Arena& A, int k, uint32_t idx) {
  const ArenaLayout& B = *A.BL;
  if (B.ih_sid[k]) return SparseGet(SparseMap{(unsigned long long*)(A.base0 + B.sparse), B.sparse_mask}, SparseKey(B.ih_sid[k], idx));
  return ((const uint32_t*)(A.base0 + B.ih_tab[k]))[idx];
}
// Mixer directory: pool id of the weight set of mixer m's gate context idx (0 = none yet)
GMX_DEV inline uint32_t DirGet(const Arena& A, int m, uint32_t idx) {
  const ArenaLayout& L = *A.L;
  if (!GMX_IS_OV(L)) return A.at<uint32_t>(L.mix_dir[m])[idx];
  unsigned long long e;
  SparseFind(A.map(), SparseKey(OV_SID_MIX + m, idx), &e);
  return e ? (uint32_t)e : ((const uint32_t*)(A.base0 + A.BL->mix_dir[m]))[idx];
}
GMX_DEV inline void DirSet(StreamSmem& s, const Arena& A, int m, uint32_t idx, uint32_t id) {
  const ArenaLayout& L = *A.L;
  if (!GMX_IS_OV(L)) A.at<uint32_t>(L.mix_dir[m])[idx] = id;
  else SparseSet(A.map(), &s.sparse_used, L.sparse_limit, &s.error, SparseKey(OV_SID_MIX + m, idx), id);
}
// Pool record of weight set `id` (float4 units): the model's pool below base_sets in overlay mode (read-only)
GMX_DEV inline float4* PoolRec(const Arena& A, uint32_t id) {
  const ArenaLayout& L = *A.L;
  const uint32_t stride4 = L.mix_set_stride / 4;
  if (GMX_IS_OV(L) && id < L.base_sets) return (float4*)(A.base0 + A.BL->mix_pool) + (size_t)id * stride4;
  return A.at<float4>(L.mix_pool) + (size_t)(id - (GMX_IS_OV(L) ? L.base_sets : 0u)) * stride4;
}
GMX_DEV inline uint32_t HistByte(const Arena& A, uint32_t pos) {
  const ArenaLayout& L = *A.L;
  if (GMX_IS_OV(L)) return pos < L.base_hist ? (A.base0 + A.BL->history)[pos] : A.at<uint8_t>(L.history)[pos - L.base_hist];
  return A.at<uint8_t>(L.history)[pos];
}
// One staged weight set goes back to its pool record; all 32 lanes with the same arguments. `idx` = the gate context it
// belongs to. A set that has not learned since it was staged is skipped (its record is current). Overlay mode: a changed
// set of the model moves to a fresh local record and the overlay directory follows.
GMX_DEV inline void WriteBackSet(StreamSmem& s, const Arena& A, int m, uint32_t old, uint32_t idx, int lane) {
#if GMX_OVERLAY
  if (!old || !s.set_dirty[m]) return;
#else
  if (!old) return;   // (compress / decompress learn every bit: a staged set is always dirty, no bookkeeping there)
#endif
  const ArenaLayout& L = *A.L;
  uint32_t id = old;
  if (GMX_IS_OV(L) && old < L.base_sets) {
    if (lane == 0) {
      id = atomicAdd(&s.pool_next, 1u);
      if (id >= L.mix_pool_sets) { SetError(s, GMX_ERR_MIXER_POOL); id = 0; }
      else DirSet(s, A, m, idx, id);
    }
    id = __shfl_sync(0xffffffffu, id, 0);
    if (!id) return;
  }
  if (lane <= (MixerNW(m) + 3) / 4) {
    float4* rec = PoolRec(A, id);
    if (lane == 0) rec[0] = make_float4(u2f(s.set_steps[m]), 0.0f, 0.0f, 0.0f);
    else rec[lane] = ((const float4*)(s.w + WOff(m)))[lane - 1];
  }
}

// ---- role groups -----------------------------------------------------------------------------------
// A role is N consecutive threads (a multiple of 32); `id` is its named barrier (0 is left to __syncthreads).
template <int N>
GMX_DEV inline void GroupSync(int id) {
#if defined(__CUDA_ARCH__)
  if (N == 32) __syncwarp();
  else asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(N) : "memory");
#elif defined(GMX_EMU)
  if (N == 32) __syncwarp(); else cuda_emu::NamedBarrier(id, N);
#else
  (void)id;
#endif
}
enum : int { BAR_BIT = 1, BAR_LSTM = 2 };

GMX_DEV inline uint32_t VolatileLoad(const uint32_t* p) { return *(const volatile uint32_t*)p; }
GMX_DEV inline void FenceBlock() {
#if defined(__CUDA_ARCH__)
  __threadfence_block();
#endif
}
GMX_DEV inline void Backoff(unsigned ns) {
#if defined(__CUDA_ARCH__)
  __nanosleep(ns);
#elif defined(GMX_EMU)
  (void)ns; cuda_emu::Yield();
#else
  (void)ns;
#endif
}
// Inter-role hand-off. Publish: every write of the publishing thread (and, through the role barrier in front of it, of
// its role) is visible to a thread that has seen the counter. WaitAtLeast gives up when the stream has failed (the
// caller then runs on garbage until its role's next stop check; all addressing stays in bounds).
GMX_DEV inline void Publish(uint32_t* cnt, uint32_t v) { FenceBlock(); *(volatile uint32_t*)cnt = v; }
GMX_DEV inline void WaitAtLeast(const StreamSmem& s, const uint32_t* cnt, uint32_t need, unsigned ns) {
  while (VolatileLoad(cnt) < need) {
    if (VolatileLoad(&s.error)) break;
    Backoff(ns);
    if (ns < 4000u) ns += ns;   // a role that is far ahead sleeps longer and longer: polling must not eat issue slots
  }
  FenceBlock();
}
// 4-byte asynchronous global -> shared copy (LDGSTS): no register round trip, so many can be in flight.
GMX_DEV inline void CpAsync4(void* smem_dst, const void* gmem_src) {
#if defined(__CUDA_ARCH__)
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
#elif defined(GMX_EMU_DEFER_CP)
  cuda_emu::CpAsyncIssue(smem_dst, gmem_src, 4);
#else
  *(uint32_t*)smem_dst = *(const uint32_t*)gmem_src;
#endif
}
GMX_DEV inline void CpAsync16(void* smem_dst, const void* gmem_src) {   // both 16-byte aligned
#if defined(__CUDA_ARCH__)
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
#elif defined(GMX_EMU_DEFER_CP)
  cuda_emu::CpAsyncIssue(smem_dst, gmem_src, 16);
#else
  memcpy(smem_dst, gmem_src, 16);
#endif
}
GMX_DEV inline void CpAsync16Hint(void* smem_dst, const void* gmem_src, uint64_t policy) {
#if defined(__CUDA_ARCH__)
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "l"(policy) : "memory");
#elif defined(GMX_EMU_DEFER_CP)
  (void)policy; cuda_emu::CpAsyncIssue(smem_dst, gmem_src, 16);
#else
  (void)policy; memcpy(smem_dst, gmem_src, 16);
#endif
}
GMX_DEV inline float LoadHint(const float* p, uint64_t policy) {
#if defined(__CUDA_ARCH__)
  float v; asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(policy) : "memory"); return v;
#else
  (void)policy; return *p;
#endif
}
GMX_DEV inline float4 LoadHint4(const float4* p, uint64_t policy) {
#if defined(__CUDA_ARCH__)
  float4 v;
  asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy) : "memory");
  return v;
#else
  (void)policy; return *p;
#endif
}
GMX_DEV inline void StoreHint(float* p, float v, uint64_t policy) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(p), "f"(v), "l"(policy) : "memory");
#else
  (void)policy; *p = v;
#endif
}
GMX_DEV inline void CpAsyncCommit() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;" ::: "memory");
#elif defined(GMX_EMU_DEFER_CP)
  cuda_emu::CpAsyncCommitGroup();
#endif
}
template <int N>
GMX_DEV inline void CpAsyncWaitGroup() {   // at most N of this thread's most recent groups still pending
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#elif defined(GMX_EMU_DEFER_CP)
  cuda_emu::CpAsyncWaitGroupN(N);
#endif
}
GMX_DEV inline void CpAsyncWaitAll() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.wait_all;" ::: "memory");
#elif defined(GMX_EMU_DEFER_CP)
  cuda_emu::CpAsyncWaitEverything();
#endif
}
GMX_DEV inline void PrefetchL2(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
GMX_DEV inline unsigned long long GlobalTimerNs() {
#if defined(__CUDA_ARCH__)
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#else
  return 0ull;
#endif
}
GMX_DEV inline uint32_t SmId() {
#if defined(__CUDA_ARCH__)
  uint32_t v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
#else
  return 0u;
#endif
}

GMX_DEV inline float4 LoadStream4(const float4* p) {
#if defined(__CUDA_ARCH__)
  return __ldcs(p);
#else
  return *p;
#endif
}
GMX_DEV inline void StoreStream4(float4* p, float4 v) {
#if defined(__CUDA_ARCH__)
  __stcs(p, v);
#else
  *p = v;
#endif
}
// L2 prefetch of `bytes` bytes at p, cooperatively by the `nthr` threads numbered t = 0..nthr-1.
GMX_DEV inline void PrefetchRange(const void* p, uint32_t bytes, int t, int nthr) {
  for (uint32_t o = (uint32_t)t * 128u; o < bytes; o += (uint32_t)nthr * 128u) PrefetchL2((const char*)p + o);
}


// ---- TMA bulk copies (cp.async.bulk, 1-D, no tensor map) and their mbarrier --------------------------------
GMX_DEV inline void MbarInit(uint64_t* mbar, uint32_t count) {
#if defined(__CUDA_ARCH__)
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#else
  *mbar = 0; (void)count;
#endif
}
GMX_DEV inline void MbarExpectTx(uint64_t* mbar, uint32_t bytes) {   // one arrival + `bytes` of pending transactions
#if defined(__CUDA_ARCH__)
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(bytes) : "memory");
#else
  (void)mbar; (void)bytes;
#endif
}
GMX_DEV inline void MbarWait(uint64_t* mbar, uint32_t parity) {
#if defined(__CUDA_ARCH__)
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
  uint32_t done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(a), "r"(parity) : "memory");
  }
#else
  (void)mbar; (void)parity;
#endif
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion is signalled on mbar
GMX_DEV inline void BulkG2S(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* mbar) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
#else
  (void)mbar; memcpy(smem_dst, gmem_src, bytes);
#endif
}
// `bytes` (multiple of 16) at a 16-byte aligned global address -> L2, one instruction
GMX_DEV inline void BulkPrefetchL2(const void* gmem_src, uint32_t bytes) {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem_src), "r"(bytes) : "memory");
#else
  (void)gmem_src; (void)bytes;
#endif
}
// orders this thread's earlier global/shared writes (generic proxy) before later bulk copies (async proxy)
GMX_DEV inline void FenceProxyAsync() {
#if defined(__CUDA_ARCH__)
  __threadfence();
  asm volatile("fence.proxy.async;" ::: "memory");
#endif
}

// Gate weights resident in shared memory (latency configurations, one CTA per SM): the dense part of the three gate
// matrices, [3][77 quads][50 cells] float4 = 184 800 B of dynamic shared memory, refreshed from the arena by three TMA
// bulk copies after every Adam step (once per 100 bytes) instead of being streamed from HBM for every byte.
enum : int { W_DENSE_Q = L_ROWQ - L_NOUT / 4, W_DENSE_F4 = 3 * W_DENSE_Q * L_CELLS, W_DENSE_BYTES = W_DENSE_F4 * 16 };
struct WeightSmem { float4* w; uint64_t* mbar; };   // w == nullptr: not resident
template <int NL>
GMX_DEV void LoadGateWeights(StreamSmem& s, const Arena& A, const WeightSmem& ws, int ltid) {
  FenceProxyAsync();            // Adam's (or InitStream's) stores of every thread of the group
  GroupSync<NL>(BAR_LSTM);
  const uint32_t parity = s.wphase & 1u;
  if (ltid == 0) {
    const float* W = A.at<float>(A.L->l_w);
    MbarExpectTx(ws.mbar, (uint32_t)W_DENSE_BYTES);
    for (int g = 0; g < 3; ++g) BulkG2S(ws.w + g * W_DENSE_Q * L_CELLS, W + LstmW(g, L_NOUT, 0), W_DENSE_Q * L_CELLS * 16, ws.mbar);
  }
  MbarWait(ws.mbar, parity);
  GroupSync<NL>(BAR_LSTM);
  if (ltid == 0) s.wphase = parity ^ 1u;
}

// ---- stream start ------------------------------------------------------------------------------
template <int NT>
GMX_DEV void FillWords(uint32_t* p, uint64_t nwords, uint32_t v, int tid) {
  for (uint64_t i = tid; i < nwords; i += NT) p[i] = v;
}

// All NT threads of the CTA, before the roles split (CTA-wide barriers are fine here).
template <int NT>
GMX_DEV void InitStream(StreamSmem& s, const Arena& A, const StreamParams& P, int tid) {
  const ArenaLayout& L = *A.L;
  if (P.tmpl_arena) {   // clone of a parked stream: arena image, then everything of StreamSmem behind the launch tables
    if (GMX_IS_OV(L)) {
      // overlay mode: only the state the stream updates densely is copied (every region is a 256-byte multiple apart from
      // its tail, all offsets are 256-byte aligned: 16-byte copies); the overlay map starts empty
      const ArenaLayout& B = *P.tmpl_layout;
      auto copy = [&](uint64_t dst_off, uint64_t src_off, uint64_t bytes) {
        const uint4* src = (const uint4*)(P.tmpl_arena + src_off);
        uint4* dst = (uint4*)(A.base + dst_off);
        for (uint64_t i = tid; i < (bytes + 15) / 16; i += NT) dst[i] = src[i];
      };
      copy(L.ind_pred, B.ind_pred, (uint64_t)NIND * 512 * 4);
      copy(L.match_pred, B.match_pred, NMATCH * 256 * 4); copy(L.match_cnt, B.match_cnt, NMATCH * 256 * 4);
      copy(L.l_w, B.l_w, (uint64_t)L_WSIZE * 4); copy(L.l_m, B.l_m, (uint64_t)L_WSIZE * 4); copy(L.l_v, B.l_v, (uint64_t)L_WSIZE * 4);
      copy(L.l_gb, B.l_gb, 8 * 3 * L_CELLS * 4);
      copy(L.l_wout, B.l_wout, (uint64_t)L_HORIZON * L_HID * L_NOUT * 4);
      copy(L.l_lin, B.l_lin, (uint64_t)L_HORIZON * (L_NIN + 1) * 4); copy(L.l_out, B.l_out, (uint64_t)L_HORIZON * L_NOUT * 4);
      copy(L.l_gstate, B.l_gstate, 3ull * L_HORIZON * L_CELLS * 4); copy(L.l_norm, B.l_norm, 3ull * L_HORIZON * L_CELLS * 4);
      copy(L.l_ivar, B.l_ivar, 3ull * L_HORIZON * 4);
      copy(L.l_tanh, B.l_tanh, (uint64_t)L_HORIZON * L_CELLS * 4); copy(L.l_ig, B.l_ig, (uint64_t)L_HORIZON * L_CELLS * 4);
      copy(L.l_last, B.l_last, (uint64_t)L_HORIZON * L_CELLS * 4);
      copy(L.p_state, B.p_state, sizeof(PpmdState));
      if (L.p_mask) {   // small model: its whole power-of-two window
        copy(L.p_heap, B.p_heap, (uint64_t)L.p_mask + 1);
      } else {   // the three live areas of the model's PPMd heap (a power-of-two window there) into the segmented private backing
        const PpmdState* bs = (const PpmdState*)(P.tmpl_arena + B.p_state);
        const uint8_t* bh = P.tmpl_arena + B.p_heap;
        uint8_t* h = A.at<uint8_t>(L.p_heap);
        const uint32_t text = bs->text_ptr, lo = bs->lo_unit - PPMD_UNITS_START, hi = PPMD_HEAP_END - bs->hi_unit;
        const uint32_t hi_base = PPMD_HEAP_END - L.p_hi_cap;
        // 4-byte words: every area starts 4-byte aligned in both backings (units are 12 bytes, caps multiples of 16)
        for (uint32_t i = tid * 4u; i < text; i += NT * 4u) *(uint32_t*)(h + i) = *(const uint32_t*)(bh + (i & B.p_mask));
        for (uint32_t i = tid * 4u; i < lo; i += NT * 4u) *(uint32_t*)(h + L.p_seg_lo + i) = *(const uint32_t*)(bh + ((PPMD_UNITS_START + i) & B.p_mask));
        for (uint32_t i = tid * 4u; i < hi; i += NT * 4u)
          *(uint32_t*)(h + L.p_seg_lo + L.p_lo_cap + (bs->hi_unit - hi_base) + i) = *(const uint32_t*)(bh + ((bs->hi_unit + i) & B.p_mask));
        // what the model never used must read as zero (mod_ppmd.cpp relies on fresh pages): text tail, units beyond lo, below hi
        for (uint32_t i = ((text + 3u) & ~3u) + tid * 4u; i < L.p_seg_lo; i += NT * 4u) *(uint32_t*)(h + i) = 0u;
        for (uint32_t i = ((lo + 3u) & ~3u) + tid * 4u; i < L.p_lo_cap + (bs->hi_unit - hi_base); i += NT * 4u) *(uint32_t*)(h + L.p_seg_lo + i) = 0u;
      }
      uint4* z = A.at<uint4>(L.sparse);
      const uint64_t n16 = ((uint64_t)L.sparse_mask + 1) / 2;
      for (uint64_t i = tid; i < n16; i += NT) z[i] = make_uint4(0u, 0u, 0u, 0u);
    } else {
      const uint4* src = (const uint4*)P.tmpl_arena;
      uint4* dst = (uint4*)A.base;
      const uint64_t n16 = L.total / 16;
      for (uint64_t i = tid; i < n16; i += NT) dst[i] = src[i];
    }
    constexpr int kFirst = (int)(sizeof(StreamTables) / 4), kWords = (int)(offsetof(StreamSmem, pkt) / 4);
    uint32_t* sw = (uint32_t*)&s;
    for (int i = kFirst + tid; i < kWords; i += NT) sw[i] = P.tmpl_state[i];
    __syncthreads();
    if (tid == 0) { s.error = 0; s.nswap = 0; s.x1 = 0; s.x2 = 0xffffffffu; s.x = 0; if (GMX_IS_OV(L)) s.sparse_used = 0; }
  } else {
    for (int k = 0; k < NIND; ++k)
      if (!L.ind_sid[k]) FillWords<NT>(A.at<uint32_t>(L.ind_tab[k]), ((uint64_t)L.ind_size[k] + 1) / 2, 0x00FF00FFu, tid);
    if (L.sparse_mask) {
      uint4* z = A.at<uint4>(L.sparse);
      const uint64_t n16 = ((uint64_t)L.sparse_mask + 1) / 2;
      for (uint64_t i = tid; i < n16; i += NT) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    FillWords<NT>(A.at<uint32_t>(L.ind_pred), NIND * 512, 0u, tid);
    for (int k = 0; k < NMATCH; ++k)
      if (!L.match_sid[k]) FillWords<NT>(A.at<uint32_t>(L.match_tab[k]), 1ull << s.T.match[k].log2, 0u, tid);
    for (int i = tid; i < NMATCH * 256; i += NT) {
      A.at<float>(L.match_pred)[i] = (float)(0.5 + ((double)(i & 255) + 0.5) / 512);  // match.cpp:19-21
      A.at<int>(L.match_cnt)[i] = 1;
    }
    for (int k = 0; k < NIH; ++k)
      if (!L.ih_sid[k]) FillWords<NT>(A.at<uint32_t>(L.ih_tab[k]), 1ull << s.T.ih[k].log2, 0u, tid);
    for (int m = 0; m < NMIX; ++m) FillWords<NT>(A.at<uint32_t>(L.mix_dir[m]), 1ull << s.T.mixer[m].log2, 0u, tid);
    // LSTM (lstm.cpp:8-43, lstm-layer.cpp:36-54,156-196)
    for (int i = tid; i < L_WSIZE; i += NT) {
      A.at<float>(L.l_w)[i] = P.lstm_init[i];
      A.at<float>(L.l_m)[i] = 0.0f;
      A.at<float>(L.l_v)[i] = 0.0f;
    }
    for (int i = tid; i < 8 * 3 * L_CELLS; i += NT) A.at<float>(L.l_gb)[i] = i < 3 * L_CELLS ? 1.0f : 0.0f;
    FillWords<NT>(A.at<uint32_t>(L.l_wout), L_HID * L_NOUT, 0u, tid);  // epoch slot 0; others are written before read
    for (int e = tid; e < L_HORIZON; e += NT) A.at<float>(L.l_lin)[e * (L_NIN + 1) + L_NIN - 1] = 1.0f;
    for (int i = tid; i < L_HORIZON * L_NOUT; i += NT) A.at<float>(L.l_out)[i] = (float)(1.0 / L_NOUT);  // lstm.cpp:19 (only a checkpoint ever shows it)
    // PPMd heap must start zeroed (mod_ppmd.cpp relies on fresh pages)
    {
      uint4* z = A.at<uint4>(L.p_heap);
      const uint64_t n16 = ((uint64_t)L.p_mask + 1) / 16;
      for (uint64_t i = tid; i < n16; i += NT) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    // shared state
    for (int i = tid; i < NPRED + 2; i += NT) s.preds[i] = 0.0f;
    for (int i = tid; i < NPRED + NL0 + 2; i += NT) s.act[i] = i >= NPRED;
    for (int i = tid; i < C_COUNT + 2; i += NT) s.ctx[i] = 0;
    for (int i = tid; i < NL0; i += NT) s.l0_out[i] = 0.0f;
    for (int i = tid; i < NL1; i += NT) s.l1_out[i] = 0.0f;
    for (int i = tid; i < 256; i += NT) { s.ppm[i] = (float)(1.0 / 256); s.lprob[i] = (float)(1.0 / 256); }
    for (int i = tid; i < 32; i += NT) s.ring[i] = 0;
    for (int i = tid; i < WTOTAL; i += NT) s.w[i] = 0.0f;
    for (int i = tid; i < NPRED + NL0 + 2; i += NT) s.xe[i] = 0.0f;
    for (int i = tid; i < NMIX; i += NT) {
      s.set_steps[i] = 0; s.max_steps[i] = 1; s.set_idx[i] = 0xFFFFFFFFu; s.set_pool[i] = 0; s.shrink[i] = 0;
    }
    for (int i = tid; i < NMATCH; i += NT) { s.m_cur[i] = 0; s.m_byte[i] = 0; s.m_bitpos[i] = 128; s.m_len[i] = 0; }
    for (int i = tid; i < NIH; i += NT) { s.ih_outer[i] = 0; s.ih_hash[i] = 0; }
    for (int i = tid; i < L_HID + 1; i += NT) s.l_hidden[i] = i == L_HID - 1 ? 1.0f : 0.0f;
    for (int i = tid; i < L_CELLS; i += NT) { s.l_state[i] = 0; s.l_state_err[i] = 0; s.l_stored_err[i] = 0; s.l_hidden_err[i] = 0; }
    for (int i = tid; i < L_HORIZON; i += NT) { s.l_hist[i] = 0; s.l_symin[i] = 0; }
    if (tid == 0) {
      s.final_out = 0; s.prob = 0.5f; s.ring_pos = 0; s.new_bit = 0; s.recent_bits = 1; s.bb = 0;
      s.first_prediction = 1; s.error = 0; s.steps = 0; s.pool_next = 1; s.hist_len = 0; s.sparse_used = 0; s.nswap = 0;
      s.l_epoch = 0; s.l_update_steps = 0; s.l_old_input = 0; s.l_fused = 0;
      s.x1 = 0; s.x2 = 0xffffffffu; s.x = 0;
    }
    __syncthreads();
    if (tid == 0) {
      Ppmd pm{A.at<PpmdState>(L.p_state), A.at<uint8_t>(L.p_heap), L.p_mask, L.p_text_cap, L.p_units_cap, s.sqp, 0, s.p_masked, L.p_seg_lo, L.p_lo_cap, L.p_hi_cap};
      pm.Init();
    }
  }
  __syncthreads();
  if (tid == 0) {   // role pipeline
    s.n_ppm = 0; s.n_ppm_used = 0; s.n_pkt = 0; s.n_done = 0; s.bit_stop = 0; s.lstm_stop = 0;
    // the byte in front of the stream: none from scratch (the reference's models see a phantom 0 at their first byte
    // boundary), the checkpoint's last byte otherwise (it is completed by the first BasicContexts::Predict)
    s.byte0 = s.first_prediction ? 0u : (uint32_t)(s.recent_bits * 2 + s.new_bit) & 0xffu;
    s.t_start_us = (uint32_t)(GlobalTimerNs() / 1000ull);
  }
  FenceProxyAsync();   // the arena this thread has just written may be read by TMA bulk copies (resident gate weights)
  __syncthreads();
}

// ==== LSTM role ===================================================================================
// All functions below are executed by the NL threads of the LSTM role (ltid = 0..NL-1) and synchronise
// with GroupSync<NL>(BAR_LSTM) only.

// Output layer of the next epoch slot: copy of the slot just used plus one SGD step (Lstm::Perceive
// lstm.cpp:81-88). `byte` is the symbol that followed the forward pass of slot `last`.
// One quad of adjacent outputs (4q .. 4q+3), rows j0 .. j1-1 of the output-layer step.
template <int UNROLL>
GMX_DEV inline void OutputStepRows(const StreamSmem& s, const float* wl, float* wc, int q, int j0, int j1, uint32_t byte) {
  constexpr int NQ = L_NOUT / 4;
  const float lr = (float)0.03;
  float le[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t i = 4 * q + c;
    const float err = i == byte ? f_sub(s.lprob[i], 1.0f) : s.lprob[i];
    le[c] = f_mul(lr, err);
  }
  const float4* wl4 = (const float4*)wl + q;
  float4* wc4 = (float4*)wc + q;
GMX_UNROLL(UNROLL)
  for (int j = j0; j < j1; ++j) {
    const float h = s.l_hidden[j];
    float4 w = LoadStream4(wl4 + j * NQ);
    w.x = f_sub(w.x, f_mul(le[0], h)); w.y = f_sub(w.y, f_mul(le[1], h));
    w.z = f_sub(w.z, f_mul(le[2], h)); w.w = f_sub(w.w, f_mul(le[3], h));
    StoreStream4(wc4 + j * NQ, w);
  }
}
template <int NL>
GMX_DEV void LstmOutputStep(StreamSmem& s, const Arena& A, uint32_t last, uint32_t cur, uint32_t byte, int ltid) {
  const ArenaLayout& L = *A.L;
  const float* wl = A.at<float>(L.l_wout) + (size_t)last * L_HID * L_NOUT;
  float* wc = A.at<float>(L.l_wout) + (size_t)cur * L_HID * L_NOUT;
  constexpr int NQ = L_NOUT / 4;   // 64 quads of adjacent outputs x 51 rows, split evenly over up to 128 threads
  if (NL >= 2 * NQ) {              // two row halves per quad
    if (ltid < 2 * NQ) { const int half = ltid / NQ; OutputStepRows<13>(s, wl, wc, ltid & (NQ - 1), half * 26, half ? L_HID : 26, byte); }
  } else if (NL >= NQ + NQ / 2) {  // 96 threads: 64 take rows 0..33 of one quad, 32 take rows 34..50 of two quads
    if (ltid < NQ) OutputStepRows<17>(s, wl, wc, ltid, 0, 34, byte);
    else if (ltid < NQ + NQ / 2) { OutputStepRows<17>(s, wl, wc, 2 * (ltid - NQ), 34, L_HID, byte); OutputStepRows<17>(s, wl, wc, 2 * (ltid - NQ) + 1, 34, L_HID, byte); }
  } else if (NL >= NQ) {
    if (ltid < NQ) OutputStepRows<17>(s, wl, wc, ltid, 0, L_HID, byte);
  } else {                         // fewer threads than quads
    static_assert(NL >= NQ || NQ % NL == 0, "a role smaller than 64 threads must divide 64");
    for (int q = ltid; q < NQ; q += NL) OutputStepRows<17>(s, wl, wc, q, 0, L_HID, byte);
  }
}

// One node of the binary interval search of a byte model (ModPPMD::Predict mod_ppmd.cpp:1662-1681, LstmModel::Predict
// lstm-model.cpp:36-47): node = recent_bits (1..255) covers [bot, top]; num = sum(mid+1..top) from 0.0f ascending,
// denom continues from num over bot..mid. Returns Logit(p); flag bit 0: denom != 0, bit 1: p != 0.5.
static GMX_DEV GMX_NOINLINE float IntervalNode(const float* probs, int node, uint32_t* flag) {
  const int level = 31 - __clz(node);
  const int width = 256 >> level;
  const int bot = (node - (1 << level)) * width;
  const int top = bot + width - 1;
  const int mid = bot + ((top - bot) / 2);
  float num = 0.0f;
#pragma unroll 4
  for (int i = mid + 1; i <= top; ++i) num = f_add(num, probs[i]);
  float denom = num;
#pragma unroll 4
  for (int i = bot; i <= mid; ++i) denom = f_add(denom, probs[i]);
  if (denom != 0.0f) {
    const float p = f_div(num, denom);
    *flag = 1u | (p == 0.5f ? 0u : 2u);
    return Logit(p);
  }
  *flag = 0u;
  return 0.0f;
}

// AHEAD (compress): the byte whose bits the bit role is going to code is known, so the eight nodes on its path are
// evaluated here, off the bit role's critical path, one node per lane (lanes 0..7 of the calling warp; bit j of the
// byte is predicted at node (1 << j) | (byte >> (8 - j))). `which`: 0 = PPMd, 1 = LSTM. Warp-collective.
GMX_DEV inline void PathNodes(BytePacket& pk, const float* probs, uint32_t byte, int which, int lane) {
  uint32_t fl = 0;
  if (lane < 8) {
    const int node = (1 << lane) | (int)(byte >> (8 - lane));
    pk.node[which][lane] = IntervalNode(probs, node, &fl);
  }
  uint32_t bits = lane < 8 ? fl << (2 * lane) : 0u;
  for (int o = 4; o > 0; o >>= 1) bits |= __shfl_xor_sync(0xffffffffu, bits, o);
  if (lane == 0) {
    if (which == 0) pk.flags = bits; else pk.flags |= bits << 16;
  }
}

// Gate pre-activations f = w[sym]; f += in[j] * w[256 + j], j ascending (lstm-layer.cpp:227-232) for the 150 gate rows:
// R rows per thread as independent sequential sums (rows t, t + NA, ... of the NA = ceil(150 / R) working threads),
// weights as 16-byte loads straight from global memory (the four consecutive input columns of a cell are one float4,
// cells adjacent: a warp load is 512 contiguous bytes), marked evict-first in L2: the gate weights of all resident
// streams are 1.7 x the L2. The caller has started L2 prefetches of the whole 185 KB (LstmForward), and every row's
// next quad is requested right after its current one is consumed, so R loads per thread are always in flight without
// a second set of registers.
// RING: the weights travel through a private ring of LSTM_STAGES x R float4 slots per thread in s.w instead (cp.async
// keeps R x LSTM_STAGES 16-byte copies in flight per thread without holding registers). Only legal when the bit path of
// the same stream is not running (s.w is its weight-set staging area, empty at a byte boundary: BitBoundaryA).
#ifndef LSTM_STAGES
#define LSTM_STAGES 5
#endif
template <int NL>
GMX_DEV void LstmGateDotsRing(StreamSmem& s, const float* W, uint32_t sym, int ltid) {
  constexpr int NROWS = 3 * L_CELLS;
  constexpr int R = (NROWS + NL - 1) / NL;
  constexpr int NA = (NROWS + R - 1) / R;
  constexpr int NQ = L_NOUT / 4 + L_CELLS / 4 + 1;
  static_assert(LSTM_STAGES * R * NA * 16 <= WTOTAL * 4, "ring does not fit the weight-set staging area");
  if (ltid >= NA) return;
  float f[R];
  const float4* w[R];
  float4* ring = (float4*)s.w;
  const uint64_t pol = PolicyEvictFirst();
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int r = ltid + k * NA < NROWS ? ltid + k * NA : 0;
    const int g = r / L_CELLS, i = r - g * L_CELLS;
    w[k] = (const float4*)W + ((size_t)g * L_ROWQ + L_NOUT / 4) * L_CELLS + i;
    f[k] = W[LstmW(g, (int)sym, i)];
  }
#pragma unroll
  for (int q = 0; q < LSTM_STAGES; ++q) {
#pragma unroll
    for (int k = 0; k < R; ++k) CpAsync16Hint(ring + (q * R + k) * NA + ltid, w[k] + q * L_CELLS, pol);
    CpAsyncCommit();
  }
  int st = 0;
#pragma unroll 1
  for (int q = 0; q < NQ; ++q) {
    CpAsyncWaitGroup<LSTM_STAGES - 1>();
    float4 a[R];
#pragma unroll
    for (int k = 0; k < R; ++k) a[k] = ring[(st * R + k) * NA + ltid];
    if (q + LSTM_STAGES < NQ) {
#pragma unroll
      for (int k = 0; k < R; ++k) CpAsync16Hint(ring + (st * R + k) * NA + ltid, w[k] + (q + LSTM_STAGES) * L_CELLS, pol);
    }
    CpAsyncCommit();   // one group per iteration, empty at the tail, keeps the wait distance constant
    st = st + 1 == LSTM_STAGES ? 0 : st + 1;
    if (q < NQ - 1) {   // layer input = [ppm 256 | hidden 50 | 1]
      const float4 x = q < L_NOUT / 4 ? ((const float4*)s.ppm)[q] : ((const float4*)s.l_hidden)[q - L_NOUT / 4];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        f[k] = f_add(f[k], f_mul(x.x, a[k].x)); f[k] = f_add(f[k], f_mul(x.y, a[k].y));
        f[k] = f_add(f[k], f_mul(x.z, a[k].z)); f[k] = f_add(f[k], f_mul(x.w, a[k].w));
      }
    } else {            // hidden 48, 49 and the bias input (1.0); the fourth column is padding
      const float h48 = s.l_hidden[L_CELLS - 2], h49 = s.l_hidden[L_CELLS - 1];
#pragma unroll
      for (int k = 0; k < R; ++k) {
        f[k] = f_add(f[k], f_mul(h48, a[k].x)); f[k] = f_add(f[k], f_mul(h49, a[k].y)); f[k] = f_add(f[k], f_mul(1.0f, a[k].z));
      }
    }
  }
  CpAsyncWaitAll();
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int r = ltid + k * NA;
    if (r < NROWS) { const int g = r / L_CELLS; s.l_gate[g][r - g * L_CELLS] = f[k]; }
  }
}

// Resident variant: the dense weights come from shared memory (WeightSmem), only the one-hot column from the arena.
template <int NL>
GMX_DEV void LstmGateDotsSmem(StreamSmem& s, const float* W, const float4* wd, uint32_t sym, int ltid) {
  constexpr int NROWS = 3 * L_CELLS;
  constexpr int R = (NROWS + NL - 1) / NL;
  constexpr int NA = (NROWS + R - 1) / R;
  constexpr int NQ = W_DENSE_Q;
  if (ltid >= NA) return;
  float f[R];
  const float4* w[R];
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int r = ltid + k * NA < NROWS ? ltid + k * NA : 0;
    const int g = r / L_CELLS, i = r - g * L_CELLS;
    w[k] = wd + (size_t)g * W_DENSE_Q * L_CELLS + i;
    f[k] = W[LstmW(g, (int)sym, i)];
  }
#pragma unroll 4
  for (int q = 0; q < NQ - 1; ++q) {
    const float4 x = q < L_NOUT / 4 ? ((const float4*)s.ppm)[q] : ((const float4*)s.l_hidden)[q - L_NOUT / 4];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float4 c = w[k][q * L_CELLS];
      f[k] = f_add(f[k], f_mul(x.x, c.x)); f[k] = f_add(f[k], f_mul(x.y, c.y));
      f[k] = f_add(f[k], f_mul(x.z, c.z)); f[k] = f_add(f[k], f_mul(x.w, c.w));
    }
  }
  {
    const float h48 = s.l_hidden[L_CELLS - 2], h49 = s.l_hidden[L_CELLS - 1];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float4 c = w[k][(NQ - 1) * L_CELLS];
      f[k] = f_add(f[k], f_mul(h48, c.x)); f[k] = f_add(f[k], f_mul(h49, c.y)); f[k] = f_add(f[k], f_mul(1.0f, c.z));
    }
  }
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int r = ltid + k * NA;
    if (r < NROWS) { const int g = r / L_CELLS; s.l_gate[g][r - g * L_CELLS] = f[k]; }
  }
}

template <int NL>
GMX_DEV void LstmGateDots(StreamSmem& s, const float* W, uint32_t sym, int ltid) {
  constexpr int NROWS = 3 * L_CELLS;
  constexpr int R = (NROWS + NL - 1) / NL;
  constexpr int NA = (NROWS + R - 1) / R;
  constexpr int NQ = L_NOUT / 4 + L_CELLS / 4 + 1;   // 64 quads of ppm, 12 of hidden 0..47, 1 of {h48, h49, bias, pad}
  if (ltid >= NA) return;
  float f[R];
  const float4* w[R];
  float4 a[R];
#if defined(__CUDA_ARCH__)
  const uint64_t pol = PolicyEvictFirst();
#define GMX_GATE_LD(p) LoadHint4(p, pol)
#else
#define GMX_GATE_LD(p) (*(p))
#endif
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int r = ltid + k * NA < NROWS ? ltid + k * NA : 0;   // surplus slots (R * NA > 150) recompute row 0 and drop it
    const int g = r / L_CELLS, i = r - g * L_CELLS;
    w[k] = (const float4*)W + ((size_t)g * L_ROWQ + L_NOUT / 4) * L_CELLS + i;
    a[k] = GMX_GATE_LD(w[k]);
    f[k] = W[LstmW(g, (int)sym, i)];
  }
#pragma unroll 1
  for (int q = 0; q < NQ - 1; ++q) {
    const float4 x = q < L_NOUT / 4 ? ((const float4*)s.ppm)[q] : ((const float4*)s.l_hidden)[q - L_NOUT / 4];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      const float4 c = a[k];
      a[k] = GMX_GATE_LD(w[k] + (q + 1) * L_CELLS);
      f[k] = f_add(f[k], f_mul(x.x, c.x)); f[k] = f_add(f[k], f_mul(x.y, c.y));
      f[k] = f_add(f[k], f_mul(x.z, c.z)); f[k] = f_add(f[k], f_mul(x.w, c.w));
    }
  }
  {   // hidden 48, 49 and the bias input (1.0); the fourth column is padding
    const float h48 = s.l_hidden[L_CELLS - 2], h49 = s.l_hidden[L_CELLS - 1];
#pragma unroll
    for (int k = 0; k < R; ++k) {
      f[k] = f_add(f[k], f_mul(h48, a[k].x)); f[k] = f_add(f[k], f_mul(h49, a[k].y)); f[k] = f_add(f[k], f_mul(1.0f, a[k].z));
    }
  }
#undef GMX_GATE_LD
#pragma unroll
  for (int k = 0; k < R; ++k) {
    const int r = ltid + k * NA;
    if (r < NROWS) { const int g = r / L_CELLS; s.l_gate[g][r - g * L_CELLS] = f[k]; }
  }
}

// ---- LSTM forward at byte boundary b (Lstm::Predict lstm.cpp:91-122, LstmLayer::ForwardPass lstm-layer.cpp:198-241).
// Precondition: s.ppm holds the normalised PPMd distribution of this byte. sym = the byte in front of it.
// known_byte >= 0 (AHEAD): the byte this distribution is about to code; its path nodes go into the packet and the
// output-layer step of Lstm::Perceive is fused (see below). Publishes n_ppm_used and n_pkt. ------------------------
template <int NL, bool PROF, bool RING = false>
GMX_DEV void LstmForward(StreamSmem& s, const Arena& A, const StreamParams& P, uint32_t b, uint32_t sym, int known_byte, int ltid, Lap<PROF>& lap,
                         const WeightSmem& ws = WeightSmem{nullptr, nullptr}, int part = 0) {
  const ArenaLayout& L = *A.L;
  const uint32_t e = s.l_epoch;
  float* lin_e = A.at<float>(L.l_lin) + e * (L_NIN + 1);
  // Lock-step generation splits the pass around the gate products: part 1 ends after the input vector (which also goes to the
  // batched product's operand planes), part 2 starts from the pre-activations the batched kernel computed for this stream.
  if (part == 1) {
    PrefetchRange(A.at<float>(L.l_wout) + (size_t)e * L_HID * L_NOUT, L_HID * L_NOUT * 4, ltid, NL);
    for (int i = ltid; i < L_NIN; i += NL) lin_e[i] = i < 256 ? s.ppm[i] : i < 306 ? s.l_hidden[i - 256] : 1.0f;
    for (int i = ltid; i < L_CELLS; i += NL) A.at<float>(L.l_last)[e * L_CELLS + i] = s.l_state[i];
    const uint32_t slot = blockIdx.x;
    for (int k = ltid; k < GG_K; k += NL) {
      const float v = k < 256 ? s.ppm[k] : k < 306 ? s.l_hidden[k - 256] : k == 306 ? 1.0f : 0.0f;
      const float hi = Tf32Hi(v);
      P.gate_x[GateXIndex(slot, k, 0)] = hi;
      P.gate_x[GateXIndex(slot, k, 1)] = f_sub(v, hi);
    }
    if (ltid == 0) P.gate_sym[slot] = sym;
    return;
  }
  if (part == 2) {
    const float* g = P.gate_g + (size_t)blockIdx.x * GG_N;
    for (int t = ltid; t < 3 * L_CELLS; t += NL) s.l_gate[t / L_CELLS][t % L_CELLS] = g[t];
  } else {
  // The pass streams 184.8 KB of gate weights and then the 52 KB output layer of this epoch slot. Ask L2 for all of it
  // now: the role has few threads, so its own loads keep only a few KB in flight; the prefetches put the rest of the
  // HBM latency behind them (the lines are consumed within this pass, long before L2 could evict them).
  {
    const float* W = A.at<float>(L.l_w);
    // (measured at 8 CTAs/SM: per-thread prefetch instructions beat one bulk prefetch by 1 %; the bulk form serves the
    // resident-weight configurations, whose roles have better things to do than issue 400 prefetches)
    if (ws.w) {
      if (ltid == 0) BulkPrefetchL2(A.at<float>(L.l_wout) + (size_t)e * L_HID * L_NOUT, L_HID * L_NOUT * 4);
    } else {
      if (!RING) for (int g = 0; g < 3; ++g) PrefetchRange(W + LstmW(g, L_NOUT, 0), W_DENSE_Q * L_CELLS * 16, ltid, NL);
      PrefetchRange(A.at<float>(L.l_wout) + (size_t)e * L_HID * L_NOUT, L_HID * L_NOUT * 4, ltid, NL);
    }
  }
  // layer_input[e] = [ppm 256 | hidden 50 | 1]  (SetInput lstm.cpp:45-50, copy :94-96)
  for (int i = ltid; i < L_NIN; i += NL) lin_e[i] = i < 256 ? s.ppm[i] : i < 306 ? s.l_hidden[i - 256] : 1.0f;
  for (int i = ltid; i < L_CELLS; i += NL) A.at<float>(L.l_last)[e * L_CELLS + i] = s.l_state[i];  // last_state_[epoch] = state_
  if (ws.w) LstmGateDotsSmem<NL>(s, A.at<float>(L.l_w), ws.w, sym, ltid);
  else if (RING) LstmGateDotsRing<NL>(s, A.at<float>(L.l_w), sym, ltid);
  else LstmGateDots<NL>(s, A.at<float>(L.l_w), sym, ltid);
  }
  GroupSync<NL>(BAR_LSTM);
  if (ltid == 0) Publish(&s.n_ppm_used, b + 1);   // the PPMd role may overwrite ppm now
  lap.mark(17);
  // ivar = 1 / sqrt(sum(norm^2)/cells + 1e-5): _Expr::sum() runs descending (lstm-layer.cpp:233-236)
  if (ltid < 3) {
    const float* nrm = s.l_gate[ltid];
    float acc = f_mul(nrm[L_CELLS - 1], nrm[L_CELLS - 1]);
    for (int i = L_CELLS - 2; i >= 0; --i) acc = f_add(acc, f_mul(nrm[i], nrm[i]));
    const float ivar = f_div(1.0f, f_sqrt(f_add(f_div(acc, (float)L_CELLS), 1e-5f)));
    s.l_red[ltid] = ivar;
    A.at<float>(L.l_ivar)[ltid * L_HORIZON + e] = ivar;
  }
  GroupSync<NL>(BAR_LSTM);
  for (int t = ltid; t < 3 * L_CELLS; t += NL) {
    const int g = t / L_CELLS, i = t - g * L_CELLS;
    const float* gb = A.at<float>(L.l_gb);
    const float nrm = f_mul(s.l_gate[g][i], s.l_red[g]);
    A.at<float>(L.l_norm)[((size_t)g * L_HORIZON + e) * L_CELLS + i] = nrm;
    float st = f_add(f_mul(nrm, gb[g * L_CELLS + i]), gb[(3 + g) * L_CELLS + i]);  // norm*gamma + beta
    st = g == 1 ? gm_tanhf(st) : Logistic(st);  // lstm-layer.cpp:205-211
    s.l_gate[g][i] = st;
    A.at<float>(L.l_gstate)[((size_t)g * L_HORIZON + e) * L_CELLS + i] = st;
  }
  GroupSync<NL>(BAR_LSTM);
  for (int i = ltid; i < L_CELLS; i += NL) {  // lstm-layer.cpp:212-217
    const float F = s.l_gate[0][i], I = s.l_gate[1][i], O = s.l_gate[2][i];
    const float ig = f_sub(1.0f, F);
    float st = f_mul(s.l_state[i], F);
    st = f_add(st, f_mul(I, ig));
    const float th = gm_tanhf(st);
    s.l_state[i] = st;
    s.l_hidden[i] = f_mul(O, th);
    A.at<float>(L.l_ig)[e * L_CELLS + i] = ig;
    A.at<float>(L.l_tanh)[e * L_CELLS + i] = th;
  }
  GroupSync<NL>(BAR_LSTM);
  // output layer: sum_j hidden[j] * Wout[e][i][j], j ascending, hidden[50] = 1 (lstm.cpp:105-113); 4 adjacent outputs
  // per work item: one 16-byte load feeds 4 sequential sums
  const float* wo = A.at<float>(L.l_wout) + (size_t)e * L_HID * L_NOUT;
  float mx = 0.0f;
  {
    constexpr int NQ = L_NOUT / 4;
    constexpr int QPT = NL >= NQ ? 1 : NQ / NL;   // quads of adjacent outputs per thread, advanced together
    if (ltid < NQ) {
      const float4* wo4 = (const float4*)wo + ltid;
      float acc[QPT][4];
#pragma unroll
      for (int k = 0; k < QPT; ++k) { acc[k][0] = 0.0f; acc[k][1] = 0.0f; acc[k][2] = 0.0f; acc[k][3] = 0.0f; }
GMX_UNROLL(QPT > 1 ? 3 : 17)
      for (int j = 0; j < L_HID; ++j) {
        const float h = s.l_hidden[j];
#pragma unroll
        for (int k = 0; k < QPT; ++k) {
          const float4 w = wo4[j * NQ + k * NL];
          acc[k][0] = f_add(acc[k][0], f_mul(h, w.x)); acc[k][1] = f_add(acc[k][1], f_mul(h, w.y));
          acc[k][2] = f_add(acc[k][2], f_mul(h, w.z)); acc[k][3] = f_add(acc[k][3], f_mul(h, w.w));
        }
      }
#pragma unroll
      for (int k = 0; k < QPT; ++k)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          s.l_err256[4 * (ltid + k * NL) + c] = acc[k][c];
          mx = acc[k][c] > mx ? acc[k][c] : mx;
        }
    }
  }
  // max over all outputs, seeded with 0 (max is order independent)
  for (int o = 16; o > 0; o >>= 1) { const float v = __shfl_xor_sync(0xffffffffu, mx, o); mx = v > mx ? v : mx; }
  if ((ltid & 31) == 0) s.l_red[4 + (ltid >> 5)] = mx;
  GroupSync<NL>(BAR_LSTM);
  mx = 0.0f;
  for (int wi = 0; wi < NL / 32; ++wi) { const float v = s.l_red[4 + wi]; mx = v > mx ? v : mx; }
  for (int i = ltid; i < L_NOUT; i += NL) s.lprob[i] = gm_expf(f_sub(s.l_err256[i], mx));
  GroupSync<NL>(BAR_LSTM);
  if (ltid == 0) {  // valarray::sum(): ascending, seeded with element 0 (lstm.cpp:118)
    float acc = s.lprob[0];
#pragma unroll 8
    for (int i = 1; i < L_NOUT; ++i) acc = f_add(acc, s.lprob[i]);
    s.l_red[3] = acc;
  }
  GroupSync<NL>(BAR_LSTM);
  const float denom = s.l_red[3];
  // lstm_prediction_context = first index of the maximum, strict > from 0 (lstm-model.cpp:26-33)
  float bv = 0.0f; int bi = 0;
  for (int i = ltid; i < L_NOUT; i += NL) {
    const float v = f_div(s.lprob[i], denom);
    s.lprob[i] = v;
    A.at<float>(L.l_out)[e * L_NOUT + i] = v;
    if (v > bv) { bv = v; bi = i; }
  }
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((ltid & 31) == 0) { s.l_red[4 + (ltid >> 5)] = bv; s.l_red[20 + (ltid >> 5)] = (float)bi; }
  GroupSync<NL>(BAR_LSTM);
  BytePacket& pk = s.pkt[b % PKT_RING];
  if (ltid < 32) {
    if (ltid == 0) {
      float v = 0.0f; int vi = 0;
      for (int wi = 0; wi < NL / 32; ++wi) {
        const float ov = s.l_red[4 + wi]; const int oi = (int)s.l_red[20 + wi];
        if (ov > v || (ov == v && ov > 0.0f && oi < vi)) { v = ov; vi = oi; }
      }
      pk.lstm_ctx = v > 0.0f ? (uint32_t)vi : 0u;
      s.l_epoch = e + 1 == L_HORIZON ? 0 : e + 1;
      s.l_fused = known_byte >= 0 && e + 1 < L_HORIZON;
    }
    if (known_byte >= 0) PathNodes(pk, s.lprob, (uint32_t)known_byte, 1, ltid);
    __syncwarp();
    if (ltid == 0) Publish(&s.n_pkt, b + 1);
  }
  static_assert(NL / 32 <= 16, "l_red slots 4..19 / 20..35 hold one partial result per warp");
  lap.mark(18);
  // AHEAD knows the byte this distribution is about to code, so the output-layer step that Lstm::Perceive performs
  // after the byte (same operands: these probabilities, this hidden state, the layer of slot e) runs now, while slot e
  // is still in L2 from the dot products above: one HBM read of the 52 KB layer per byte instead of two. Not in the
  // last slot: there Perceive runs BPTT over all 100 stored layers before it overwrites slot 0.
  if (known_byte >= 0 && e + 1 < L_HORIZON) {
    LstmOutputStep<NL>(s, A, e, e + 1, (uint32_t)known_byte, ltid);
    lap.mark(19);
  }
  GroupSync<NL>(BAR_LSTM);
}

// Truncated BPTT over the 100 stored steps + Adam (Lstm::Perceive lstm.cpp:57-79,
// LstmLayer::BackwardPass lstm-layer.cpp:252-354). Weight gradients are accumulated per weight in
// the reference's epoch order (99 -> 0) by the thread that owns the weight, then Adam is applied.
template <int NL, bool PROF>
GMX_DEV void LstmBptt(StreamSmem& s, const Arena& A, const StreamParams& P, int ltid, Lap<PROF>& lap, const WeightSmem& ws) {
  const ArenaLayout& L = *A.L;
  float* gb = A.at<float>(L.l_gb);
  const float* W = A.at<float>(L.l_w);
  float* errh = A.at<float>(L.l_errh);
  // recurrent weights W[cell j][512 + i], snapshot transposed so that lanes (= i) read them coalesced
  // (the reference snapshots the same block into transpose_ at the first epoch, lstm-layer.cpp:300-311)
  float* Wt = A.at<float>(L.l_wt);
  for (int q = ltid; q < 3 * L_CELLS * L_CELLS; q += NL) {
    const int g = q / (L_CELLS * L_CELLS), rem = q - g * (L_CELLS * L_CELLS);
    const int j = rem / L_CELLS, i = rem - j * L_CELLS;
    Wt[q] = W[LstmW(g, 512 + i, j)];
  }
  GroupSync<NL>(BAR_LSTM);
  constexpr int CPT = (L_CELLS + NL - 1) / NL;   // cells per thread in the cell-parallel phases (2 when the role is one warp)
  if (!ws.w) PrefetchRange(A.at<float>(L.l_wout) + (size_t)(L_HORIZON - 1) * L_HID * L_NOUT, L_HID * L_NOUT * 4, ltid, NL);
  else if (ltid == 0) BulkPrefetchL2(A.at<float>(L.l_wout) + (size_t)(L_HORIZON - 1) * L_HID * L_NOUT, L_HID * L_NOUT * 4);
#pragma unroll 1
  for (int ep = L_HORIZON - 1; ep >= 0; --ep) {
    // the output layer of the next (earlier) epoch: 52 KB this pass will stream one epoch from now
    if (ep > 0) {
      if (!ws.w) PrefetchRange(A.at<float>(L.l_wout) + (size_t)(ep - 1) * L_HID * L_NOUT, L_HID * L_NOUT * 4, ltid, NL);
      else if (ltid == 0) BulkPrefetchL2(A.at<float>(L.l_wout) + (size_t)(ep - 1) * L_HID * L_NOUT, L_HID * L_NOUT * 4);
    }
    const float* out_e = A.at<float>(L.l_out) + ep * L_NOUT;
    for (int i = ltid; i < L_NOUT; i += NL)
      s.l_err256[i] = (uint32_t)i == s.l_hist[ep] ? f_sub(out_e[i], 1.0f) : out_e[i];
    GroupSync<NL>(BAR_LSTM);
    {
      // hidden_error[j] += Wout[ep][i][j] * err_i, i ascending; hidden_error is 0 on entry (lstm.cpp:60-70).
      // CPT independent sequential sums per thread.
      float he[CPT];
      const float4* wo4[CPT];
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        const int c = ltid + k * NL < L_CELLS ? ltid + k * NL : 0;
        he[k] = s.l_hidden_err[c];
        wo4[k] = (const float4*)(A.at<float>(L.l_wout) + ((size_t)ep * L_HID + c) * L_NOUT);
      }
      const float4* er4 = (const float4*)s.l_err256;
GMX_UNROLL(GMX_BPTT_UNROLL)
      for (int i = 0; i < L_NOUT / 4; ++i) {
        const float4 e = er4[i];
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const float4 w = LoadStream4(wo4[k] + i);
          he[k] = f_add(he[k], f_mul(w.x, e.x)); he[k] = f_add(he[k], f_mul(w.y, e.y));
          he[k] = f_add(he[k], f_mul(w.z, e.z)); he[k] = f_add(he[k], f_mul(w.w, e.w));
        }
      }
      // LstmLayer::BackwardPass lstm-layer.cpp:256-281
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        const int i = ltid + k * NL;
        if (i < L_CELLS) {
          float stored = ep == L_HORIZON - 1 ? he[k] : f_add(s.l_stored_err[i], he[k]);
          float se = ep == L_HORIZON - 1 ? 0.0f : s.l_state_err[i];
          const float th = A.at<float>(L.l_tanh)[ep * L_CELLS + i];
          const float ig = A.at<float>(L.l_ig)[ep * L_CELLS + i];
          const float ls = A.at<float>(L.l_last)[ep * L_CELLS + i];
          const float F = A.at<float>(L.l_gstate)[((size_t)0 * L_HORIZON + ep) * L_CELLS + i];
          const float I = A.at<float>(L.l_gstate)[((size_t)1 * L_HORIZON + ep) * L_CELLS + i];
          const float O = A.at<float>(L.l_gstate)[((size_t)2 * L_HORIZON + ep) * L_CELLS + i];
          s.l_gerr[2][i] = f_mul(f_mul(f_mul(th, stored), O), f_sub(1.0f, O));
          se = f_add(se, f_mul(f_mul(stored, O), f_sub(1.0f, f_mul(th, th))));
          s.l_gerr[1][i] = f_mul(f_mul(se, ig), f_sub(1.0f, f_mul(I, I)));
          s.l_gerr[0][i] = f_mul(f_mul(f_mul(f_sub(ls, I), se), F), ig);
          s.l_hidden_err[i] = 0.0f;
          if (ep > 0) { se = f_mul(se, F); stored = 0.0f; }
          s.l_state_err[i] = se;
          s.l_stored_err[i] = stored;
        }
      }
    }
    if (ltid == 0 && ep == 0 && s.l_update_steps < (uint32_t)L_UPDATE_LIMIT) s.l_update_steps++;
    GroupSync<NL>(BAR_LSTM);
    // per gate (lstm-layer.cpp:313-318): beta_u += err; gamma_u += err*norm; err *= gamma*ivar
    for (int t = ltid; t < 3 * L_CELLS; t += NL) {
      const int g = t / L_CELLS, i = t - g * L_CELLS;
      const float err = s.l_gerr[g][i];
      const float nrm = A.at<float>(L.l_norm)[((size_t)g * L_HORIZON + ep) * L_CELLS + i];
      const float gu = ep == L_HORIZON - 1 ? 0.0f : gb[(6 * 3 + g) * L_CELLS + i];
      const float bu = ep == L_HORIZON - 1 ? 0.0f : gb[(7 * 3 + g) * L_CELLS + i];
      gb[(7 * 3 + g) * L_CELLS + i] = f_add(bu, err);
      gb[(6 * 3 + g) * L_CELLS + i] = f_add(gu, f_mul(err, nrm));
      s.l_gerr[g][i] = f_mul(err, f_mul(gb[g * L_CELLS + i], A.at<float>(L.l_ivar)[g * L_HORIZON + ep]));
    }
    GroupSync<NL>(BAR_LSTM);
    // err -= (sum(err*norm)/cells) * norm; the sum is an _Expr::sum(): descending (lstm-layer.cpp:319-321).
    // One lane per gate forms the sum (it is the same for every cell of the gate).
    if (ltid < 3) {
      const int g = ltid;
      const float* nrm = A.at<float>(L.l_norm) + ((size_t)g * L_HORIZON + ep) * L_CELLS;
      float acc = f_mul(s.l_gerr[g][L_CELLS - 1], nrm[L_CELLS - 1]);
#pragma unroll 7
      for (int k = L_CELLS - 2; k >= 0; --k) acc = f_add(acc, f_mul(s.l_gerr[g][k], nrm[k]));
      s.l_red[36 + g] = f_div(acc, (float)L_CELLS);
    }
    GroupSync<NL>(BAR_LSTM);
    for (int t = ltid; t < 3 * L_CELLS; t += NL) {
      const int g = t / L_CELLS, i = t - g * L_CELLS;
      const float nrm = A.at<float>(L.l_norm)[((size_t)g * L_HORIZON + ep) * L_CELLS + i];
      const float ne = f_sub(s.l_gerr[g][i], f_mul(s.l_red[36 + g], nrm));
      s.l_gerr[g][i] = ne;
      errh[((size_t)g * L_HORIZON + ep) * L_CELLS + i] = ne;
    }
    GroupSync<NL>(BAR_LSTM);
    for (int i = ltid; i < L_CELLS; i += NL) {
      float stored = s.l_stored_err[i];
      if (ep > 0) {  // stored_error[i] += sum_j err[j] * W[j][512 + i], gates in order (lstm-layer.cpp:331-339)
        for (int g = 0; g < 3; ++g) {
          const float* wt = Wt + (size_t)g * L_CELLS * L_CELLS + i;
          float f = 0.0f;
#pragma unroll 10
          for (int j = 0; j < L_CELLS; ++j) f = f_add(f, f_mul(s.l_gerr[g][j], wt[j * L_CELLS]));
          stored = f_add(stored, f);
        }
      }
      // ClipGradients(+-10) on state_error, stored_error, hidden_error(=0) (lstm-layer.cpp:243-250,292-294)
      float se = s.l_state_err[i];
      se = se < -10.0f ? -10.0f : se > 10.0f ? 10.0f : se;
      stored = stored < -10.0f ? -10.0f : stored > 10.0f ? 10.0f : stored;
      s.l_state_err[i] = se;
      s.l_stored_err[i] = stored;
    }
    GroupSync<NL>(BAR_LSTM);
  }
  lap.mark(20);
  // Weight gradients + Adam (lstm-layer.cpp:340-353, :12-34). Every gradient is accumulated by one
  // thread in the reference's epoch order (99 -> 0).
  const float* ad = P.adam + 4 * s.l_update_steps;
  const float alpha = ad[0], d1 = ad[1], d2 = ad[2];
  const float beta1 = (float)0.025, beta2 = (float)0.9999, eps = 1e-6f;
  const float omb1 = f_sub(1.0f, beta1), omb2 = f_sub(1.0f, beta2);
  float* Wm = A.at<float>(L.l_w);
  float* M = A.at<float>(L.l_m);
  float* V = A.at<float>(L.l_v);
  const float* lin = A.at<float>(L.l_lin);
  // one Adam step of one weight (lstm-layer.cpp:12-34)
  auto adam1 = [&](float& w, float& m, float& v, float grad) {
    m = f_add(f_mul(m, beta1), f_mul(omb1, grad));
    v = f_add(f_mul(v, beta2), f_mul(f_mul(omb2, grad), grad));
    w = f_sub(w, f_mul(alpha, f_div(f_div(m, d1), f_sqrt(f_add(f_div(v, d2), eps)))));
  };
  float4* W4 = (float4*)Wm; float4* M4 = (float4*)M; float4* V4 = (float4*)V;
  // (a) one-hot rows: row r only receives err of the epochs whose input symbol was r. Per-symbol epoch
  // lists (descending) are threaded through the scratch buffer. One thread updates the four rows of a quad
  // for one cell = one float4 of W, m and v.
  uint8_t* head = (uint8_t*)s.l_err256;          // [256] first (largest) epoch of a symbol, 0xFF = none
  uint8_t* nxt = head + 256;                     // [100] next smaller epoch with the same symbol
  for (int i = ltid; i < 256; i += NL) head[i] = 0xFF;
  GroupSync<NL>(BAR_LSTM);
  if (ltid == 0)
#pragma unroll 1
    for (int ep = 0; ep < L_HORIZON; ++ep) { const int sy = s.l_symin[ep]; nxt[ep] = head[sy]; head[sy] = (uint8_t)ep; }
  GroupSync<NL>(BAR_LSTM);
#pragma unroll 1
  for (int q = ltid; q < 3 * (L_NOUT / 4) * L_CELLS; q += NL) {
    const int g = q / ((L_NOUT / 4) * L_CELLS), rem = q - g * ((L_NOUT / 4) * L_CELLS);
    const int rq = rem / L_CELLS, i = rem - rq * L_CELLS;
    const float* eh = errh + (size_t)g * L_HORIZON * L_CELLS + i;
    float grad[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gsum = 0.0f;
      for (int ep = head[4 * rq + k]; ep != 0xFF; ep = nxt[ep]) gsum = f_add(gsum, eh[ep * L_CELLS]);
      grad[k] = gsum;
    }
    const size_t f4 = ((size_t)g * L_ROWQ + rq) * L_CELLS + i;
    float4 w = W4[f4], m = M4[f4], v = V4[f4];
    adam1(w.x, m.x, v.x, grad[0]); adam1(w.y, m.y, v.y, grad[1]); adam1(w.z, m.z, v.z, grad[2]); adam1(w.w, m.w, v.w, grad[3]);
    W4[f4] = w; M4[f4] = m; V4[f4] = v;
  }
  // (b) dense rows: grad[r][i] = sum_ep err[ep][i] * in[ep][r] as a register-tiled product, 4 rows (one quad) x 2
  // cells per thread (one 16-byte and one 8-byte load feed 8 multiply-adds), then two float4 Adam updates. A small role
  // (<= 64 threads) takes the three gates of a tile together: the layer-input quad is loaded once for all three and
  // 24 independent sums are in flight per thread.
  constexpr int RG = (L_NIN + 3) / 4, IP = L_CELLS / 2;
  constexpr int GPT = NL <= 64 ? 3 : 1;
#pragma unroll 1
  for (int id = ltid; id < (3 / GPT) * RG * IP; id += NL) {
    const int ip = id % IP, rg = (id / IP) % RG, g0 = id / (IP * RG);
    const float2* e2 = (const float2*)(errh + (size_t)g0 * L_HORIZON * L_CELLS + 2 * ip);
    const float4* x4 = (const float4*)(lin + 4 * rg);
    float a[GPT][4][2];
#pragma unroll
    for (int g = 0; g < GPT; ++g)
#pragma unroll
      for (int r = 0; r < 4; ++r) { a[g][r][0] = 0.0f; a[g][r][1] = 0.0f; }
GMX_UNROLL(GPT > 1 ? 2 : 4)
    for (int ep = L_HORIZON - 1; ep >= 0; --ep) {
      const float4 x = x4[ep * ((L_NIN + 1) / 4)];
#pragma unroll
      for (int g = 0; g < GPT; ++g) {
        const float2 e = e2[(size_t)g * (L_HORIZON * L_CELLS / 2) + ep * (L_CELLS / 2)];
        a[g][0][0] = f_add(a[g][0][0], f_mul(e.x, x.x)); a[g][0][1] = f_add(a[g][0][1], f_mul(e.y, x.x));
        a[g][1][0] = f_add(a[g][1][0], f_mul(e.x, x.y)); a[g][1][1] = f_add(a[g][1][1], f_mul(e.y, x.y));
        a[g][2][0] = f_add(a[g][2][0], f_mul(e.x, x.z)); a[g][2][1] = f_add(a[g][2][1], f_mul(e.y, x.z));
        a[g][3][0] = f_add(a[g][3][0], f_mul(e.x, x.w)); a[g][3][1] = f_add(a[g][3][1], f_mul(e.y, x.w));
      }
    }
    const bool pad = 4 * rg + 3 >= L_NIN;   // the last quad's fourth column does not exist (stays 0)
#pragma unroll
    for (int g = 0; g < GPT; ++g)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const size_t f4 = ((size_t)(g0 + g) * L_ROWQ + L_NOUT / 4 + rg) * L_CELLS + 2 * ip + c;
        float4 w = W4[f4], m = M4[f4], v = V4[f4];
        adam1(w.x, m.x, v.x, a[g][0][c]); adam1(w.y, m.y, v.y, a[g][1][c]); adam1(w.z, m.z, v.z, a[g][2][c]);
        if (!pad) adam1(w.w, m.w, v.w, a[g][3][c]);
        W4[f4] = w; M4[f4] = m; V4[f4] = v;
      }
  }
  for (int t = ltid; t < 2 * 3 * L_CELLS; t += NL) {  // gamma then beta (lstm-layer.cpp:349-352)
    const int which = t / (3 * L_CELLS), k = t - which * 3 * L_CELLS;  // k = g*CELLS + i
    float* w = gb + (which == 0 ? 0 : 3) * L_CELLS + k;
    float* pm = gb + ((which == 0 ? 2 : 4) * 3) * L_CELLS + k;
    float* pv = gb + ((which == 0 ? 3 : 5) * 3) * L_CELLS + k;
    const float grad = gb[((which == 0 ? 6 : 7) * 3) * L_CELLS + k];
    float m = f_mul(*pm, beta1);
    m = f_add(m, f_mul(omb1, grad));
    float v = f_mul(*pv, beta2);
    v = f_add(v, f_mul(f_mul(omb2, grad), grad));
    *pm = m; *pv = v;
    *w = f_sub(*w, f_mul(alpha, f_div(f_div(m, d1), f_sqrt(f_add(f_div(v, d2), eps)))));
  }
  GroupSync<NL>(BAR_LSTM);
  if (ws.w) LoadGateWeights<NL>(s, A, ws, ltid);   // the resident copy follows the Adam step
  lap.mark(21);
}

// Lstm::Perceive (lstm.cpp:52-89) after the last bit of `byte`.
template <int NL, bool PROF>
GMX_DEV void LstmPerceive(StreamSmem& s, const Arena& A, const StreamParams& P, uint32_t byte, int ltid, Lap<PROF>& lap,
                          const WeightSmem& ws = WeightSmem{nullptr, nullptr}) {
  const uint32_t cur = s.l_epoch;
  const uint32_t last = cur == 0 ? L_HORIZON - 1 : cur - 1;
  if (ltid == 0) { s.l_old_input = s.l_hist[last]; s.l_hist[last] = (uint8_t)byte; }
  GroupSync<NL>(BAR_LSTM);
  if (cur == 0) {
    // input symbol of epoch ep = byte perceived before it (lstm.cpp:71-74)
    for (int ep = ltid; ep < L_HORIZON; ep += NL) s.l_symin[ep] = ep == 0 ? (uint8_t)s.l_old_input : s.l_hist[ep - 1];
    GroupSync<NL>(BAR_LSTM);
    LstmBptt<NL, PROF>(s, A, P, ltid, lap, ws);
  }
  if (!s.l_fused) { LstmOutputStep<NL>(s, A, last, cur, byte, ltid); lap.mark(19); }
  GroupSync<NL>(BAR_LSTM);
}

// ==== PPMd role (one warp) ===========================================================================
// Byte boundary b: ModPPMD::Predict byte part (mod_ppmd.cpp:1651-1661): update the model with the byte in front of the
// boundary, build the symbol distribution, normalise it in place (ppm_predictions = max(sqp, 1) / sum, valarray::sum()
// ascending), in AHEAD mode evaluate the known byte's path nodes, publish.
template <bool PROF>
GMX_DEV void PpmdStep(StreamSmem& s, const Arena& A, uint32_t b, uint32_t last_byte, int known_byte, bool ahead, int lane, Lap<PROF>& lap) {
  const ArenaLayout& L = *A.L;
  Ppmd pm{A.at<PpmdState>(L.p_state), A.at<uint8_t>(L.p_heap), L.p_mask, L.p_text_cap, L.p_units_cap, s.sqp, lane, s.p_masked, L.p_seg_lo, L.p_lo_cap, L.p_hi_cap};
  pm.UpdateByte(last_byte);
  // sqp/ppm is one buffer: the LSTM role must be done with the previous distribution (and, in lockstep modes, the bit
  // role with its interval nodes: guaranteed by the caller having waited for byte b-1 to be known)
  lap.mark(25);
  WaitAtLeast(s, &s.n_ppm_used, b, 200);
  lap.mark(24);
  if (!pm.S->error) pm.PrepareByte();
  if (pm.S->error) SetError(s, GMX_ERR_PPMD_ARENA);
  __syncwarp();
  lap.mark(25);
  for (int i = lane; i < 256; i += 32) { float v = (float)s.sqp[i]; if (v < 1.0f) v = 1.0f; s.ppm[i] = v; }
  __syncwarp();
  float sum = 0.0f;
  if (lane == 0) {
    sum = s.ppm[0];
#pragma unroll 8
    for (int i = 1; i < 256; ++i) sum = f_add(sum, s.ppm[i]);
  }
  sum = __shfl_sync(0xffffffffu, sum, 0);
  for (int i = lane; i < 256; i += 32) s.ppm[i] = f_div(s.ppm[i], sum);
  __syncwarp();
  if (ahead) {
    // packet slot b % PKT_RING was last used by byte b - PKT_RING
    if (b >= (uint32_t)PKT_RING) { lap.mark(26); WaitAtLeast(s, &s.n_done, b - PKT_RING + 1, 1000); lap.mark(24); }
    PathNodes(s.pkt[b % PKT_RING], s.ppm, (uint32_t)known_byte, 0, lane);
    __syncwarp();
  }
  if (lane == 0) Publish(&s.n_ppm, b + 1);
  lap.mark(26);
}

// ==== bit role =======================================================================================
// Executed by the NB threads of the bit role (btid = 0..NB-1), synchronised with GroupSync<NB>(BAR_BIT).

// BasicContexts::Predict (basic-contexts.cpp:21-40), one thread.
GMX_DEV inline void Bookkeeping(StreamSmem& s) {
  if (s.first_prediction) {
    s.first_prediction = 0;
  } else {
    int rb = s.recent_bits * 2 + s.new_bit;
    if (rb >= 256) {  // ByteUpdate basic-contexts.cpp:5-19
      const uint32_t lb = rb - 256;
      s.ctx[C_LAST_BYTE] = lb;
      const uint32_t rp = (s.ring_pos + 1) & 31u;
      s.ring_pos = rp;
      s.ring[rp] = (uint8_t)lb;
      for (int i = 1; i < 10; ++i) s.ctx[C_RB1 + i - 1] = RecentByte(s, i);
      rb = 1;
    }
    s.recent_bits = rb;
    s.ctx[C_BIT_CONTEXT] = rb - 1;
    s.ctx[C_LBPR] = (s.ctx[C_LAST_BYTE] << 8) + (rb - 1);
    s.ctx[C_SLPR] = (s.ctx[C_RB1] << 8) + (rb - 1);
  }
  s.bb = s.recent_bits == 1;
}

// Everything of the bit role that only happens when a new byte has been perceived (recent_bits == 1), in two parts:
// A needs nothing from the byte models (it can run beside them), B needs the packet of byte boundary b.
template <int NB>
GMX_DEV void BitBoundaryA(StreamSmem& s, const Arena& A, int btid) {
  const ArenaLayout& L = *A.L;
  const uint32_t last_byte = s.ctx[C_LAST_BYTE];
  // (0) Nearly every gate context changes with the byte: all staged weight sets go back to the pool at once and the
  // swap phase of this bit fetches all 33 (one warp per set, one float4 per lane).
  {
    const int lane = btid & 31;
#pragma unroll 1
    for (int m = btid >> 5; m < NMIX; m += NB / 32) WriteBackSet(s, A, m, s.set_pool[m], s.set_idx[m], lane);
  }
  // (1) contexts: 9 intervals, 20 hashed skip contexts, 9 indirect-hash tables
#pragma unroll 1
  for (int t = btid; t < 9 + 20 + NIH; t += NB) {
    if (t < 9) {  // IntervalContext::Predict interval-context.cpp:17-23
      const IntervalSpec sp = s.T.interval[t];
      s.ctx[C_IV0 + t] = sp.mask & ((s.ctx[C_IV0 + t] << sp.shift) + (last_byte >> sp.div_log2));
    } else if (t < 29) {  // SkipContext::Predict skip-context.cpp:9-19
      const int id = t - 9;
      const SkipSpec sp = s.T.skip[id];
      uint64_t c = 0;
      for (int k = 0; k < sp.n; ++k) c = (c << 8) + RecentByte(s, sp.b[k]);
      s.ctx[id < 5 ? C_H2 + id : C_SK0 + (id - 5)] = Murmur64(c);
    } else {  // IndirectHash::Predict indirect-hash.cpp:16-31
      const int k = t - 29;
      const IHSpec sp = s.T.ih[k];
      const uint32_t mask = (1u << sp.log2) - 1;
      const uint64_t inner_mod = 1ull << (8 * (sp.inner_order - 1)), outer_mod = 1ull << (8 * (sp.outer_order - 1));
      const uint64_t oc = ((s.ih_outer[k] % outer_mod) << 8) + last_byte;
      s.ih_outer[k] = oc;
      const uint32_t oh = Murmur64(oc);
      const uint32_t sid = L.ih_sid[k];
      if (sid) {
        const SparseMap M = A.map();
        const uint32_t key = SparseKey(sid, s.ih_hash[k] & mask);
        unsigned long long e;
        const uint32_t pos = SparseFind(M, key, &e);
        const uint32_t cur = e ? (uint32_t)e : GMX_IS_OV(L) ? BaseIH(A, k, s.ih_hash[k] & mask) : 0u;
        SparsePut(M, &s.sparse_used, L.sparse_limit, &s.error, key, pos, e != 0ull, (uint32_t)((((uint64_t)cur % inner_mod) << 8) + last_byte));
        unsigned long long e2;
        SparseFind(M, SparseKey(sid, oh & mask), &e2);
        s.ctx[C_IH0 + k] = Murmur32(e2 ? (uint32_t)e2 : GMX_IS_OV(L) ? BaseIH(A, k, oh & mask) : 0u);
      } else {
        uint32_t* tab = A.at<uint32_t>(L.ih_tab[k]);
        uint32_t* slot = tab + (s.ih_hash[k] & mask);
        *slot = (uint32_t)((((uint64_t)*slot % inner_mod) << 8) + last_byte);
        s.ctx[C_IH0 + k] = Murmur32(tab[oh & mask]);
      }
      s.ih_hash[k] = oh;
    }
  }
}
// known_byte >= 0 (compress): the byte about to be coded. Every table slot and every weight set its 8 bits will touch is
// then already determined (bit_context of bit j = (1 << j) - 1 + (byte >> (8 - j))): their lines are requested from L2
// now, so that the per-bit lookups and weight-set swaps of this byte find them there instead of waiting for HBM. These
// are not extra requests, only earlier ones (the round-1 prefetches guessed both values of the next bit). Wins when a
// stream has its SM to itself (+1 %), loses when other resident streams already hide the latency (-2 % at 8 CTAs/SM):
// LAT kernels only.
template <int NB, bool LAT = false>
GMX_DEV void BitBoundaryB(StreamSmem& s, const Arena& A, uint32_t b, int btid, int known_byte = -1) {
  const ArenaLayout& L = *A.L;
  if (btid == 0) s.ctx[C_LSTM] = s.pkt[b % PKT_RING].lstm_ctx;
  GroupSync<NB>(BAR_BIT);   // (also orders part A's shared-memory writes when the same threads ran it)
  // Indirect row bases ((ctx << 8) % M, so that slot = (base + bit_context) % M); nothing is staged any more
  for (int k = btid; k < NIND; k += NB) {
    const uint32_t M = L.ind_size[k];
    const uint32_t base = (s.ctx[s.T.ind[k].ctx] << 8) % M;
    s.ind_base[k] = base;
#if !defined(GMX_NO_BYTE_PREFETCH)
    if (LAT && known_byte >= 0) {   // (compiled into the one-CTA-per-SM kernels only: A/B in profiles/r02_ab_byte_prefetch.txt)
      const uint32_t sid = L.ind_sid[k];
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        uint32_t slot = base + ((1u << j) - 1u + ((uint32_t)known_byte >> (8 - j)));
        if (slot >= M) slot -= M;
        if (sid) PrefetchL2(A.map().tab + (SparseHash(SparseKey(sid, slot)) & L.sparse_mask));
        else if (j == 0 || j == 6) PrefetchL2(A.at<uint16_t>(L.ind_tab[k]) + slot);   // a dense row is 510 bytes: bits 0-5 share a line
      }
    }
#endif
  }
  for (int m = btid; m < NMIX; m += NB) { s.set_idx[m] = 0xFFFFFFFFu; s.set_pool[m] = 0; s.set_dirty[m] = 0; }
#if !defined(GMX_NO_BYTE_PREFETCH)
  if (LAT && known_byte >= 0 && !GMX_IS_OV(L)) {
    // weight sets: the 27 byte-gated mixers' one set, the 4 bit-gated mixers' eight (the two longest-match gates depend on
    // the lookups): directory entry, then the record's lines. Work items = (mixer, bit).
#pragma unroll 1
    for (int t = btid; t < NMIX * 8; t += NB) {
      const int m = t >> 3, j = t & 7;
      const int cid = s.T.mixer[m].ctx;
      const bool bit_level = cid == C_SLPR || cid == C_LBPR || cid == C_BIT_CONTEXT;
      if (cid == C_LONGEST || (!bit_level && j)) continue;
      const uint32_t bc = (1u << j) - 1u + ((uint32_t)known_byte >> (8 - j));
      const uint32_t c = cid == C_BIT_CONTEXT ? bc : cid == C_LBPR ? (s.ctx[C_LAST_BYTE] << 8) + bc : cid == C_SLPR ? (s.ctx[C_RB1] << 8) + bc : s.ctx[cid];
      const uint32_t nid = A.at<uint32_t>(L.mix_dir[m])[c & ((1u << s.T.mixer[m].log2) - 1)];
      if (nid) {
        const char* rec = (const char*)(A.at<float4>(L.mix_pool) + (size_t)nid * (L.mix_set_stride / 4));
        const int bytes = (MixerNW(m) + 4) * 4;
        for (int o = 0; o < bytes; o += 128) PrefetchL2(rec + o);
      }
    }
  }
#endif
  GroupSync<NB>(BAR_BIT);
}

// Indirect::Learn (x41), Match::Learn (x6) and the history append of BasicContexts::Learn for one bit, one model per
// work item t = 0 .. NIND + NMATCH. They depend on the bit and on what the lookup phase left in shared memory, not on
// the mixer outputs.
enum : int { LEARN_TABLE_ITEMS = NIND + NMATCH + 1 };
GMX_DEV inline void LearnTables(StreamSmem& s, const Arena& A, int bit, int t) {
  const ArenaLayout& L = *A.L;
  const float fbit = (float)bit;
  const int cur = s.recent_bits * 2 + bit;
  const bool byte_done = cur >= 256;
  const uint32_t longest = s.ctx[C_LONGEST];
  // history length after BasicContexts::Learn (basic-contexts.cpp:42-54)
  const uint32_t hist_after = s.hist_len + ((byte_done && longest < 2) ? 1u : 0u);
  if (t < NIND) {  // Indirect::Learn indirect.cpp:47-70
    const int k = t;
    const float lr = s.T.ind[k].slow_lr ? f_div(1.0f, 200.0f) : (float)0.02;
    float* pr = A.at<float>(L.ind_pred) + k * 512;
    const uint32_t e = s.ind_state[k];
    uint32_t ns = e & 0xff;
    const uint32_t rm = e >> 8;
    if (ns == 255) ns = 0;
    const float a = s.ind_pa[k];
    const uint64_t keep = PolicyEvictLast();
    StoreHint(pr + ns, f_add(a, f_mul(f_sub(fbit, Logistic(a)), lr)), keep);
    const float b = s.ind_pb[k];
    StoreHint(pr + 256 + rm, f_add(b, f_mul(f_sub(fbit, Logistic(b)), lr)), keep);
    // RunMap::Next run-map.cpp:3-21
    uint32_t nrm;
    if (bit == 0) nrm = rm < 127 ? rm + 1 : rm >= 128 ? 1 : rm;
    else nrm = rm < 128 ? 128 : rm < 255 ? rm + 1 : rm;
    const uint32_t nst = s.T.nonstationary[ns * 2 + bit] | (nrm << 8);
    const uint32_t sid = L.ind_sid[k];
    if (sid) {
      uint32_t slot = s.ind_base[k] + s.ctx[C_BIT_CONTEXT];
      if (slot >= L.ind_size[k]) slot -= L.ind_size[k];
      SparsePut(A.map(), &s.sparse_used, L.sparse_limit, &s.error, SparseKey(sid, slot), s.ind_slot[k], s.ind_found[k] != 0, nst);
    } else {
      A.at<uint16_t>(L.ind_tab[k])[s.ind_slot[k]] = (uint16_t)nst;
    }
  } else if (t < NIND + NMATCH) {  // Match::Learn match.cpp:76-109
    const int k = t - NIND;
    const uint32_t len = s.m_len[k];
    if (len > 2) {
      const int hit = bit == ((s.m_byte[k] & s.m_bitpos[k]) != 0);
      int* cnt = A.at<int>(L.match_cnt) + k * 256 + len;
      float* mp = A.at<float>(L.match_pred) + k * 256 + len;
      float rate = (float)(1.0 / 400);
      const int c = *cnt;
      if (c < 400) { *cnt = c + 1; rate = (float)d_div(1.0, (double)(c + 1)); }
      const float v = *mp;
      *mp = f_add(v, f_mul(f_sub((float)hit, v), rate));
    }
    if (s.recent_bits >= 128 && longest < 2) {
      const uint32_t idx = s.ctx[s.T.match[k].ctx] & ((1u << s.T.match[k].log2) - 1);
      if (L.match_sid[k]) SparseSet(A.map(), &s.sparse_used, L.sparse_limit, &s.error, SparseKey(L.match_sid[k], idx), hist_after - 1);
      else A.at<uint32_t>(L.match_tab[k])[idx] = hist_after - 1;
    }
  } else if (t == NIND + NMATCH && byte_done && longest < 2) {
    if (s.hist_len >= L.history_cap) SetError(s, GMX_ERR_HISTORY_CAP);
    else A.at<uint8_t>(L.history)[s.hist_len - (GMX_IS_OV(L) ? L.base_hist : 0u)] = (uint8_t)cur;
  }
}

// Mixer gate selection (mixer.cpp:29-37) of mixer m whose gate context has the value c: if another weight set is needed,
// queue the swap (the old set goes back to the pool, the new one - if it exists - is staged).
GMX_DEV inline void GateSelect(StreamSmem& s, const Arena& A, int m, uint32_t c) {
  const uint32_t idx = c & ((1u << s.T.mixer[m].log2) - 1);
  if (idx != s.set_idx[m]) {
    const uint32_t q = atomicAdd(&s.nswap, 1u);
    s.swap_m[q] = (uint8_t)m;
    s.swap_old[q] = s.set_pool[m];
    s.swap_oldidx[q] = s.set_idx[m];
    s.swap_new[q] = DirGet(A, m, idx);
    s.set_idx[m] = idx;
  }
}

// ---- Predictor::Predict (predictor.cpp:360-376), the part behind the bookkeeping and the byte boundary. ----------
// path_bit >= 0 (AHEAD): index j of this bit inside its byte (0 = MSB); the byte models' predictions then come from
// the packet of byte boundary b. Otherwise (lockstep) one work item per byte model evaluates its interval node.
// learn_bit >= 0 (compress): the bit that is about to be coded; the table models then learn it on the threads that have
// nothing to do while the first warp evaluates the mixer network (they depend on the bit and on the lookups, not on the
// mixers), and the caller passes tables_done to LearnBit.
#if defined(__CUDACC__)
#define GMX_DEV_INLINE __device__ __forceinline__
#else
#define GMX_DEV_INLINE inline
#endif
// ahead: the weight update of the CURRENT bit still reads the active flags (bit 0 of the act bytes), so the flags of the next bit
// go to bit 1 and PredictBit shifts them down when that bit starts.
GMX_DEV inline void SetAct(StreamSmem& s, int pi, bool v, bool ahead) {
  s.act[pi] = ahead ? (uint8_t)((s.act[pi] & 1u) | (v ? 2u : 0u)) : (uint8_t)v;
}
// The table lookups of one bit (41 x Indirect::Predict, 6 x Match::Predict), work item t = model. bitctx / new_bit / bb are those of
// the bit being predicted: PredictBit passes the stream's current ones; in compress (hybrid order), where the byte is known, the
// lookups of the NEXT bit run under this bit's mixer network with the values Bookkeeping is going to produce.
GMX_DEV_INLINE void TableLookup(StreamSmem& s, const Arena& A, int t, uint32_t bitctx, int new_bit, bool bb, bool zero_inactive, bool ahead) {
  const ArenaLayout& L = *A.L;
  if (t < NIND) {  // Indirect::Predict indirect.cpp:28-45
    const int k = t;
    const uint32_t M = L.ind_size[k];
    uint32_t slot = s.ind_base[k] + bitctx;
    if (slot >= M) slot -= M;
    uint32_t e;
    const uint32_t sid = L.ind_sid[k];
    if (sid) {  // absent == never written == {ns 255, rm 0} (overlay mode: == what the model's table holds)
      unsigned long long ent;
      const uint32_t index = slot;
      slot = SparseFind(A.map(), SparseKey(sid, slot), &ent);
      e = ent ? (uint32_t)ent & 0xffffu : GMX_IS_OV(L) ? BaseInd(A, k, index) : 0x00ffu;
      s.ind_found[k] = ent != 0ull;
    } else {
      e = A.at<uint16_t>(L.ind_tab[k])[slot];
    }
    s.ind_slot[k] = slot;
    s.ind_state[k] = (uint16_t)e;
    const uint32_t ns = e & 0xff, rm = e >> 8;
    const float* pr = A.at<float>(L.ind_pred) + k * 512;
    const int pi = s.T.ind[k].pred;
    // both entries are read unconditionally: Indirect::Learn updates exactly these two (a never-seen state learns as
    // state 0, indirect.cpp:52-54) and takes them from shared memory instead of two more dependent global loads
    const uint64_t keep = PolicyEvictLast();
    const float pa = LoadHint(pr + (ns == 255 ? 0 : ns), keep), pb = LoadHint(pr + 256 + rm, keep);
    s.ind_pa[k] = pa; s.ind_pb[k] = pb;
    if (ns != 255) { s.preds[pi] = pa; SetAct(s, pi, pa != 0.0f, ahead); }
    else { SetAct(s, pi, false, ahead); if (zero_inactive) s.preds[pi] = 0.0f; }
    if (rm != 0) { s.preds[pi + 1] = pb; SetAct(s, pi + 1, pb != 0.0f, ahead); }
    else { SetAct(s, pi + 1, false, ahead); if (zero_inactive) s.preds[pi + 1] = 0.0f; }
  } else {  // Match::Predict match.cpp:25-74
    const int k = t - NIND;
    uint32_t len = s.m_len[k];
    const uint32_t cb = s.m_byte[k];
    uint32_t bp = s.m_bitpos[k];
    const int hit = new_bit == ((cb & bp) != 0);
    if (hit) { if (len < 255) ++len; } else len = 0;
    bp >>= 1;
    uint32_t cbyte = cb;
    if (bb) {
      uint32_t cm = s.m_cur[k];
      if (s.hist_len != 0 && cm == s.hist_len - 1) len = 0;
      if (len < 8) {
        const uint32_t idx = s.ctx[s.T.match[k].ctx] & ((1u << s.T.match[k].log2) - 1);
        if (L.match_sid[k]) {
          unsigned long long ent;
          SparseFind(A.map(), SparseKey(L.match_sid[k], idx), &ent);
          cm = ent ? (uint32_t)ent : GMX_IS_OV(L) ? BaseMatch(A, k, idx) : 0u;
        } else {
          cm = A.at<uint32_t>(L.match_tab[k])[idx];
        }
      } else ++cm;
      if (s.hist_len != 0) {
        if (cm >= s.hist_len) { SetError(s, GMX_ERR_MATCH_RANGE); cm = 0; }
        cbyte = HistByte(A, cm);
      }
      s.m_cur[k] = cm;
      bp = 128;
    }
    s.m_len[k] = (uint8_t)len; s.m_byte[k] = (uint8_t)cbyte; s.m_bitpos[k] = (uint8_t)bp;
    const int pi = P_MATCH0 + k;
    if (len > 2) {
      const float mp = A.at<float>(L.match_pred)[k * 256 + len];
      const float p = (cbyte & bp) ? mp : f_sub(1.0f, mp);
      s.preds[pi] = Logit(p);
      SetAct(s, pi, p != 0.5f, ahead);
    } else {
      SetAct(s, pi, false, ahead);
      if (zero_inactive) s.preds[pi] = 0.0f;
    }
  }
}

enum : int { BAR_LOOK = 3 };
// look (LOOK only): bit 0 = this bit's table lookups were already done under the previous bit's network, bit 1 = do the next bit's.
template <int NB, bool PROF, bool LAT = false, bool LOOK = false>
GMX_DEV void PredictBit(StreamSmem& s, const Arena& A, const StreamParams& P, uint32_t b, int path_bit, int btid, Lap<PROF>& lap, int learn_bit = -1, int look = 0) {
  const ArenaLayout& L = *A.L;
  const uint32_t bitctx = s.ctx[C_BIT_CONTEXT];
  const bool zero_inactive = s.analysis != 0;  // predictor.cpp:362-365
  // Work items of the first phase: 41 Indirect + 6 Match lookups, the two byte models' nodes, and the gate selection of
  // the 31 mixers whose gate context does not come out of this phase (their directory loads overlap the table probes).
#pragma unroll 1
  for (int t = btid; t < NIND + NMATCH + 2 + NMIX; t += NB) {
    if (t >= NIND + NMATCH + 2) {
      const int m = t - (NIND + NMATCH + 2);
      if (s.T.mixer[m].ctx != C_LONGEST) GateSelect(s, A, m, s.ctx[s.T.mixer[m].ctx]);
    } else if (t < NIND + NMATCH) {
      if (!(look & 1)) TableLookup(s, A, t, bitctx, s.new_bit, s.bb != 0, zero_inactive, false);
      else if (t < NIND) { const int pi = s.T.ind[t].pred; s.act[pi] >>= 1; s.act[pi + 1] >>= 1; }   // looked up one bit ahead
      else s.act[P_MATCH0 + t - NIND] >>= 1;
    } else {  // per-bit part of ModPPMD / LstmModel::Predict (mod_ppmd.cpp:1662-1681, lstm-model.cpp:36-47)
      const int which = t - (NIND + NMATCH);
      uint32_t fl;
      float val;
      if (path_bit >= 0) {
        const BytePacket& pk = s.pkt[b % PKT_RING];
        val = pk.node[which][path_bit];
        fl = (pk.flags >> (16 * which + 2 * path_bit)) & 3u;
      } else {
        val = IntervalNode(which ? s.lprob : s.ppm, s.recent_bits, &fl);
      }
      if (fl & 1) { s.preds[which] = val; s.act[which] = (fl >> 1) & 1; }
      else { s.act[which] = 0; if (zero_inactive) s.preds[which] = 0.0f; }
    }
  }
  GroupSync<NB>(BAR_BIT);
  lap.mark(3);
  // longest_match = max(match_length / 32) (match.cpp:71-73) and the two mixers gated by it; the layer-0 input vector:
  // the active predictions, inactive ones as +0 (Mixer::Predict sums the active ones in index order; a +-0 product leaves
  // the running sum unchanged, the sum itself is never -0)
#pragma unroll 1
  for (int t = btid; t < 3 + NPRED; t += NB) {
    if (t < 3) {
      uint32_t c = 0;
#pragma unroll 1
      for (int k = 0; k < NMATCH; ++k) { const uint32_t v = s.m_len[k] >> 5; c = v > c ? v : c; }
      if (t == 0) s.ctx[C_LONGEST] = c;
      else GateSelect(s, A, t == 1 ? 6 : NL0 + 6, c);
    } else {
      const int i = t - 3;
      s.xe[i] = s.act[i] ? s.preds[i] : 0.0f;
    }
  }
  GroupSync<NB>(BAR_BIT);
  lap.mark(4);
  // Swap staged weight sets, one queued mixer per round, one float4 per lane: first all write-backs, then all fetches
  // as asynchronous copies, all queued sets in flight together (zero weights == no set yet: the dot product of zeros is
  // +0, exactly the reference's "data == nullptr" output). Pool record = {steps, 0, 0, 0 | weights...}.
  {
    const uint32_t nswap = s.nswap;
    const int lane = btid & 31;
#pragma unroll 1
    for (uint32_t r = btid >> 5; r < nswap; r += NB / 32) WriteBackSet(s, A, s.swap_m[r], s.swap_old[r], s.swap_oldidx[r], lane);
    __syncwarp();   // the same warp re-reads swap_* and overwrites the staged sets below
#pragma unroll 1
    for (uint32_t r = btid >> 5; r < nswap; r += NB / 32) {
      const int m = s.swap_m[r];
      const uint32_t nid = s.swap_new[r];
      if (lane <= (MixerNW(m) + 3) / 4) {
        const float4* rec = nid ? PoolRec(A, nid) : nullptr;
        if (lane == 0) {
          if (nid) CpAsync4(&s.set_steps[m], rec); else s.set_steps[m] = 0u;
          s.set_pool[m] = nid;
          s.set_dirty[m] = 0;
        } else {
          float4* dst = (float4*)(s.w + WOff(m)) + (lane - 1);
          if (nid) CpAsync16(dst, rec + lane); else *dst = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
      }
    }
    CpAsyncWaitAll();
  }
  GroupSync<NB>(BAR_BIT);
  lap.mark(5);
  // Mixer::Predict (mixer.cpp:51-106): the role's first warp, one lane per neuron, sequential sums, the serial
  // same-layer chain is resolved by warp shuffles in neuron order.
  if (btid < 32) {
    const int lane = btid;
    if (lane == 0) s.nswap = 0;
    const float* w = s.w + (lane < NL0 ? lane : 0) * WSTRIDE0;
    float acc = 0.0f;
    {
      const float4* x4 = (const float4*)s.xe;
      const float4* w4 = (const float4*)w;
GMX_UNROLL(LAT ? NPRED / 4 : 2)   // a stream alone on its SM (LAT) is not instruction-cache bound: straight-line code there
      for (int q = 0; q < NPRED / 4; ++q) {
        const float4 x = x4[q], v = w4[q];
        acc = f_add(acc, f_mul(x.x, v.x)); acc = f_add(acc, f_mul(x.y, v.y));
        acc = f_add(acc, f_mul(x.z, v.z)); acc = f_add(acc, f_mul(x.w, v.w));
      }
#pragma unroll
      for (int i = NPRED / 4 * 4; i < NPRED; ++i) acc = f_add(acc, f_mul(s.xe[i], w[i]));
    }
    lap.mark(6);
    __syncwarp();   // converged warp: the shuffles below take their fast path
    {
      // serial chain of the layer: neuron j's finished output feeds every later neuron (mixer.cpp:60-70). Rolled (a
      // handful of instructions that stay in the instruction cache); the chain weight of the next step is loaded from
      // shared memory while the current step's shuffle is in flight.
      if (LAT) {   // the 23 chain weights in registers: one step is shuffle -> mul -> add
        float cw[NL0 - 1];
#pragma unroll
        for (int j = 0; j < NL0 - 1; ++j) cw[j] = w[NPRED + j];
#pragma unroll
        for (int j = 0; j < NL0 - 1; ++j) {
          const float oj = __shfl_sync(0xffffffffu, acc, j);
          if (lane > j && lane < NL0) acc = f_add(acc, f_mul(oj, cw[j]));
        }
      } else {
        float c = w[NPRED];
#pragma unroll 1
        for (int j = 0; j < NL0 - 1; ++j) {
          const float cn = w[NPRED + j + 1];   // j = 22 reads the pad word behind the set: never used
          const float oj = __shfl_sync(0xffffffffu, acc, j);
          if (lane > j && lane < NL0) acc = f_add(acc, f_mul(oj, c));
          c = cn;
        }
      }
    }
    if (lane < NL0) { s.l0_out[lane] = acc; s.xe[NPRED + lane] = acc; }
    __syncwarp();
    lap.mark(7);
    const float skip = s.preds[P_LSTM];
    const float* w1 = s.w + NL0 * WSTRIDE0 + (lane < NL1 ? lane : 0) * WSTRIDE1;
    acc = 0.0f;
    {
      const float4* x4 = (const float4*)s.l0_out;
      const float4* w4 = (const float4*)w1;
GMX_UNROLL(LAT ? NL0 / 4 : 1)
      for (int q = 0; q < NL0 / 4; ++q) {
        const float4 x = x4[q], v = w4[q];
        acc = f_add(acc, f_mul(x.x, v.x)); acc = f_add(acc, f_mul(x.y, v.y));
        acc = f_add(acc, f_mul(x.z, v.z)); acc = f_add(acc, f_mul(x.w, v.w));
      }
    }
    __syncwarp();
    // a layer-1 output is complete (skip connection added last, weights[num_layer0 + output_index],
    // mixer.cpp:77-83) before the next layer-1 neuron reads it
GMX_UNROLL(LAT ? NL1 : 1)
    for (int j = 0; j < NL1; ++j) {
      if (lane == j) acc = f_add(acc, f_mul(skip, w1[NL0 + j]));
      if (j < NL1 - 1) {
        const float oj = __shfl_sync(0xffffffffu, acc, j);
        if (lane > j && lane < NL1) acc = f_add(acc, f_mul(oj, w1[NL0 + j]));
      }
    }
    if (lane < NL1) s.l1_out[lane] = acc;
    __syncwarp();
    lap.mark(8);
    if (lane == 0) {
      const float* w2 = s.w + NL0 * WSTRIDE0 + NL1 * WSTRIDE1;
      float p = 0.0f;
      const float4* x4 = (const float4*)s.l0_out;   // l0_out[24] then l1_out[8]
      const float4* w4 = (const float4*)w2;
GMX_UNROLL(LAT ? (NL0 + NL1) / 4 : 1)
      for (int q = 0; q < (NL0 + NL1) / 4; ++q) {
        const float4 x = x4[q], v = w4[q];
        p = f_add(p, f_mul(x.x, v.x)); p = f_add(p, f_mul(x.y, v.y));
        p = f_add(p, f_mul(x.z, v.z)); p = f_add(p, f_mul(x.w, v.w));
      }
      p = f_add(p, f_mul(skip, w2[NL0 + NL1]));
      s.final_out = p;
      float prob = Logistic(p);
      const float eps = (float)0.0001;
      const float hi = f_sub(1.0f, eps);
      if (prob < eps) prob = eps; else if (prob > hi) prob = hi;
      s.prob = prob;
    }
    lap.mark(9);
  } else if (learn_bit >= 0) {
#pragma unroll 1
    for (int t = btid - 32; t < LEARN_TABLE_ITEMS; t += NB - 32) LearnTables(s, A, learn_bit, t);
    if (LOOK && (look & 2)) {
      // every table model has learned this bit (the barrier also orders their sparse-map insertions); nothing later in this bit
      // reads the lookup state or the table models' predictions (the network and the weight update read xe and bit 0 of act), so
      // the next bit's lookups - memory round trips - run now, beside the first warp's serial mixer network
      GroupSync<NB - 32>(BAR_LOOK);
      const uint32_t next_bitctx = (uint32_t)(s.recent_bits * 2 + learn_bit) - 1u;
#pragma unroll 1
      for (int t = btid - 32; t < NIND + NMATCH; t += NB - 32) TableLookup(s, A, t, next_bitctx, learn_bit, false, zero_inactive, true);
    }
  }
  GroupSync<NB>(BAR_BIT);
  lap.mark(10);
}

// ---- coder (encoder.cpp:8-34, decoder.cpp:3-39) ---------------------------------------------------
GMX_DEV inline uint32_t Discretize(float p) { return (uint32_t)f_add(1.0f, f_mul(65534.0f, p)); }

GMX_DEV inline void PutByte(StreamSmem& s, uint8_t* out, uint32_t b) {
  if (s.out_pos < s.out_cap) out[s.out_pos] = (uint8_t)b; else SetError(s, GMX_ERR_OUTPUT_CAP);
  s.out_pos++;
}
GMX_DEV inline uint32_t GetByte(StreamSmem& s, const uint8_t* in) {  // Decoder::ReadByte: 0 past the end
  const uint32_t b = s.in_pos < s.in_len ? in[s.in_pos] : 0u;
  s.in_pos++;
  return b;
}

// One Encoder::Encode step (encoder.cpp:11-31) with the probability PredictBit left in s.prob.
GMX_DEV inline void EncodeBit(StreamSmem& s, uint8_t* out, int bit) {
  const uint32_t p16 = Discretize(s.prob);
  const uint32_t r = s.x2 - s.x1;
  const uint32_t xmid = s.x1 + (r >> 16) * p16 + (((r & 0xffff) * p16) >> 16);
  if (bit) s.x2 = xmid; else s.x1 = xmid + 1;
  while (((s.x1 ^ s.x2) & 0xff000000u) == 0) { PutByte(s, out, s.x2 >> 24); s.x1 <<= 8; s.x2 = (s.x2 << 8) + 255; }
}

// ---- Predictor::Learn (predictor.cpp:383-387), bit role part (LstmModel::Learn belongs to the LSTM role) -----------
// known_bit >= 0 (compress): the caller has not stored s.new_bit and not run the coder yet; both happen here, as one
// more work item of the first phase, which saves the barrier between coding and learning. code_out = the stream's
// output slice. Returns with s.bit_stop refreshed (role-uniform).
template <int NB, bool PROF, bool LAT = false>
// next_bookkeeping: BasicContexts::Predict of the NEXT bit (Bookkeeping) runs here, on the thread that closes this bit, instead of
// as a one-thread phase with its own barrier in front of the next PredictBit (the hybrid compress order; never after the last
// bit of a stream - the parked state is the one before the next Predict).
GMX_DEV void LearnBit(StreamSmem& s, const Arena& A, const StreamParams& P, int btid, Lap<PROF>& lap, int known_bit = -1, uint8_t* code_out = nullptr,
                      bool tables_done = false, bool next_bookkeeping = false) {
  const ArenaLayout& L = *A.L;
  const int bit = known_bit >= 0 ? known_bit : s.new_bit;
  const float fbit = (float)bit;
  const int cur = s.recent_bits * 2 + bit;
  const bool byte_done = cur >= 256;
  const uint32_t longest = s.ctx[C_LONGEST];
  // history length after BasicContexts::Learn (basic-contexts.cpp:42-54)
  const uint32_t hist_after = s.hist_len + ((byte_done && longest < 2) ? 1u : 0u);
#pragma unroll 1
  for (int t = btid; t < NMIX + LEARN_TABLE_ITEMS + 1; t += NB) {
    if (t < NMIX) {  // Mixer::Learn scalars (mixer.cpp:108-127)
      const int m = t;
      if (s.set_pool[m] == 0) {  // FindOrCreateMixerData mixer.cpp:39-49
        const uint32_t id = atomicAdd(&s.pool_next, 1u);
        if (id >= L.mix_pool_sets) { SetError(s, GMX_ERR_MIXER_POOL); }
        else { s.set_pool[m] = id; DirSet(s, A, m, s.set_idx[m], id); }
      }
      const uint32_t st = s.steps < P.decay_len ? s.steps : P.decay_len - 1;
      float decay = P.decay[st];
      const uint32_t dsteps = s.set_steps[m];
      uint32_t mx = s.max_steps[m];
      decay = (float)d_mul((double)decay, d_sub(1.5, d_div((double)dsteps, (double)mx)));
      const float out = m < NL0 ? s.l0_out[m] : m < NL0 + NL1 ? s.l1_out[m - NL0] : s.final_out;
      const float p = Logistic(out);
      s.upd[m] = f_mul(f_mul(decay, s.T.mixer[m].lr), f_sub(p, fbit));
      const uint32_t nsteps = dsteps + 1;
      s.set_steps[m] = nsteps;
      if (nsteps > mx) s.max_steps[m] = nsteps;
      s.shrink[m] = (nsteps & 1023u) == 0;
    } else if (t < NMIX + LEARN_TABLE_ITEMS) {
      if (!tables_done) LearnTables(s, A, bit, t - NMIX);
    } else if (known_bit >= 0) {   // Encoder::Encode encoder.cpp:8-34 (+ Perceive: predictor.cpp:378-381)
      EncodeBit(s, code_out, bit);
      s.new_bit = bit;
    }
  }
  GroupSync<NB>(BAR_BIT);
  lap.mark(11);
  // Mixer weight updates (mixer.cpp:128-175): w -= update * x over exactly the inputs used by
  // Predict, then the (1 - 3e-6) shrink every 1024 steps of the set.
GMX_UNROLL(LAT ? 3 : 1)
  for (int m = btid >> 5; m < NMIX; m += NB / 32) {
    const int lane = btid & 31;
    const int nw = MixerNW(m);
    const float upd = s.upd[m];
    const bool shrink = s.shrink[m] != 0;
    const float keep = f_sub(1.0f, 3.0e-6f);
    float* w = s.w + WOff(m);
#if GMX_OVERLAY
    if (lane == 0) s.set_dirty[m] = 1;
#endif
    if (m < NL0) {
      // Layer 0: inputs = [90 predictions | outputs of the earlier layer-0 neurons] = s.xe[0 .. nw) (inactive
      // predictions are +0 there and are skipped through the `use` bytes); one float4 per lane covers them all.
      const int left = nw - 4 * lane;   // inputs of this lane's quad that exist
      if (left > 0) {
        float4 v = ((float4*)w)[lane];
        const float4 x = ((const float4*)s.xe)[lane];
        uint32_t on = ((const uint32_t*)s.act)[lane];
        if (left < 4) on &= 0x00ffffffu >> (8 * (3 - left));
        // (bit 0 of an act byte: bit 1 may already hold the next bit's flag, TableLookup ahead)
        if (on & 0x00000001u) v.x = f_sub(v.x, f_mul(upd, x.x));
        if (on & 0x00000100u) v.y = f_sub(v.y, f_mul(upd, x.y));
        if (on & 0x00010000u) v.z = f_sub(v.z, f_mul(upd, x.z));
        if (on & 0x01000000u) v.w = f_sub(v.w, f_mul(upd, x.w));
        if (shrink) {   // applies to every weight of the set (mixer.cpp:170-174), used or not; pad lanes hold 0
          v.x = f_mul(v.x, keep); v.y = f_mul(v.y, keep); v.z = f_mul(v.z, keep); v.w = f_mul(v.w, keep);
        }
        ((float4*)w)[lane] = v;
      }
    } else {
      const int nin = m < NL0 + NL1 ? NL0 + (m - NL0) : NL0 + NL1;  // inputs before the skip connection
#pragma unroll 1
      for (int i = lane; i < nw; i += 32) {
        const float x = i < NL0 ? s.l0_out[i] : i < nin ? s.l1_out[i - NL0] : s.preds[P_LSTM];
        float v = f_sub(w[i], f_mul(upd, x));
        if (shrink) v = f_mul(v, keep);
        w[i] = v;
      }
    }
  }
  if (btid == 0) {
    s.steps++; s.hist_len = hist_after; s.bit_stop = VolatileLoad(&s.error) != 0;
    if (next_bookkeeping) Bookkeeping(s);   // (nothing in the weight update reads the contexts it advances)
  }
  GroupSync<NB>(BAR_BIT);
  lap.mark(12);
}

static GMX_DEV GMX_NOINLINE void Trace(StreamSmem& s, const StreamParams& P, uint64_t bit_index) {
  if (P.bit_trace) {
    const uint32_t p16 = Discretize(s.prob);
    P.bit_trace[bit_index] = (uint64_t)f2u(s.prob) | ((uint64_t)p16 << 32);
  }
  if (P.pred_trace) {
    float* t = P.pred_trace + bit_index * 126;
    for (int i = 0; i < NPRED; ++i) t[i] = s.preds[i];
    uint32_t mask[3] = {0, 0, 0};
    for (int i = 0; i < NPRED; ++i) if (s.act[i]) mask[i >> 5] |= 1u << (i & 31);
    for (int i = 0; i < 3; ++i) t[NPRED + i] = u2f(mask[i]);
    for (int i = 0; i < NL0; ++i) t[93 + i] = s.l0_out[i];
    for (int i = 0; i < NL1; ++i) t[117 + i] = s.l1_out[i];
    t[125] = s.final_out;
  }
}

GMX_DEV inline void WriteUsage(const StreamSmem& s, const Arena& A, const StreamParams& P, uint32_t sid) {
  if (!P.usage) return;
  const PpmdState* ps = A.at<PpmdState>(A.L->p_state);
  uint32_t* u = P.usage + 8 * (size_t)sid;
  u[0] = s.sparse_used; u[1] = s.pool_next;
  u[2] = (ps->lo_unit - PPMD_UNITS_START) + (PPMD_HEAP_END - ps->hi_unit); u[3] = s.hist_len;
  u[4] = SmId(); u[5] = s.t_start_us; u[6] = (uint32_t)(GlobalTimerNs() / 1000ull); u[7] = 0;
}

// Parks the stream's shared-memory state in global memory (with the arena it is the whole stream: what
// Predictor::WriteCheckpoint serialises, checkpoint.h FromArena).
template <int NT>
GMX_DEV void ParkState(const StreamSmem& s, const StreamParams& P, uint32_t sid, int tid) {
  if (!P.final_state) return;
  constexpr int kWords = (int)(sizeof(StreamSmem) / 4);
  uint32_t* dst = P.final_state + (size_t)sid * kWords;
  const uint32_t* sw = (const uint32_t*)&s;
  for (int i = tid; i < kWords; i += NT) dst[i] = sw[i];
}

// ---- what a stream asks of the kernel ----------------------------------------------------------------
enum : int { MODE_COMPRESS = 0, MODE_DECOMPRESS = 1, MODE_GENERATE = 2 };
struct StreamJob {
  const uint8_t* in;       // compress: the stream; decompress: the coded bytes; generate: the prompt
  uint8_t* out;
  uint64_t n_in;
  uint32_t n_bytes;        // compress / decompress: bytes of the stream; generate: bytes to sample
  uint32_t n_preset;       // generate: bytes 0 .. n_preset-1 come from the prompt and are learned
};

// ==== AHEAD pipeline (compress): three roles, each over all bytes of the stream =========================
// PPMd role: byte boundaries 0 .. n-1.
template <bool PROF>
GMX_DEV void PpmdRole(StreamSmem& s, const Arena& A, const StreamJob& J, ProfSmem* prof, int lane) {
  Lap<PROF> lap;
  lap.start(prof, lane == 0);
#pragma unroll 1
  for (uint32_t b = 0; b < J.n_bytes; ++b) {
    if (__shfl_sync(0xffffffffu, VolatileLoad(&s.error), 0)) break;   // warp-uniform stop check
    PpmdStep<PROF>(s, A, b, b ? J.in[b - 1] : s.byte0, (int)J.in[b], true, lane, lap);
  }
}

// LSTM role: forward(b), Perceive(byte b) back to back.
template <int NL, bool PROF>
GMX_DEV void LstmRole(StreamSmem& s, const Arena& A, const StreamParams& P, const StreamJob& J, ProfSmem* prof, int ltid, const WeightSmem& ws) {
  Lap<PROF> lap;
  lap.start(prof, ltid == 0);
  if (ws.w) LoadGateWeights<NL>(s, A, ws, ltid);
#pragma unroll 1
  for (uint32_t b = 0; b < J.n_bytes; ++b) {
    if (ltid == 0) s.lstm_stop = VolatileLoad(&s.error) != 0;
    GroupSync<NL>(BAR_LSTM);
    if (s.lstm_stop) break;
    WaitAtLeast(s, &s.n_ppm, b + 1, 200);
    lap.mark(16);
    LstmForward<NL, PROF>(s, A, P, b, b ? J.in[b - 1] : s.byte0, (int)J.in[b], ltid, lap, ws);
    LstmPerceive<NL, PROF>(s, A, P, J.in[b], ltid, lap, ws);
  }
}

// Bit role: runner_utils::Compress (runner-utils.cpp:43-67); the 5-byte header of RunCompression (:109) is written by
// the kernel entry.
template <int NB, bool PROF, bool LAT = false>
GMX_DEV void BitRoleCompress(StreamSmem& s, const Arena& A, const StreamParams& P, const StreamJob& J, uint32_t sid, ProfSmem* prof, int btid) {
  Lap<PROF> lap;
  lap.start(prof, btid == 0);
  const bool tracing = sid == 0 && (P.bit_trace || P.pred_trace);
#pragma unroll 1
  for (uint32_t pos = 0; pos < J.n_bytes; ++pos) {
    const uint32_t c = J.in[pos];
#pragma unroll 1
    for (int j = 7; j >= 0; --j) {
      const int bit = (c >> j) & 1;
      if (btid == 0) Bookkeeping(s);
      GroupSync<NB>(BAR_BIT);
      lap.mark(0);
      if (s.bb) {
        BitBoundaryA<NB>(s, A, btid);
        lap.mark(2);
        WaitAtLeast(s, &s.n_pkt, pos + 1, 100);
        lap.mark(1);
        BitBoundaryB<NB, LAT>(s, A, pos, btid, (int)c);
        lap.mark(2);
      }
      PredictBit<NB, PROF, LAT>(s, A, P, pos, 7 - j, btid, lap, NB > 32 ? bit : -1);
      if (tracing) {   // debug/parity traces of stream 0 read the blackboard before the mixers learn
        if (btid == 0) Trace(s, P, (uint64_t)pos * 8 + (7 - j));
        GroupSync<NB>(BAR_BIT);
        lap.mark(13);
      }
      LearnBit<NB, PROF, LAT>(s, A, P, btid, lap, bit, J.out, NB > 32);   // codes the bit and learns it
      if (s.bit_stop) return;
    }
    if (btid == 0) Publish(&s.n_done, pos + 1);
  }
}

// Two-role variant (WL = 0): the PPMd warp runs ahead, the other NB threads do the LSTM and the bit path one after the
// other with all of them (forward(b), the 8 bits of byte b, Perceive(b)). PPMd is the one phase a single warp has to
// itself in the serial order (its neighbours would idle at a barrier for ~16 % of the byte); everything else keeps the
// full width.
template <int NB, bool PROF, bool LAT = false>
GMX_DEV void BitLstmRoleCompress(StreamSmem& s, const Arena& A, const StreamParams& P, const StreamJob& J, uint32_t sid, ProfSmem* prof, int btid,
                                 const WeightSmem& ws) {
  Lap<PROF> lap;
  lap.start(prof, btid == 0);
  if (ws.w) LoadGateWeights<NB>(s, A, ws, btid);
  const bool tracing = sid == 0 && (P.bit_trace || P.pred_trace);
#pragma unroll 1
  for (uint32_t pos = 0; pos < J.n_bytes; ++pos) {
    const uint32_t c = J.in[pos];
#pragma unroll 1
    for (int j = 7; j >= 0; --j) {
      const int bit = (c >> j) & 1;
      if (btid == 0) Bookkeeping(s);
      GroupSync<NB>(BAR_BIT);
      lap.mark(0);
      if (s.bb) {
        BitBoundaryA<NB>(s, A, btid);
        lap.mark(2);
        WaitAtLeast(s, &s.n_ppm, pos + 1, 100);
        GroupSync<NB>(BAR_BIT);   // every staged weight set is back in the pool: s.w is free for the forward pass's ring
        lap.mark(1);
        LstmForward<NB, PROF, true>(s, A, P, pos, pos ? J.in[pos - 1] : s.byte0, (int)c, btid, lap, ws);
        BitBoundaryB<NB, LAT>(s, A, pos, btid, (int)c);
        lap.mark(2);
      }
      PredictBit<NB, PROF, LAT>(s, A, P, pos, 7 - j, btid, lap, bit);
      if (tracing) {
        if (btid == 0) Trace(s, P, (uint64_t)pos * 8 + (7 - j));
        GroupSync<NB>(BAR_BIT);
        lap.mark(13);
      }
      LearnBit<NB, PROF, LAT>(s, A, P, btid, lap, bit, J.out, true);
      if (s.bit_stop) return;
    }
    LstmPerceive<NB, PROF>(s, A, P, c, btid, lap, ws);
    if (btid == 0) Publish(&s.n_done, pos + 1);
  }
}

// Hybrid order (SERIAL with WL = 0): the PPMd warp is ahead by ONE byte and is not lost to the LSTM. Per byte: all NT threads
// run the LSTM forward pass; then the PPMd warp prepares the NEXT byte (UpdateByte with this byte, PrepareByte, normalisation,
// path nodes - the phase that is one warp wide by nature) while the other NT - 32 threads run this byte's eight bit steps (whose
// phases have at most 47 work items or are serial chains); then all NT threads run Lstm::Perceive. Against the phase-serial
// order this hides PPMd behind the bit path; against the two-role variant the LSTM phases, whose time is inversely proportional
// to their threads, keep the full CTA. Hand-overs: two CTA barriers per byte (forward: the distribution and the freed weight-set
// staging area; Perceive: the byte is done), packets of consecutive bytes alternate ring slots.
template <int NT, bool PROF, bool LAT = false>
GMX_DEV void HybridCompress(StreamSmem& s, const Arena& A, const StreamParams& P, const StreamJob& J, uint32_t sid, ProfSmem* prof, int tid,
                            const WeightSmem& ws) {
  constexpr int NB = NT - 32;
  const bool bitw = tid < NB;
  const int lane = tid - NB;
  Lap<PROF> lap, plap;
  lap.start(prof, tid == 0);        // bit + LSTM slots
  plap.start(prof, tid == NB);      // PPMd slots (24..26); the PPMd warp's share of the LSTM phases is not counted twice
  Lap<PROF> nolap;
  nolap.start(prof, false);
  if (ws.w) LoadGateWeights<NT>(s, A, ws, tid);
  const bool tracing = sid == 0 && (P.bit_trace || P.pred_trace);
  if (!bitw && J.n_bytes) PpmdStep<PROF>(s, A, 0, s.byte0, (int)J.in[0], true, lane, plap);
#pragma unroll 1
  for (uint32_t pos = 0; pos < J.n_bytes; ++pos) {
    const uint32_t c = J.in[pos];
    if (bitw) {
      if (pos == 0 || tracing) {   // later bits: folded into the previous bit's LearnBit
        if (tid == 0) Bookkeeping(s);
        GroupSync<NB>(BAR_BIT);
      }
      lap.mark(0);
      BitBoundaryA<NB>(s, A, tid);   // (the first bit of a byte: s.bb is set)
      lap.mark(2);
    }
    __syncthreads();   // the distribution of this byte is published; every staged weight set is back in the pool (s.w = the forward pass's ring)
    if (bitw) lap.mark(1);
    LstmForward<NT, PROF, true>(s, A, P, pos, pos ? J.in[pos - 1] : s.byte0, (int)c, tid, bitw ? lap : nolap, ws);
    // (Measured and not kept, profiles/r02_hybrid_compress.md: the next byte's gate products on the PPMd warp - one warp needs
    // longer for them than the bit path lasts, 3.64 instead of 4.58 MB/s; the fused output-layer step on the PPMd warp - no change.)
    if (!bitw) {
      if (pos + 1 < J.n_bytes && !__shfl_sync(0xffffffffu, VolatileLoad(&s.error), 0))
        PpmdStep<PROF>(s, A, pos + 1, c, (int)J.in[pos + 1], true, lane, plap);
    } else {
      BitBoundaryB<NB, LAT>(s, A, pos, tid, (int)c);
      lap.mark(2);
#pragma unroll 1
      for (int j = 7; j >= 0; --j) {
        const int bit = (c >> j) & 1;
        if (j != 7 && tracing) {
          if (tid == 0) Bookkeeping(s);
          GroupSync<NB>(BAR_BIT);
          lap.mark(0);
        }
        const int look = tracing ? 0 : ((j != 7 ? 1 : 0) | (j != 0 ? 2 : 0));   // table lookups of bits 1..7 run one bit ahead
        PredictBit<NB, PROF, LAT, true>(s, A, P, pos, 7 - j, tid, lap, bit, look);
        if (tracing) {
          if (tid == 0) Trace(s, P, (uint64_t)pos * 8 + (7 - j));
          GroupSync<NB>(BAR_BIT);
          lap.mark(13);
        }
        LearnBit<NB, PROF, LAT>(s, A, P, tid, lap, bit, J.out, true, !tracing && !(j == 0 && pos + 1 == J.n_bytes));
        if (s.bit_stop) break;
      }
    }
    __syncthreads();
    if (VolatileLoad(&s.error) != 0) return;   // uniform: nothing between the barrier and Perceive's first barrier sets it
    if (bitw) lap.mark(1);
    LstmPerceive<NT, PROF>(s, A, P, c, tid, bitw ? lap : nolap, ws);
    if (tid == 0) Publish(&s.n_done, pos + 1);
  }
}

// ==== LOCKSTEP (decompress, generation, the Predictor facade): a byte is only known when its last bit is decided, so
// nothing can run ahead; the whole CTA (NT threads) walks the phases one after the other, every phase with all the
// threads it can use, through the same step functions as the roles above. ===================================

// Predictor::Predict of one bit. known_byte >= 0 (serial compress): the byte being coded; path_bit = index of this bit in
// it. The byte models then leave their eight path nodes in packet 0 at the byte boundary.
// part (lock-step generation, always at a byte boundary): 1 = up to the LSTM input vector, 2 = from the gate pre-activations on.
template <int NT, bool PROF, bool LAT = false>
GMX_DEV void SerialPredict(StreamSmem& s, const Arena& A, const StreamParams& P, int tid, Lap<PROF>& lap, const WeightSmem& ws, int known_byte = -1,
                           int path_bit = -1, int learn_bit = -1, int part = 0) {
  if (part != 2) {
    if (tid == 0) Bookkeeping(s);
    __syncthreads();
    lap.mark(0);
  }
  if (part == 2 || s.bb) {
    const uint32_t last = s.ctx[C_LAST_BYTE];
    if (part != 2) {
      // PPMd on the last warp while the other warps do the byte contexts (with b = 0 no wait inside can block)
      if (tid >= NT - 32) PpmdStep<PROF>(s, A, 0, last, known_byte, known_byte >= 0, tid - (NT - 32), lap);
      else BitBoundaryA<NT - 32>(s, A, tid);
      __syncthreads();
      lap.mark(2);
    }
    LstmForward<NT, PROF, true>(s, A, P, 0, last, known_byte, tid, lap, ws, part);
    if (part == 1) return;
    BitBoundaryB<NT, LAT>(s, A, 0, tid, known_byte);
  } else if (part == 1) {
    if (tid == 0) SetError(s, GMX_ERR_INTERNAL);   // lock-step hand-over off a byte boundary
    return;
  }
  PredictBit<NT, PROF, LAT>(s, A, P, 0, path_bit, tid, lap, learn_bit);
}
// Predictor::Learn of one bit (s.new_bit, or known_bit which is then also coded into code_out first).
template <int NT, bool PROF, bool LAT = false>
GMX_DEV void SerialLearn(StreamSmem& s, const Arena& A, const StreamParams& P, int tid, Lap<PROF>& lap, const WeightSmem& ws, int known_bit = -1,
                         uint8_t* code_out = nullptr) {
  const int cur = s.recent_bits * 2 + (known_bit >= 0 ? known_bit : s.new_bit);
  LearnBit<NT, PROF, LAT>(s, A, P, tid, lap, known_bit, code_out, known_bit >= 0);   // (SerialPredict ran the table models' Learn already when it knew the bit)
  if (cur >= 256) LstmPerceive<NT, PROF>(s, A, P, (uint32_t)(cur - 256), tid, lap, ws);   // LstmModel::Learn lstm-model.cpp:50-59
}

// Predictor::RunAnalysis(bit) (predictor.cpp:471-504) with UpdateEntropy (:437-469), between Predict and Learn of a bit: the
// blackboard holds this bit's predictions (inactive ones zeroed: analysis is on), s.steps == bits_seen. One thread per column.
// log2 is the one operation here that is not bit-pinned to the host libm (CUDA's is within 1 ulp); the reference rounds its
// result to float before it enters the double accumulator, which absorbs that.
static GMX_DEV GMX_NOINLINE void AnalysisStep(StreamSmem& s, const Arena& A, const StreamParams& P, int bit, int tid) {
  if (tid < AN_COLS) {
    const float v = tid < 2 ? s.preds[tid] : tid < AN_COLS - 1 ? s.preds[P_IND0 + 2 * 17 + (tid - 2)] : s.final_out;
    float prob = Logistic(v);
    const float eps = (float)0.01, hi = f_sub(1.0f, eps);
    if (prob < eps) prob = eps; else if (prob > hi) prob = hi;
    const float e = (float)log2((double)(bit ? prob : f_sub(1.0f, prob)));
    const double alpha = 0.00001;
    const double acc = d_add(d_mul(d_sub(1.0, alpha), P.an_entropy[tid]), d_mul(alpha, (double)e));
    P.an_entropy[tid] = acc;
    const uint64_t bits_seen = s.steps;
    if (bits_seen > 0 && bits_seen % P.an_freq == 0 && bits_seen / P.an_freq <= P.an_max_rows) {
      AnalysisRow& row = P.an_rows[bits_seen / P.an_freq - 1];
      row.neg_entropy[tid] = -acc;
      if (tid == 0) {
        const PpmdState* ps = A.at<PpmdState>(A.L->p_state);
        uint64_t used = (uint64_t)PPMD_HEAP_END - (ps->hi_unit - ps->lo_unit) - (ps->units_start - ps->text_ptr);   // GetUsedMemory mod_ppmd.cpp:142-149
        for (uint32_t i = 0; i < PPMD_N_INDEXES; ++i) used -= (uint64_t)ps->indx2units[i] * ps->bl_stamp[i] * 12u;
        row.bits_seen = bits_seen; row.ppmd_used = used; row.history = s.hist_len;
      }
    }
  }
  __syncthreads();
}

// runner_utils::Compress (runner-utils.cpp:43-67) without the role pipeline: all phases with all threads. With a full
// wave of resident streams per SM the other streams already hide this stream's latencies, and every phase having all
// threads beats the pipeline's fixed split of them (kernels.h: configurations).
template <int NT, bool PROF, bool LAT = false>
GMX_DEV void SerialCompress(StreamSmem& s, const Arena& A, const StreamParams& P, const StreamJob& J, uint32_t sid, ProfSmem* prof, int tid,
                            const WeightSmem& ws) {
  Lap<PROF> lap;
  lap.start(prof, tid == 0);
  if (ws.w) LoadGateWeights<NT>(s, A, ws, tid);
  const bool tracing = sid == 0 && (P.bit_trace || P.pred_trace);
#pragma unroll 1
  for (uint32_t pos = 0; pos < J.n_bytes; ++pos) {
    const uint32_t c = J.in[pos];
#pragma unroll 1
    for (int j = 7; j >= 0; --j) {
      SerialPredict<NT, PROF, LAT>(s, A, P, tid, lap, ws, (int)c, 7 - j, (c >> j) & 1);
      if (P.an_rows) AnalysisStep(s, A, P, (c >> j) & 1, tid);
      if (tracing) {
        if (tid == 0) Trace(s, P, (uint64_t)pos * 8 + (7 - j));
        __syncthreads();
      }
      SerialLearn<NT, PROF, LAT>(s, A, P, tid, lap, ws, (c >> j) & 1, J.out);
      if (s.bit_stop) return;
    }
  }
}

// runner_utils::Decompress (runner-utils.cpp:69-86); Decoder::Decode decoder.cpp:19-39. Analysis is never on.
template <int NT, bool PROF, bool LAT = false>
GMX_DEV void SerialDecompress(StreamSmem& s, const Arena& A, const StreamParams& P, const StreamJob& J, ProfSmem* prof, int tid, const WeightSmem& ws) {
  Lap<PROF> lap;
  lap.start(prof, tid == 0);
  if (ws.w) LoadGateWeights<NT>(s, A, ws, tid);
#pragma unroll 1
  for (uint32_t pos = 0; pos < J.n_bytes; ++pos) {
#pragma unroll 1
    for (int j = 7; j >= 0; --j) {
      SerialPredict<NT, PROF, LAT>(s, A, P, tid, lap, ws);
      if (tid == 0) {
        const uint32_t p16 = Discretize(s.prob);
        const uint32_t r = s.x2 - s.x1;
        const uint32_t xmid = s.x1 + (r >> 16) * p16 + (((r & 0xffff) * p16) >> 16);
        int bit = 0;
        if (s.x <= xmid) { bit = 1; s.x2 = xmid; } else s.x1 = xmid + 1;
        s.new_bit = bit;
        while (((s.x1 ^ s.x2) & 0xff000000u) == 0) { s.x1 <<= 8; s.x2 = (s.x2 << 8) + 255; s.x = (s.x << 8) + GetByte(s, J.in); }
        if (j == 0) J.out[pos] = (uint8_t)((s.recent_bits * 2 + bit) & 0xff);
      }
      __syncthreads();
      SerialLearn<NT, PROF, LAT>(s, A, P, tid, lap, ws);
      if (s.bit_stop) return;
    }
  }
}

// runner_utils::RunGeneration (runner-utils.cpp:158-221): the prompt (all but its last byte) is consumed WITH
// learning, then n_bytes bytes are sampled bit by bit without Learn: prob = Logistic(Logit(prob) / temperature),
// bit = r < prob with r the next rand()/RAND_MAX draw, Perceive(bit), Predict().
template <int NT, bool PROF, bool LAT = false>
GMX_DEV void SerialGenerate(StreamSmem& s, const Arena& A, const StreamParams& P, const StreamJob& J, const float* ru, ProfSmem* prof, int tid,
                            const WeightSmem& ws) {
  Lap<PROF> lap;
  lap.start(prof, tid == 0);
  if (ws.w) LoadGateWeights<NT>(s, A, ws, tid);
#pragma unroll 1
  for (uint32_t pos = 0; pos < J.n_preset; ++pos) {   // :187-194
    const uint32_t c = J.in[pos];
#pragma unroll 1
    for (int j = 7; j >= 0; --j) {
      SerialPredict<NT, PROF, LAT>(s, A, P, tid, lap, ws);
      if (tid == 0) s.new_bit = (c >> j) & 1;
      __syncthreads();
      SerialLearn<NT, PROF, LAT>(s, A, P, tid, lap, ws);
      if (s.bit_stop) return;
    }
  }
  if (P.lockstep) {   // the sampling phase belongs to GenStepKernel + the batched gate product
    SerialPredict<NT, PROF, LAT>(s, A, P, tid, lap, ws, -1, -1, -1, 1);
    return;
  }
  SerialPredict<NT, PROF, LAT>(s, A, P, tid, lap, ws);   // :198
#pragma unroll 1
  for (uint32_t i = 0; i < J.n_bytes; ++i) {
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      if (tid == 0) {
        const float r = ru[(size_t)i * 8 + j];
        const float prob = Logistic(f_div(Logit(s.prob), P.temperature));
        s.new_bit = r < prob ? 1 : 0;
        if (j == 7) J.out[i] = (uint8_t)((s.recent_bits * 2 + s.new_bit) & 0xff);
        s.bit_stop = VolatileLoad(&s.error) != 0;
      }
      __syncthreads();
      if (s.bit_stop) return;
      SerialPredict<NT, PROF, LAT>(s, A, P, tid, lap, ws);
    }
  }
}

// Copies the launch's read-only tables into shared memory (StreamTables).
template <int NT>
GMX_DEV void StageTables(StreamSmem& s, const StreamParams& P, int tid) {
  const uint32_t* src = (const uint32_t*)P.layout;
  uint32_t* dst = (uint32_t*)&s.T.L;
  for (int i = tid; i < (int)(sizeof(ArenaLayout) / 4); i += NT) dst[i] = src[i];
  for (int i = tid; i < NIND; i += NT) s.T.ind[i] = kInd[i];
  for (int i = tid; i < 20; i += NT) s.T.skip[i] = kSkip[i];
  for (int i = tid; i < 9; i += NT) s.T.interval[i] = kInterval[i];
  for (int i = tid; i < NIH; i += NT) s.T.ih[i] = kIH[i];
  for (int i = tid; i < NMATCH; i += NT) s.T.match[i] = kMatch[i];
  for (int i = tid; i < NMIX; i += NT) s.T.mixer[i] = kMixer[i];
  for (int i = tid; i < 512; i += NT) s.T.nonstationary[i] = kNonstationary[i];
  __syncthreads();
}

// ---- kernel entry: persistent CTAs, one stream at a time, ids from an atomic queue ---------------
// WB warps of bit role (threads 0 .. 32 WB - 1), WL warps of LSTM role, one PPMd warp.
// SERIAL: compress without the role pipeline (the same NT threads walk the phases together, as the lockstep modes do).
// WS: the dense gate weights are resident in W_DENSE_BYTES + 16 bytes of dynamic shared memory (one CTA per SM).
template <int WB, int WL, int MODE, int MINB, bool PROF, bool SERIAL = false, bool WS = false>
__global__ void __launch_bounds__(32 * (WB + WL + 1), MINB) StreamKernel(StreamParams P) {
  constexpr int NB = 32 * WB, NL = 32 * WL, NT = NB + NL + 32;
  __shared__ StreamSmem s;
#if defined(__CUDACC__)
  extern __shared__ __align__(128) unsigned char dyn_smem[];
#else
  static unsigned char dyn_smem[WS ? W_DENSE_BYTES + 16 : 16] __attribute__((aligned(128)));
#endif
  WeightSmem ws{WS ? (float4*)dyn_smem : nullptr, WS ? (uint64_t*)(dyn_smem + W_DENSE_BYTES) : nullptr};
  __shared__ ProfHolder<PROF> prof_mem;
  __shared__ uint32_t next_stream;
  __shared__ StreamJob job;
  const int tid = (int)threadIdx.x;
  ProfSmem* prof = prof_mem.get();
  if (WS && tid == 0) { MbarInit(ws.mbar, 1); s.wphase = 0; }
  StageTables<NT>(s, P, tid);
  Arena A{P.arenas + (uint64_t)blockIdx.x * P.arena_stride, &s.T.L, P.tmpl_arena, P.tmpl_layout};
  for (uint32_t round = 0;; ++round) {
    const bool lockstep = MODE == MODE_GENERATE && P.lockstep;   // block b runs stream stream_base + b in arena b, once
    if (!lockstep && tid == 0) next_stream = atomicAdd(P.queue, 1u);
    __syncthreads();
    const uint32_t q = lockstep ? (round == 0 ? P.stream_base + blockIdx.x : P.n_streams) : next_stream;
    __syncthreads();
    if (q >= P.n_streams) break;
    const uint32_t sid = P.ids ? P.ids[q] : q;
    long long t_init = 0;
    if (PROF && tid == 0) { for (int i = 0; i < GMX_PROF_SLOTS; ++i) prof->acc[i] = 0; t_init = GMX_CLOCK(); }
    InitStream<NT>(s, A, P, tid);
    if (tid == 0) {
      const uint8_t* in = P.in + P.in_off[sid];
      const uint64_t n = P.in_off[sid + 1] - P.in_off[sid];
      job.in = in; job.n_in = n; job.n_preset = 0;
      if (MODE == MODE_COMPRESS) {
        uint8_t* out = P.out + P.out_off[sid];
        job.out = out; job.n_bytes = (uint32_t)n;
        s.out_pos = 0; s.out_cap = P.out_off[sid + 1] - P.out_off[sid];
        const uint64_t announced = P.part ? P.part_total : n;
        s.analysis = P.analysis >= 0 ? P.analysis : (8 * announced / 1000) > 0;  // EnableAnalysis(8*n/1000) -> predictions zeroed every bit
        if (!P.part || P.part_header) for (int i = 4; i >= 0; --i) PutByte(s, out, (uint32_t)(announced >> (8 * i)) & 0xff);
        if (P.part && P.coder_in) { s.x1 = P.coder_in[0]; s.x2 = P.coder_in[1]; }
      } else if (MODE == MODE_DECOMPRESS && P.part && !P.part_header) {   // a later part of a stream: the decoder continues
        job.out = P.out + P.out_off[sid];
        s.in_pos = 0; s.in_len = n;
        s.out_pos = 0; s.out_cap = P.part_total;
        s.analysis = P.analysis >= 0 ? P.analysis : 0;
        if (P.part_total > P.out_off[sid + 1] - P.out_off[sid]) SetError(s, GMX_ERR_OUTPUT_CAP);
        if (P.coder_in) { s.x1 = P.coder_in[0]; s.x2 = P.coder_in[1]; s.x = P.coder_in[2]; }
        job.n_bytes = (uint32_t)P.part_total;
      } else if (MODE == MODE_DECOMPRESS) {   // ReadHeader runner-utils.cpp:29-36, Decoder::Decoder decoder.cpp:3-9
        job.out = P.out + P.out_off[sid];
        s.in_pos = 0; s.in_len = n;
        s.out_pos = 0; s.out_cap = P.out_off[sid + 1] - P.out_off[sid];
        s.analysis = P.analysis >= 0 ? P.analysis : 0;
        uint64_t len = 0;
        for (int i = 0; i <= 4; ++i) len = (len << 8) + GetByte(s, in);
        if (P.part && P.part_total <= len) len = P.part_total;   // first part of a stream coded in parts: only this many bytes now
        if (s.in_len < 5 || len > s.out_cap) { SetError(s, s.in_len < 5 ? GMX_ERR_BAD_HEADER : GMX_ERR_OUTPUT_CAP); len = 0; }
        s.out_cap = len;  // number of bytes to produce
        for (int i = 0; i < 4; ++i) s.x = (s.x << 8) + (GetByte(s, in) & 0xff);
        job.n_bytes = (uint32_t)len;
      } else {
        job.out = P.out + (size_t)sid * P.gen_bytes;
        job.n_preset = n ? (uint32_t)n - 1 : 0; job.n_bytes = P.gen_bytes;
        s.analysis = P.analysis >= 0 ? P.analysis : (8 * (uint64_t)P.gen_bytes / 1000) > 0;   // :177
      }
      if (PROF) prof->acc[14] += (unsigned long long)(GMX_CLOCK() - t_init);
    }
    __syncthreads();
    const bool failed_early = s.error != 0;
    if (!failed_early) {
      if (MODE == MODE_COMPRESS && SERIAL && WL == 0) {
        HybridCompress<NT, PROF, MINB == 1>(s, A, P, job, sid, prof, tid, ws);
      } else if (MODE == MODE_COMPRESS && SERIAL) {
        SerialCompress<NT, PROF, MINB == 1>(s, A, P, job, sid, prof, tid, ws);
      } else if (MODE == MODE_COMPRESS && WL == 0) {
        if (tid < NB) BitLstmRoleCompress<NB, PROF, MINB == 1>(s, A, P, job, sid, prof, tid, ws);
        else PpmdRole<PROF>(s, A, job, prof, tid - NB);
      } else if (MODE == MODE_COMPRESS) {
        if (tid < NB) BitRoleCompress<NB, PROF, MINB == 1>(s, A, P, job, sid, prof, tid);
        else if (tid < NB + NL) LstmRole<(NL > 0 ? NL : 32), PROF>(s, A, P, job, prof, tid - NB, ws);
        else PpmdRole<PROF>(s, A, job, prof, tid - NB - NL);
      } else if (MODE == MODE_DECOMPRESS) {
        SerialDecompress<NT, PROF, MINB == 1>(s, A, P, job, prof, tid, ws);
      } else {
        SerialGenerate<NT, PROF, MINB == 1>(s, A, P, job, P.rand_u + (size_t)sid * P.rand_stride, prof, tid, ws);
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (P.part && P.coder_out) { P.coder_out[0] = s.x1; P.coder_out[1] = s.x2; P.coder_out[2] = s.x; P.coder_out[3] = (uint32_t)s.in_pos; }
      if (MODE == MODE_COMPRESS) {   // Encoder::Flush encoder.cpp:27-34
        uint8_t* out = job.out;
        if (!P.part || P.part_last) {
          while (((s.x1 ^ s.x2) & 0xff000000u) == 0) { PutByte(s, out, s.x2 >> 24); s.x1 <<= 8; s.x2 = (s.x2 << 8) + 255; }
          PutByte(s, out, s.x2 >> 24);
        }
        P.out_len[sid] = s.out_pos;
      } else if (MODE == MODE_DECOMPRESS) {
        P.out_len[sid] = s.error ? 0 : job.n_bytes;
      } else {
        P.out_len[sid] = s.error ? 0 : P.gen_bytes;
      }
      P.status[sid] = s.error;
      WriteUsage(s, A, P, sid);
      if (PROF && P.prof) for (int i = 0; i < GMX_PROF_SLOTS; ++i) P.prof[(size_t)sid * GMX_PROF_SLOTS + i] = prof->acc[i];
    }
    __syncthreads();
    ParkState<NT>(s, P, lockstep ? blockIdx.x : sid, tid);
    __syncthreads();
  }
}

// ---- lock-step generation: one sampled byte of every stream per launch ----------------------------------------------------
// Block b = stream stream_base + b in arena b, state parked in P.park between launches. A launch starts where the previous one
// (or the prompt launch of StreamKernel<MODE_GENERATE> with P.lockstep) stopped: the byte models have run for byte
// `byte_index`, the batched kernel has left this stream's 150 gate pre-activations in P.gate_g. It finishes the LSTM forward
// pass and the byte boundary, samples the byte's eight bits exactly as SerialGenerate does (runner-utils.cpp:198-215) and -
// unless this was the last byte - runs the byte models for the next byte up to the next gate product.
struct GenStepParams {
  StreamParams P;
  uint32_t n_slots, byte_index, last;
};
template <int WB, int WL, int MINB>
__global__ void __launch_bounds__(32 * (WB + WL + 1), MINB) GenStepKernel(GenStepParams Q) {
  constexpr int NB = 32 * WB, NL = 32 * WL, NT = NB + NL + 32;
  constexpr int kWords = (int)(sizeof(StreamSmem) / 4);
  __shared__ StreamSmem s;
  const StreamParams& P = Q.P;
  const int tid = (int)threadIdx.x;
  const uint32_t slot = blockIdx.x;
  if (slot >= Q.n_slots) return;
  const uint32_t sid = P.stream_base + slot;
  static_assert(sizeof(StreamSmem) % 16 == 0, "the parked state moves as 16-byte words");
  uint4* sw = (uint4*)&s;
  uint4* parked = (uint4*)(P.park + (size_t)slot * kWords);
  for (int i = tid; i < kWords / 4; i += NT) sw[i] = parked[i];
  __syncthreads();
  Arena A{P.arenas + (uint64_t)slot * P.arena_stride, &s.T.L, P.tmpl_arena, P.tmpl_layout};
  Lap<false> lap;
  lap.start(nullptr, false);
  const WeightSmem ws{nullptr, nullptr};
  if (s.error == 0) {
    SerialPredict<NT, false>(s, A, P, tid, lap, ws, -1, -1, -1, 2);
    const float* ru = P.rand_u + (size_t)sid * P.rand_stride + (size_t)Q.byte_index * 8;
    uint8_t* out = P.out + (size_t)sid * P.gen_bytes;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      if (tid == 0) {
        const float r = ru[j];
        const float prob = Logistic(f_div(Logit(s.prob), P.temperature));
        s.new_bit = r < prob ? 1 : 0;
        if (j == 7) out[Q.byte_index] = (uint8_t)((s.recent_bits * 2 + s.new_bit) & 0xff);
        s.bit_stop = VolatileLoad(&s.error) != 0;
      }
      __syncthreads();
      if (s.bit_stop) break;
      if (j < 7) SerialPredict<NT, false>(s, A, P, tid, lap, ws);
      else if (!Q.last) SerialPredict<NT, false>(s, A, P, tid, lap, ws, -1, -1, -1, 1);
    }
  }
  __syncthreads();
  for (int i = tid; i < kWords / 4; i += NT) parked[i] = sw[i];
  if (tid == 0) {
    P.status[sid] = s.error;
    P.out_len[sid] = s.error ? 0 : P.gen_bytes;
    if (Q.last) WriteUsage(s, A, P, sid);
  }
}

// ---- single-stream stepping (the Predictor facade: reference src/predictor.h:20-38) ---------------
// One launch per Predict() / Learn() call of ONE stream; the shared-memory state of the stream is
// parked in global memory between launches. Perceive(bit) travels with the next launch. The three roles run one after
// the other here (CTA-wide barriers), through the same step functions as the stream kernels.
enum : int { STEP_INIT = 0, STEP_PREDICT = 1, STEP_LEARN = 2 };
struct StepParams {
  StreamParams P;          // arenas/layout/tables of the one stream (n_streams, in/out unused)
  uint32_t* state;         // sizeof(StreamSmem) bytes
  int op, has_bit, bit, analysis;
  float* prob_out;         // STEP_PREDICT: what Predictor::Predict() returns
  uint32_t* status_out;    // stream error code after the step
};

template <int WB, int WL>
__global__ void __launch_bounds__(32 * (WB + WL + 1)) StepKernel(StepParams Q) {
  constexpr int NB = 32 * WB, NL = 32 * WL, NT = NB + NL + 32;
  __shared__ StreamSmem s;
  const int tid = (int)threadIdx.x;
  uint32_t* sw = (uint32_t*)&s;
  constexpr int kWords = (int)(sizeof(StreamSmem) / 4);
  Lap<false> lap;
  lap.start(nullptr, false);
  if (Q.op == STEP_INIT) {
    StageTables<NT>(s, Q.P, tid);
    Arena A{Q.P.arenas, &s.T.L};
    InitStream<NT>(s, A, Q.P, tid);
    if (tid == 0) s.analysis = Q.analysis;
  } else {
    for (int i = tid; i < kWords; i += NT) sw[i] = Q.state[i];
    __syncthreads();
    Arena A{Q.P.arenas, &s.T.L};
    if (tid == 0 && Q.has_bit) s.new_bit = Q.bit;   // Predictor::Perceive predictor.cpp:378-381
    if (tid == 0 && Q.analysis >= 0) s.analysis = Q.analysis;
    __syncthreads();
    if (Q.op == STEP_PREDICT) {
      SerialPredict<NT, false>(s, A, Q.P, tid, lap, WeightSmem{nullptr, nullptr});
      __syncthreads();
      if (tid == 0) *Q.prob_out = s.prob;
    } else {
      SerialLearn<NT, false>(s, A, Q.P, tid, lap, WeightSmem{nullptr, nullptr});
    }
  }
  __syncthreads();
  for (int i = tid; i < kWords; i += NT) Q.state[i] = sw[i];
  if (tid == 0) *Q.status_out = s.error;
}

}  // namespace gmx
#endif  // GMIX_B200_STREAM_KERNEL_CUH_
