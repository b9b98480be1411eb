// One CTA per stream: the complete gmix bit step (reference src/predictor.cpp:360-387 and every
// Model::Predict/Learn it calls) plus the binary arithmetic coder (src/coder/*.cpp), for
// compress and decompress. Persistent CTAs pull stream ids from an atomic queue and reuse their
// arena (global-memory workspace) across streams.
//
// Exactness rules: every fp32 operation goes through
// gmx::f_* (single IEEE rounding, never contracted), every dot product is accumulated by ONE
// thread in the reference's index order, libm calls go through gmx::gm_* (dmath.cuh). Parallelism
// inside a stream is only across independent quantities (the 41 Indirect models, the 24 layer-0
// mixers, the 150 LSTM gate rows, the 2697 mixer weights, ...).
//
// The file is plain CUDA C++ that also compiles for the host under tests/emu/cuda_emu.h (a fiber
// based SIMT emulator used by the CPU test-suite to debug the kernel logic without a GPU). The
// shipped library only contains the nvcc build.
#ifndef GMIX_B200_STREAM_KERNEL_CUH_
#define GMIX_B200_STREAM_KERNEL_CUH_
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
};

// Shared-memory staging of the 33 selected weight sets: layer-0 sets (<= 113 weights) at stride 132 words,
// layer-1/final sets (<= 33 weights) at stride 36 words. Both strides are 16-byte multiples (float4
// loads) and = 4 mod 32 banks, which makes lane-per-neuron float4 reads conflict free.
enum : int { WSTRIDE0 = 132, WSTRIDE1 = 36, WTOTAL = 24 * WSTRIDE0 + 9 * WSTRIDE1 };
#define GMX_PRAGMA_(x) _Pragma(#x)
#define GMX_UNROLL(n) GMX_PRAGMA_(unroll n)
#if defined(GMX_LSTM_STREAM)
#define GMX_LSTM_LOAD(p) LoadStream4(p)   // gate weights as streaming (evict-first) traffic
#else
#define GMX_LSTM_LOAD(p) (*(p))
#endif
#ifndef GMX_BPTT_UNROLL
#define GMX_BPTT_UNROLL 4
#endif
// L2 prefetches one step ahead (next-bit sparse slots, next-bit / next-byte weight sets, the output layer before
// Perceive and between BPTT epochs) all LOSE 0.5-2.4 % each with 8 CTAs per SM (A/B in profiles/r01_s3_ab.md): the other
// resident streams already cover the latency and the extra requests only queue in front of demand loads. They stay
// available as GMX_PF_* build flags for low-occupancy use.
// GMX_NEW_FRONT (experimental, OFF): gate selection and weight-set swaps on the last warp during the lookup phase
// (MixerFrontEarly / MixerFrontLate). +1.3 % at 8 CTAs/SM; bit-exact under the CPU emulator (both cp.async timings,
// three thread orders) and in the GPU checks it was given (golden vectors, 300 x 8 KiB against the oracle), but the
// complete GPU suite has not been run with it yet (profiles/r01_s3_ab.md), so the product build keeps the old phases.
#if !defined(GMX_NEW_FRONT) && !defined(GMX_OLD_FRONT)
#define GMX_OLD_FRONT 1
#endif
#if defined(GMX_CAND_STAGE) && !defined(GMX_OLD_FRONT)
#define GMX_OLD_FRONT 1       // candidate staging belongs to the old gate-selection / swap phases
#endif
#if !defined(GMX_CAND_STAGE)
#define GMX_NO_CAND_STAGE 1   // staging both candidate weight sets one bit ahead wins 11 % at 1 CTA/SM and loses 2 % at 8 (A/B in profiles/)
#endif
#ifndef LSTM_STAGES
#define LSTM_STAGES 5       // float4 pairs in flight per gate-row pair of the LSTM forward pass (ring in s.w: 5 x 2 x 75 x 16 B = 12 KB)
#endif
#ifndef GMX_LSTM_UNROLL
#define GMX_LSTM_UNROLL 2   // quads in flight per gate row of the LSTM forward pass (register budget: 64)
#endif
enum : int { GMX_PROF_SLOTS = 24 };
enum : int { NBITMIX = 4, CAND_Q = 68 };   // bit-gated mixers; float4s per staged candidate: (1 header + ceil(weights / 4)) summed over the four
// Phase slots: 0 byte contexts+PPMd, 1 ppm normalise, 2 LSTM forward, 3 interval nodes, 4 indirect/match
// lookups, 5 mixer set swap, 6 mixer predict, 7 coder, 8 learn scalars+indirect, 9 mixer weight update,
// 10 LSTM output-layer step, 11 BPTT epochs, 12 BPTT weight grads+Adam, 13 stream init, 14 bit bookkeeping,
// 16 gate selection, 17 layer-0 dot products, 18 layer-0 chain, 19 layer 1, 20 final neuron (6 = what is left
// of the predict phase: the barrier).
#if defined(__CUDA_ARCH__)
#define GMX_CLOCK() clock64()
#else
#define GMX_CLOCK() 0ll
#endif
#define GMX_PROF(slot)                                                        \
  do {                                                                        \
    if (PROF && tid == 0) {                                                   \
      const long long t_ = GMX_CLOCK();                                       \
      s.prof[slot] += (unsigned long long)(t_ - s.prof_t);                    \
      s.prof_t = t_;                                                          \
    }                                                                         \
  } while (0)  // smem stride of one staged mixer weight set (odd: conflict-free lanes)

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
  uint64_t total;             // arena bytes
};

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
  uint32_t* final_state;     // optional: the stream's StreamSmem is parked here at its end (n_streams x sizeof(StreamSmem)), or null
  int32_t analysis;          // -1: what the reference runner does for this mode; 0/1: forced (Predictor::EnableAnalysis)
  // generation (runner_utils::RunGeneration runner-utils.cpp:158-221): `in` holds the prompts, out[sid * gen_bytes ..] the samples
  uint32_t gen_bytes; float temperature;
  const float* rand_u; uint64_t rand_stride;   // rand()/RAND_MAX draws, one per generated bit; stream sid reads rand_u[sid * rand_stride + k]
};

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

// ---- per-stream state staged in shared memory ---------------------------------------------------
struct StreamSmem {
  StreamTables T;
  // blackboard (ShortTermMemory)
  alignas(16) float preds[NPRED + 2];
  alignas(4) uint8_t act[NPRED + NL0 + 2];   // prediction i is active; entries 90.. (layer-0 outputs) are always 1
  uint32_t ctx[C_COUNT + 2];
  alignas(16) float l0_out[NL0]; float l1_out[NL1], final_out, prob;   // l0_out | l1_out contiguous (final mixer input)
  alignas(16) float ppm[256], lprob[256];   // byte distributions of PPMd and LSTM
  uint8_t ring[32];            // last bytes (the reference keeps 1000, short-term-memory.h:23; only 10 are ever read)
  uint32_t ring_pos;
  int32_t new_bit, recent_bits, bb, first_prediction, analysis;
  uint32_t error;
  uint32_t steps;                      // Mixer::steps_ (identical for all 33 mixers)
  // mixers
  alignas(16) float w[WTOTAL];
  alignas(16) float xe[NPRED + NL0 + 2];    // layer-0 input vector: predictions (inactive ones zeroed) | layer-0 outputs
  uint32_t set_steps[NMIX], max_steps[NMIX], set_idx[NMIX], set_pool[NMIX];
  uint32_t swap_old[NMIX], swap_new[NMIX], nswap;   // queued set swaps of this bit
  uint8_t swap_m[NMIX + 3], shrink[NMIX + 3];
  uint8_t swap_staged[NMIX + 3];             // queued swap takes its new set from cand_w (1 + candidate) instead of the pool
  uint32_t cand_id[NBITMIX][2], cand_valid;  // pool ids of the staged candidates (0 = no set yet); valid for the next bit only
  alignas(16) float cand_w[2 * CAND_Q * 4];  // [candidate bit][record images {steps,0,0,0 | weights} of the four mixers]
  float upd[NMIX];
  uint32_t pool_next;
  // indirect
  uint32_t ind_base[NIND], ind_slot[NIND];   // ind_slot: dense slot, or position in the sparse map
  uint16_t ind_state[NIND + 1];
  uint8_t ind_found[NIND + 3];               // sparse tables: entry exists at ind_slot
  float ind_pa[NIND], ind_pb[NIND];          // the two logit-map entries Learn will update, as read by Predict
  uint32_t sparse_used;
  uint32_t tables_done;                      // LearnTables of this bit already ran during PredictBit (compress)
  uint32_t t_start_us;
  // match
  uint32_t m_cur[NMATCH]; uint8_t m_byte[NMATCH], m_bitpos[NMATCH], m_len[NMATCH];
  uint32_t hist_len;
  // indirect hash
  uint64_t ih_outer[NIH]; uint32_t ih_hash[NIH];
  // LSTM
  alignas(16) float l_hidden[L_HID + 1]; float l_state[L_CELLS], l_state_err[L_CELLS], l_stored_err[L_CELLS], l_hidden_err[L_CELLS];
  float l_gate[3][L_CELLS], l_gerr[3][L_CELLS];
  float l_red[16];
  union alignas(16) {            // never live at the same time:
    uint32_t sqp[256];           //   PPMd symbol pseudo-probabilities (byte boundary, before the LSTM forward pass)
    float l_err256[256];         //   LSTM scratch (forward pass, Perceive, BPTT)
  };
  uint32_t p_masked[8];          // PPMd: bit sym = CharMask[sym] == EscCount (ppmd.cuh)
  uint8_t l_hist[L_HORIZON], l_symin[L_HORIZON];
  uint32_t l_epoch, l_update_steps, l_old_input, l_fused;
  // coder
  uint32_t x1, x2, x;
  uint64_t out_pos, out_cap, in_pos, in_len;
  // phase profiler (lap timer driven by thread 0)
  unsigned long long prof[GMX_PROF_SLOTS];
  long long prof_t;
};

struct SparseMap { unsigned long long* tab; uint32_t mask; };

struct Arena {
  uint8_t* base;
  const ArenaLayout* L;
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
#if (defined(GMX_SPARSE_EVICT_LAST) || defined(GMX_SPARSE_EVICT_FIRST)) && defined(__CUDA_ARCH__)
    unsigned long long e;
#if defined(GMX_SPARSE_EVICT_LAST)
    { const uint64_t pol = PolicyEvictLast(); asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(e) : "l"(M.tab + pos), "l"(pol) : "memory"); }
#else
    { const uint64_t pol = PolicyEvictFirst(); asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(e) : "l"(M.tab + pos), "l"(pol) : "memory"); }
#endif
#else
    const unsigned long long e = M.tab[pos];
#endif
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
  if (atomicAdd(used, 1u) >= limit) { *error = GMX_ERR_SPARSE_FULL; return; }
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

// The four mixers whose gate context contains bit_context select another weight set every bit. Both sets the next
// bit can select are copied into shared memory one bit ahead (cand_w), so the swap on the per-bit critical path is a
// shared-memory copy instead of two dependent global loads (directory entry, then record).
GMX_DEV inline int BitMixer(int slot) { return slot == 0 ? 2 : slot == 1 ? 11 : slot == 2 ? NL0 + 2 : NL0 + 5; }   // gates: SLPR, LBPR, BIT_CONTEXT x2
GMX_DEV inline int BitSlot(int m) { return m == 2 ? 0 : m == 11 ? 1 : m == NL0 + 2 ? 2 : m == NL0 + 5 ? 3 : -1; }
GMX_DEV inline int CandOff(int slot) { return slot == 0 ? 0 : slot == 1 ? 24 : slot == 2 ? 51 : 59; }            // 24 + 27 + 8 + 9 = 68

GMX_DEV inline void BlockSync() { __syncthreads(); }
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

// Cache-policy helpers. The LSTM gate weights (184 KB per stream) are re-read every byte and are the
// only large per-stream data worth keeping in the 126 MB L2; the output-layer copies (52 KB read + 52 KB
// written per byte out of a 5.2 MB ring) are pure streaming traffic and are marked evict-first.
GMX_DEV inline void PrefetchL2Keep(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(p));
#else
  (void)p;
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
GMX_DEV inline void PrefetchRangeKeep(const void* p, uint32_t bytes, int t, int nthr) {
  for (uint32_t o = (uint32_t)t * 128u; o < bytes; o += (uint32_t)nthr * 128u) PrefetchL2Keep((const char*)p + o);
}

// L2 prefetch of `bytes` bytes at p, cooperatively by the `nthr` threads numbered t = 0..nthr-1.
GMX_DEV inline void PrefetchRange(const void* p, uint32_t bytes, int t, int nthr) {
  for (uint32_t o = (uint32_t)t * 128u; o < bytes; o += (uint32_t)nthr * 128u) PrefetchL2((const char*)p + o);
}

// ---- stream start ------------------------------------------------------------------------------
template <int NT>
GMX_DEV void FillWords(uint32_t* p, uint64_t nwords, uint32_t v, int tid) {
  for (uint64_t i = tid; i < nwords; i += NT) p[i] = v;
}

template <int NT, bool PROF>
GMX_DEV void InitStream(StreamSmem& s, const Arena& A, const StreamParams& P, int tid) {
  const ArenaLayout& L = *A.L;
  if (tid == 0) { s.prof_t = GMX_CLOCK(); s.t_start_us = (uint32_t)(GlobalTimerNs() / 1000ull); }
  if (P.tmpl_arena) {   // clone of a parked stream: arena image, then everything of StreamSmem behind the launch tables
    const uint4* src = (const uint4*)P.tmpl_arena;
    uint4* dst = (uint4*)A.base;
    const uint64_t n16 = L.total / 16;
    for (uint64_t i = tid; i < n16; i += NT) dst[i] = src[i];
    constexpr int kFirst = (int)(sizeof(StreamTables) / 4), kWords = (int)(sizeof(StreamSmem) / 4);
    uint32_t* sw = (uint32_t*)&s;
    for (int i = kFirst + tid; i < kWords; i += NT) sw[i] = P.tmpl_state[i];
    BlockSync();
    if (tid == 0) {
      s.error = 0; s.nswap = 0; s.cand_valid = 0; s.tables_done = 0; s.x1 = 0; s.x2 = 0xffffffffu; s.x = 0;
      for (int i = 0; i < GMX_PROF_SLOTS; ++i) s.prof[i] = 0;
      s.prof_t = GMX_CLOCK(); s.t_start_us = (uint32_t)(GlobalTimerNs() / 1000ull);
    }
    BlockSync();
    GMX_PROF(13);
    return;
  }
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
    s.first_prediction = 1; s.error = 0; s.steps = 0; s.pool_next = 1; s.hist_len = 0; s.sparse_used = 0; s.nswap = 0; s.cand_valid = 0; s.tables_done = 0;
    s.l_epoch = 0; s.l_update_steps = 0; s.l_old_input = 0; s.l_fused = 0;
    s.x1 = 0; s.x2 = 0xffffffffu; s.x = 0;
    for (int i = 0; i < GMX_PROF_SLOTS; ++i) s.prof[i] = 0;
  }
  BlockSync();
  if (tid == 0) {
    Ppmd pm{A.at<PpmdState>(L.p_state), A.at<uint8_t>(L.p_heap), L.p_mask, L.p_text_cap, L.p_units_cap, s.sqp, 0, s.p_masked};
    pm.Init();
  }
  BlockSync();
  GMX_PROF(13);
}

// Output layer of the next epoch slot: copy of the slot just used plus one SGD step (Lstm::Perceive
// lstm.cpp:81-88). `byte` is the symbol that followed the forward pass of slot `last`.
template <int NT>
GMX_DEV void LstmOutputStep(StreamSmem& s, const Arena& A, uint32_t last, uint32_t cur, uint32_t byte, int tid) {
  const ArenaLayout& L = *A.L;
  const float* wl = A.at<float>(L.l_wout) + (size_t)last * L_HID * L_NOUT;
  float* wc = A.at<float>(L.l_wout) + (size_t)cur * L_HID * L_NOUT;
  const float lr = (float)0.03;
  {
    const int q = tid & (L_NOUT / 4 - 1), half = tid / (L_NOUT / 4);   // outputs 4q..4q+3, rows of half `half`
    float le[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t i = 4 * q + k;
      const float err = i == byte ? f_sub(s.lprob[i], 1.0f) : s.lprob[i];
      le[k] = f_mul(lr, err);
    }
    const float4* wl4 = (const float4*)wl + q;
    float4* wc4 = (float4*)wc + q;
    const int j0 = half * ((L_HID + NT / (L_NOUT / 4) - 1) / (NT / (L_NOUT / 4)));
    const int j1 = j0 + (L_HID + NT / (L_NOUT / 4) - 1) / (NT / (L_NOUT / 4)) < L_HID ? j0 + (L_HID + NT / (L_NOUT / 4) - 1) / (NT / (L_NOUT / 4)) : L_HID;
#pragma unroll 13
    for (int j = j0; j < j1; ++j) {
      const float h = s.l_hidden[j];
#if defined(GMX_WOUT_PLAIN)
      float4 w = wl4[j * (L_NOUT / 4)];
#else
      float4 w = LoadStream4(wl4 + j * (L_NOUT / 4));
#endif
      w.x = f_sub(w.x, f_mul(le[0], h)); w.y = f_sub(w.y, f_mul(le[1], h));
      w.z = f_sub(w.z, f_mul(le[2], h)); w.w = f_sub(w.w, f_mul(le[3], h));
#if defined(GMX_WOUT_PLAIN)
      wc4[j * (L_NOUT / 4)] = w;
#else
      StoreStream4(wc4 + j * (L_NOUT / 4), w);
#endif
    }
  }
}

// ---- LSTM forward at a byte boundary (Lstm::Predict lstm.cpp:91-122, LstmLayer::ForwardPass
// lstm-layer.cpp:198-241). Precondition: s.ppm holds the normalised PPMd distribution. --------------
template <int NT>
GMX_DEV void LstmForward(StreamSmem& s, const Arena& A, const StreamParams& P, int known_byte, int tid) {
  const ArenaLayout& L = *A.L;
  const uint32_t e = s.l_epoch;
  const uint32_t sym = s.ctx[C_LAST_BYTE];
  float* lin_e = A.at<float>(L.l_lin) + e * (L_NIN + 1);
  // The pass streams 184 KB of gate weights and then the 52 KB output layer of this epoch slot from HBM.
  // Request them all now (a few thousand cycles ahead of use: long enough to cover the HBM latency,
  // short enough that the lines are still in L2 when the dot products reach them).
  {
    const float* W = A.at<float>(L.l_w);
#if defined(GMX_PF_GATES)   // off by default: with the asynchronous ring below the extra L2 traffic costs more than it hides
    for (int g = 0; g < 3; ++g) PrefetchRangeKeep(W + LstmW(g, L_NOUT, 0), (L_ROWQ - L_NOUT / 4) * L_CELLS * 16, tid, NT);
#endif
#if !defined(GMX_NO_PF_WOUT)
    PrefetchRange(A.at<float>(L.l_wout) + (size_t)e * L_HID * L_NOUT, L_HID * L_NOUT * 4, tid, NT);
#endif
    (void)W;
  }
  // layer_input[e] = [ppm 256 | hidden 50 | 1]  (SetInput lstm.cpp:45-50, copy :94-96)
  for (int i = tid; i < L_NIN; i += NT) {
    const float v = i < 256 ? s.ppm[i] : i < 306 ? s.l_hidden[i - 256] : 1.0f;
    lin_e[i] = v;
  }
  if (tid < L_CELLS) A.at<float>(L.l_last)[e * L_CELLS + tid] = s.l_state[tid];  // last_state_[epoch] = state_
  BlockSync();
  // gate pre-activations: f = w[sym]; f += in[j] * w[256 + j], j ascending (lstm-layer.cpp:227-232).
  // 150 rows on 75 threads, two independent sequential sums per thread (twice the loads in flight).
  if (tid < 3 * L_CELLS / 2) {
    const int t0 = tid, t1 = tid + 3 * L_CELLS / 2;
    const int g0 = t0 / L_CELLS, i0 = t0 - g0 * L_CELLS, g1 = t1 / L_CELLS, i1 = t1 - g1 * L_CELLS;
    const float* W = A.at<float>(L.l_w);
    float f0 = W[LstmW(g0, (int)sym, i0)], f1 = W[LstmW(g1, (int)sym, i1)];
    // quad q of this row = input columns 4q..4q+3 = one 16-byte load
    const float4* w0 = (const float4*)W + ((size_t)g0 * L_ROWQ + L_NOUT / 4) * L_CELLS + i0;
    const float4* w1 = (const float4*)W + ((size_t)g1 * L_ROWQ + L_NOUT / 4) * L_CELLS + i1;
#if !defined(GMX_NO_LSTM_STAGE)
    // The 77 quads of both rows stream through a private ring of LSTM_STAGES x 2 float4 slots in s.w (free at this
    // point, ByteBoundary (0)): cp.async keeps 2 x LSTM_STAGES 16-byte copies in flight per thread without holding
    // registers, 2.5x what register-staged loads reach under the 64-register cap. No barrier: a thread only ever
    // touches its own slots.
    constexpr int NQ = L_NOUT / 4 + L_CELLS / 4 + 1, HALF = 3 * L_CELLS / 2;
    float4* ring = (float4*)s.w;
#if !defined(GMX_NO_GATES_EVICT_FIRST)
    const uint64_t pol = PolicyEvictFirst();
#define GMX_RING_CP(dst, src) CpAsync16Hint(dst, src, pol)
#else
#define GMX_RING_CP(dst, src) CpAsync16(dst, src)
#endif
    static_assert(LSTM_STAGES * 2 * HALF * 16 <= WTOTAL * 4, "ring does not fit the weight-set staging area");
#pragma unroll
    for (int q = 0; q < LSTM_STAGES; ++q) {
      GMX_RING_CP(ring + (q * 2 + 0) * HALF + tid, w0 + q * L_CELLS);
      GMX_RING_CP(ring + (q * 2 + 1) * HALF + tid, w1 + q * L_CELLS);
      CpAsyncCommit();
    }
    int st = 0;
#pragma unroll 1
    for (int q = 0; q < NQ; ++q) {
      CpAsyncWaitGroup<LSTM_STAGES - 1>();
      const float4 a = ring[(st * 2 + 0) * HALF + tid], b = ring[(st * 2 + 1) * HALF + tid];
      if (q + LSTM_STAGES < NQ) {
        GMX_RING_CP(ring + (st * 2 + 0) * HALF + tid, w0 + (q + LSTM_STAGES) * L_CELLS);
        GMX_RING_CP(ring + (st * 2 + 1) * HALF + tid, w1 + (q + LSTM_STAGES) * L_CELLS);
      }
      CpAsyncCommit();   // one group per iteration, empty at the tail, keeps the wait distance constant
      st = st + 1 == LSTM_STAGES ? 0 : st + 1;
      if (q < NQ - 1) {   // layer input = [ppm 256 | hidden 50 | 1]
        const float4 x = q < L_NOUT / 4 ? ((const float4*)s.ppm)[q] : ((const float4*)s.l_hidden)[q - L_NOUT / 4];
        f0 = f_add(f0, f_mul(x.x, a.x)); f1 = f_add(f1, f_mul(x.x, b.x));
        f0 = f_add(f0, f_mul(x.y, a.y)); f1 = f_add(f1, f_mul(x.y, b.y));
        f0 = f_add(f0, f_mul(x.z, a.z)); f1 = f_add(f1, f_mul(x.z, b.z));
        f0 = f_add(f0, f_mul(x.w, a.w)); f1 = f_add(f1, f_mul(x.w, b.w));
      } else {            // hidden 48, 49 and the bias input (1.0); the fourth column is padding
        const float h48 = s.l_hidden[L_CELLS - 2], h49 = s.l_hidden[L_CELLS - 1];
        f0 = f_add(f0, f_mul(h48, a.x)); f1 = f_add(f1, f_mul(h48, b.x));
        f0 = f_add(f0, f_mul(h49, a.y)); f1 = f_add(f1, f_mul(h49, b.y));
        f0 = f_add(f0, f_mul(1.0f, a.z)); f1 = f_add(f1, f_mul(1.0f, b.z));
      }
    }
    CpAsyncWaitAll();
#else
    const float4* x4 = (const float4*)s.ppm;
GMX_UNROLL(GMX_LSTM_UNROLL)
    for (int q = 0; q < L_NOUT / 4; ++q) {           // layer input = [ppm 256 | hidden 50 | 1]
      const float4 x = x4[q], a = GMX_LSTM_LOAD(w0 + q * L_CELLS), b = GMX_LSTM_LOAD(w1 + q * L_CELLS);
      f0 = f_add(f0, f_mul(x.x, a.x)); f1 = f_add(f1, f_mul(x.x, b.x));
      f0 = f_add(f0, f_mul(x.y, a.y)); f1 = f_add(f1, f_mul(x.y, b.y));
      f0 = f_add(f0, f_mul(x.z, a.z)); f1 = f_add(f1, f_mul(x.z, b.z));
      f0 = f_add(f0, f_mul(x.w, a.w)); f1 = f_add(f1, f_mul(x.w, b.w));
    }
    w0 += (L_NOUT / 4) * L_CELLS; w1 += (L_NOUT / 4) * L_CELLS;
    x4 = (const float4*)s.l_hidden;
GMX_UNROLL(GMX_LSTM_UNROLL)
    for (int q = 0; q < L_CELLS / 4; ++q) {          // hidden 0..47
      const float4 x = x4[q], a = w0[q * L_CELLS], b = w1[q * L_CELLS];
      f0 = f_add(f0, f_mul(x.x, a.x)); f1 = f_add(f1, f_mul(x.x, b.x));
      f0 = f_add(f0, f_mul(x.y, a.y)); f1 = f_add(f1, f_mul(x.y, b.y));
      f0 = f_add(f0, f_mul(x.z, a.z)); f1 = f_add(f1, f_mul(x.z, b.z));
      f0 = f_add(f0, f_mul(x.w, a.w)); f1 = f_add(f1, f_mul(x.w, b.w));
    }
    {                                                // hidden 48, 49 and the bias input (1.0); the fourth column is padding
      const float4 a = w0[(L_CELLS / 4) * L_CELLS], b = w1[(L_CELLS / 4) * L_CELLS];
      const float h48 = s.l_hidden[L_CELLS - 2], h49 = s.l_hidden[L_CELLS - 1];
      f0 = f_add(f0, f_mul(h48, a.x)); f1 = f_add(f1, f_mul(h48, b.x));
      f0 = f_add(f0, f_mul(h49, a.y)); f1 = f_add(f1, f_mul(h49, b.y));
      f0 = f_add(f0, f_mul(1.0f, a.z)); f1 = f_add(f1, f_mul(1.0f, b.z));
    }
#endif
    s.l_gate[g0][i0] = f0;
    s.l_gate[g1][i1] = f1;
  }
  BlockSync();
  // ivar = 1 / sqrt(sum(norm^2)/cells + 1e-5): _Expr::sum() runs descending (lstm-layer.cpp:233-236)
  if (tid < 3) {
    const float* nrm = s.l_gate[tid];
    float acc = f_mul(nrm[L_CELLS - 1], nrm[L_CELLS - 1]);
    for (int i = L_CELLS - 2; i >= 0; --i) acc = f_add(acc, f_mul(nrm[i], nrm[i]));
    const float ivar = f_div(1.0f, f_sqrt(f_add(f_div(acc, (float)L_CELLS), 1e-5f)));
    s.l_red[tid] = ivar;
    A.at<float>(L.l_ivar)[tid * L_HORIZON + e] = ivar;
  }
  BlockSync();
  for (int t = tid; t < 3 * L_CELLS; t += NT) {
    const int g = t / L_CELLS, i = t - g * L_CELLS;
    const float* gb = A.at<float>(L.l_gb);
    const float nrm = f_mul(s.l_gate[g][i], s.l_red[g]);
    A.at<float>(L.l_norm)[((size_t)g * L_HORIZON + e) * L_CELLS + i] = nrm;
    float st = f_add(f_mul(nrm, gb[g * L_CELLS + i]), gb[(3 + g) * L_CELLS + i]);  // norm*gamma + beta
    st = g == 1 ? gm_tanhf(st) : Logistic(st);  // lstm-layer.cpp:205-211
    s.l_gate[g][i] = st;
    A.at<float>(L.l_gstate)[((size_t)g * L_HORIZON + e) * L_CELLS + i] = st;
  }
  BlockSync();
  if (tid < L_CELLS) {  // lstm-layer.cpp:212-217
    const int i = tid;
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
  BlockSync();
  // output layer: sum_j hidden[j] * Wout[e][i][j], j ascending, hidden[50] = 1 (lstm.cpp:105-113)
  const float* wo = A.at<float>(L.l_wout) + (size_t)e * L_HID * L_NOUT;
  float mx = 0.0f;
  if (tid < L_NOUT / 4) {  // 4 adjacent outputs per thread: one 16-byte load feeds 4 sequential sums
    const float4* wo4 = (const float4*)wo + tid;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
#pragma unroll 17
    for (int j = 0; j < L_HID; ++j) {
      const float h = s.l_hidden[j];
#if defined(GMX_WOUT_STREAM)
      const float4 w = LoadStream4(wo4 + j * (L_NOUT / 4));
#else
      const float4 w = wo4[j * (L_NOUT / 4)];
#endif
      a0 = f_add(a0, f_mul(h, w.x)); a1 = f_add(a1, f_mul(h, w.y));
      a2 = f_add(a2, f_mul(h, w.z)); a3 = f_add(a3, f_mul(h, w.w));
    }
    s.l_err256[4 * tid + 0] = a0; s.l_err256[4 * tid + 1] = a1; s.l_err256[4 * tid + 2] = a2; s.l_err256[4 * tid + 3] = a3;
    mx = a0 > mx ? a0 : mx; mx = a1 > mx ? a1 : mx; mx = a2 > mx ? a2 : mx; mx = a3 > mx ? a3 : mx;
  }
  // max over all outputs, seeded with 0 (max is order independent)
  for (int o = 16; o > 0; o >>= 1) { const float v = __shfl_xor_sync(0xffffffffu, mx, o); mx = v > mx ? v : mx; }
  if ((tid & 31) == 0) s.l_red[4 + (tid >> 5)] = mx;
  BlockSync();
  mx = 0.0f;
  for (int wi = 0; wi < NT / 32; ++wi) { const float v = s.l_red[4 + wi]; mx = v > mx ? v : mx; }
  for (int i = tid; i < L_NOUT; i += NT) s.lprob[i] = gm_expf(f_sub(s.l_err256[i], mx));
  BlockSync();
  if (tid == 0) {  // valarray::sum(): ascending, seeded with element 0 (lstm.cpp:118)
    float acc = s.lprob[0];
    for (int i = 1; i < L_NOUT; ++i) acc = f_add(acc, s.lprob[i]);
    s.l_red[3] = acc;
  }
  BlockSync();
  const float denom = s.l_red[3];
  for (int i = tid; i < L_NOUT; i += NT) {
    const float v = f_div(s.lprob[i], denom);
    s.lprob[i] = v;
    A.at<float>(L.l_out)[e * L_NOUT + i] = v;
  }
  if (tid == 0) { s.l_epoch = e + 1 == L_HORIZON ? 0 : e + 1; s.l_fused = known_byte >= 0 && e + 1 < L_HORIZON; }
  BlockSync();
  // Compress knows the byte this distribution is about to code, so the output-layer step that
  // Lstm::Perceive performs after the byte (same operands: these probabilities, this hidden state, the
  // layer of slot e) can run now, while slot e is still in L2 from the dot products above: one HBM read of
  // the 52 KB layer per byte instead of two. Not in the last slot: there Perceive runs BPTT over all 100
  // stored layers before it overwrites slot 0.
  if (known_byte >= 0 && e + 1 < L_HORIZON) {
    LstmOutputStep<NT>(s, A, e, e + 1, (uint32_t)known_byte, tid);
    BlockSync();
  }
}

// Truncated BPTT over the 100 stored steps + Adam (Lstm::Perceive lstm.cpp:57-79,
// LstmLayer::BackwardPass lstm-layer.cpp:252-354). Weight gradients are accumulated per weight in
// the reference's epoch order (99 -> 0) by the thread that owns the weight, then Adam is applied.
template <int NT, bool PROF>
GMX_DEV void LstmBptt(StreamSmem& s, const Arena& A, const StreamParams& P, int tid) {
  const ArenaLayout& L = *A.L;
  float* gb = A.at<float>(L.l_gb);
  const float* W = A.at<float>(L.l_w);
  float* errh = A.at<float>(L.l_errh);
  // recurrent weights W[cell j][512 + i], snapshot transposed so that lanes (= i) read them coalesced
  // (the reference snapshots the same block into transpose_ at the first epoch, lstm-layer.cpp:300-311)
  float* Wt = A.at<float>(L.l_wt);
  for (int q = tid; q < 3 * L_CELLS * L_CELLS; q += NT) {
    const int g = q / (L_CELLS * L_CELLS), rem = q - g * (L_CELLS * L_CELLS);
    const int j = rem / L_CELLS, i = rem - j * L_CELLS;
    Wt[q] = W[LstmW(g, 512 + i, j)];
  }
  BlockSync();
  for (int ep = L_HORIZON - 1; ep >= 0; --ep) {
#if defined(GMX_PF_BPTT)
    if (ep > 0) PrefetchRange(A.at<float>(L.l_wout) + (size_t)(ep - 1) * L_HID * L_NOUT, L_HID * L_NOUT * 4, tid, NT);
#endif
    const float* out_e = A.at<float>(L.l_out) + ep * L_NOUT;
    for (int i = tid; i < L_NOUT; i += NT)
      s.l_err256[i] = (uint32_t)i == s.l_hist[ep] ? f_sub(out_e[i], 1.0f) : out_e[i];
    BlockSync();
    if (tid < L_CELLS) {
      // hidden_error[j] += Wout[ep][i][j] * err_i, i ascending; hidden_error is 0 on entry (lstm.cpp:60-70)
      const float* wo = A.at<float>(L.l_wout) + ((size_t)ep * L_HID + tid) * L_NOUT;
      float he = s.l_hidden_err[tid];
      const float4* wo4 = (const float4*)wo;
      const float4* er4 = (const float4*)s.l_err256;
GMX_UNROLL(GMX_BPTT_UNROLL)
      for (int i = 0; i < L_NOUT / 4; ++i) {
        const float4 w = LoadStream4(wo4 + i), e = er4[i];
        he = f_add(he, f_mul(w.x, e.x)); he = f_add(he, f_mul(w.y, e.y));
        he = f_add(he, f_mul(w.z, e.z)); he = f_add(he, f_mul(w.w, e.w));
      }
      // LstmLayer::BackwardPass lstm-layer.cpp:256-281
      const int i = tid;
      float stored = ep == L_HORIZON - 1 ? he : f_add(s.l_stored_err[i], he);
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
    if (tid == 0 && ep == 0 && s.l_update_steps < (uint32_t)L_UPDATE_LIMIT) s.l_update_steps++;
    BlockSync();
    // per gate (lstm-layer.cpp:313-318): beta_u += err; gamma_u += err*norm; err *= gamma*ivar
    for (int t = tid; t < 3 * L_CELLS; t += NT) {
      const int g = t / L_CELLS, i = t - g * L_CELLS;
      const float err = s.l_gerr[g][i];
      const float nrm = A.at<float>(L.l_norm)[((size_t)g * L_HORIZON + ep) * L_CELLS + i];
      const float gu = ep == L_HORIZON - 1 ? 0.0f : gb[(6 * 3 + g) * L_CELLS + i];
      const float bu = ep == L_HORIZON - 1 ? 0.0f : gb[(7 * 3 + g) * L_CELLS + i];
      gb[(7 * 3 + g) * L_CELLS + i] = f_add(bu, err);
      gb[(6 * 3 + g) * L_CELLS + i] = f_add(gu, f_mul(err, nrm));
      s.l_gerr[g][i] = f_mul(err, f_mul(gb[g * L_CELLS + i], A.at<float>(L.l_ivar)[g * L_HORIZON + ep]));
    }
    BlockSync();
    // err -= (sum(err*norm)/cells) * norm; the sum is an _Expr::sum(): descending (lstm-layer.cpp:319-321).
    // One lane per gate forms the sum (it is the same for every cell of the gate).
    if (tid < 3) {
      const int g = tid;
      const float* nrm = A.at<float>(L.l_norm) + ((size_t)g * L_HORIZON + ep) * L_CELLS;
      float acc = f_mul(s.l_gerr[g][L_CELLS - 1], nrm[L_CELLS - 1]);
#pragma unroll 7
      for (int k = L_CELLS - 2; k >= 0; --k) acc = f_add(acc, f_mul(s.l_gerr[g][k], nrm[k]));
      s.l_red[8 + g] = f_div(acc, (float)L_CELLS);
    }
    BlockSync();
    for (int t = tid; t < 3 * L_CELLS; t += NT) {
      const int g = t / L_CELLS, i = t - g * L_CELLS;
      const float nrm = A.at<float>(L.l_norm)[((size_t)g * L_HORIZON + ep) * L_CELLS + i];
      const float ne = f_sub(s.l_gerr[g][i], f_mul(s.l_red[8 + g], nrm));
      s.l_gerr[g][i] = ne;
      errh[((size_t)g * L_HORIZON + ep) * L_CELLS + i] = ne;
    }
    BlockSync();
    if (tid < L_CELLS) {
      const int i = tid;
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
    BlockSync();
  }
  GMX_PROF(11);
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
  for (int i = tid; i < 256; i += NT) head[i] = 0xFF;
  BlockSync();
  if (tid == 0)
#pragma unroll 1
    for (int ep = 0; ep < L_HORIZON; ++ep) { const int sy = s.l_symin[ep]; nxt[ep] = head[sy]; head[sy] = (uint8_t)ep; }
  BlockSync();
#pragma unroll 1
  for (int q = tid; q < 3 * (L_NOUT / 4) * L_CELLS; q += NT) {
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
  // cells per thread (one 16-byte and one 8-byte load feed 8 multiply-adds), then two float4 Adam updates.
  constexpr int RG = (L_NIN + 3) / 4, IP = L_CELLS / 2;
#pragma unroll 1
  for (int id = tid; id < 3 * RG * IP; id += NT) {
    const int ip = id % IP, rg = (id / IP) % RG, g = id / (IP * RG);
    const float2* e2 = (const float2*)(errh + (size_t)g * L_HORIZON * L_CELLS + 2 * ip);
    const float4* x4 = (const float4*)(lin + 4 * rg);
    float a[4][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}, {0.0f, 0.0f}, {0.0f, 0.0f}};
#pragma unroll 4
    for (int ep = L_HORIZON - 1; ep >= 0; --ep) {
#if defined(GMX_BPTT_STREAM) && defined(__CUDA_ARCH__)
      const float2 e = __ldcs(e2 + ep * (L_CELLS / 2));
      const float4 x = __ldcs(x4 + ep * ((L_NIN + 1) / 4));
#else
      const float2 e = e2[ep * (L_CELLS / 2)];
      const float4 x = x4[ep * ((L_NIN + 1) / 4)];
#endif
      a[0][0] = f_add(a[0][0], f_mul(e.x, x.x)); a[0][1] = f_add(a[0][1], f_mul(e.y, x.x));
      a[1][0] = f_add(a[1][0], f_mul(e.x, x.y)); a[1][1] = f_add(a[1][1], f_mul(e.y, x.y));
      a[2][0] = f_add(a[2][0], f_mul(e.x, x.z)); a[2][1] = f_add(a[2][1], f_mul(e.y, x.z));
      a[3][0] = f_add(a[3][0], f_mul(e.x, x.w)); a[3][1] = f_add(a[3][1], f_mul(e.y, x.w));
    }
    const bool pad = 4 * rg + 3 >= L_NIN;   // the last quad's fourth column does not exist (stays 0)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const size_t f4 = ((size_t)g * L_ROWQ + L_NOUT / 4 + rg) * L_CELLS + 2 * ip + c;
#if defined(GMX_ADAM_STREAM)
      float4 w = LoadStream4(W4 + f4), m = LoadStream4(M4 + f4), v = LoadStream4(V4 + f4);
#else
      float4 w = W4[f4], m = M4[f4], v = V4[f4];
#endif
      adam1(w.x, m.x, v.x, a[0][c]); adam1(w.y, m.y, v.y, a[1][c]); adam1(w.z, m.z, v.z, a[2][c]);
      if (!pad) adam1(w.w, m.w, v.w, a[3][c]);
#if defined(GMX_ADAM_STREAM)
      StoreStream4(W4 + f4, w); StoreStream4(M4 + f4, m); StoreStream4(V4 + f4, v);
#else
      W4[f4] = w; M4[f4] = m; V4[f4] = v;
#endif
    }
  }
  for (int t = tid; t < 2 * 3 * L_CELLS; t += NT) {  // gamma then beta (lstm-layer.cpp:349-352)
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
  BlockSync();
  GMX_PROF(12);
}

// Lstm::Perceive (lstm.cpp:52-89) on the 8th bit of a byte.
template <int NT, bool PROF>
GMX_DEV void LstmPerceive(StreamSmem& s, const Arena& A, const StreamParams& P, uint32_t byte, int tid) {
  const ArenaLayout& L = *A.L;
  const uint32_t cur = s.l_epoch;
  const uint32_t last = cur == 0 ? L_HORIZON - 1 : cur - 1;
  if (tid == 0) { s.l_old_input = s.l_hist[last]; s.l_hist[last] = (uint8_t)byte; }
  BlockSync();
  if (cur == 0) {
    // input symbol of epoch ep = byte perceived before it (lstm.cpp:71-74)
    for (int ep = tid; ep < L_HORIZON; ep += NT) s.l_symin[ep] = ep == 0 ? (uint8_t)s.l_old_input : s.l_hist[ep - 1];
    BlockSync();
    LstmBptt<NT, PROF>(s, A, P, tid);
  }
  if (!s.l_fused) LstmOutputStep<NT>(s, A, last, cur, byte, tid);
  BlockSync();
  GMX_PROF(10);
}

// One node of the binary interval search of a byte model (ModPPMD::Predict mod_ppmd.cpp:1662-1681, LstmModel::Predict
// lstm-model.cpp:36-47): node = recent_bits (1..255) covers [bot, top]; num = sum(mid+1..top) from 0.0f ascending,
// denom continues from num over bot..mid. Evaluated per bit by one lane per model while the Indirect lanes wait for
// their table loads (256 sequential adds at the first bit of a byte, 2 at the last). Returns Logit(p); flag bit 0:
// denom != 0, bit 1: p != 0.5.
GMX_DEV inline float IntervalNode(const float* probs, int node, uint32_t* flag) {
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

// L2 prefetch of the pool record a mixer's gate will select. which = 0: with the contexts as they are now
// (byte-level gates, called at the byte boundary before the swap); which = 1/2: the bit-level gates'
// contexts after the next bit turns out 0/1 (called one bit ahead).
GMX_DEV inline void PrefetchMixerSet(const StreamSmem& s, const Arena& A, int m, int which) {
  const ArenaLayout& L = *A.L;
  const int cid = s.T.mixer[m].ctx;
  const bool bit_level = cid == C_SLPR || cid == C_LBPR || cid == C_BIT_CONTEXT;
  uint32_t c;
  if (which == 0) {
    if (bit_level || cid == C_LONGEST || cid == C_ZERO || cid == C_LSTM) return;
    c = s.ctx[cid];
  } else {
    const uint32_t bc = s.ctx[C_BIT_CONTEXT];
    if (!bit_level || bc >= 127) return;
    const uint32_t nbc = 2 * bc + which;  // bit_context after the next bit
    c = cid == C_BIT_CONTEXT ? nbc : cid == C_LBPR ? (s.ctx[C_LAST_BYTE] << 8) + nbc : (s.ctx[C_RB1] << 8) + nbc;
  }
  const uint32_t nid = A.at<uint32_t>(L.mix_dir[m])[c & ((1u << s.T.mixer[m].log2) - 1)];
  if (nid) {
    const float* rec = A.at<float>(L.mix_pool) + (size_t)nid * L.mix_set_stride;
    const int bytes = (MixerNW(m) + 4) * 4;
    for (int o = 0; o < bytes; o += 128) PrefetchL2((const char*)rec + o);
  }
}

// ---- everything that only happens when a new byte has been perceived (recent_bits == 1) ----------
template <int NT, bool PROF>
GMX_DEV void ByteBoundary(StreamSmem& s, const Arena& A, const StreamParams& P, int known_byte, int tid) {
  const ArenaLayout& L = *A.L;
  const uint32_t last_byte = s.ctx[C_LAST_BYTE];
#if !defined(GMX_NO_LSTM_STAGE)
  // (0) Nearly every gate context changes with the byte, so all staged weight sets go back to the pool now instead of
  // at the swap a moment later: until then s.w is free, and the LSTM forward pass uses it as the landing zone of its
  // asynchronous gate-weight copies (LstmForward). Warps 0..2; the last warp starts PPMd right away.
  if (tid < NT - 32) {
    float4* pool = A.at<float4>(L.mix_pool);
    const uint32_t stride4 = L.mix_set_stride / 4;
    const int lane = tid & 31;
#pragma unroll 1
    for (int m = tid >> 5; m < NMIX; m += NT / 32 - 1) {
      const uint32_t old = s.set_pool[m];
      if (old && lane <= (MixerNW(m) + 3) / 4) {
        float4* rec = pool + (size_t)old * stride4;
        if (lane == 0) rec[0] = make_float4(u2f(s.set_steps[m]), 0.0f, 0.0f, 0.0f);
        else rec[lane] = ((const float4*)(s.w + WOff(m)))[lane - 1];
      }
    }
  }
#endif
  // (1) contexts: intervals, hashed skip contexts, indirect-hash tables; PPMd on its own thread.
  if (tid < 9) {  // IntervalContext::Predict interval-context.cpp:17-23
    const IntervalSpec sp = s.T.interval[tid];
    s.ctx[C_IV0 + tid] = sp.mask & ((s.ctx[C_IV0 + tid] << sp.shift) + (last_byte >> sp.div_log2));
  } else if (tid >= 32 && tid < 52) {  // SkipContext::Predict skip-context.cpp:9-19
    const SkipSpec sp = s.T.skip[tid - 32];
    uint64_t c = 0;
    for (int k = 0; k < sp.n; ++k) c = (c << 8) + RecentByte(s, sp.b[k]);
    const int id = tid - 32;
    s.ctx[id < 5 ? C_H2 + id : C_SK0 + (id - 5)] = Murmur64(c);
  } else if (tid >= 64 && tid < 64 + NIH) {  // IndirectHash::Predict indirect-hash.cpp:16-31
    const int k = tid - 64;
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
      SparsePut(M, &s.sparse_used, L.sparse_limit, &s.error, key, pos, e != 0ull,
                (uint32_t)((((uint64_t)(uint32_t)e % inner_mod) << 8) + last_byte));
      s.ctx[C_IH0 + k] = Murmur32(SparseGet(M, SparseKey(sid, oh & mask)));
    } else {
      uint32_t* tab = A.at<uint32_t>(L.ih_tab[k]);
      uint32_t* slot = tab + (s.ih_hash[k] & mask);
      *slot = (uint32_t)((((uint64_t)*slot % inner_mod) << 8) + last_byte);
      s.ctx[C_IH0 + k] = Murmur32(tab[oh & mask]);
    }
    s.ih_hash[k] = oh;
  } else if (tid >= NT - 32) {  // ModPPMD::Predict byte part, mod_ppmd.cpp:1651-1654 (the last warp, collectively)
    Ppmd pm{A.at<PpmdState>(L.p_state), A.at<uint8_t>(L.p_heap), L.p_mask, L.p_text_cap, L.p_units_cap, s.sqp, tid - (NT - 32), s.p_masked};
    pm.UpdateByte(last_byte);
    if (!pm.S->error) pm.PrepareByte();
    if (pm.S->error) s.error = GMX_ERR_PPMD_ARENA;
  }
  BlockSync();
  GMX_PROF(0);
#if !defined(GMX_NO_LSTM_STAGE)
  if (tid < NMIX) { s.set_idx[tid] = 0xFFFFFFFFu; s.set_pool[tid] = 0; }   // nothing is staged any more: the swap of this bit fetches all 33
#endif
  // (2) ppm_predictions = max(sqp, 1) / sum, valarray::sum() ascending (mod_ppmd.cpp:1655-1661)
  for (int i = tid; i < 256; i += NT) { float v = (float)s.sqp[i]; if (v < 1.0f) v = 1.0f; s.ppm[i] = v; }
  // Indirect row bases ((ctx << 8) % M, so that slot = (base + bit_context) % M) and L2 prefetch of
  // the 255-slot row every table will touch during this byte.
  if (tid >= 80 && tid < 80 + NIND) {  // (the row keyed by lstm_prediction_context is redone below)
    const int k = tid - 80;
    const uint32_t M = L.ind_size[k];
    const uint32_t base = (s.ctx[s.T.ind[k].ctx] << 8) % M;
    s.ind_base[k] = base;
  }
  BlockSync();
  if (tid == 0) {
    float acc = s.ppm[0];
    for (int i = 1; i < 256; ++i) acc = f_add(acc, s.ppm[i]);
    s.l_red[3] = acc;
  } else {
    // prefetch: 41 tables x 5 lines of 128 B cover the 510-byte row
#if defined(GMX_NO_PF_IND_ROW)
    for (int t = NIND * 5; t < NIND * 5; ++t) {
#else
    for (int t = tid - 1; t < NIND * 5; t += NT - 1) {
#endif
      const int k = t / 5, ln = t - k * 5;
      if (L.ind_sid[k]) {  // sparse: the probe start of the row's first slot (bit_context 0)
        if (ln == 0) PrefetchL2(A.map().tab + (SparseHash(SparseKey(L.ind_sid[k], s.ind_base[k])) & L.sparse_mask));
        continue;
      }
      const uint32_t M = L.ind_size[k];
      uint32_t slot = s.ind_base[k] + ln * 64;
      if (slot >= M) slot -= M;
      PrefetchL2(A.at<uint16_t>(L.ind_tab[k]) + slot);
    }
    // weight sets the byte-level gates will select at this boundary: bring their pool records to L2
#if defined(GMX_PF_MIX_BYTE)
    if (tid >= 64 && tid < 64 + NMIX) PrefetchMixerSet(s, A, tid - 64, 0);
#endif
  }
  BlockSync();
  {
    const float sum = s.l_red[3];
    for (int i = tid; i < 256; i += NT) s.ppm[i] = f_div(s.ppm[i], sum);
  }
  BlockSync();
  // (3) LSTM forward (LstmModel::Predict byte part, lstm-model.cpp:19-34)
  GMX_PROF(1);
  LstmForward<NT>(s, A, P, known_byte, tid);
  GMX_PROF(2);
  // lstm_prediction_context = first index of the maximum, strict > from 0 (lstm-model.cpp:26-33)
  {
    float bv = 0.0f; int bi = 0;
    for (int i = tid; i < 256; i += NT) if (s.lprob[i] > bv) { bv = s.lprob[i]; bi = i; }
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((tid & 31) == 0) { s.l_red[4 + (tid >> 5)] = bv; s.l_err256[tid >> 5] = (float)bi; }
  }
  BlockSync();
  if (tid == 0) {
    float bv = 0.0f; int bi = 0;
    for (int wi = 0; wi < NT / 32; ++wi) {
      const float ov = s.l_red[4 + wi]; const int oi = (int)s.l_err256[wi];
      if (ov > bv || (ov == bv && ov > 0.0f && oi < bi)) { bv = ov; bi = oi; }
    }
    s.ctx[C_LSTM] = bv > 0.0f ? bi : 0;
    for (int k = 0; k < NIND; ++k)
      if (s.T.ind[k].ctx == C_LSTM) s.ind_base[k] = (s.ctx[C_LSTM] << 8) % L.ind_size[k];
  }
  BlockSync();
  GMX_PROF(3);
}

// Indirect::Learn (x41), Match::Learn (x6) and the history append of BasicContexts::Learn for one bit, one model per lane
// t = 0 .. NIND + NMATCH. They depend on the bit and on what the lookup phase left in shared memory, not on the mixer
// outputs, so compress (which knows the bit before it is coded) runs them on the upper warps WHILE warp 0 evaluates the
// mixer network (PredictBit); everything else runs them at the start of LearnBit.
enum : int { LEARN_TABLE_LANES = NIND + NMATCH + 1 };
#if defined(GMX_LT_NOINLINE)
static GMX_DEV GMX_NOINLINE void LearnTables(StreamSmem& s, const Arena& A, int bit, int t) {
#else
GMX_DEV inline void LearnTables(StreamSmem& s, const Arena& A, int bit, int t) {
#endif
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
#if !defined(GMX_NO_PRED_EVICT_LAST)
    const uint64_t keep = PolicyEvictLast();
    StoreHint(pr + ns, f_add(a, f_mul(f_sub(fbit, Logistic(a)), lr)), keep);
    const float b = s.ind_pb[k];
    StoreHint(pr + 256 + rm, f_add(b, f_mul(f_sub(fbit, Logistic(b)), lr)), keep);
#else
    pr[ns] = f_add(a, f_mul(f_sub(fbit, Logistic(a)), lr));
    const float b = s.ind_pb[k];
    pr[256 + rm] = f_add(b, f_mul(f_sub(fbit, Logistic(b)), lr));
#endif
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
    if (s.hist_len >= L.history_cap) s.error = GMX_ERR_HISTORY_CAP;
    else A.at<uint8_t>(L.history)[s.hist_len] = (uint8_t)cur;
  }
}

// ---- mixer front end on the last warp ------------------------------------------------------------------------
// Gate selection (mixer.cpp:29-37) and the weight-set swaps it triggers need nothing from this bit's table lookups,
// except for the two mixers gated by longest_match (m = 6 and m = NL0 + 6, which wait for Match::Predict). The last
// warp therefore runs them DURING the lookup phase: directory loads, write-backs and the asynchronous fetches of
// the new sets overlap the sparse-map probes of the other warps, and the copies are only awaited after the barrier.
GMX_DEV inline bool LongestGated(int m) { return m == 6 || m == NL0 + 6; }
GMX_DEV inline int FrontMixer(int lane) { return lane < 6 ? lane : lane < 29 ? lane + 1 : lane + 2; }   // the 31 others, lane 0..30

// One weight set, all 32 lanes with the same arguments: staged set -> its pool record `old` (0 = none), then record
// `nid` (0 = no set yet: zeros) -> staging area, asynchronously.
GMX_DEV inline void SwapSet(StreamSmem& s, const Arena& A, int m, uint32_t old, uint32_t nid, int lane) {
  const ArenaLayout& L = *A.L;
  float4* pool = A.at<float4>(L.mix_pool);
  const uint32_t stride4 = L.mix_set_stride / 4;
  float4* w4 = (float4*)(s.w + WOff(m));
  if (lane <= (MixerNW(m) + 3) / 4) {
    if (old) {
      float4* rec = pool + (size_t)old * stride4;
      if (lane == 0) rec[0] = make_float4(u2f(s.set_steps[m]), 0.0f, 0.0f, 0.0f);
      else rec[lane] = w4[lane - 1];
    }
    const float4* rec = pool + (size_t)nid * stride4;
    if (lane == 0) {
#if defined(GMX_NF_L2) && defined(__CUDA_ARCH__)
      if (nid) s.set_steps[m] = __ldcg((const uint32_t*)rec); else s.set_steps[m] = 0u;   // experiment: header through L2, not L1
#else
      if (nid) CpAsync4(&s.set_steps[m], rec); else s.set_steps[m] = 0u;
#endif
      s.set_pool[m] = nid;
    } else {
      if (nid) CpAsync16(w4 + (lane - 1), rec + lane); else w4[lane - 1] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
  }
}

// Lookup phase, last warp: the 31 mixers not gated by longest_match.
GMX_DEV inline void MixerFrontEarly(StreamSmem& s, const Arena& A, int lane) {
  const ArenaLayout& L = *A.L;
  int m = 0;
  uint32_t old = 0, nid = 0;
  bool changed = false;
  if (lane < NMIX - 2) {
    m = FrontMixer(lane);
    const uint32_t idx = s.ctx[s.T.mixer[m].ctx] & ((1u << s.T.mixer[m].log2) - 1);
    changed = idx != s.set_idx[m];
    if (changed) {
      old = s.set_pool[m];
#if defined(GMX_NF_L2) && defined(__CUDA_ARCH__)
      nid = __ldcg(A.at<uint32_t>(L.mix_dir[m]) + idx);
#else
      nid = A.at<uint32_t>(L.mix_dir[m])[idx];
#endif
      s.set_idx[m] = idx;
    }
  }
  unsigned todo = __ballot_sync(0xffffffffu, changed);
  while (todo) {
    const int src = __ffs((int)todo) - 1;
    todo &= todo - 1;
    SwapSet(s, A, __shfl_sync(0xffffffffu, m, src), __shfl_sync(0xffffffffu, old, src), __shfl_sync(0xffffffffu, nid, src), lane);
  }
}

// After the lookup barrier, last warp: longest_match context, its two mixers, then wait for every copy of this bit.
GMX_DEV inline void MixerFrontLate(StreamSmem& s, const Arena& A, int lane) {
  const ArenaLayout& L = *A.L;
  uint32_t c = 0;   // longest_match = max(match_length / 32) (match.cpp:71-73)
#pragma unroll 1
  for (int k = 0; k < NMATCH; ++k) { const uint32_t v = s.m_len[k] >> 5; c = v > c ? v : c; }
#pragma unroll 1
  for (int i = 0; i < 2; ++i) {
    const int m = i ? NL0 + 6 : 6;
    const uint32_t idx = c & ((1u << s.T.mixer[m].log2) - 1);
    const bool changed = idx != s.set_idx[m];            // same answer on every lane
    const uint32_t old = s.set_pool[m];
    __syncwarp();
    if (changed) {
      const uint32_t nid = A.at<uint32_t>(L.mix_dir[m])[idx];
      if (lane == 0) s.set_idx[m] = idx;
      SwapSet(s, A, m, old, nid, lane);
    }
    __syncwarp();
  }
  if (lane == 0) s.ctx[C_LONGEST] = c;
  CpAsyncWaitAll();
}

// ---- Predictor::Predict (predictor.cpp:360-376) --------------------------------------------------
template <int NT, bool PROF>
// known_byte: the byte whose bits are being predicted if the caller knows it (compress), else -1.
GMX_DEV void PredictBit(StreamSmem& s, const Arena& A, const StreamParams& P, int known_byte, int tid) {
  const ArenaLayout& L = *A.L;
  if (tid == 0) {  // BasicContexts::Predict basic-contexts.cpp:21-40
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
  BlockSync();
  GMX_PROF(14);
  if (s.bb) ByteBoundary<NT, PROF>(s, A, P, known_byte, tid);
  const uint32_t bitctx = s.ctx[C_BIT_CONTEXT];
  const bool zero_inactive = s.analysis != 0;  // predictor.cpp:362-365
  if (tid < NIND) {  // Indirect::Predict indirect.cpp:28-45
    const int k = tid;
    const uint32_t M = L.ind_size[k];
    uint32_t slot = s.ind_base[k] + bitctx;
    if (slot >= M) slot -= M;
    uint32_t e;
    const uint32_t sid = L.ind_sid[k];
    if (sid) {  // absent == never written == {ns 255, rm 0}
      unsigned long long ent;
      slot = SparseFind(A.map(), SparseKey(sid, slot), &ent);
      e = ent ? (uint32_t)ent & 0xffffu : 0x00ffu;
      s.ind_found[k] = ent != 0ull;
    } else {
      e = A.at<uint16_t>(L.ind_tab[k])[slot];
    }
    s.ind_slot[k] = slot;
    s.ind_state[k] = (uint16_t)e;
#if !defined(GMX_PF_IND_NEXT)
    if (false) {
#else
    if (sid && bitctx < 127) {  // both slots the next bit can select: start their probes' sectors towards L2
#endif
      uint32_t nslot = s.ind_base[k] + 2 * bitctx + 1;
      if (nslot >= M) nslot -= M;
      const SparseMap Mp = A.map();
      PrefetchL2(Mp.tab + (SparseHash(SparseKey(sid, nslot)) & Mp.mask));
      if (++nslot >= M) nslot -= M;
      PrefetchL2(Mp.tab + (SparseHash(SparseKey(sid, nslot)) & Mp.mask));
    }
    const uint32_t ns = e & 0xff, rm = e >> 8;
    const float* pr = A.at<float>(L.ind_pred) + k * 512;
    const int pi = s.T.ind[k].pred;
    // both entries are read unconditionally: Indirect::Learn updates exactly these two (a never-seen state learns as
    // state 0, indirect.cpp:52-54) and takes them from shared memory instead of two more dependent global loads
#if !defined(GMX_NO_PRED_EVICT_LAST)
    const uint64_t keep = PolicyEvictLast();
    const float pa = LoadHint(pr + (ns == 255 ? 0 : ns), keep), pb = LoadHint(pr + 256 + rm, keep);
#else
    const float pa = pr[ns == 255 ? 0 : ns], pb = pr[256 + rm];
#endif
    s.ind_pa[k] = pa; s.ind_pb[k] = pb;
    if (ns != 255) { s.preds[pi] = pa; s.act[pi] = pa != 0.0f; }
    else { s.act[pi] = 0; if (zero_inactive) s.preds[pi] = 0.0f; }
    if (rm != 0) { s.preds[pi + 1] = pb; s.act[pi + 1] = pb != 0.0f; }
    else { s.act[pi + 1] = 0; if (zero_inactive) s.preds[pi + 1] = 0.0f; }
  } else if (tid >= 64 && tid < 64 + NMATCH) {  // Match::Predict match.cpp:25-74
    const int k = tid - 64;
    uint32_t len = s.m_len[k];
    const uint32_t cb = s.m_byte[k];
    uint32_t bp = s.m_bitpos[k];
    const int hit = s.new_bit == ((cb & bp) != 0);
    if (hit) { if (len < 255) ++len; } else len = 0;
    bp >>= 1;
    uint32_t cbyte = cb;
    if (s.bb) {
      uint32_t cm = s.m_cur[k];
      if (s.hist_len != 0 && cm == s.hist_len - 1) len = 0;
      if (len < 8) {
        const uint32_t idx = s.ctx[s.T.match[k].ctx] & ((1u << s.T.match[k].log2) - 1);
        cm = L.match_sid[k] ? SparseGet(A.map(), SparseKey(L.match_sid[k], idx)) : A.at<uint32_t>(L.match_tab[k])[idx];
      } else ++cm;
      if (s.hist_len != 0) {
        if (cm >= s.hist_len) { s.error = GMX_ERR_MATCH_RANGE; cm = 0; }
        cbyte = A.at<uint8_t>(L.history)[cm];
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
      s.act[pi] = p != 0.5f;
    } else {
      s.act[pi] = 0;
      if (zero_inactive) s.preds[pi] = 0.0f;
    }
  }
#if !defined(GMX_OLD_FRONT)
  else if (tid >= NT - 32) {
    MixerFrontEarly(s, A, tid - (NT - 32));
  } else if (tid == 70 || tid == 71) {  // per-bit part of ModPPMD / LstmModel::Predict
    const int which = tid - 70;
#else
  else if (tid == 96 || tid == 97) {  // per-bit part of ModPPMD / LstmModel::Predict
    const int which = tid - 96;
#endif
    uint32_t fl;
    const float val = IntervalNode(which ? s.lprob : s.ppm, s.recent_bits, &fl);
    if (fl & 1) { s.preds[which] = val; s.act[which] = (fl >> 1) & 1; }
    else { s.act[which] = 0; if (zero_inactive) s.preds[which] = 0.0f; }
  }
  BlockSync();
  GMX_PROF(4);
#if !defined(GMX_OLD_FRONT)
  if (tid >= NT - 32) {
    MixerFrontLate(s, A, tid - (NT - 32));
  } else if (tid >= 64 && tid < 96) {
    // layer-0 input vector: the active predictions, inactive ones as +0 (Mixer::Predict sums the active ones
    // in index order; a +-0 product leaves the running sum unchanged, the sum itself is never -0)
    for (int i = tid - 64; i < NPRED; i += 32) s.xe[i] = s.act[i] ? s.preds[i] : 0.0f;
  }
  BlockSync();
  GMX_PROF(16);
#else
  // Mixer gate selection (mixer.cpp:29-37): which weight set does each mixer need for this bit?
  if (tid < NMIX) {
    const int m = tid;
    uint32_t c;
    if (s.T.mixer[m].ctx == C_LONGEST) {  // longest_match = max(match_length / 32) (match.cpp:71-73)
      c = 0;
#pragma unroll 1
      for (int k = 0; k < NMATCH; ++k) { const uint32_t v = s.m_len[k] >> 5; c = v > c ? v : c; }
    } else {
      c = s.ctx[s.T.mixer[m].ctx];
    }
    const uint32_t idx = c & ((1u << s.T.mixer[m].log2) - 1);
    if (idx != s.set_idx[m]) {  // queue the swap: old set goes back to the pool, new one (if any) is staged
      const uint32_t q = atomicAdd(&s.nswap, 1u);
      s.swap_m[q] = (uint8_t)m;
      s.swap_old[q] = s.set_pool[m];
      const int slot = BitSlot(m);
      if (slot >= 0 && s.cand_valid) {   // staged one bit ago for both values of the bit that has just been perceived
        s.swap_new[q] = s.cand_id[slot][s.new_bit];
        s.swap_staged[q] = (uint8_t)(1 + s.new_bit);
      } else {
        s.swap_new[q] = A.at<uint32_t>(L.mix_dir[m])[idx];
        s.swap_staged[q] = 0;
      }
      s.set_idx[m] = idx;
    }
  } else if (tid == 40) {
    uint32_t c = 0;
    for (int k = 0; k < NMATCH; ++k) { const uint32_t v = s.m_len[k] >> 5; c = v > c ? v : c; }
    s.ctx[C_LONGEST] = c;
  } else if (tid >= 96 && tid < 128) {
#if defined(GMX_NO_CAND_STAGE)
    // the weight sets the bit-level gates can select for the NEXT bit (both values of the bit): L2 prefetch only
#if defined(GMX_PF_MIX_BIT)
    for (int j = tid - 96; j < 2 * NMIX; j += 32) PrefetchMixerSet(s, A, j >> 1, 1 + (j & 1));
#endif
#else
    CpAsyncWaitAll();   // the candidate copies this warp issued during the previous bit's mixer phase have landed
#endif
  } else if (tid >= 64 && tid < 96) {
    // layer-0 input vector: the active predictions, inactive ones as +0 (Mixer::Predict sums the active ones
    // in index order; a +-0 product leaves the running sum unchanged, the sum itself is never -0)
    for (int i = tid - 64; i < NPRED; i += 32) s.xe[i] = s.act[i] ? s.preds[i] : 0.0f;
  }
  BlockSync();
  GMX_PROF(16);
  // Swap staged weight sets, one queued mixer per round, one weight per thread: first all write-backs,
  // then all fetches (zero weights == no set yet: the dot product of zeros is +0, exactly the
  // reference's "data == nullptr" output).
  {
    // Pool record = {steps, 0, 0, 0 | weights...}: header and weights are 16-byte aligned, a set moves as
    // at most 29 float4. One warp per queued set, one float4 per lane.
    const uint32_t nswap = s.nswap;
    float4* pool = A.at<float4>(L.mix_pool);
    const uint32_t stride4 = L.mix_set_stride / 4;
    const int lane = tid & 31;
#pragma unroll 1
    for (uint32_t r = tid >> 5; r < nswap; r += NT / 32) {
      const int m = s.swap_m[r];
      const uint32_t old = s.swap_old[r];
      if (old && lane <= (MixerNW(m) + 3) / 4) {
        float4* rec = pool + (size_t)old * stride4;
        if (lane == 0) rec[0] = make_float4(u2f(s.set_steps[m]), 0.0f, 0.0f, 0.0f);
        else rec[lane] = ((const float4*)(s.w + WOff(m)))[lane - 1];
      }
    }
    __syncwarp();   // the same warp re-reads swap_* and overwrites the staged sets below
#pragma unroll 1
    for (uint32_t r = tid >> 5; r < nswap; r += NT / 32) {   // asynchronous copies: all queued sets in flight together
      const int m = s.swap_m[r];
      const uint32_t nid = s.swap_new[r];
      const int staged = s.swap_staged[r];
      if (staged && nid) {   // record image already in shared memory
        if (lane <= (MixerNW(m) + 3) / 4) {
          const float4* img = (const float4*)s.cand_w + (staged - 1) * CAND_Q + CandOff(BitSlot(m));
          if (lane == 0) { s.set_steps[m] = f2u(img[0].x); s.set_pool[m] = nid; }
          else ((float4*)(s.w + WOff(m)))[lane - 1] = img[lane];
        }
      } else if (lane <= (MixerNW(m) + 3) / 4) {
        const float4* rec = pool + (size_t)nid * stride4;
        if (lane == 0) {
          if (nid) CpAsync4(&s.set_steps[m], rec); else s.set_steps[m] = 0u;
          s.set_pool[m] = nid;
        } else {
          float4* dst = (float4*)(s.w + WOff(m)) + (lane - 1);
#if defined(GMX_MIX_EVICT_LAST)
          if (nid) CpAsync16Hint(dst, rec + lane, PolicyEvictLast()); else *dst = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#else
          if (nid) CpAsync16(dst, rec + lane); else *dst = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#endif
        }
      }
    }
    CpAsyncWaitAll();
  }
  BlockSync();
  GMX_PROF(5);
#endif
  // Mixer::Predict (mixer.cpp:51-106): warp 0, one lane per neuron, sequential sums, the serial
  // same-layer chain is resolved by warp shuffles in neuron order.
  if (tid < 32) {
    const int lane = tid;
    if (lane == 0) s.nswap = 0;
    const float* w = s.w + (lane < NL0 ? lane : 0) * WSTRIDE0;
    float acc = 0.0f;
    {
      const float4* x4 = (const float4*)s.xe;
      const float4* w4 = (const float4*)w;
#pragma unroll 2
      for (int q = 0; q < NPRED / 4; ++q) {
        const float4 x = x4[q], v = w4[q];
        acc = f_add(acc, f_mul(x.x, v.x)); acc = f_add(acc, f_mul(x.y, v.y));
        acc = f_add(acc, f_mul(x.z, v.z)); acc = f_add(acc, f_mul(x.w, v.w));
      }
#pragma unroll
      for (int i = NPRED / 4 * 4; i < NPRED; ++i) acc = f_add(acc, f_mul(s.xe[i], w[i]));
    }
    GMX_PROF(17);
    __syncwarp();   // converged warp: the shuffles below take their fast path
    {
      // serial chain of the layer: neuron j's finished output feeds every later neuron (mixer.cpp:60-70).
#if defined(GMX_CHAIN_UNROLLED)
      // The 23 chain weights sit in registers so that one step is shuffle -> mul -> add.
      float cw[NL0 - 1];
#pragma unroll
      for (int j = 0; j < NL0 - 1; ++j) cw[j] = w[NPRED + j];
#pragma unroll
      for (int j = 0; j < NL0 - 1; ++j) {
        const float oj = __shfl_sync(0xffffffffu, acc, j);
        if (lane > j && lane < NL0) acc = f_add(acc, f_mul(oj, cw[j]));
      }
#else
      // Rolled (a handful of instructions that stay in the instruction cache); the chain weight of the next step is
      // loaded from shared memory while the current step's shuffle is in flight.
      float c = w[NPRED];
#pragma unroll 1
      for (int j = 0; j < NL0 - 1; ++j) {
        const float cn = w[NPRED + j + 1];   // j = 22 reads the pad word behind the set: never used
        const float oj = __shfl_sync(0xffffffffu, acc, j);
        if (lane > j && lane < NL0) acc = f_add(acc, f_mul(oj, c));
        c = cn;
      }
#endif
    }
    if (lane < NL0) { s.l0_out[lane] = acc; s.xe[NPRED + lane] = acc; }
    __syncwarp();
    GMX_PROF(18);
    const float skip = s.preds[P_LSTM];
    const float* w1 = s.w + NL0 * WSTRIDE0 + (lane < NL1 ? lane : 0) * WSTRIDE1;
    acc = 0.0f;
    {
      const float4* x4 = (const float4*)s.l0_out;
      const float4* w4 = (const float4*)w1;
#if defined(GMX_MIX_UNROLL)
#pragma unroll
#else
#pragma unroll 1
#endif
      for (int q = 0; q < NL0 / 4; ++q) {
        const float4 x = x4[q], v = w4[q];
        acc = f_add(acc, f_mul(x.x, v.x)); acc = f_add(acc, f_mul(x.y, v.y));
        acc = f_add(acc, f_mul(x.z, v.z)); acc = f_add(acc, f_mul(x.w, v.w));
      }
    }
    __syncwarp();
    // a layer-1 output is complete (skip connection added last, weights[num_layer0 + output_index],
    // mixer.cpp:77-83) before the next layer-1 neuron reads it
#pragma unroll 1
    for (int j = 0; j < NL1; ++j) {
      if (lane == j) acc = f_add(acc, f_mul(skip, w1[NL0 + j]));
      if (j < NL1 - 1) {
        const float oj = __shfl_sync(0xffffffffu, acc, j);
        if (lane > j && lane < NL1) acc = f_add(acc, f_mul(oj, w1[NL0 + j]));
      }
    }
    if (lane < NL1) s.l1_out[lane] = acc;
    __syncwarp();
    GMX_PROF(19);
    if (lane == 0) {
      const float* w2 = s.w + NL0 * WSTRIDE0 + NL1 * WSTRIDE1;
      float p = 0.0f;
      const float4* x4 = (const float4*)s.l0_out;   // l0_out[24] then l1_out[8]
      const float4* w4 = (const float4*)w2;
#if defined(GMX_MIX_UNROLL)
#pragma unroll
#else
#pragma unroll 1
#endif
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
    GMX_PROF(20);
  }
#if defined(GMX_NO_CAND_STAGE) && defined(GMX_LEARN_OVERLAP)   // off: measured 2 % slower at 8 CTAs/SM (profiles/r01_s3_ab.md)
  else if (known_byte >= 0 && tid >= 64 && tid < 64 + LEARN_TABLE_LANES) {
    // compress: the bit about to be coded is known, so the table models learn it now, under the mixer network
    const int pos = 31 - __clz(s.recent_bits);   // bits of this byte already perceived
    LearnTables(s, A, (known_byte >> (7 - pos)) & 1, tid - 64);
    if (tid == 64) s.tables_done = 1;
  }
#endif
#if !defined(GMX_NO_CAND_STAGE)
  else if (tid >= NT - 32) {
    // Meanwhile the last warp stages, for the four bit-gated mixers, the sets both values of this bit lead to: lanes
    // 0..7 read the eight directory entries, then all lanes start the asynchronous copies of the record images. The
    // copies are awaited by this warp in the next bit's gate-selection phase, a whole Learn step away.
    const int lane = tid - (NT - 32);
    const uint32_t bc = s.ctx[C_BIT_CONTEXT];
    const bool ok = bc < 127;   // the next bit belongs to the same byte
    if (lane == 0) s.cand_valid = ok;
    if (ok) {
      uint32_t nid = 0;
      if (lane < 2 * NBITMIX) {
        const int slot = lane >> 1, which = lane & 1, m = BitMixer(slot);
        const uint32_t nbc = 2 * bc + 1 + which;   // bit_context after the next bit (basic-contexts.cpp:32-36)
        const int cid = s.T.mixer[m].ctx;
        const uint32_t c = cid == C_BIT_CONTEXT ? nbc : cid == C_LBPR ? (s.ctx[C_LAST_BYTE] << 8) + nbc : (s.ctx[C_RB1] << 8) + nbc;
        nid = A.at<uint32_t>(L.mix_dir[m])[c & ((1u << s.T.mixer[m].log2) - 1)];
        s.cand_id[slot][which] = nid;
      }
      const float4* pool = A.at<float4>(L.mix_pool);
      const uint32_t stride4 = L.mix_set_stride / 4;
#pragma unroll 1
      for (int i0 = 0; i0 < 2 * CAND_Q; i0 += 32) {
        const int i = i0 + lane;
        const int which = i >= CAND_Q, o = i - which * CAND_Q;
        const int slot = o < 24 ? 0 : o < 51 ? 1 : o < 59 ? 2 : 3;
        const uint32_t id = __shfl_sync(0xffffffffu, nid, (2 * slot + which) & 31);
        if (i < 2 * CAND_Q && id) CpAsync16((float4*)s.cand_w + i, pool + (size_t)id * stride4 + (o - CandOff(slot)));
      }
    }
  }
#endif
  BlockSync();
  GMX_PROF(6);
}

// ---- coder (encoder.cpp:8-34, decoder.cpp:3-39) ---------------------------------------------------
GMX_DEV inline uint32_t Discretize(float p) { return (uint32_t)f_add(1.0f, f_mul(65534.0f, p)); }

GMX_DEV inline void PutByte(StreamSmem& s, uint8_t* out, uint32_t b) {
  if (s.out_pos < s.out_cap) out[s.out_pos] = (uint8_t)b; else s.error = GMX_ERR_OUTPUT_CAP;
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

// ---- Predictor::Learn (predictor.cpp:383-387) ----------------------------------------------------
template <int NT, bool PROF>
// known_bit >= 0 (compress): the caller has not stored s.new_bit and not run the coder yet; both happen here, on lanes
// that are otherwise idle in the first phase, which saves the barrier between coding and learning. code_out = the
// stream's output slice.
GMX_DEV void LearnBit(StreamSmem& s, const Arena& A, const StreamParams& P, int tid, int known_bit = -1, uint8_t* code_out = nullptr) {
  const ArenaLayout& L = *A.L;
  const int bit = known_bit >= 0 ? known_bit : s.new_bit;
  const float fbit = (float)bit;
  const int cur = s.recent_bits * 2 + bit;
  const bool byte_done = cur >= 256;
  const uint32_t longest = s.ctx[C_LONGEST];
  // history length after BasicContexts::Learn (basic-contexts.cpp:42-54)
  const uint32_t hist_after = s.hist_len + ((byte_done && longest < 2) ? 1u : 0u);
  if (tid < NMIX) {  // Mixer::Learn scalars (mixer.cpp:108-127)
    const int m = tid;
    if (s.set_pool[m] == 0) {  // FindOrCreateMixerData mixer.cpp:39-49
      const uint32_t id = atomicAdd(&s.pool_next, 1u);
      if (id >= L.mix_pool_sets) { s.error = GMX_ERR_MIXER_POOL; }
      else { s.set_pool[m] = id; A.at<uint32_t>(L.mix_dir[m])[s.set_idx[m]] = id; }
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
  } else if (tid >= 64 && tid < 64 + LEARN_TABLE_LANES) {
    if (!s.tables_done) LearnTables(s, A, bit, tid - 64);
  } else if (tid == NT - 1 && known_bit >= 0) {   // Encoder::Encode encoder.cpp:8-34 (+ Perceive: predictor.cpp:378-381)
    EncodeBit(s, code_out, bit);
    s.new_bit = bit;
  }
  BlockSync();
  GMX_PROF(8);
  // Mixer weight updates (mixer.cpp:128-175): w -= update * x over exactly the inputs used by
  // Predict, then the (1 - 3e-6) shrink every 1024 steps of the set.
#pragma unroll 1
  for (int m = tid >> 5; m < NMIX; m += NT / 32) {
    const int lane = tid & 31;
    const int nw = MixerNW(m);
    const float upd = s.upd[m];
    const bool shrink = s.shrink[m] != 0;
    const float keep = f_sub(1.0f, 3.0e-6f);
    float* w = s.w + WOff(m);
    if (m < NL0) {
      // Layer 0: inputs = [90 predictions | outputs of the earlier layer-0 neurons] = s.xe[0 .. nw) (inactive
      // predictions are +0 there and are skipped through the `use` bytes); one float4 per lane covers them all.
      const int left = nw - 4 * lane;   // inputs of this lane's quad that exist
      if (left > 0) {
        float4 v = ((float4*)w)[lane];
        const float4 x = ((const float4*)s.xe)[lane];
        uint32_t on = ((const uint32_t*)s.act)[lane];
        if (left < 4) on &= 0x00ffffffu >> (8 * (3 - left));
        if (on & 0x000000ffu) v.x = f_sub(v.x, f_mul(upd, x.x));
        if (on & 0x0000ff00u) v.y = f_sub(v.y, f_mul(upd, x.y));
        if (on & 0x00ff0000u) v.z = f_sub(v.z, f_mul(upd, x.z));
        if (on & 0xff000000u) v.w = f_sub(v.w, f_mul(upd, x.w));
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
  if (tid == 0) { s.steps++; s.hist_len = hist_after; s.tables_done = 0; }
  BlockSync();
  GMX_PROF(9);
  if (byte_done) LstmPerceive<NT, PROF>(s, A, P, (uint32_t)(cur - 256), tid);  // LstmModel::Learn lstm-model.cpp:50-59
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

// runner_utils::Compress (runner-utils.cpp:43-67) incl. the 5-byte header of RunCompression (:109).
template <int NT, bool PROF>
GMX_DEV void CompressStream(StreamSmem& s, const Arena& A, const StreamParams& P, uint32_t sid, int tid) {
  const uint8_t* in = P.in + P.in_off[sid];
  const uint64_t n = P.in_off[sid + 1] - P.in_off[sid];
  uint8_t* out = P.out + P.out_off[sid];
  InitStream<NT, PROF>(s, A, P, tid);
  if (tid == 0) {
    s.out_pos = 0; s.out_cap = P.out_off[sid + 1] - P.out_off[sid];
    s.analysis = P.analysis >= 0 ? P.analysis : (8 * n / 1000) > 0;  // EnableAnalysis(8*n/1000) -> predictions zeroed every bit
    for (int i = 4; i >= 0; --i) PutByte(s, out, (uint32_t)(n >> (8 * i)) & 0xff);
  }
  BlockSync();
  const bool tracing = sid == 0;
#pragma unroll 1
  for (uint64_t pos = 0; pos < n; ++pos) {
    const uint32_t c = in[pos];
#pragma unroll 1
    for (int j = 7; j >= 0; --j) {
      const int bit = (c >> j) & 1;
      PredictBit<NT, PROF>(s, A, P, (int)c, tid);
      if (tracing && (P.bit_trace || P.pred_trace)) {   // debug/parity traces of stream 0 read the blackboard before Learn touches it
        if (tid == 0) Trace(s, P, pos * 8 + (7 - j));
        BlockSync();
      }
      GMX_PROF(7);
      LearnBit<NT, PROF>(s, A, P, tid, bit, out);   // codes the bit (lane NT-1) and learns it
      if (s.error) break;
    }
    if (s.error) break;
  }
  BlockSync();
  if (tid == 0) {
    while (((s.x1 ^ s.x2) & 0xff000000u) == 0) { PutByte(s, out, s.x2 >> 24); s.x1 <<= 8; s.x2 = (s.x2 << 8) + 255; }
    PutByte(s, out, s.x2 >> 24);
    P.out_len[sid] = s.out_pos;
    P.status[sid] = s.error;
    WriteUsage(s, A, P, sid);
    if (PROF && P.prof) for (int i = 0; i < GMX_PROF_SLOTS; ++i) P.prof[(size_t)sid * GMX_PROF_SLOTS + i] = s.prof[i];
  }
  BlockSync();
  ParkState<NT>(s, P, sid, tid);
  BlockSync();
}

// runner_utils::Decompress (runner-utils.cpp:69-86) + ReadHeader (:29-36). Analysis is never on.
template <int NT, bool PROF>
GMX_DEV void DecompressStream(StreamSmem& s, const Arena& A, const StreamParams& P, uint32_t sid, int tid) {
  const uint8_t* in = P.in + P.in_off[sid];
  uint8_t* out = P.out + P.out_off[sid];
  InitStream<NT, PROF>(s, A, P, tid);
  if (tid == 0) {
    s.in_pos = 0; s.in_len = P.in_off[sid + 1] - P.in_off[sid];
    s.out_pos = 0; s.out_cap = P.out_off[sid + 1] - P.out_off[sid];
    s.analysis = P.analysis >= 0 ? P.analysis : 0;
    uint64_t len = 0;
    for (int i = 0; i <= 4; ++i) len = (len << 8) + GetByte(s, in);
    if (s.in_len < 5 || len > s.out_cap) { s.error = s.in_len < 5 ? GMX_ERR_BAD_HEADER : GMX_ERR_OUTPUT_CAP; len = 0; }
    s.out_cap = len;  // number of bytes to produce
    for (int i = 0; i < 4; ++i) s.x = (s.x << 8) + (GetByte(s, in) & 0xff);
  }
  BlockSync();
  const uint64_t n = s.out_cap;
#pragma unroll 1
  for (uint64_t pos = 0; pos < n; ++pos) {
#pragma unroll 1
    for (int j = 7; j >= 0; --j) {
      PredictBit<NT, PROF>(s, A, P, -1, tid);
      if (tid == 0) {  // Decoder::Decode decoder.cpp:19-39
        const uint32_t p16 = Discretize(s.prob);
        const uint32_t r = s.x2 - s.x1;
        const uint32_t xmid = s.x1 + (r >> 16) * p16 + (((r & 0xffff) * p16) >> 16);
        int bit = 0;
        if (s.x <= xmid) { bit = 1; s.x2 = xmid; } else s.x1 = xmid + 1;
        s.new_bit = bit;
        while (((s.x1 ^ s.x2) & 0xff000000u) == 0) { s.x1 <<= 8; s.x2 = (s.x2 << 8) + 255; s.x = (s.x << 8) + GetByte(s, in); }
        if (j == 0) out[pos] = (uint8_t)((s.recent_bits * 2 + bit) & 0xff);
      }
      BlockSync();
      GMX_PROF(7);
      LearnBit<NT, PROF>(s, A, P, tid);
      if (s.error) break;
    }
    if (s.error) break;
  }
  BlockSync();
  if (tid == 0) {
    P.out_len[sid] = s.error ? 0 : n; P.status[sid] = s.error;
    WriteUsage(s, A, P, sid);
    if (PROF && P.prof) for (int i = 0; i < GMX_PROF_SLOTS; ++i) P.prof[(size_t)sid * GMX_PROF_SLOTS + i] = s.prof[i];
  }
  BlockSync();
  ParkState<NT>(s, P, sid, tid);
  BlockSync();
}

// runner_utils::RunGeneration (runner-utils.cpp:158-221): the prompt (all but its last byte) is consumed WITH
// learning, then gen_bytes bytes are sampled bit by bit without Learn: prob = Logistic(Logit(prob) / temperature),
// bit = r < prob with r the next rand()/RAND_MAX draw, Perceive(bit), Predict().
template <int NT, bool PROF>
GMX_DEV void GenerateStream(StreamSmem& s, const Arena& A, const StreamParams& P, uint32_t sid, int tid) {
  const uint8_t* in = P.in + P.in_off[sid];
  const uint64_t n = P.in_off[sid + 1] - P.in_off[sid];
  uint8_t* out = P.out + (size_t)sid * P.gen_bytes;
  const float* ru = P.rand_u + (size_t)sid * P.rand_stride;
  InitStream<NT, PROF>(s, A, P, tid);
  if (tid == 0) s.analysis = P.analysis >= 0 ? P.analysis : (8 * (uint64_t)P.gen_bytes / 1000) > 0;   // :177
  BlockSync();
#pragma unroll 1
  for (uint64_t pos = 0; pos + 1 < n; ++pos) {   // :187-194
    const uint32_t c = in[pos];
#pragma unroll 1
    for (int j = 7; j >= 0; --j) {
      PredictBit<NT, PROF>(s, A, P, -1, tid);
      if (tid == 0) s.new_bit = (c >> j) & 1;
      BlockSync();
      LearnBit<NT, PROF>(s, A, P, tid);
      if (s.error) break;
    }
    if (s.error) break;
  }
  if (!s.error) {
    PredictBit<NT, PROF>(s, A, P, -1, tid);   // :198
#pragma unroll 1
    for (uint32_t i = 0; i < P.gen_bytes && !s.error; ++i) {
#pragma unroll 1
      for (int j = 0; j < 8; ++j) {
        if (tid == 0) {
          const float r = ru[(size_t)i * 8 + j];
          const float prob = Logistic(f_div(Logit(s.prob), P.temperature));
          s.new_bit = r < prob ? 1 : 0;
          if (j == 7) out[i] = (uint8_t)((s.recent_bits * 2 + s.new_bit) & 0xff);
        }
        BlockSync();
        PredictBit<NT, PROF>(s, A, P, -1, tid);
      }
    }
  }
  BlockSync();
  if (tid == 0) {
    P.out_len[sid] = s.error ? 0 : P.gen_bytes; P.status[sid] = s.error;
    WriteUsage(s, A, P, sid);
  }
  BlockSync();
  ParkState<NT>(s, P, sid, tid);
  BlockSync();
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
  BlockSync();
}

// ---- kernel entry: persistent CTAs, one stream at a time, ids from an atomic queue ---------------
enum : int { MODE_COMPRESS = 0, MODE_DECOMPRESS = 1, MODE_GENERATE = 2 };

template <int NT, int MODE, int MINB, bool PROF>
__global__ void __launch_bounds__(NT, MINB) StreamKernel(StreamParams P) {
  __shared__ StreamSmem s;
  __shared__ uint32_t next_stream;
  const int tid = (int)threadIdx.x;
  StageTables<NT>(s, P, tid);
  Arena A{P.arenas + (uint64_t)blockIdx.x * P.arena_stride, &s.T.L};
  for (;;) {
    if (tid == 0) next_stream = atomicAdd(P.queue, 1u);
    BlockSync();
    const uint32_t q = next_stream;
    BlockSync();
    if (q >= P.n_streams) break;
    const uint32_t sid = P.ids ? P.ids[q] : q;
    if (MODE == MODE_COMPRESS) CompressStream<NT, PROF>(s, A, P, sid, tid);
    else if (MODE == MODE_DECOMPRESS) DecompressStream<NT, PROF>(s, A, P, sid, tid);
    else GenerateStream<NT, PROF>(s, A, P, sid, tid);
  }
}

// ---- single-stream stepping (the Predictor facade: reference src/predictor.h:20-38) ---------------
// One launch per Predict() / Learn() call of ONE stream; the shared-memory state of the stream is
// parked in global memory between launches. Perceive(bit) travels with the next launch.
enum : int { STEP_INIT = 0, STEP_PREDICT = 1, STEP_LEARN = 2 };
struct StepParams {
  StreamParams P;          // arenas/layout/tables of the one stream (n_streams, in/out unused)
  uint32_t* state;         // sizeof(StreamSmem) bytes
  int op, has_bit, bit, analysis;
  float* prob_out;         // STEP_PREDICT: what Predictor::Predict() returns
  uint32_t* status_out;    // stream error code after the step
};

template <int NT>
__global__ void __launch_bounds__(NT) StepKernel(StepParams Q) {
  __shared__ StreamSmem s;
  const int tid = (int)threadIdx.x;
  uint32_t* sw = (uint32_t*)&s;
  constexpr int kWords = (int)(sizeof(StreamSmem) / 4);
  if (Q.op == STEP_INIT) {
    StageTables<NT>(s, Q.P, tid);
    Arena A{Q.P.arenas, &s.T.L};
    InitStream<NT, false>(s, A, Q.P, tid);
    if (tid == 0) s.analysis = Q.analysis;
  } else {
    for (int i = tid; i < kWords; i += NT) sw[i] = Q.state[i];
    BlockSync();
    Arena A{Q.P.arenas, &s.T.L};
    if (tid == 0 && Q.has_bit) s.new_bit = Q.bit;   // Predictor::Perceive predictor.cpp:378-381
    if (tid == 0 && Q.analysis >= 0) s.analysis = Q.analysis;
    BlockSync();
    if (Q.op == STEP_PREDICT) {
      PredictBit<NT, false>(s, A, Q.P, -1, tid);
      if (tid == 0) *Q.prob_out = s.prob;
    } else {
      LearnBit<NT, false>(s, A, Q.P, tid);
    }
  }
  CpAsyncWaitAll();   // candidate weight sets staged by PredictBit must have landed before shared memory is parked
  BlockSync();
  for (int i = tid; i < kWords; i += NT) Q.state[i] = sw[i];
  if (tid == 0) *Q.status_out = s.error;
}

}  // namespace gmx
#endif  // GMIX_B200_STREAM_KERNEL_CUH_
