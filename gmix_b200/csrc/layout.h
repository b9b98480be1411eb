// Host-side sizing of one stream arena (the global-memory workspace a CTA reuses across the
// streams it processes) and of the host-precomputed libm tables. Pure C++ (no CUDA): used by the
// C-ABI library and by the CPU emulation tests.
#ifndef GMIX_B200_LAYOUT_H_
#define GMIX_B200_LAYOUT_H_
#include <math.h>
#include <stdlib.h>

#include <vector>

#include "stream_kernel.cuh"

namespace gmx {

inline uint64_t AlignUp(uint64_t x, uint64_t a) { return (x + a - 1) / a * a; }

inline uint64_t Pow2Ceil(uint64_t x) { uint64_t p = 1; while (p < x) p <<= 1; return p; }

// What a stream that starts from a checkpoint already occupies (checkpoint.h: Count): added on top of what
// max_len new bytes can need.
struct Preload {
  uint64_t sparse_entries = 0, mixer_sets = 0, ppmd_unit_bytes = 0, ppmd_text_bytes = 0, history_bytes = 0, steps = 0;
  uint64_t ppmd_lo_bytes = 0, ppmd_hi_bytes = 0;   // the two unit areas separately (their sum is ppmd_unit_bytes)
};

// Geometry of a PPMd heap window of P = 2^k bytes (ppmd.cuh): virtual offset v lives at heap[v & (P - 1)].
// text_room = bytes below the low units; units_room = the gap the low (upward) and high (downward) unit areas share.
// valid = that gap is one piece above the text area.
struct PpmdWindow { bool valid; uint64_t text_room, units_room; };
inline PpmdWindow PpmdWindowOf(uint64_t P) {
  const uint64_t us = PPMD_UNITS_START % P;
  const uint64_t he = PPMD_HEAP_END % P == 0 ? P : PPMD_HEAP_END % P;
  PpmdWindow w;
  w.valid = us < he;
  w.text_room = us;
  w.units_room = w.valid ? he - us : 0;
  return w;
}

// max_len = longest stream (in uncompressed bytes) the arena must hold.
// roomy = false: the shared sparse map and the mixer weight-set pool are sized for what text-like data
// touches (a stream that needs more ends with GMX_ERR_SPARSE_FULL / GMX_ERR_MIXER_POOL and the host
// re-runs it in a roomy arena); roomy = true: both are sized for the worst case of max_len bytes.
// force_dense: every table as a dense array whatever the stream length (what long streams get automatically); used by
// tests to exercise that path on short inputs.
inline ArenaLayout MakeLayout(uint64_t max_len, bool roomy = false, const Preload* pre = nullptr, bool force_dense = false) {
  const Preload none;
  if (!pre) pre = &none;
  static const IndirectSpec ind[NIND] = {GMX_INDIRECT_SPECS};
  static const IHSpec ih[NIH] = {GMX_IH_SPECS};
  static const MatchSpec mt[NMATCH] = {GMX_MATCH_SPECS};
  static const MixerSpec mx[NMIX] = {GMX_MIXER_SPECS};
  ArenaLayout L;
  memset(&L, 0, sizeof(L));
  uint64_t off = 0;
  auto take = [&](uint64_t bytes) { uint64_t o = off; off = AlignUp(off + bytes, 256); return o; };
  // Which big tables go into the shared sparse map? Worst-case entry count vs their dense size.
  uint64_t worst = 64 + pre->sparse_entries, dense_bytes = 0;
  uint32_t next_sid = 1;
  for (int k = 0; k < NIND; ++k) {
    L.ind_size[k] = (1u << ind[k].log2) * 256 + 1;  // indirect.cpp:15-19
    if (ind[k].log2 >= 15) {
      L.ind_sid[k] = (uint8_t)next_sid++;
      worst += 8 * max_len < L.ind_size[k] ? 8 * max_len : L.ind_size[k];  // one new slot per bit at most
      dense_bytes += ((uint64_t)L.ind_size[k] + 1) / 2 * 4;
    }
  }
  for (int k = 0; k < NMATCH; ++k)
    if (mt[k].log2 >= 21) {
      L.match_sid[k] = (uint8_t)next_sid++;
      worst += max_len < (1ull << mt[k].log2) ? max_len : (1ull << mt[k].log2);  // one store per byte
      dense_bytes += 4ull << mt[k].log2;
    }
  for (int k = 0; k < NIH; ++k)
    if (ih[k].log2 >= 24) {
      L.ih_sid[k] = (uint8_t)next_sid++;
      worst += max_len + 1 < (1ull << ih[k].log2) ? max_len + 1 : (1ull << ih[k].log2);
      dense_bytes += 4ull << ih[k].log2;
    }
  // Text touches ~5.2e5 * (len / 4096)^0.72 slots (measured on the synthetic-text chunks: 517 K at 4 KiB,
  // 1.45 M at 16 KiB, 3.6 M at 64 KiB); 35 % head room on top, load limit 3/4.
  const double typical = 5.2e5 * pow((double)(max_len > 4096 ? max_len : 4096) / 4096.0, 0.72) * 1.35;
  uint64_t cap = Pow2Ceil(roomy ? worst * 4 / 3 + 64 : (uint64_t)(typical * 4.0 / 3.0) + pre->sparse_entries * 4 / 3 + 64);
  const uint64_t cap_worst = Pow2Ceil(worst * 4 / 3 + 64);
  if (cap > cap_worst) cap = cap_worst;
  if (force_dense || cap * 8 >= dense_bytes || cap > (1ull << 31)) {  // long streams: the dense tables are smaller
    for (int k = 0; k < NIND; ++k) L.ind_sid[k] = 0;
    for (int k = 0; k < NMATCH; ++k) L.match_sid[k] = 0;
    for (int k = 0; k < NIH; ++k) L.ih_sid[k] = 0;
  } else {
    L.sparse_mask = (uint32_t)(cap - 1);
    L.sparse_limit = (uint32_t)(cap / 4 * 3);
    L.sparse = take(cap * 8);
  }
  for (int k = 0; k < NIND; ++k)
    if (!L.ind_sid[k]) L.ind_tab[k] = take(((uint64_t)L.ind_size[k] + 1) / 2 * 4);
  L.ind_pred = take((uint64_t)NIND * 512 * 4);
  for (int k = 0; k < NMATCH; ++k)
    if (!L.match_sid[k]) L.match_tab[k] = take((4ull << mt[k].log2));
  L.match_pred = take(NMATCH * 256 * 4);
  L.match_cnt = take(NMATCH * 256 * 4);
  L.history_cap = max_len + 8 + pre->history_bytes;
  L.history = take(L.history_cap);
  for (int k = 0; k < NIH; ++k)
    if (!L.ih_sid[k]) L.ih_tab[k] = take(4ull << ih[k].log2);
  // Weight-set pool: a mixer can create at most one set per distinct gate context it ever sees:
  // min(table size, bytes + 1) for byte-level contexts, min(table size, bits + 1) otherwise.
  uint64_t sets = 1 + pre->mixer_sets;
  for (int m = 0; m < NMIX; ++m) {
    L.mix_dir[m] = take(4ull << mx[m].log2);
    const uint64_t t = 1ull << mx[m].log2;
    const bool bit_level = mx[m].ctx == C_SLPR || mx[m].ctx == C_LBPR || mx[m].ctx == C_BIT_CONTEXT || mx[m].ctx == C_LONGEST;
    const uint64_t seen = bit_level ? 8 * max_len + 1 : max_len + 1;
    sets += t < seen ? t : seen;
  }
  if (!roomy) {  // text creates ~0.35 sets per byte
    const uint64_t typical = 8192 + max_len * 3 / 4 + pre->mixer_sets;
    if (typical < sets) sets = typical;
  }
  L.mix_pool_sets = (uint32_t)sets;
  L.mix_set_stride = 120;  // 4 header words {steps, 0, 0, 0} + up to 116 weight slots (29 float4), 16-byte multiple
  L.mix_pool = take(sets * L.mix_set_stride * 4);
  const uint64_t wsz = (uint64_t)L_WSIZE * 4;
  L.l_w = take(wsz); L.l_m = take(wsz); L.l_v = take(wsz);
  L.l_gb = take(8 * 3 * L_CELLS * 4);
  L.l_wout = take((uint64_t)L_HORIZON * L_HID * L_NOUT * 4);
  L.l_lin = take((uint64_t)L_HORIZON * (L_NIN + 1) * 4);
  L.l_out = take((uint64_t)L_HORIZON * L_NOUT * 4);
  L.l_gstate = take(3ull * L_HORIZON * L_CELLS * 4);
  L.l_norm = take(3ull * L_HORIZON * L_CELLS * 4);
  L.l_ivar = take(3ull * L_HORIZON * 4);
  L.l_tanh = take((uint64_t)L_HORIZON * L_CELLS * 4);
  L.l_ig = take((uint64_t)L_HORIZON * L_CELLS * 4);
  L.l_last = take((uint64_t)L_HORIZON * L_CELLS * 4);
  L.l_errh = take(3ull * L_HORIZON * L_CELLS * 4);
  L.l_wt = take(3ull * L_CELLS * L_CELLS * 4);
  L.p_state = take(sizeof(PpmdState));
  // PPMd heap window (ppmd.cuh): smallest 2^k whose unit window (2^k - units_start mod 2^k) holds the
  // units this stream can need (~19 B/byte on text; 400 B/byte worst-case sizing)
  // and whose bottom holds the text area.
  {
    const uint64_t want_units = (roomy ? 400 : 90) * max_len + (256u << 10) + pre->ppmd_unit_bytes;
    const uint64_t want_text = max_len + 64 + pre->ppmd_text_bytes;
    // Inside the window the low units grow up from units_start mod P and the high units grow down from
    // heap_end mod P (= the top of the window while P divides 2000 MiB, i.e. up to 16 MiB). The kernel's test
    // "low + high <= units_cap" is only right when the two areas share ONE gap that does not contain the text at
    // the bottom of the window: units_start mod P < heap_end mod P. Window sizes where the gap would wrap through 0
    // (32 .. 256 MiB) are skipped; 512 MiB (214 MiB of units) and 1 GiB (726 MiB) are valid again.
    uint64_t P = 1ull << 20;
    for (;; P <<= 1) {
      const PpmdWindow w = PpmdWindowOf(P);
      if (w.valid && ((w.text_room >= want_text && w.units_room >= want_units) || P >= (1ull << 30))) break;
    }
    const PpmdWindow w = PpmdWindowOf(P);
    const uint64_t text_room = w.text_room, units_room = w.units_room;
    L.p_mask = (uint32_t)(P - 1);
    L.p_text_cap = (uint32_t)(want_text < text_room ? want_text : text_room);
    uint64_t units = units_room / 48 * 48;
    const uint64_t units_max = 800ull << 20;   // must stay below half of the virtual units area
    if (units > units_max) units = units_max / 48 * 48;
    L.p_units_cap = (uint32_t)units;
    L.p_heap = take(P);
  }
  L.total = off;
  return L;
}

// Arena of a stream that starts from a loaded model in OVERLAY mode (ArenaLayout::ov, stream_kernel.cuh): the model's
// tables, weight-set pool and history are shared read-only; this arena holds the overlay map, the local pool / history and
// private copies of the densely updated state. base = the model's layout, pre = what the model holds, learn_bytes = bytes
// the stream learns at most (the prompt), new_bytes = bytes it adds in total (prompt + samples). Sized for the worst case:
// overlay entries = every table write (41 Indirect states per learned bit, 6 Match pointers per learned byte, one
// IndirectHash update per table and byte, one directory entry per new or changed weight set); weight sets = at most 27
// byte-gated + 6 x 8 bit-gated or longest-match-gated ones change per learned byte.
inline ArenaLayout MakeOverlayLayout(const ArenaLayout& base, const Preload& pre, uint64_t learn_bytes, uint64_t new_bytes, bool force_segmented = false) {
  ArenaLayout L;
  memset(&L, 0, sizeof(L));
  uint64_t off = 0;
  auto take = [&](uint64_t bytes) { uint64_t o = off; off = AlignUp(off + bytes, 256); return o; };
  L.ov = 1;
  for (int k = 0; k < NIND; ++k) { L.ind_size[k] = base.ind_size[k]; L.ind_sid[k] = (uint8_t)(OV_SID_IND + k); }
  for (int k = 0; k < NMATCH; ++k) L.match_sid[k] = (uint8_t)(OV_SID_MATCH + k);
  for (int k = 0; k < NIH; ++k) L.ih_sid[k] = (uint8_t)(OV_SID_IH + k);
  const uint64_t sets = 75 * learn_bytes + 64;
  const uint64_t entries = learn_bytes * (8 * NIND + NMATCH) + new_bytes * NIH + sets + 64;
  const uint64_t cap = Pow2Ceil(entries * 4 / 3 + 64);
  L.sparse_mask = (uint32_t)(cap - 1);
  L.sparse_limit = (uint32_t)(cap / 4 * 3);
  L.sparse = take(cap * 8);
  L.ind_pred = take((uint64_t)NIND * 512 * 4);
  L.match_pred = take(NMATCH * 256 * 4);
  L.match_cnt = take(NMATCH * 256 * 4);
  L.base_hist = (uint32_t)pre.history_bytes;
  L.history_cap = pre.history_bytes + learn_bytes + 8;
  L.history = take(learn_bytes + 8);
  L.base_sets = (uint32_t)(pre.mixer_sets + 1);
  L.mix_pool_sets = (uint32_t)(L.base_sets + sets);
  L.mix_set_stride = base.mix_set_stride;
  L.mix_pool = take(sets * L.mix_set_stride * 4);
  const uint64_t wsz = (uint64_t)L_WSIZE * 4;
  L.l_w = take(wsz); L.l_m = take(wsz); L.l_v = take(wsz);
  L.l_gb = take(8 * 3 * L_CELLS * 4);
  L.l_wout = take((uint64_t)L_HORIZON * L_HID * L_NOUT * 4);
  L.l_lin = take((uint64_t)L_HORIZON * (L_NIN + 1) * 4);
  L.l_out = take((uint64_t)L_HORIZON * L_NOUT * 4);
  L.l_gstate = take(3ull * L_HORIZON * L_CELLS * 4);
  L.l_norm = take(3ull * L_HORIZON * L_CELLS * 4);
  L.l_ivar = take(3ull * L_HORIZON * 4);
  L.l_tanh = take((uint64_t)L_HORIZON * L_CELLS * 4);
  L.l_ig = take((uint64_t)L_HORIZON * L_CELLS * 4);
  L.l_last = take((uint64_t)L_HORIZON * L_CELLS * 4);
  L.l_errh = take(3ull * L_HORIZON * L_CELLS * 4);
  L.l_wt = take(3ull * L_CELLS * L_CELLS * 4);
  L.p_state = take(sizeof(PpmdState));
  // PPMd: the model's power-of-two window copied whole while it is small (one AND per heap access); beyond 32 MiB a
  // segmented private backing (ppmd.cuh; two compares per access, measured 25 % slower generation on a small model): what
  // the model uses of each area + room for new_bytes more (one text byte per input byte; units: the same worst-case
  // allowance MakeLayout's roomy class uses, per area)
  if (!force_segmented && (uint64_t)base.p_mask + 1 <= (32ull << 20)) {
    L.p_mask = base.p_mask; L.p_text_cap = base.p_text_cap; L.p_units_cap = base.p_units_cap;
    L.p_heap = take((uint64_t)L.p_mask + 1);
  } else {
    const uint64_t grow = 400 * new_bytes + (64u << 10);
    L.p_mask = 0;
    L.p_text_cap = (uint32_t)AlignUp(pre.ppmd_text_bytes + new_bytes + 64, 16);
    L.p_seg_lo = L.p_text_cap;
    L.p_lo_cap = (uint32_t)AlignUp(pre.ppmd_lo_bytes + grow, 48);
    L.p_hi_cap = (uint32_t)AlignUp(pre.ppmd_hi_bytes + grow, 48);
    L.p_units_cap = L.p_lo_cap + L.p_hi_cap;
    L.p_heap = take((uint64_t)L.p_seg_lo + L.p_lo_cap + L.p_hi_cap + 64);
  }
  L.total = off;
  return L;
}

// decay[s] = (float)(0.9 / pow(0.0000001 * s + 0.8, 0.8)) — mixer.cpp:111, a function of the mixer's
// global step counter only; evaluated with the HOST libm exactly as the reference does.
inline void FillDecayTable(std::vector<float>& t, uint64_t n) {
  const uint64_t old = t.size();
  if (n <= old) return;
  t.resize(n);
  for (uint64_t s = old; s < n; ++s) {
    unsigned long long steps = s;
    float decay = 0.9 / pow(0.0000001 * steps + 0.8, 0.8);
    t[s] = decay;
  }
}

// Adam scalars per update step t = 1..3000 (lstm-layer.cpp:12-34): alpha, 1 - beta1^t, 1 - beta2^t,
// with the reference's exact expression types (powf below the limit, double pow at the limit).
inline void FillAdamTable(std::vector<float>& t) {
  t.assign(4 * (L_UPDATE_LIMIT + 1), 0.0f);
  const float beta1 = 0.025, beta2 = 0.9999;
  const float learning_rate = 0.03;
  const unsigned long long update_limit = L_UPDATE_LIMIT;
  for (unsigned long long steps = 1; steps <= update_limit; ++steps) {
    float tt = steps;
    float alpha, d1, d2;
    if (tt < update_limit) {
      alpha = learning_rate * 0.1f / sqrtf(5e-5f * tt + 1.0f);
      d1 = (float)(1.0f - powf(beta1, tt));
      d2 = (float)(1.0f - powf(beta2, tt));
    } else {
      alpha = learning_rate * 0.1f / sqrtf(5e-5f * update_limit + 1.0f);
      d1 = (float)(1.0f - pow((double)beta1, (double)update_limit));
      d2 = (float)(1.0f - pow((double)beta2, (double)update_limit));
    }
    t[4 * steps + 0] = alpha; t[4 * steps + 1] = d1; t[4 * steps + 2] = d2;
  }
}

// Initial LSTM gate weights (lstm-layer.cpp:176-195): srand(0xDEADBEEF) (predictor.cpp:18), then
// per cell i, per column j, one glibc rand() draw for each of the three gates in turn; forget-gate
// bias column = 1. Stored in the device layout: out[LstmW(g, j, i)] (spec.cuh).
inline void FillLstmInit(std::vector<float>& out) {
  out.assign((size_t)L_WSIZE, 0.0f);
  srand(0xDEADBEEF);
  float val = sqrtf(6.0f / float(256 + 256));
  float low = -val;
  float range = 2 * val;
  for (int i = 0; i < L_CELLS; ++i) {
    for (int j = 0; j < L_ROW; ++j) {
      for (int g = 0; g < 3; ++g) {
        float r = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
        out[LstmW(g, j, i)] = low + r * range;
      }
    }
    out[LstmW(0, L_ROW - 1, i)] = 1;
  }
}

}  // namespace gmx
#endif
