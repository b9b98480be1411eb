// The reference's checkpoint format (`<path>.short` + `<path>.long`, reference src/predictor.cpp:389-420)
// <-> one stream's device state (arena image + parked StreamSmem). Pure host C++: used by the C-ABI
// library (host.cu) and by the CPU emulation tests.
//
// `Image` holds a checkpoint field for field in the order the reference serialises it (SURVEY.md
// appendix B): every model's WriteToDisk in construction order, then ShortTermMemory
// (memory/short-term-memory.cpp:3-59), and LongTermMemory (memory/long-term-memory.cpp:6-108). Big tables
// are kept as sorted key/value lists whatever encoding (sparse or dense) the file used, so a parsed
// image costs memory proportional to what the model has seen, not to the 1.1 GB of dense tables.
//
//   Parse      .short/.long bytes -> Image            (reads files written by the reference, unchanged)
//   Serialize  Image -> .short/.long bytes            (Parse followed by Serialize is byte-identical)
//   ToArena    Image -> arena image + StreamSmem      (a stream that continues exactly where the checkpoint stopped)
//   FromArena  arena image + StreamSmem -> Image      (what Predictor::WriteCheckpoint would write for that stream)
//
// Only byte-boundary checkpoints are supported (the reference writes them after whole files). Fields the
// reference writes but never reads back before overwriting (per-byte / per-BPTT scratch: PPMd SQ/trF/trT/
// AuxUnit/saved_pc, NeuronLayer error_/update_/transpose_, analysis entropy) survive Parse -> Serialize
// untouched; FromArena fills them with neutral values (documented at each field).
#ifndef GMIX_B200_CHECKPOINT_H_
#define GMIX_B200_CHECKPOINT_H_
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "layout.h"

namespace gmx {
namespace ckpt {

enum : int { kEntropy = NPRED + NMIX, kRot = 1000, kTr = L_HID };  // 123 analysis slots; rotating_history; transpose_ rows

struct ByteReader {
  const uint8_t* p; uint64_t n, pos; bool ok;
  ByteReader(const void* data, uint64_t len) : p((const uint8_t*)data), n(len), pos(0), ok(true) {}
  void raw(void* dst, uint64_t bytes) {
    if (!ok || bytes > n - pos) { ok = false; memset(dst, 0, bytes); return; }
    memcpy(dst, p + pos, bytes); pos += bytes;
  }
  template <typename T> T get() { T v; raw(&v, sizeof(T)); return v; }
  template <typename T> void arr(T* dst, uint64_t count) { raw(dst, count * sizeof(T)); }
  const uint8_t* take(uint64_t bytes) {  // zero-copy view
    if (!ok || bytes > n - pos) { ok = false; return nullptr; }
    const uint8_t* r = p + pos; pos += bytes; return r;
  }
};
struct ByteWriter {
  std::vector<uint8_t>* v;
  void raw(const void* src, uint64_t bytes) { const uint8_t* s = (const uint8_t*)src; v->insert(v->end(), s, s + bytes); }
  template <typename T> void put(T x) { raw(&x, sizeof(T)); }
  template <typename T> void arr(const T* src, uint64_t count) { raw(src, count * sizeof(T)); }
};

struct IndEntry { uint32_t key; uint8_t ns, rm; };

struct Image {
  // ---- .short -------------------------------------------------------------------------------------
  uint8_t first_prediction = 1;                         // BasicContexts (contexts/basic-contexts.cpp:56-58)
  // ModPPMD (models/mod_ppmd.cpp:1684-1689, ppmd_Model::WriteToDisk :1384-1482)
  int32_t ppm_top = 255, ppm_mid = 127, ppm_bot = 0;
  struct BlkNode { uint32_t stamp, next; } blist[PPMD_N_INDEXES + 1];
  uint32_t glue_count = 0, glue_count1 = 0;
  uint64_t sa_size = PPMD_HEAP_END;
  uint64_t p_text = 0, units_start = PPMD_UNITS_START, lo_unit = 0, hi_unit = 0, aux_unit = 0, found_state = 0, max_context = 0, saved_pc = 0;
  int32_t order_fall = 0; uint32_t esc_count = 0; uint32_t char_mask[256];
  int32_t bsumm = 0, run_length = 0, init_rl = 0, num_masked = 0, prev_success = 0;
  uint16_t bin_summ[25][64];
  struct See { uint8_t count, shift; uint16_t summ; } see2[23][32], dummy_see2;
  struct QSym { uint16_t freq, sym, total; } sq[1024];   // serialised as freq, sym, total (:1432-1436)
  uint32_t sq_ptr = 0; uint32_t sqp[256], trf[256], trt[256]; uint32_t cxt = 0, y = 1;
  // the three parts of the 2000 MiB heap that are ever written: text [0, p_text), low units
  // [units_start, lo_unit), high units [hi_unit, heap end); everything else is zero
  std::vector<uint8_t> heap_text, heap_lo, heap_hi;
  // LstmModel (models/lstm-model.cpp:61-67), Lstm (lstm.cpp:124-140), LstmLayer (lstm-layer.cpp:356-374)
  int32_t l_top = 255, l_mid = 127, l_bot = 0; float l_probs[256];
  uint32_t l_input_history[L_HORIZON]; float l_hidden[L_HID], l_hidden_error[L_CELLS];
  std::vector<float> l_layer_input, l_output;            // [100][307], [100][256]
  uint32_t l_epoch = 0;
  float ll_state[L_CELLS], ll_state_error[L_CELLS], ll_stored_error[L_CELLS];
  std::vector<float> ll_tanh, ll_ig, ll_last;            // [100][50]
  uint32_t ll_epoch = 0; uint64_t ll_update_steps = 0;
  struct Neuron {                                        // NeuronLayer::WriteToDisk lstm-layer.cpp:62-91
    float error[L_CELLS], ivar[L_HORIZON];
    float gamma[L_CELLS], gamma_u[L_CELLS], gamma_m[L_CELLS], gamma_v[L_CELLS], beta[L_CELLS], beta_u[L_CELLS], beta_m[L_CELLS], beta_v[L_CELLS];
    std::vector<float> state, update, m, v, transpose, norm;   // [100][50], [50][563] x3, [51][50], [100][50]
  } nl[3];
  struct MatchState { uint64_t cur; uint8_t byte, bitpos, len; } match[NMATCH];   // models/match.cpp:111-116
  struct IH { std::vector<std::pair<uint32_t, uint32_t>> e; uint64_t outer; uint32_t hash; } ih[NIH];  // contexts/indirect-hash.cpp:33-54
  struct MixerState { uint64_t steps, max_steps, contexts_seen; } mixer[NMIX];     // mixer/mixer.cpp:178-182
  // ShortTermMemory (memory/short-term-memory.cpp:3-59)
  float predictions[NPRED]; int32_t new_bit = 0, recent_bits = 1;
  uint32_t bit_context = 0, last_byte = 0, always_zero = 0, h3 = 0, h4 = 0, h5 = 0, h6 = 0;
  uint32_t ih_ctx[NIH], interval[9], skip[15], lbpr = 0, slpr = 0;   // skip[] in FILE order (differs from construction order)
  float l0_out[NL0], l1_out[NL1], final_out = 0;
  uint32_t longest = 0; uint64_t bits_seen = 0; double entropy[kEntropy]; uint32_t lstm_ctx = 0;
  uint8_t rot[kRot]; uint32_t rot_pos = 0; uint32_t recent_bytes[10];
  // ---- .long --------------------------------------------------------------------------------------
  struct Ind { std::vector<IndEntry> e; float ns_pred[256], rm_pred[256]; } ind[NIND];
  struct Mix { uint32_t input_size = 0; std::vector<uint32_t> ctx; std::vector<uint64_t> steps; std::vector<float> w; } mix[NMIX];
  std::vector<float> wout, wgate;                        // [100][256][51], [3][50][563]
  std::vector<uint8_t> history;
  struct Mt { std::vector<std::pair<uint32_t, uint64_t>> e; float pred[256]; int32_t cnt[256]; } mt[NMATCH];
};

// ShortTermMemory serialises the 15 skip contexts in declaration order, the predictor constructs them in
// another (predictor.cpp:122-185): file position -> construction index (= C_SK0 + index here).
static const int kSkipFileToModel[15] = {0, 1, 2, 3, 4, 5, 6, 8, 9, 7, 10, 11, 12, 13, 14};

inline const IndirectSpec* IndSpecs() { static const IndirectSpec t[NIND] = {GMX_INDIRECT_SPECS}; return t; }
inline const IHSpec* IHSpecs() { static const IHSpec t[NIH] = {GMX_IH_SPECS}; return t; }
inline const MatchSpec* MatchSpecs() { static const MatchSpec t[NMATCH] = {GMX_MATCH_SPECS}; return t; }
inline const MixerSpec* MixerSpecs() { static const MixerSpec t[NMIX] = {GMX_MIXER_SPECS}; return t; }
inline uint32_t IndSize(int k) { return (1u << IndSpecs()[k].log2) * 256 + 1; }

// ---- Parse ------------------------------------------------------------------------------------------
inline bool ParseHeap(ByteReader& r, Image* im, std::string* err) {
  const int32_t nseq = r.get<int32_t>();
  if (!r.ok || nseq < 0) { *err = "bad PPMd zero-run count"; return false; }
  std::vector<uint64_t> cnt(nseq), start(nseq);
  for (int i = 0; i < nseq; ++i) { cnt[i] = r.get<uint64_t>(); start[i] = r.get<uint64_t>(); }
  if (im->units_start != PPMD_UNITS_START || im->sa_size != PPMD_HEAP_END || im->p_text > im->units_start ||
      im->lo_unit < im->units_start || im->hi_unit > im->sa_size || im->lo_unit > im->hi_unit) {
    *err = "PPMd heap geometry differs from Init(20, 2000, 1, 0) (a checkpoint taken after the out-of-memory path ran is not supported)";
    return false;
  }
  im->heap_text.assign(im->p_text, 0);
  im->heap_lo.assign(im->lo_unit - im->units_start, 0);
  im->heap_hi.assign(im->sa_size - im->hi_unit, 0);
  auto place = [&](uint64_t v, const uint8_t* src, uint64_t len) -> bool {  // literal bytes [v, v+len)
    struct Seg { uint64_t lo, hi; uint8_t* dst; } seg[3] = {{0, im->p_text, im->heap_text.data()},
                                                           {im->units_start, im->lo_unit, im->heap_lo.data()},
                                                           {im->hi_unit, im->sa_size, im->heap_hi.data()}};
    uint64_t covered = 0;
    for (auto& s : seg) {
      const uint64_t a = std::max(v, s.lo), b = std::min(v + len, s.hi);
      if (a < b) { memcpy(s.dst + (a - s.lo), src + (a - v), b - a); covered += b - a; }
    }
    if (covered == len) return true;
    for (uint64_t i = 0; i < len; ++i) {  // bytes outside the three areas must be zero
      const uint64_t q = v + i;
      const bool in = q < im->p_text || (q >= im->units_start && q < im->lo_unit) || q >= im->hi_unit;
      if (!in && src[i] != 0) return false;
    }
    return true;
  };
  uint64_t pos = 0;
  for (int i = 0; i <= nseq; ++i) {
    const uint64_t end = i < nseq ? start[i] : im->sa_size;
    if (end < pos || end > im->sa_size) { *err = "bad PPMd zero run"; return false; }
    const uint8_t* src = r.take(end - pos);
    if (!r.ok) { *err = "truncated PPMd heap"; return false; }
    if (end > pos && !place(pos, src, end - pos)) { *err = "PPMd heap has data outside the text/unit areas"; return false; }
    pos = i < nseq ? start[i] + cnt[i] : end;
  }
  return true;
}

inline bool Parse(const void* short_blob, uint64_t short_len, const void* long_blob, uint64_t long_len, Image* im, std::string* err) {
  ByteReader r(short_blob, short_len);
  im->first_prediction = r.get<uint8_t>();
  im->ppm_top = r.get<int32_t>(); im->ppm_mid = r.get<int32_t>(); im->ppm_bot = r.get<int32_t>();
  for (auto& b : im->blist) { b.stamp = r.get<uint32_t>(); b.next = r.get<uint32_t>(); }
  im->glue_count = r.get<uint32_t>(); im->glue_count1 = r.get<uint32_t>(); im->sa_size = r.get<uint64_t>();
  im->p_text = r.get<uint64_t>(); im->units_start = r.get<uint64_t>(); im->lo_unit = r.get<uint64_t>(); im->hi_unit = r.get<uint64_t>();
  im->aux_unit = r.get<uint64_t>(); im->found_state = r.get<uint64_t>(); im->max_context = r.get<uint64_t>(); im->saved_pc = r.get<uint64_t>();
  im->order_fall = r.get<int32_t>(); im->esc_count = r.get<uint32_t>(); r.arr(im->char_mask, 256);
  im->bsumm = r.get<int32_t>(); im->run_length = r.get<int32_t>(); im->init_rl = r.get<int32_t>();
  im->num_masked = r.get<int32_t>(); im->prev_success = r.get<int32_t>();
  r.arr(&im->bin_summ[0][0], 25 * 64);
  for (int i = 0; i < 23; ++i) for (int j = 0; j < 32; ++j) { auto& s = im->see2[i][j]; s.count = r.get<uint8_t>(); s.shift = r.get<uint8_t>(); s.summ = r.get<uint16_t>(); }
  im->dummy_see2.count = r.get<uint8_t>(); im->dummy_see2.shift = r.get<uint8_t>(); im->dummy_see2.summ = r.get<uint16_t>();
  for (auto& q : im->sq) { q.freq = r.get<uint16_t>(); q.sym = r.get<uint16_t>(); q.total = r.get<uint16_t>(); }
  im->sq_ptr = r.get<uint32_t>();
  for (int i = 0; i < 256; ++i) { im->sqp[i] = r.get<uint32_t>(); im->trf[i] = r.get<uint32_t>(); im->trt[i] = r.get<uint32_t>(); }
  im->cxt = r.get<uint32_t>(); im->y = r.get<uint32_t>();
  if (!r.ok) { *err = "truncated .short (PPMd header)"; return false; }
  if (!ParseHeap(r, im, err)) return false;
  im->l_top = r.get<int32_t>(); im->l_mid = r.get<int32_t>(); im->l_bot = r.get<int32_t>(); r.arr(im->l_probs, 256);
  r.arr(im->l_input_history, L_HORIZON); r.arr(im->l_hidden, L_HID); r.arr(im->l_hidden_error, L_CELLS);
  im->l_layer_input.resize((size_t)L_HORIZON * L_NIN); r.arr(im->l_layer_input.data(), im->l_layer_input.size());
  im->l_output.resize((size_t)L_HORIZON * L_NOUT); r.arr(im->l_output.data(), im->l_output.size());
  im->l_epoch = r.get<uint32_t>();
  r.arr(im->ll_state, L_CELLS); r.arr(im->ll_state_error, L_CELLS); r.arr(im->ll_stored_error, L_CELLS);
  const size_t hc = (size_t)L_HORIZON * L_CELLS;
  im->ll_tanh.resize(hc); im->ll_ig.resize(hc); im->ll_last.resize(hc);
  r.arr(im->ll_tanh.data(), hc); r.arr(im->ll_ig.data(), hc); r.arr(im->ll_last.data(), hc);
  im->ll_epoch = r.get<uint32_t>(); im->ll_update_steps = r.get<uint64_t>();
  for (auto& n : im->nl) {
    r.arr(n.error, L_CELLS); r.arr(n.ivar, L_HORIZON);
    r.arr(n.gamma, L_CELLS); r.arr(n.gamma_u, L_CELLS); r.arr(n.gamma_m, L_CELLS); r.arr(n.gamma_v, L_CELLS);
    r.arr(n.beta, L_CELLS); r.arr(n.beta_u, L_CELLS); r.arr(n.beta_m, L_CELLS); r.arr(n.beta_v, L_CELLS);
    const size_t wsz = (size_t)L_CELLS * L_ROW;
    n.state.resize(hc); n.update.resize(wsz); n.m.resize(wsz); n.v.resize(wsz); n.transpose.resize((size_t)kTr * L_CELLS); n.norm.resize(hc);
    r.arr(n.state.data(), hc); r.arr(n.update.data(), wsz); r.arr(n.m.data(), wsz); r.arr(n.v.data(), wsz);
    r.arr(n.transpose.data(), n.transpose.size()); r.arr(n.norm.data(), hc);
  }
  if (!r.ok) { *err = "truncated .short (LSTM)"; return false; }
  for (auto& m : im->match) { m.cur = r.get<uint64_t>(); m.byte = r.get<uint8_t>(); m.bitpos = r.get<uint8_t>(); m.len = r.get<uint8_t>(); }
  for (int k = 0; k < NIH; ++k) {
    auto& h = im->ih[k];
    const uint32_t tsize = 1u << IHSpecs()[k].log2;
    const uint32_t n = r.get<uint32_t>();
    h.e.clear();
    if (n < tsize / 2) {
      h.e.resize(n);
      for (auto& kv : h.e) { kv.first = r.get<uint32_t>(); kv.second = r.get<uint32_t>(); }
      for (auto& kv : h.e) if (kv.first >= tsize) { *err = "IndirectHash key out of range"; return false; }
    } else {
      const uint32_t* t = (const uint32_t*)r.take((uint64_t)tsize * 4);
      if (!r.ok) { *err = "truncated .short (IndirectHash table)"; return false; }
      for (uint32_t i = 0; i < tsize; ++i) { uint32_t v; memcpy(&v, t + i, 4); if (v) h.e.emplace_back(i, v); }
    }
    h.outer = r.get<uint64_t>(); h.hash = r.get<uint32_t>();
  }
  for (auto& m : im->mixer) { m.steps = r.get<uint64_t>(); m.max_steps = r.get<uint64_t>(); m.contexts_seen = r.get<uint64_t>(); }
  r.arr(im->predictions, NPRED);
  im->new_bit = r.get<int32_t>(); im->recent_bits = r.get<int32_t>();
  im->bit_context = r.get<uint32_t>(); im->last_byte = r.get<uint32_t>(); im->always_zero = r.get<uint32_t>();
  im->h3 = r.get<uint32_t>(); im->h4 = r.get<uint32_t>(); im->h5 = r.get<uint32_t>(); im->h6 = r.get<uint32_t>();
  r.arr(im->ih_ctx, NIH); r.arr(im->interval, 9); r.arr(im->skip, 15);
  im->lbpr = r.get<uint32_t>(); im->slpr = r.get<uint32_t>();
  r.arr(im->l0_out, NL0); r.arr(im->l1_out, NL1); im->final_out = r.get<float>();
  im->longest = r.get<uint32_t>(); im->bits_seen = r.get<uint64_t>(); r.arr(im->entropy, kEntropy);
  im->lstm_ctx = r.get<uint32_t>(); r.arr(im->rot, kRot); im->rot_pos = r.get<uint32_t>(); r.arr(im->recent_bytes, 10);
  if (!r.ok || r.pos != r.n) { *err = r.ok ? "trailing bytes in .short" : "truncated .short"; return false; }

  ByteReader q(long_blob, long_len);
  for (int k = 0; k < NIND; ++k) {
    auto& d = im->ind[k];
    const uint32_t tsize = IndSize(k);
    const uint32_t n = q.get<uint32_t>();
    d.e.clear();
    if (n < tsize / 3) {
      d.e.resize(n);
      for (auto& e : d.e) { e.key = q.get<uint32_t>(); e.ns = q.get<uint8_t>(); e.rm = q.get<uint8_t>(); }
      for (auto& e : d.e) if (e.key >= tsize) { *err = "Indirect key out of range"; return false; }
    } else {
      const uint8_t* ns = q.take(tsize);
      const uint8_t* rm = q.take(tsize);
      if (!q.ok) { *err = "truncated .long (Indirect table)"; return false; }
      for (uint32_t i = 0; i < tsize; ++i) if (ns[i] != 255 || rm[i] != 0) d.e.push_back(IndEntry{i, ns[i], rm[i]});
    }
    q.arr(d.ns_pred, 256); q.arr(d.rm_pred, 256);
  }
  for (int m = 0; m < NMIX; ++m) {
    auto& d = im->mix[m];
    const uint32_t n = q.get<uint32_t>();
    d.input_size = q.get<uint32_t>();
    if (!q.ok || (n && d.input_size != (uint32_t)MixerWeights(m)) || n > (1u << MixerSpecs()[m].log2)) { *err = "mixer table does not match the model graph"; return false; }
    d.ctx.resize(n); d.steps.resize(n); d.w.resize((size_t)n * d.input_size);
    for (uint32_t i = 0; i < n; ++i) {
      d.ctx[i] = q.get<uint32_t>(); d.steps[i] = q.get<uint64_t>();
      q.arr(d.w.data() + (size_t)i * d.input_size, d.input_size);
      if (d.ctx[i] >> MixerSpecs()[m].log2) { *err = "mixer context out of range"; return false; }
    }
  }
  im->wout.resize((size_t)L_HORIZON * L_NOUT * L_HID); q.arr(im->wout.data(), im->wout.size());
  im->wgate.resize((size_t)3 * L_CELLS * L_ROW); q.arr(im->wgate.data(), im->wgate.size());
  const uint64_t hl = q.get<uint64_t>();
  if (!q.ok || hl > q.n - q.pos) { *err = "truncated .long (history)"; return false; }
  im->history.resize(hl); q.arr(im->history.data(), hl);
  for (int k = 0; k < NMATCH; ++k) {
    auto& d = im->mt[k];
    const uint32_t tsize = 1u << MatchSpecs()[k].log2;
    const uint32_t n = q.get<uint32_t>();
    d.e.clear();
    auto ptr5 = [](const uint8_t* b) { return (uint64_t)b[0] | ((uint64_t)b[1] << 8) | ((uint64_t)b[2] << 16) | ((uint64_t)b[3] << 24) | ((uint64_t)b[4] << 32); };
    if (n < (5.0 / 9.0) * tsize) {
      d.e.resize(n);
      for (auto& kv : d.e) { kv.first = q.get<uint32_t>(); uint8_t b[5]; q.arr(b, 5); kv.second = ptr5(b); if (kv.first >= tsize) { *err = "Match key out of range"; return false; } }
    } else {
      const uint8_t* t = q.take((uint64_t)tsize * 5);
      if (!q.ok) { *err = "truncated .long (Match table)"; return false; }
      for (uint32_t i = 0; i < tsize; ++i) { const uint64_t v = ptr5(t + (size_t)i * 5); if (v) d.e.emplace_back(i, v); }
    }
    q.arr(d.pred, 256); q.arr(d.cnt, 256);
  }
  if (!q.ok || q.pos != q.n) { *err = q.ok ? "trailing bytes in .long" : "truncated .long"; return false; }
  return true;
}

// ---- Serialize --------------------------------------------------------------------------------------
// The reference scans all 2000 MiB for runs of more than 100 zero bytes (:1446-1481); the same run list
// falls out of scanning the three backed areas with the untouched gaps between them taken as zeros.
inline void SerializeHeap(const Image& im, ByteWriter& w) {
  struct Piece { uint64_t start, len; const uint8_t* data; };   // data == nullptr: zeros
  const Piece pieces[5] = {{0, im.p_text, im.heap_text.data()},
                           {im.p_text, im.units_start - im.p_text, nullptr},
                           {im.units_start, im.lo_unit - im.units_start, im.heap_lo.data()},
                           {im.lo_unit, im.hi_unit - im.lo_unit, nullptr},
                           {im.hi_unit, im.sa_size - im.hi_unit, im.heap_hi.data()}};
  std::vector<uint64_t> counts, starts;
  uint64_t run = 0, run_start = 0;
  auto end_run = [&]() { if (run > 100) { counts.push_back(run); starts.push_back(run_start); } run = 0; };
  for (const Piece& pc : pieces) {
    if (!pc.len) continue;
    if (!pc.data) { if (!run) run_start = pc.start; run += pc.len; continue; }
    for (uint64_t i = 0; i < pc.len; ++i) {
      if (pc.data[i] == 0) { if (!run) run_start = pc.start + i; ++run; }
      else if (run) end_run();
    }
  }
  // a run that reaches the end of the heap is never closed by the reference's loop, hence never listed
  w.put<int32_t>((int32_t)counts.size());
  for (size_t i = 0; i < counts.size(); ++i) { w.put<uint64_t>(counts[i]); w.put<uint64_t>(starts[i]); }
  size_t sp = 0;
  for (const Piece& pc : pieces) {
    uint64_t i = 0;
    while (i < pc.len) {
      const uint64_t v = pc.start + i;
      while (sp < starts.size() && starts[sp] + counts[sp] <= v) ++sp;
      if (sp < starts.size() && v >= starts[sp]) { i = starts[sp] + counts[sp] - pc.start; continue; }  // inside a listed run
      const uint64_t lim = std::min(pc.len, sp < starts.size() ? starts[sp] - pc.start : pc.len);
      if (pc.data) w.raw(pc.data + i, lim - i);
      else w.v->insert(w.v->end(), lim - i, (uint8_t)0);
      i = lim;
    }
  }
}


inline void Serialize(const Image& im, std::vector<uint8_t>* short_blob, std::vector<uint8_t>* long_blob) {
  short_blob->clear(); long_blob->clear();
  ByteWriter w{short_blob};
  w.put<uint8_t>(im.first_prediction);
  w.put<int32_t>(im.ppm_top); w.put<int32_t>(im.ppm_mid); w.put<int32_t>(im.ppm_bot);
  for (auto& b : im.blist) { w.put<uint32_t>(b.stamp); w.put<uint32_t>(b.next); }
  w.put<uint32_t>(im.glue_count); w.put<uint32_t>(im.glue_count1); w.put<uint64_t>(im.sa_size);
  w.put<uint64_t>(im.p_text); w.put<uint64_t>(im.units_start); w.put<uint64_t>(im.lo_unit); w.put<uint64_t>(im.hi_unit);
  w.put<uint64_t>(im.aux_unit); w.put<uint64_t>(im.found_state); w.put<uint64_t>(im.max_context); w.put<uint64_t>(im.saved_pc);
  w.put<int32_t>(im.order_fall); w.put<uint32_t>(im.esc_count); w.arr(im.char_mask, 256);
  w.put<int32_t>(im.bsumm); w.put<int32_t>(im.run_length); w.put<int32_t>(im.init_rl); w.put<int32_t>(im.num_masked); w.put<int32_t>(im.prev_success);
  w.arr(&im.bin_summ[0][0], 25 * 64);
  for (int i = 0; i < 23; ++i) for (int j = 0; j < 32; ++j) { auto& s = im.see2[i][j]; w.put<uint8_t>(s.count); w.put<uint8_t>(s.shift); w.put<uint16_t>(s.summ); }
  w.put<uint8_t>(im.dummy_see2.count); w.put<uint8_t>(im.dummy_see2.shift); w.put<uint16_t>(im.dummy_see2.summ);
  for (auto& q : im.sq) { w.put<uint16_t>(q.freq); w.put<uint16_t>(q.sym); w.put<uint16_t>(q.total); }
  w.put<uint32_t>(im.sq_ptr);
  for (int i = 0; i < 256; ++i) { w.put<uint32_t>(im.sqp[i]); w.put<uint32_t>(im.trf[i]); w.put<uint32_t>(im.trt[i]); }
  w.put<uint32_t>(im.cxt); w.put<uint32_t>(im.y);
  SerializeHeap(im, w);
  w.put<int32_t>(im.l_top); w.put<int32_t>(im.l_mid); w.put<int32_t>(im.l_bot); w.arr(im.l_probs, 256);
  w.arr(im.l_input_history, L_HORIZON); w.arr(im.l_hidden, L_HID); w.arr(im.l_hidden_error, L_CELLS);
  w.arr(im.l_layer_input.data(), im.l_layer_input.size()); w.arr(im.l_output.data(), im.l_output.size());
  w.put<uint32_t>(im.l_epoch);
  w.arr(im.ll_state, L_CELLS); w.arr(im.ll_state_error, L_CELLS); w.arr(im.ll_stored_error, L_CELLS);
  w.arr(im.ll_tanh.data(), im.ll_tanh.size()); w.arr(im.ll_ig.data(), im.ll_ig.size()); w.arr(im.ll_last.data(), im.ll_last.size());
  w.put<uint32_t>(im.ll_epoch); w.put<uint64_t>(im.ll_update_steps);
  for (auto& n : im.nl) {
    w.arr(n.error, L_CELLS); w.arr(n.ivar, L_HORIZON);
    w.arr(n.gamma, L_CELLS); w.arr(n.gamma_u, L_CELLS); w.arr(n.gamma_m, L_CELLS); w.arr(n.gamma_v, L_CELLS);
    w.arr(n.beta, L_CELLS); w.arr(n.beta_u, L_CELLS); w.arr(n.beta_m, L_CELLS); w.arr(n.beta_v, L_CELLS);
    w.arr(n.state.data(), n.state.size()); w.arr(n.update.data(), n.update.size()); w.arr(n.m.data(), n.m.size());
    w.arr(n.v.data(), n.v.size()); w.arr(n.transpose.data(), n.transpose.size()); w.arr(n.norm.data(), n.norm.size());
  }
  for (auto& m : im.match) { w.put<uint64_t>(m.cur); w.put<uint8_t>(m.byte); w.put<uint8_t>(m.bitpos); w.put<uint8_t>(m.len); }
  for (int k = 0; k < NIH; ++k) {
    auto& h = im.ih[k];
    const uint32_t tsize = 1u << IHSpecs()[k].log2;
    const uint32_t n = (uint32_t)h.e.size();
    w.put<uint32_t>(n);
    if (n < tsize / 2) {
      for (auto& kv : h.e) { w.put<uint32_t>(kv.first); w.put<uint32_t>(kv.second); }
    } else {
      std::vector<uint32_t> t(tsize, 0);
      for (auto& kv : h.e) t[kv.first] = kv.second;
      w.arr(t.data(), tsize);
    }
    w.put<uint64_t>(h.outer); w.put<uint32_t>(h.hash);
  }
  for (auto& m : im.mixer) { w.put<uint64_t>(m.steps); w.put<uint64_t>(m.max_steps); w.put<uint64_t>(m.contexts_seen); }
  w.arr(im.predictions, NPRED);
  w.put<int32_t>(im.new_bit); w.put<int32_t>(im.recent_bits);
  w.put<uint32_t>(im.bit_context); w.put<uint32_t>(im.last_byte); w.put<uint32_t>(im.always_zero);
  w.put<uint32_t>(im.h3); w.put<uint32_t>(im.h4); w.put<uint32_t>(im.h5); w.put<uint32_t>(im.h6);
  w.arr(im.ih_ctx, NIH); w.arr(im.interval, 9); w.arr(im.skip, 15);
  w.put<uint32_t>(im.lbpr); w.put<uint32_t>(im.slpr);
  w.arr(im.l0_out, NL0); w.arr(im.l1_out, NL1); w.put<float>(im.final_out);
  w.put<uint32_t>(im.longest); w.put<uint64_t>(im.bits_seen); w.arr(im.entropy, kEntropy);
  w.put<uint32_t>(im.lstm_ctx); w.arr(im.rot, kRot); w.put<uint32_t>(im.rot_pos); w.arr(im.recent_bytes, 10);

  ByteWriter q{long_blob};
  for (int k = 0; k < NIND; ++k) {
    auto& d = im.ind[k];
    const uint32_t tsize = IndSize(k);
    uint32_t n = 0;
    for (auto& e : d.e) n += e.ns != 255;   // "seen" = nonstationary state != 255 (long-term-memory.cpp:10-14)
    q.put<uint32_t>(n);
    if (n < tsize / 3) {
      for (auto& e : d.e) if (e.ns != 255) { q.put<uint32_t>(e.key); q.put<uint8_t>(e.ns); q.put<uint8_t>(e.rm); }
    } else {
      std::vector<uint8_t> ns(tsize, 255), rm(tsize, 0);
      for (auto& e : d.e) { ns[e.key] = e.ns; rm[e.key] = e.rm; }
      q.arr(ns.data(), tsize); q.arr(rm.data(), tsize);
    }
    q.arr(d.ns_pred, 256); q.arr(d.rm_pred, 256);
  }
  for (int m = 0; m < NMIX; ++m) {
    auto& d = im.mix[m];
    q.put<uint32_t>((uint32_t)d.ctx.size()); q.put<uint32_t>(d.ctx.empty() ? 0u : d.input_size);
    for (size_t i = 0; i < d.ctx.size(); ++i) {
      q.put<uint32_t>(d.ctx[i]); q.put<uint64_t>(d.steps[i]); q.arr(d.w.data() + i * d.input_size, d.input_size);
    }
  }
  q.arr(im.wout.data(), im.wout.size()); q.arr(im.wgate.data(), im.wgate.size());
  q.put<uint64_t>((uint64_t)im.history.size()); q.arr(im.history.data(), im.history.size());
  for (int k = 0; k < NMATCH; ++k) {
    auto& d = im.mt[k];
    const uint32_t tsize = 1u << MatchSpecs()[k].log2;
    const uint32_t n = (uint32_t)d.e.size();
    q.put<uint32_t>(n);
    auto put5 = [&](uint64_t v) { uint8_t b[5] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24), (uint8_t)(v >> 32)}; q.arr(b, 5); };
    if (n < (5.0 / 9.0) * tsize) {
      for (auto& kv : d.e) { q.put<uint32_t>(kv.first); put5(kv.second); }
    } else {
      std::vector<uint8_t> t((size_t)tsize * 5, 0);
      for (auto& kv : d.e) for (int b = 0; b < 5; ++b) t[(size_t)kv.first * 5 + b] = (uint8_t)(kv.second >> (8 * b));
      q.arr(t.data(), t.size());
    }
    q.arr(d.pred, 256); q.arr(d.cnt, 256);
  }
}

// ---- sizing ---------------------------------------------------------------------------------------
// What a checkpoint already occupies, for MakeLayout(max_new_bytes, roomy, &preload).
inline Preload Count(const Image& im) {
  Preload p;
  for (int k = 0; k < NIND; ++k) if (IndSpecs()[k].log2 >= 15) p.sparse_entries += im.ind[k].e.size();
  for (int k = 0; k < NMATCH; ++k) if (MatchSpecs()[k].log2 >= 21) p.sparse_entries += im.mt[k].e.size();
  for (int k = 0; k < NIH; ++k) if (IHSpecs()[k].log2 >= 24) p.sparse_entries += im.ih[k].e.size();
  for (int m = 0; m < NMIX; ++m) p.mixer_sets += im.mix[m].ctx.size();
  p.ppmd_unit_bytes = im.heap_lo.size() + im.heap_hi.size();
  p.ppmd_lo_bytes = im.heap_lo.size(); p.ppmd_hi_bytes = im.heap_hi.size();
  p.ppmd_text_bytes = im.heap_text.size();
  p.history_bytes = im.history.size();
  p.steps = im.mixer[0].steps;
  return p;
}

// ---- host mirror of the device sparse map (stream_kernel.cuh SparseKey/SparseHash/SparseFind) ------------
inline uint32_t HostSparseHash(uint32_t k) { k ^= k >> 16; k *= 0x85ebca6bu; k ^= k >> 13; k *= 0xc2b2ae35u; k ^= k >> 16; return k; }
inline void HostSparseInsert(uint64_t* tab, uint32_t mask, uint32_t sid, uint32_t index, uint32_t value) {
  const uint32_t key = (sid << 25) | index;
  uint32_t pos = HostSparseHash(key) & mask;
  while (tab[pos] != 0 && (uint32_t)(tab[pos] >> 32) != key) pos = (pos + 1) & mask;
  tab[pos] = ((uint64_t)key << 32) | value;
}

inline uint32_t F2U(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float U2F(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

// ---- Image -> arena + state -----------------------------------------------------------------------
// `arena` (L.total bytes) and `st` are fully overwritten. st->T is filled with the model-graph tables
// and L exactly as StageTables does on the device.
inline bool ToArena(const Image& im, const ArenaLayout& L, uint8_t* arena, StreamSmem* st, std::string* err) {
  const IndirectSpec* ind = IndSpecs(); const IHSpec* ih = IHSpecs(); const MatchSpec* mt = MatchSpecs(); const MixerSpec* mx = MixerSpecs();
  if (!im.first_prediction && im.recent_bits * 2 + im.new_bit < 256) { *err = "checkpoint was not taken at a byte boundary"; return false; }
  memset(arena, 0, L.total);
  memset((void*)st, 0, sizeof(StreamSmem));
  StreamSmem& s = *st;
  s.T.L = L;
  { static const SkipSpec sk[20] = {GMX_SKIP_SPECS}; static const IntervalSpec iv[9] = {GMX_INTERVAL_SPECS};
    memcpy(s.T.ind, ind, sizeof(s.T.ind)); memcpy(s.T.skip, sk, sizeof(s.T.skip)); memcpy(s.T.interval, iv, sizeof(s.T.interval));
    memcpy(s.T.ih, ih, sizeof(s.T.ih)); memcpy(s.T.match, mt, sizeof(s.T.match)); memcpy(s.T.mixer, mx, sizeof(s.T.mixer));
    static const uint8_t ns[512] = {
#include "nonstationary.inc"
    };
    memcpy(s.T.nonstationary, ns, 512); }
  auto at = [&](uint64_t off) { return arena + off; };
  uint64_t* sparse = (uint64_t*)at(L.sparse);
  uint32_t sparse_used = 0;
  auto sparse_put = [&](uint32_t sid, uint32_t index, uint32_t value) -> bool {
    if (sparse_used >= L.sparse_limit) return false;
    ++sparse_used;
    HostSparseInsert(sparse, L.sparse_mask, sid, index, value);
    return true;
  };
  // Indirect (long-term-memory.h:11-25)
  for (int k = 0; k < NIND; ++k) {
    if (L.ind_sid[k]) {
      for (auto& e : im.ind[k].e) if (!sparse_put(L.ind_sid[k], e.key, (uint32_t)e.ns | ((uint32_t)e.rm << 8))) { *err = "sparse map too small for the checkpoint"; return false; }
    } else {
      uint16_t* t = (uint16_t*)at(L.ind_tab[k]);
      const uint64_t n = ((uint64_t)L.ind_size[k] + 1) / 2 * 2;
      for (uint64_t i = 0; i < n; ++i) t[i] = 0x00FF;
      for (auto& e : im.ind[k].e) t[e.key] = (uint16_t)(e.ns | (e.rm << 8));
    }
    float* pr = (float*)at(L.ind_pred) + k * 512;
    memcpy(pr, im.ind[k].ns_pred, 1024); memcpy(pr + 256, im.ind[k].rm_pred, 1024);
  }
  // Match
  for (int k = 0; k < NMATCH; ++k) {
    for (auto& kv : im.mt[k].e) {
      if (kv.second >> 32) { *err = "history pointer beyond 4 GiB"; return false; }
      if (L.match_sid[k]) { if (!sparse_put(L.match_sid[k], kv.first, (uint32_t)kv.second)) { *err = "sparse map too small for the checkpoint"; return false; } }
      else ((uint32_t*)at(L.match_tab[k]))[kv.first] = (uint32_t)kv.second;
    }
    memcpy((float*)at(L.match_pred) + k * 256, im.mt[k].pred, 1024);
    memcpy((int32_t*)at(L.match_cnt) + k * 256, im.mt[k].cnt, 1024);
    if (im.match[k].cur >> 32) { *err = "match cursor beyond 4 GiB"; return false; }
    s.m_cur[k] = (uint32_t)im.match[k].cur; s.m_byte[k] = im.match[k].byte; s.m_bitpos[k] = im.match[k].bitpos; s.m_len[k] = im.match[k].len;
  }
  if (im.history.size() > L.history_cap) { *err = "history larger than the arena's history area"; return false; }
  memcpy(at(L.history), im.history.data(), im.history.size());
  s.hist_len = (uint32_t)im.history.size();
  // IndirectHash
  for (int k = 0; k < NIH; ++k) {
    for (auto& kv : im.ih[k].e) {
      if (L.ih_sid[k]) { if (!sparse_put(L.ih_sid[k], kv.first, kv.second)) { *err = "sparse map too small for the checkpoint"; return false; } }
      else ((uint32_t*)at(L.ih_tab[k]))[kv.first] = kv.second;
    }
    s.ih_outer[k] = im.ih[k].outer; s.ih_hash[k] = im.ih[k].hash;
  }
  s.sparse_used = sparse_used;
  // Mixers: every weight set becomes a pool record {steps, 0, 0, 0 | weights}; nothing is staged in shared memory
  uint32_t next = 1;
  for (int m = 0; m < NMIX; ++m) {
    const auto& d = im.mix[m];
    uint32_t* dir = (uint32_t*)at(L.mix_dir[m]);
    for (size_t i = 0; i < d.ctx.size(); ++i) {
      if (next >= L.mix_pool_sets) { *err = "mixer weight-set pool too small for the checkpoint"; return false; }
      if (d.steps[i] >> 32) { *err = "mixer set step count beyond 2^32"; return false; }
      float* rec = (float*)at(L.mix_pool) + (size_t)next * L.mix_set_stride;
      rec[0] = U2F((uint32_t)d.steps[i]);
      memcpy(rec + 4, d.w.data() + i * d.input_size, (size_t)d.input_size * 4);
      dir[d.ctx[i]] = next++;
    }
    if (im.mixer[m].steps != im.mixer[0].steps || (im.mixer[m].steps >> 32) || (im.mixer[m].max_steps >> 32)) { *err = "mixer step counters out of range"; return false; }
    s.max_steps[m] = (uint32_t)im.mixer[m].max_steps;
    s.set_idx[m] = 0xFFFFFFFFu;
  }
  s.pool_next = next;
  s.steps = (uint32_t)im.mixer[0].steps;
  // LSTM (layouts: ArenaLayout comments in stream_kernel.cuh)
  {
    float* W = (float*)at(L.l_w); float* M = (float*)at(L.l_m); float* V = (float*)at(L.l_v);
    float* gb = (float*)at(L.l_gb);
    for (int g = 0; g < 3; ++g) {
      const auto& n = im.nl[g];
      for (int i = 0; i < L_CELLS; ++i)
        for (int j = 0; j < L_ROW; ++j) {
          const size_t dst = LstmW(g, j, i), src = (size_t)i * L_ROW + j;
          W[dst] = im.wgate[(size_t)g * L_CELLS * L_ROW + src]; M[dst] = n.m[src]; V[dst] = n.v[src];
        }
      const float* q8[8] = {n.gamma, n.beta, n.gamma_m, n.gamma_v, n.beta_m, n.beta_v, n.gamma_u, n.beta_u};
      for (int q = 0; q < 8; ++q) memcpy(gb + ((size_t)q * 3 + g) * L_CELLS, q8[q], L_CELLS * 4);
      memcpy((float*)at(L.l_gstate) + (size_t)g * L_HORIZON * L_CELLS, n.state.data(), (size_t)L_HORIZON * L_CELLS * 4);
      memcpy((float*)at(L.l_norm) + (size_t)g * L_HORIZON * L_CELLS, n.norm.data(), (size_t)L_HORIZON * L_CELLS * 4);
      memcpy((float*)at(L.l_ivar) + (size_t)g * L_HORIZON, n.ivar, L_HORIZON * 4);
    }
    float* wo = (float*)at(L.l_wout);
    for (int e = 0; e < L_HORIZON; ++e)
      for (int i = 0; i < L_NOUT; ++i)
        for (int j = 0; j < L_HID; ++j) wo[((size_t)e * L_HID + j) * L_NOUT + i] = im.wout[((size_t)e * L_NOUT + i) * L_HID + j];
    float* lin = (float*)at(L.l_lin);
    for (int e = 0; e < L_HORIZON; ++e) memcpy(lin + (size_t)e * (L_NIN + 1), im.l_layer_input.data() + (size_t)e * L_NIN, L_NIN * 4);
    memcpy(at(L.l_out), im.l_output.data(), (size_t)L_HORIZON * L_NOUT * 4);
    memcpy(at(L.l_tanh), im.ll_tanh.data(), (size_t)L_HORIZON * L_CELLS * 4);
    memcpy(at(L.l_ig), im.ll_ig.data(), (size_t)L_HORIZON * L_CELLS * 4);
    memcpy(at(L.l_last), im.ll_last.data(), (size_t)L_HORIZON * L_CELLS * 4);
    if (im.l_epoch >= L_HORIZON || im.ll_epoch != im.l_epoch || im.ll_update_steps > L_UPDATE_LIMIT) { *err = "LSTM counters out of range"; return false; }
    memcpy(s.l_hidden, im.l_hidden, L_HID * 4);
    memcpy(s.l_state, im.ll_state, L_CELLS * 4); memcpy(s.l_state_err, im.ll_state_error, L_CELLS * 4);
    memcpy(s.l_stored_err, im.ll_stored_error, L_CELLS * 4); memcpy(s.l_hidden_err, im.l_hidden_error, L_CELLS * 4);
    for (int e = 0; e < L_HORIZON; ++e) s.l_hist[e] = (uint8_t)im.l_input_history[e];
    s.l_epoch = im.l_epoch; s.l_update_steps = (uint32_t)im.ll_update_steps;
    memcpy(s.lprob, im.l_probs, 1024);
  }
  // PPMd
  {
    PpmdState* P = (PpmdState*)at(L.p_state);
    PpmdFillTables(P);
    for (int i = 0; i <= (int)PPMD_N_INDEXES; ++i) { P->bl_stamp[i] = im.blist[i].stamp; P->bl_next[i] = im.blist[i].next; }
    P->text_ptr = (uint32_t)im.p_text; P->units_start = (uint32_t)im.units_start; P->lo_unit = (uint32_t)im.lo_unit; P->hi_unit = (uint32_t)im.hi_unit;
    P->order_fall = im.order_fall; P->bsumm = im.bsumm; P->run_length = im.run_length; P->init_rl = im.init_rl;
    P->num_masked = im.num_masked; P->prev_success = im.prev_success;
    P->found_state = im.found_state < im.sa_size ? (uint32_t)im.found_state : 0u;   // a null FoundState is written as -HeapStart
    P->max_context = (uint32_t)im.max_context; P->esc_count = im.esc_count; P->error = 0;
    memcpy(P->char_mask, im.char_mask, sizeof(P->char_mask));
    for (int i = 0; i < 256; ++i) if (im.char_mask[i] == im.esc_count) s.p_masked[i >> 5] |= 1u << (i & 31);
    memcpy(P->bin_summ, im.bin_summ, sizeof(P->bin_summ));
    for (int i = 0; i < 23; ++i) for (int j = 0; j < 32; ++j) { P->see2[i][j].summ = im.see2[i][j].summ; P->see2[i][j].shift = im.see2[i][j].shift; P->see2[i][j].count = im.see2[i][j].count; }
    P->dummy_see2.summ = im.dummy_see2.summ; P->dummy_see2.shift = im.dummy_see2.shift; P->dummy_see2.count = im.dummy_see2.count;
    if (im.heap_text.size() > L.p_text_cap || im.heap_lo.size() + im.heap_hi.size() > L.p_units_cap) { *err = "PPMd heap window too small for the checkpoint"; return false; }
    uint8_t* heap = at(L.p_heap);
    for (size_t i = 0; i < im.heap_text.size(); ++i) heap[i & L.p_mask] = im.heap_text[i];
    for (size_t i = 0; i < im.heap_lo.size(); ++i) heap[(PPMD_UNITS_START + i) & L.p_mask] = im.heap_lo[i];
    for (size_t i = 0; i < im.heap_hi.size(); ++i) heap[((uint32_t)im.hi_unit + i) & L.p_mask] = im.heap_hi[i];
  }
  // ShortTermMemory
  memcpy(s.preds, im.predictions, NPRED * 4);
  for (int i = NPRED; i < NPRED + NL0 + 2; ++i) s.act[i] = 1;
  s.ctx[C_LAST_BYTE] = im.last_byte; s.ctx[C_BIT_CONTEXT] = im.bit_context;
  s.ctx[C_H3] = im.h3; s.ctx[C_H4] = im.h4; s.ctx[C_H5] = im.h5; s.ctx[C_H6] = im.h6;   // C_H2 is not serialised (recomputed each byte)
  s.ctx[C_LBPR] = im.lbpr; s.ctx[C_SLPR] = im.slpr;
  for (int i = 1; i < 10; ++i) s.ctx[C_RB1 + i - 1] = im.recent_bytes[i];
  s.ctx[C_LSTM] = im.lstm_ctx; s.ctx[C_LONGEST] = im.longest;
  for (int i = 0; i < 9; ++i) s.ctx[C_IV0 + i] = im.interval[i];
  for (int i = 0; i < 15; ++i) s.ctx[C_SK0 + kSkipFileToModel[i]] = im.skip[i];
  for (int i = 0; i < NIH; ++i) s.ctx[C_IH0 + i] = im.ih_ctx[i];
  memcpy(s.l0_out, im.l0_out, NL0 * 4); memcpy(s.l1_out, im.l1_out, NL1 * 4);
  s.final_out = im.final_out; s.prob = 0.5f;
  for (int i = 0; i < 256; ++i) s.ppm[i] = (float)(1.0 / 256);   // ppm_predictions is not serialised (short-term-memory.h ctor value)
  if (im.rot_pos >= kRot) { *err = "rotating history position out of range"; return false; }
  s.ring_pos = 0;
  for (int ago = 0; ago < 32; ++ago) s.ring[(0u - (uint32_t)ago) & 31u] = im.rot[(im.rot_pos + kRot - ago) % kRot];
  s.new_bit = im.new_bit; s.recent_bits = im.recent_bits; s.first_prediction = im.first_prediction;
  s.x1 = 0; s.x2 = 0xffffffffu;
  return true;
}

// ---- arena + state -> Image -----------------------------------------------------------------------
// `arena`/`st` are a stream parked at a byte boundary (after Learn of a byte's last bit). bits_seen is
// reconstructed as steps - 1 (every Predict of compress/decompress/train is followed by Learn).
inline bool FromArena(const ArenaLayout& L, const uint8_t* arena, const StreamSmem& s, Image* out, std::string* err) {
  Image& im = *out;
  const IndirectSpec* ind = IndSpecs(); const IHSpec* ih = IHSpecs(); const MatchSpec* mt = MatchSpecs(); const MixerSpec* mx = MixerSpecs();
  (void)ind; (void)mx;
  if (!s.first_prediction && s.recent_bits * 2 + s.new_bit < 256) { *err = "stream is not at a byte boundary"; return false; }
  auto at = [&](uint64_t off) { return arena + off; };
  // one pass over the sparse map distributes its entries to their tables
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> by_sid(64);
  if (L.sparse_mask) {
    const uint64_t* tab = (const uint64_t*)at(L.sparse);
    for (uint64_t i = 0; i <= L.sparse_mask; ++i) if (tab[i]) { const uint32_t key = (uint32_t)(tab[i] >> 32); by_sid[key >> 25].emplace_back(key & 0x1ffffffu, (uint32_t)tab[i]); }
    for (auto& v : by_sid) std::sort(v.begin(), v.end());
  }
  for (int k = 0; k < NIND; ++k) {
    auto& d = im.ind[k];
    d.e.clear();
    if (L.ind_sid[k]) { for (auto& kv : by_sid[L.ind_sid[k]]) d.e.push_back(IndEntry{kv.first, (uint8_t)kv.second, (uint8_t)(kv.second >> 8)}); }
    else {
      const uint16_t* t = (const uint16_t*)at(L.ind_tab[k]);
      for (uint32_t i = 0; i < L.ind_size[k]; ++i) if (t[i] != 0x00FF) d.e.push_back(IndEntry{i, (uint8_t)t[i], (uint8_t)(t[i] >> 8)});
    }
    const float* pr = (const float*)at(L.ind_pred) + k * 512;
    memcpy(d.ns_pred, pr, 1024); memcpy(d.rm_pred, pr + 256, 1024);
  }
  for (int k = 0; k < NMATCH; ++k) {
    auto& d = im.mt[k];
    d.e.clear();
    if (L.match_sid[k]) { for (auto& kv : by_sid[L.match_sid[k]]) if (kv.second) d.e.emplace_back(kv.first, (uint64_t)kv.second); }
    else { const uint32_t* t = (const uint32_t*)at(L.match_tab[k]); for (uint32_t i = 0; i < (1u << mt[k].log2); ++i) if (t[i]) d.e.emplace_back(i, (uint64_t)t[i]); }
    memcpy(d.pred, (const float*)at(L.match_pred) + k * 256, 1024);
    memcpy(d.cnt, (const int32_t*)at(L.match_cnt) + k * 256, 1024);
    im.match[k].cur = s.m_cur[k]; im.match[k].byte = s.m_byte[k]; im.match[k].bitpos = s.m_bitpos[k]; im.match[k].len = s.m_len[k];
  }
  im.history.assign(at(L.history), at(L.history) + s.hist_len);
  for (int k = 0; k < NIH; ++k) {
    auto& d = im.ih[k];
    d.e.clear();
    if (L.ih_sid[k]) { for (auto& kv : by_sid[L.ih_sid[k]]) if (kv.second) d.e.emplace_back(kv.first, kv.second); }
    else { const uint32_t* t = (const uint32_t*)at(L.ih_tab[k]); for (uint32_t i = 0; i < (1u << ih[k].log2); ++i) if (t[i]) d.e.emplace_back(i, t[i]); }
    d.outer = s.ih_outer[k]; d.hash = s.ih_hash[k];
  }
  for (int m = 0; m < NMIX; ++m) {
    auto& d = im.mix[m];
    const int nw = MixerWeights(m);
    d.input_size = nw; d.ctx.clear(); d.steps.clear(); d.w.clear();
    const uint32_t* dir = (const uint32_t*)at(L.mix_dir[m]);
    for (uint32_t c = 0; c < (1u << MixerSpecs()[m].log2); ++c) {
      const uint32_t id = dir[c];
      if (!id) continue;
      d.ctx.push_back(c);
      if (id == s.set_pool[m]) {   // the set staged in shared memory is newer than its pool record
        d.steps.push_back(s.set_steps[m]);
        const float* w = s.w + (m < NL0 ? m * WSTRIDE0 : NL0 * WSTRIDE0 + (m - NL0) * WSTRIDE1);
        d.w.insert(d.w.end(), w, w + nw);
      } else {
        const float* rec = (const float*)at(L.mix_pool) + (size_t)id * L.mix_set_stride;
        d.steps.push_back(F2U(rec[0]));
        d.w.insert(d.w.end(), rec + 4, rec + 4 + nw);
      }
    }
    im.mixer[m].steps = s.steps; im.mixer[m].max_steps = s.max_steps[m]; im.mixer[m].contexts_seen = d.ctx.size();
  }
  {
    const float* W = (const float*)at(L.l_w); const float* M = (const float*)at(L.l_m); const float* V = (const float*)at(L.l_v);
    const float* gb = (const float*)at(L.l_gb);
    const size_t hc = (size_t)L_HORIZON * L_CELLS, wsz = (size_t)L_CELLS * L_ROW;
    im.wgate.resize(3 * wsz);
    for (int g = 0; g < 3; ++g) {
      auto& n = im.nl[g];
      n.state.resize(hc); n.update.assign(wsz, 0.0f); n.m.resize(wsz); n.v.resize(wsz); n.transpose.resize((size_t)kTr * L_CELLS); n.norm.resize(hc);
      for (int i = 0; i < L_CELLS; ++i)
        for (int j = 0; j < L_ROW; ++j) {
          const size_t src = LstmW(g, j, i), dst = (size_t)i * L_ROW + j;
          im.wgate[g * wsz + dst] = W[src]; n.m[dst] = M[src]; n.v[dst] = V[src];
        }
      float* q8[8] = {n.gamma, n.beta, n.gamma_m, n.gamma_v, n.beta_m, n.beta_v, n.gamma_u, n.beta_u};
      for (int q = 0; q < 8; ++q) memcpy(q8[q], gb + ((size_t)q * 3 + g) * L_CELLS, L_CELLS * 4);
      memcpy(n.state.data(), (const float*)at(L.l_gstate) + g * hc, hc * 4);
      memcpy(n.norm.data(), (const float*)at(L.l_norm) + g * hc, hc * 4);
      memcpy(n.ivar, (const float*)at(L.l_ivar) + (size_t)g * L_HORIZON, L_HORIZON * 4);
      // scratch the reference rebuilds before reading: error_ (set per BPTT epoch), update_ (zeroed at the first
      // BPTT epoch), transpose_ (re-snapshot at the first BPTT epoch; the current weights are written here)
      memset(n.error, 0, sizeof(n.error));
      for (int j = 0; j < kTr; ++j) for (int i = 0; i < L_CELLS; ++i)
        n.transpose[(size_t)j * L_CELLS + i] = s.l_update_steps ? im.wgate[g * wsz + (size_t)i * L_ROW + 512 + j] : 0.0f;
    }
    const float* wo = (const float*)at(L.l_wout);
    im.wout.resize((size_t)L_HORIZON * L_NOUT * L_HID);
    for (int e = 0; e < L_HORIZON; ++e)
      for (int i = 0; i < L_NOUT; ++i)
        for (int j = 0; j < L_HID; ++j) im.wout[((size_t)e * L_NOUT + i) * L_HID + j] = wo[((size_t)e * L_HID + j) * L_NOUT + i];
    const float* lin = (const float*)at(L.l_lin);
    im.l_layer_input.resize((size_t)L_HORIZON * L_NIN);
    for (int e = 0; e < L_HORIZON; ++e) memcpy(im.l_layer_input.data() + (size_t)e * L_NIN, lin + (size_t)e * (L_NIN + 1), L_NIN * 4);
    im.l_output.resize((size_t)L_HORIZON * L_NOUT); memcpy(im.l_output.data(), at(L.l_out), im.l_output.size() * 4);
    im.ll_tanh.resize(hc); im.ll_ig.resize(hc); im.ll_last.resize(hc);
    memcpy(im.ll_tanh.data(), at(L.l_tanh), hc * 4); memcpy(im.ll_ig.data(), at(L.l_ig), hc * 4); memcpy(im.ll_last.data(), at(L.l_last), hc * 4);
    memcpy(im.l_hidden, s.l_hidden, L_HID * 4);
    memcpy(im.ll_state, s.l_state, L_CELLS * 4); memcpy(im.ll_state_error, s.l_state_err, L_CELLS * 4);
    memcpy(im.ll_stored_error, s.l_stored_err, L_CELLS * 4); memcpy(im.l_hidden_error, s.l_hidden_err, L_CELLS * 4);
    for (int e = 0; e < L_HORIZON; ++e) im.l_input_history[e] = s.l_hist[e];
    im.l_epoch = im.ll_epoch = s.l_epoch; im.ll_update_steps = s.l_update_steps;
    memcpy(im.l_probs, s.lprob, 1024);
  }
  {
    const PpmdState* P = (const PpmdState*)at(L.p_state);
    for (int i = 0; i <= (int)PPMD_N_INDEXES; ++i) { im.blist[i].stamp = P->bl_stamp[i]; im.blist[i].next = P->bl_next[i]; }
    im.glue_count = im.glue_count1 = 0; im.sa_size = PPMD_HEAP_END;
    im.p_text = P->text_ptr; im.units_start = P->units_start; im.lo_unit = P->lo_unit; im.hi_unit = P->hi_unit;
    im.aux_unit = 0;                       // only used by the out-of-memory path (the reference writes an uninitialised pointer here)
    im.found_state = P->found_state; im.max_context = P->max_context;
    im.saved_pc = 0;                       // RestoreModelRare scratch
    im.order_fall = P->order_fall; im.esc_count = P->esc_count; memcpy(im.char_mask, P->char_mask, sizeof(im.char_mask));
    im.bsumm = P->bsumm; im.run_length = P->run_length; im.init_rl = P->init_rl; im.num_masked = P->num_masked; im.prev_success = P->prev_success;
    memcpy(im.bin_summ, P->bin_summ, sizeof(im.bin_summ));
    for (int i = 0; i < 23; ++i) for (int j = 0; j < 32; ++j) { im.see2[i][j].summ = P->see2[i][j].summ; im.see2[i][j].shift = P->see2[i][j].shift; im.see2[i][j].count = P->see2[i][j].count; }
    im.dummy_see2.summ = P->dummy_see2.summ; im.dummy_see2.shift = P->dummy_see2.shift; im.dummy_see2.count = P->dummy_see2.count;
    // per-byte outputs of ppmd_PrepareByte, rebuilt from SQ_ptr = 0 at the next byte boundary before any read
    memset(im.sq, 0, sizeof(im.sq)); im.sq_ptr = 0; memset(im.sqp, 0, sizeof(im.sqp)); memset(im.trf, 0, sizeof(im.trf)); memset(im.trt, 0, sizeof(im.trt));
    im.cxt = 0; im.y = 1;
    const uint8_t* heap = at(L.p_heap);
    im.heap_text.resize(im.p_text); im.heap_lo.resize(im.lo_unit - im.units_start); im.heap_hi.resize(im.sa_size - im.hi_unit);
    for (size_t i = 0; i < im.heap_text.size(); ++i) im.heap_text[i] = heap[i & L.p_mask];
    for (size_t i = 0; i < im.heap_lo.size(); ++i) im.heap_lo[i] = heap[(PPMD_UNITS_START + i) & L.p_mask];
    for (size_t i = 0; i < im.heap_hi.size(); ++i) im.heap_hi[i] = heap[((uint32_t)im.hi_unit + i) & L.p_mask];
  }
  im.first_prediction = (uint8_t)s.first_prediction;
  // interval-search state of both byte models after the 8th Predict of a byte: the 2-wide interval of the last bit
  if (s.first_prediction) { im.ppm_top = im.l_top = 255; im.ppm_mid = im.l_mid = 127; im.ppm_bot = im.l_bot = 0; }
  else {
    const int lb = (s.recent_bits * 2) & 0xFE;
    im.ppm_bot = im.l_bot = lb; im.ppm_mid = im.l_mid = lb; im.ppm_top = im.l_top = lb | 1;
  }
  memcpy(im.predictions, s.preds, NPRED * 4);
  im.new_bit = s.new_bit; im.recent_bits = s.recent_bits;
  im.bit_context = s.ctx[C_BIT_CONTEXT]; im.last_byte = s.ctx[C_LAST_BYTE]; im.always_zero = 0;
  im.h3 = s.ctx[C_H3]; im.h4 = s.ctx[C_H4]; im.h5 = s.ctx[C_H5]; im.h6 = s.ctx[C_H6];
  for (int i = 0; i < NIH; ++i) im.ih_ctx[i] = s.ctx[C_IH0 + i];
  for (int i = 0; i < 9; ++i) im.interval[i] = s.ctx[C_IV0 + i];
  for (int i = 0; i < 15; ++i) im.skip[i] = s.ctx[C_SK0 + kSkipFileToModel[i]];
  im.lbpr = s.ctx[C_LBPR]; im.slpr = s.ctx[C_SLPR];
  memcpy(im.l0_out, s.l0_out, NL0 * 4); memcpy(im.l1_out, s.l1_out, NL1 * 4); im.final_out = s.final_out;
  im.longest = s.ctx[C_LONGEST];
  im.bits_seen = s.first_prediction ? 0 : (uint64_t)s.steps - 1;
  for (auto& e : im.entropy) e = -1;       // analysis accumulators (predictor.cpp:37-38 initial value)
  im.lstm_ctx = s.ctx[C_LSTM];
  // the reference keeps the last 1000 bytes; the device keeps 32 (only 10 are ever read): older ring slots are 0
  memset(im.rot, 0, sizeof(im.rot));
  const uint64_t nbytes = s.first_prediction ? 0 : s.steps / 8;   // ByteUpdate ran nbytes - 1 times before this boundary
  const uint64_t updates = nbytes ? nbytes - 1 : 0;
  im.rot_pos = (uint32_t)(updates % kRot);
  for (uint32_t ago = 0; ago < 32 && ago < updates; ++ago) im.rot[(im.rot_pos + kRot - ago) % kRot] = s.ring[(s.ring_pos - ago) & 31u];
  im.recent_bytes[0] = s.ring[s.ring_pos & 31u];
  for (int i = 1; i < 10; ++i) im.recent_bytes[i] = s.ctx[C_RB1 + i - 1];
  return true;
}

}  // namespace ckpt
}  // namespace gmx
#endif
