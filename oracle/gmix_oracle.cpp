// TEST INFRASTRUCTURE ONLY — see gmix_oracle.h. Plain single-threaded C++ restatement of the
// reference's per-bit path, written from the behaviour described in SURVEY.md section 3.6 and
// the reference sources cited per function (paths relative to /root/reference/src).
// Build: g++ -std=c++17 -O2 -ffp-contract=off (strict IEEE, no contraction), host glibc libm.
#include "gmix_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "oracle_ppmd.h"
#include "oracle_lstm.h"

namespace {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

// ---- mixer/sigmoid.cpp:5-13 -------------------------------------------------------------------
inline float Logistic(float x) { return 1 / (1 + expf(-x)); }
inline float Logit(float p) {
  if (p < 0.0001) p = 0.0001;          // compare in double, assign rounded to float
  else if (p > 0.9999) p = 0.9999;
  return logf(p / (1 - p));
}

// ---- contexts/murmur-hash.cpp:94-146 (MurmurHash3_x86_32, public algorithm by A. Appleby) -------
inline u32 rotl(u32 x, int r) { return (x << r) | (x >> (32 - r)); }
u32 Murmur32(const u8* key, int len, u32 seed) {
  u32 h = seed;
  int nb = len / 4;
  for (int i = 0; i < nb; ++i) {
    u32 k;
    memcpy(&k, key + 4 * i, 4);
    k *= 0xcc9e2d51u; k = rotl(k, 15); k *= 0x1b873593u;
    h ^= k; h = rotl(h, 13); h = h * 5 + 0xe6546b64u;
  }
  const u8* tail = key + 4 * nb;
  u32 k = 0;
  switch (len & 3) {
    case 3: k ^= (u32)tail[2] << 16; /* fallthrough */
    case 2: k ^= (u32)tail[1] << 8;  /* fallthrough */
    case 1: k ^= tail[0]; k *= 0xcc9e2d51u; k = rotl(k, 15); k *= 0x1b873593u; h ^= k;
  }
  h ^= (u32)len;
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  return h;
}
inline u32 Murmur64(u64 v) { return Murmur32((const u8*)&v, 8, 0xDEADBEEFu); }
inline u32 MurmurU32(u32 v) { return Murmur32((const u8*)&v, 4, 0xDEADBEEFu); }

// ---- contexts/nonstationary.cpp:3-58, contexts/run-map.cpp:3-21 ----------------------------------
const u8 kNonstationary[512] = {
#include "nonstationary.inc"
};
u8 kRunMap[512];
void InitRunMap() {
  for (int i = 0; i < 512; ++i) {
    int s = i / 2;
    if (i % 2 == 0) { if (s < 127) ++s; else if (s >= 128) s = 1; }
    else { if (s < 128) s = 128; else if (s < 255) ++s; }
    kRunMap[i] = (u8)s;
  }
}

// ---- model graph constants (predictor.cpp:54-358) -------------------------
enum CtxId {
  C_ZERO, C_LAST_BYTE, C_BIT_CONTEXT, C_H2, C_H3, C_H4, C_H5, C_H6, C_LBPR, C_SLPR,
  C_RB1, C_RB2, C_RB3, C_RB4, C_RB5, C_RB6, C_RB7, C_RB8, C_RB9, C_LSTM, C_LONGEST,
  C_IV0, /* 9 interval contexts */ C_SK0 = C_IV0 + 9, /* 15 skip contexts */
  C_IH0 = C_SK0 + 15, /* 9 indirect-hash contexts */ C_COUNT = C_IH0 + 9
};

struct IndirectSpec { int ctx; int log2_size; float lr; };
struct SkipSpec { int n; int bytes[6]; };
const SkipSpec kHashSkips[5] = {  // last 2..6 bytes (predictor.cpp:84-107)
    {2, {0, 1}}, {3, {0, 1, 2}}, {4, {0, 1, 2, 3}}, {5, {0, 1, 2, 3, 4}}, {6, {0, 1, 2, 3, 4, 5}}};
const SkipSpec kSkips[15] = {  // predictor.cpp:122-185
    {2, {1, 2}}, {3, {1, 2, 3}}, {2, {0, 2}}, {3, {0, 2, 3}}, {4, {1, 2, 3, 4}},
    {2, {0, 3}}, {2, {0, 4}}, {2, {0, 5}}, {4, {0, 2, 3, 4}}, {3, {0, 3, 4}},
    {2, {0, 6}}, {2, {0, 7}}, {4, {0, 1, 3, 4}}, {3, {0, 4, 5}}, {4, {0, 1, 2, 4}}};
// skip context order in ShortTermMemory naming: skip_1_2, skip_1_2_3, skip_0_2, skip_0_2_3,
// skip_1_2_3_4, skip_0_3, skip_0_4, skip_0_5, skip_0_2_3_4, skip_0_3_4, skip_0_6, skip_0_7,
// skip_0_1_3_4, skip_0_4_5, skip_0_1_2_4  -> C_SK0 + index above; skip_0_2 = C_SK0 + 2.
struct IntervalSpec { int div; int shift; u32 mask; };
const IntervalSpec kIntervals[9] = {  // predictor.cpp:54-76, interval-context.cpp:3-15
    {16, 4, 0xF}, {16, 4, 0xFF}, {16, 4, 0xFFF}, {32, 3, 0x7}, {32, 3, 0x3F}, {32, 3, 0xFFF},
    {64, 2, 0xF}, {64, 2, 0xFF}, {64, 2, 0xFFF}};
struct IHSpec { int outer_order; int log2_size; int inner_order; };
const IHSpec kIH[9] = {{1, 8, 1}, {1, 8, 2}, {1, 8, 3}, {2, 16, 1}, {2, 16, 2},
                       {2, 16, 3}, {3, 24, 1}, {4, 24, 2}, {4, 24, 3}};  // predictor.cpp:210-249
const int kIHIndirectLog2[9] = {8, 16, 15, 8, 16, 15, 8, 16, 15};
struct MatchSpec { int ctx; int log2_size; };
const MatchSpec kMatch[6] = {{C_LAST_BYTE, 8}, {C_H2, 16}, {C_H3, 24}, {C_H4, 21}, {C_H5, 21}, {C_H6, 21}};
struct MixerSpec { int ctx; double lr; int log2_size; };
const MixerSpec kMixL0[24] = {  // predictor.cpp:254-325
    {C_LAST_BYTE, 0.005, 8}, {C_RB3, 0.0055, 8}, {C_SLPR, 0.003, 16}, {C_H4, 0.0045, 15},
    {C_IH0 + 6, 0.006, 8}, {C_RB1, 0.004, 8}, {C_LONGEST, 0.0005, 3}, {C_H2, 0.0035, 16},
    {C_RB2, 0.0065, 8}, {C_H3, 0.0025, 15}, {C_LAST_BYTE, 0.001, 8}, {C_LBPR, 0.002, 16},
    {C_IV0 + 0, 0.005, 4}, {C_IV0 + 1, 0.0045, 8}, {C_IV0 + 2, 0.0055, 12}, {C_IV0 + 3, 0.004, 3},
    {C_IV0 + 4, 0.0035, 6}, {C_SK0 + 2, 0.006, 16}, {C_IV0 + 5, 0.003, 12}, {C_IV0 + 6, 0.0065, 4},
    {C_IV0 + 7, 0.003, 8}, {C_IV0 + 8, 0.0025, 12}, {C_LSTM, 0.002, 8}, {C_ZERO, 0.0005, 0}};
const MixerSpec kMixL1[8] = {  // predictor.cpp:328-351
    {C_RB1, 0.0045, 8}, {C_ZERO, 0.0035, 0}, {C_BIT_CONTEXT, 0.003, 8}, {C_RB2, 0.002, 8},
    {C_LAST_BYTE, 0.0025, 8}, {C_BIT_CONTEXT, 0.00001, 8}, {C_LONGEST, 0.0008, 3}, {C_ZERO, 0.0004, 0}};
const MixerSpec kMixFinal = {C_ZERO, 0.0005, 0};  // predictor.cpp:355-357

const int NPRED = 90, NL0 = 24, NL1 = 8;

struct IndirectMem {  // long-term-memory.h:11-25, indirect.cpp:15-25
  int ctx; float lr; u32 size;
  std::vector<u8> ns, rm;
  float ns_pred[256], rm_pred[256];
};
struct MixerSet { u64 steps; std::vector<float> w; };
struct MixerMem {  // mixer.h, long-term-memory.h:27-41
  int ctx; float lr; int layer; int out_index; int nweights; u32 table_size;
  u64 steps, max_steps, contexts_seen;
  std::vector<MixerSet*> table;
};
struct MatchMem {  // match.h, long-term-memory.h:43-55
  int ctx; u32 size;
  std::vector<u8> table;  // size*5
  float pred[256]; int counts[256];
  u64 cur_match; u8 cur_byte, bit_pos, match_length;
};
struct IndirectHashMem {  // indirect-hash.h
  std::vector<u32> table; u64 outer_ctx, outer_mod, inner_mod; u32 outer_hash;
};

}  // namespace

struct gmo_predictor {
  // ---- ShortTermMemory (memory/short-term-memory.h:19-174) ----
  float predictions[NPRED];
  u32 active_mask[3];  // active_models is always in ascending index order (model order == index order)
  int new_bit = 0, recent_bits = 1;
  u32 ctx[C_COUNT];
  float ppm_predictions[256];
  float l0_out[NL0], l1_out[NL1], final_out = 0;
  u64 bits_seen = 0;
  u8 ring[1000]; u32 ring_pos = 0;
  bool first_prediction = true;  // basic-contexts.h
  bool analysis = false;
  // ---- LongTermMemory + model-private state ----
  std::vector<IndirectMem> indirect;
  std::vector<MixerMem> mixers;
  std::vector<MatchMem> match;
  IndirectHashMem ih[9];
  std::vector<u8> history;
  oracle_ppmd::Model ppmd; int ppmd_top = 255, ppmd_mid = 127, ppmd_bot = 0;
  oracle_lstm::Lstm lstm; int lstm_top = 255, lstm_mid = 127, lstm_bot = 0; float lstm_probs[256];

  gmo_predictor();
  ~gmo_predictor();
  void AddIndirect(int c, int log2, float lr) {
    indirect.emplace_back();
    IndirectMem& m = indirect.back();
    m.ctx = c; m.lr = lr; m.size = (1u << log2) * 256 + 1;
    m.ns.assign(m.size, 255); m.rm.assign(m.size, 0);
    for (int i = 0; i < 256; ++i) m.ns_pred[i] = m.rm_pred[i] = 0;
  }
  void AddMixer(const MixerSpec& s, int layer, int idx) {
    mixers.emplace_back();
    MixerMem& m = mixers.back();
    m.ctx = s.ctx; m.lr = (float)s.lr; m.layer = layer; m.out_index = idx;
    m.table_size = 1u << s.log2_size; m.table.assign(m.table_size, nullptr);
    m.steps = 0; m.max_steps = 1; m.contexts_seen = 0;
    // mixer.cpp:17-26 (one model with a skip connection: the LSTM, lstm-model.cpp:14)
    m.nweights = layer == 0 ? NPRED + idx : layer == 1 ? NL0 + idx + 1 : NL0 + NL1 + 1;
  }
  u32 RecentByte(int ago) const {  // short-term-memory.cpp:215-219
    int pos = (int)ring_pos - ago; if (pos < 0) pos += 1000; return ring[pos];
  }
  void SetPrediction(float p, int i) {  // short-term-memory.cpp:187-191
    predictions[i] = Logit(p);
    if (p == 0.5) return;
    active_mask[i >> 5] |= 1u << (i & 31);
  }
  void SetLogit(float p, int i) {  // short-term-memory.cpp:193-197
    predictions[i] = p;
    if (p == 0) return;
    active_mask[i >> 5] |= 1u << (i & 31);
  }
  bool Active(int i) const { return (active_mask[i >> 5] >> (i & 31)) & 1; }
  void BitIntervalPrediction(const float* probs, int& top, int& mid, int& bot, bool byte_boundary, int idx);
  float Predict();
  void Learn();
};

gmo_predictor::gmo_predictor() {
  static bool once = false;
  if (!once) { InitRunMap(); once = true; }
  srand(0xDEADBEEF);  // predictor.cpp:18
  memset(predictions, 0, sizeof(predictions));
  memset(active_mask, 0, sizeof(active_mask));
  memset(ctx, 0, sizeof(ctx));
  memset(ring, 0, sizeof(ring));
  memset(l0_out, 0, sizeof(l0_out));
  memset(l1_out, 0, sizeof(l1_out));
  for (int i = 0; i < 256; ++i) ppm_predictions[i] = (float)(1.0 / 256);
  for (int i = 0; i < 256; ++i) lstm_probs[i] = (float)(1.0 / 256);
  ppmd.Init();
  lstm.Init();  // consumes rand() exactly like lstm-layer.cpp:176-195
  // predictor.cpp:78-120
  const float lr = 0.02;
  AddIndirect(C_LAST_BYTE, 8, lr);
  AddIndirect(C_H2, 16, lr);
  AddIndirect(C_H3, 15, lr);
  AddIndirect(C_H3, 16, lr);
  AddIndirect(C_H4, 15, lr);
  AddIndirect(C_H5, 15, lr);
  AddIndirect(C_H6, 15, lr);
  for (int i = 1; i < 10; ++i) AddIndirect(C_RB1 + i - 1, 8, lr);
  AddIndirect(C_LSTM, 8, lr);
  for (int i = 0; i < 15; ++i) AddIndirect(C_SK0 + i, 16, lr);  // predictor.cpp:122-185
  // predictor.cpp:187-208
  for (int i = 0; i < 6; ++i) {
    match.emplace_back();
    MatchMem& m = match.back();
    m.ctx = kMatch[i].ctx; m.size = 1u << kMatch[i].log2_size;
    m.table.assign((size_t)m.size * 5, 0);
    for (int k = 0; k < 256; ++k) { m.pred[k] = 0.5 + (k + 0.5) / 512; m.counts[k] = 1; }  // match.cpp:19-22
    m.cur_match = 0; m.cur_byte = 0; m.bit_pos = 128; m.match_length = 0;
  }
  // predictor.cpp:210-249
  for (int i = 0; i < 9; ++i) {
    ih[i].table.assign(1u << kIH[i].log2_size, 0);
    ih[i].outer_ctx = 0; ih[i].outer_hash = 0;
    ih[i].outer_mod = 1ull << (8 * (kIH[i].outer_order - 1));
    ih[i].inner_mod = 1ull << (8 * (kIH[i].inner_order - 1));
    AddIndirect(C_IH0 + i, kIHIndirectLog2[i], 1.0f / 200);
  }
  for (int i = 0; i < NL0; ++i) AddMixer(kMixL0[i], 0, i);
  for (int i = 0; i < NL1; ++i) AddMixer(kMixL1[i], 1, i);
  AddMixer(kMixFinal, 2, 0);
}

gmo_predictor::~gmo_predictor() {
  for (auto& m : mixers) for (auto* s : m.table) delete s;
}

// mod_ppmd.cpp:1662-1681 and lstm-model.cpp:23-47: binary-search interval over the byte distribution.
void gmo_predictor::BitIntervalPrediction(const float* probs, int& top, int& mid, int& bot,
                                          bool byte_boundary, int idx) {
  if (byte_boundary) { top = 255; bot = 0; }
  else if (new_bit) bot = mid + 1;
  else top = mid;
  mid = bot + ((top - bot) / 2);
  float num = 0.0f;
  for (int i = mid + 1; i <= top; ++i) num = num + probs[i];
  float denom = num;
  for (int i = bot; i <= mid; ++i) denom = denom + probs[i];
  if (denom != 0) SetPrediction(num / denom, idx);
}

float gmo_predictor::Predict() {  // predictor.cpp:360-376
  memset(active_mask, 0, sizeof(active_mask));
  if (analysis) memset(predictions, 0, sizeof(predictions));

  // BasicContexts::Predict, basic-contexts.cpp:21-40
  if (first_prediction) {
    first_prediction = false;
  } else {
    ++bits_seen;
    recent_bits += recent_bits + new_bit;
    if (recent_bits >= 256) {  // ByteUpdate, basic-contexts.cpp:5-19
      ctx[C_LAST_BYTE] = recent_bits - 256;
      if (++ring_pos == 1000) ring_pos = 0;
      ring[ring_pos] = (u8)ctx[C_LAST_BYTE];
      // recent_bytes[0] == last byte; recent_bytes[1..9] are contexts C_RB1..C_RB9
      for (int i = 1; i < 10; ++i) ctx[C_RB1 + i - 1] = RecentByte(i);
      recent_bits = 1;
    }
    ctx[C_BIT_CONTEXT] = recent_bits - 1;
    ctx[C_LBPR] = (ctx[C_LAST_BYTE] << 8) + ctx[C_BIT_CONTEXT];
    ctx[C_SLPR] = (ctx[C_RB1] << 8) + ctx[C_BIT_CONTEXT];
    ctx[C_LONGEST] = 0;
  }
  const bool bb = recent_bits == 1;  // "byte boundary" (true on the very first call too)
  const u32 last_byte = ctx[C_LAST_BYTE];
  const u32 bitctx = ctx[C_BIT_CONTEXT];

  // IntervalContext::Predict, interval-context.cpp:17-23
  if (bb) for (int i = 0; i < 9; ++i) {
    const IntervalSpec& s = kIntervals[i];
    ctx[C_IV0 + i] = s.mask & ((ctx[C_IV0 + i] << s.shift) + last_byte / s.div);
  }

  // ModPPMD::Predict, mod_ppmd.cpp:1649-1682
  if (bb) {
    ppmd.UpdateByte(last_byte);
    ppmd.PrepareByte();
    for (int i = 0; i < 256; ++i) {
      ppm_predictions[i] = (float)ppmd.sqp[i];
      if (ppm_predictions[i] < 1) ppm_predictions[i] = 1;
    }
    float sum = ppm_predictions[0];  // valarray::sum(): ascending, seeded with element 0
    for (int i = 1; i < 256; ++i) sum += ppm_predictions[i];
    for (int i = 0; i < 256; ++i) ppm_predictions[i] = ppm_predictions[i] / sum;
  }
  BitIntervalPrediction(ppm_predictions, ppmd_top, ppmd_mid, ppmd_bot, bb, 0);

  // LstmModel::Predict, lstm-model.cpp:17-48
  if (bb) {
    lstm.SetInput(ppm_predictions);
    const float* probs = lstm.Predict(last_byte);
    memcpy(lstm_probs, probs, sizeof(lstm_probs));
    float max_pred = 0;
    ctx[C_LSTM] = 0;
    for (int i = 0; i < 256; ++i) if (lstm_probs[i] > max_pred) { max_pred = lstm_probs[i]; ctx[C_LSTM] = i; }
  }
  BitIntervalPrediction(lstm_probs, lstm_top, lstm_mid, lstm_bot, bb, 1);

  // SkipContext::Predict, skip-context.cpp:9-19 (all skip contexts depend only on the ring)
  if (bb) {
    for (int i = 0; i < 5; ++i) {
      u64 c = 0;
      for (int k = 0; k < kHashSkips[i].n; ++k) c = (c << 8) + RecentByte(kHashSkips[i].bytes[k]);
      ctx[C_H2 + i] = Murmur64(c);
    }
    for (int i = 0; i < 15; ++i) {
      u64 c = 0;
      for (int k = 0; k < kSkips[i].n; ++k) c = (c << 8) + RecentByte(kSkips[i].bytes[k]);
      ctx[C_SK0 + i] = Murmur64(c);
    }
    // IndirectHash::Predict, indirect-hash.cpp:16-31
    for (int i = 0; i < 9; ++i) {
      IndirectHashMem& h = ih[i];
      u32& inner = h.table[h.outer_hash % h.table.size()];
      inner = (u32)(((inner % h.inner_mod) << 8) + last_byte);
      h.outer_ctx = ((h.outer_ctx % h.outer_mod) << 8) + last_byte;
      h.outer_hash = Murmur64(h.outer_ctx);
      ctx[C_IH0 + i] = MurmurU32(h.table[h.outer_hash % h.table.size()]);
    }
  }

  // Indirect::Predict, indirect.cpp:28-45. Prediction indices: indirect k -> 2+2k, 3+2k for the
  // first 32 (AddIndirect + AddSkip), 72+2(k-32) for the 9 double-indirect ones (after the 6 Match).
  auto indirect_predict = [&](int k, int pidx) {
    IndirectMem& m = indirect[k];
    u32 slot = ((ctx[m.ctx] << 8) + bitctx) % m.size;
    int s = m.ns[slot];
    if (s != 255) SetLogit(m.ns_pred[s], pidx);
    int r = m.rm[slot];
    if (r != 0) SetLogit(m.rm_pred[r], pidx + 1);
  };
  for (int k = 0; k < 32; ++k) indirect_predict(k, 2 + 2 * k);

  // Match::Predict, match.cpp:25-74
  for (int k = 0; k < 6; ++k) {
    MatchMem& m = match[k];
    int hit = new_bit == ((m.cur_byte & m.bit_pos) != 0);
    if (hit) { if (m.match_length < 255) ++m.match_length; } else m.match_length = 0;
    m.bit_pos /= 2;
    if (bb) {
      if (m.cur_match == (u64)history.size() - 1) m.match_length = 0;
      if (m.match_length < 8) {
        const u8* it = &m.table[(size_t)(ctx[m.ctx] % m.size) * 5];
        m.cur_match = it[0] + (1 << 8) * it[1] + (1 << 16) * it[2] + ((u64)it[3] << 24) + ((u64)it[4] << 32);
      } else {
        ++m.cur_match;
      }
      if (!history.empty()) {
        if (m.cur_match >= history.size()) { fprintf(stderr, "oracle: match pointer out of range\n"); abort(); }
        m.cur_byte = history[m.cur_match];
      }
      m.bit_pos = 128;
    }
    if (m.match_length > 2) {
      float p = (m.cur_byte & m.bit_pos) ? m.pred[m.match_length] : 1 - m.pred[m.match_length];
      SetPrediction(p, 66 + k);
    }
    u32 mc = m.match_length / 32;
    if (mc > ctx[C_LONGEST]) ctx[C_LONGEST] = mc;
  }
  for (int k = 32; k < 41; ++k) indirect_predict(k, 72 + 2 * (k - 32));

  // Mixer::Predict, mixer.cpp:51-106
  for (MixerMem& m : mixers) {
    MixerSet* d = m.table[ctx[m.ctx] % m.table_size];
    float p = 0;
    if (d) {
      const float* w = d->w.data();
      if (m.layer == 0) {
        for (int i = 0; i < NPRED; ++i) if (Active(i)) p += predictions[i] * w[i];
        for (int i = 0; i < m.out_index; ++i) p += l0_out[i] * w[NPRED + i];
      } else if (m.layer == 1) {
        for (int i = 0; i < NL0; ++i) p += l0_out[i] * w[i];
        for (int i = 0; i < m.out_index; ++i) p += l1_out[i] * w[NL0 + i];
        p += predictions[1] * w[NL0 + m.out_index];  // skip connection: LSTM prediction
      } else {
        for (int i = 0; i < NL0; ++i) p += l0_out[i] * w[i];
        for (int i = 0; i < NL1; ++i) p += l1_out[i] * w[NL0 + i];
        p += predictions[1] * w[NL0 + NL1];
      }
    }
    if (m.layer == 2) final_out = p; else if (m.layer == 1) l1_out[m.out_index] = p; else l0_out[m.out_index] = p;
  }

  float prob = Logistic(final_out);
  float eps = 0.0001;
  if (prob < eps) prob = eps; else if (prob > 1 - eps) prob = 1 - eps;
  return prob;
}

void gmo_predictor::Learn() {  // predictor.cpp:383-387
  const int cur = recent_bits * 2 + new_bit;
  // BasicContexts::Learn, basic-contexts.cpp:42-54
  if (cur >= 256 && ctx[C_LONGEST] < 2) history.push_back((u8)cur);
  // LstmModel::Learn, lstm-model.cpp:50-59
  if (cur >= 256) lstm.Perceive(cur - 256);
  // Indirect::Learn, indirect.cpp:47-70
  const u32 bitctx = ctx[C_BIT_CONTEXT];
  for (IndirectMem& m : indirect) {
    u32 slot = ((ctx[m.ctx] << 8) + bitctx) % m.size;
    int s = m.ns[slot];
    if (s == 255) s = 0;
    m.ns_pred[s] += (new_bit - Logistic(m.ns_pred[s])) * m.lr;
    m.ns[slot] = kNonstationary[s * 2 + new_bit];
    int r = m.rm[slot];
    m.rm_pred[r] += (new_bit - Logistic(m.rm_pred[r])) * m.lr;
    m.rm[slot] = kRunMap[r * 2 + new_bit];
  }
  // Match::Learn, match.cpp:76-109
  for (MatchMem& m : match) {
    if (m.match_length > 2) {
      int hit = new_bit == ((m.cur_byte & m.bit_pos) != 0);
      float rate = (float)(1.0 / 400);
      if (m.counts[m.match_length] < 400) { ++m.counts[m.match_length]; rate = 1.0 / m.counts[m.match_length]; }
      m.pred[m.match_length] += (hit - m.pred[m.match_length]) * rate;
    }
    if (recent_bits >= 128) {
      if (ctx[C_LONGEST] >= 2) continue;
      u8* loc = &m.table[(size_t)(ctx[m.ctx] % m.size) * 5];
      u64 pos = (u64)history.size() - 1;
      loc[0] = pos; loc[1] = pos >> 8; loc[2] = pos >> 16; loc[3] = pos >> 24; loc[4] = pos >> 32;
    }
  }
  // Mixer::Learn, mixer.cpp:108-176
  for (MixerMem& m : mixers) {
    MixerSet*& slot = m.table[ctx[m.ctx] % m.table_size];
    if (!slot) { ++m.contexts_seen; slot = new MixerSet{0, std::vector<float>(m.nweights, 0.0f)}; }
    MixerSet* d = slot;
    float decay = 0.9 / pow(0.0000001 * m.steps + 0.8, 0.8);
    decay *= 1.5 - ((1.0 * d->steps) / m.max_steps);
    float out = m.layer == 2 ? final_out : m.layer == 1 ? l1_out[m.out_index] : l0_out[m.out_index];
    float p = Logistic(out);
    float update = decay * m.lr * (p - new_bit);
    ++m.steps; ++d->steps;
    if (d->steps > m.max_steps) m.max_steps = d->steps;
    float* w = d->w.data();
    if (m.layer == 0) {
      for (int i = 0; i < NPRED; ++i) if (Active(i)) w[i] -= update * predictions[i];
      for (int i = 0; i < m.out_index; ++i) w[NPRED + i] -= update * l0_out[i];
    } else if (m.layer == 1) {
      for (int i = 0; i < NL0; ++i) w[i] -= update * l0_out[i];
      for (int i = 0; i < m.out_index; ++i) w[NL0 + i] -= update * l1_out[i];
      w[NL0 + m.out_index] -= update * predictions[1];
    } else {
      for (int i = 0; i < NL0; ++i) w[i] -= update * l0_out[i];
      for (int i = 0; i < NL1; ++i) w[NL0 + i] -= update * l1_out[i];
      w[NL0 + NL1] -= update * predictions[1];
    }
    if ((d->steps & 1023) == 0) for (int i = 0; i < m.nweights; ++i) w[i] *= 1.0f - 3.0e-6f;
  }
}

// ---- coder/encoder.cpp:8-34, coder/decoder.cpp:3-39 ----------------------------------------------
namespace {
inline u32 Discretize(float p) { return 1 + 65534 * p; }
struct ByteSink { u8* out; u64 cap, n; bool overflow; void put(u32 b) { if (n < cap) out[n] = (u8)b; else overflow = true; ++n; } };
}  // namespace

extern "C" {

gmo_predictor* gmo_new(void) { return new gmo_predictor(); }
void gmo_free(gmo_predictor* p) { delete p; }
void gmo_set_analysis(gmo_predictor* p, int a) { p->analysis = a != 0; }
float gmo_predict(gmo_predictor* p) { return p->Predict(); }
void gmo_perceive(gmo_predictor* p, int bit) { p->new_bit = bit; }
void gmo_learn(gmo_predictor* p) { p->Learn(); }
void gmo_peek(gmo_predictor* p, float* preds, uint32_t* mask, float* l0, float* l1, float* fin) {
  memcpy(preds, p->predictions, sizeof(p->predictions));
  memcpy(mask, p->active_mask, sizeof(p->active_mask));
  memcpy(l0, p->l0_out, sizeof(p->l0_out));
  memcpy(l1, p->l1_out, sizeof(p->l1_out));
  *fin = p->final_out;
}
void gmo_peek_bytes(gmo_predictor* p, float* ppm, float* lstm) {
  memcpy(ppm, p->ppm_predictions, 1024);
  memcpy(lstm, p->lstm_probs, 1024);
}

int gmo_compress_trace(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_len,
                       float* probs, uint32_t* p16s) {
  ByteSink s{out, cap, 0, false};
  for (int i = 4; i >= 0; --i) s.put((u8)(n >> (8 * i)));  // WriteHeader, runner-utils.cpp:22-27
  gmo_predictor* p = new gmo_predictor();
  p->analysis = (8 * n / 1000) > 0;  // runner-utils.cpp:47 + predictor.cpp:362-365
  u32 x1 = 0, x2 = 0xffffffff;
  for (u64 pos = 0; pos < n; ++pos) {
    char c = (char)in[pos];
    for (int j = 7; j >= 0; --j) {
      int bit = (c >> j) & 1;
      float prob = p->Predict();
      const u32 p16 = Discretize(prob);
      if (probs) probs[pos * 8 + (7 - j)] = prob;
      if (p16s) p16s[pos * 8 + (7 - j)] = p16;
      const u32 xmid = x1 + ((x2 - x1) >> 16) * p16 + (((x2 - x1) & 0xffff) * p16 >> 16);
      if (bit) x2 = xmid; else x1 = xmid + 1;
      while (((x1 ^ x2) & 0xff000000) == 0) { s.put(x2 >> 24); x1 <<= 8; x2 = (x2 << 8) + 255; }
      p->new_bit = bit;
      p->Learn();
    }
  }
  while (((x1 ^ x2) & 0xff000000) == 0) { s.put(x2 >> 24); x1 <<= 8; x2 = (x2 << 8) + 255; }  // Flush
  s.put(x2 >> 24);
  delete p;
  *out_len = s.n;
  return s.overflow ? -1 : 0;
}

int gmo_compress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_len) {
  return gmo_compress_trace(in, n, out, cap, out_len, nullptr, nullptr);
}

int gmo_decompress(const uint8_t* in, uint64_t n, uint8_t* out, uint64_t cap, uint64_t* out_len) {
  u64 rd = 0;
  auto get = [&]() -> u32 { return rd < n ? in[rd++] : (rd++, 0u); };  // Decoder::ReadByte: 0 past EOF
  u64 len = 0;
  for (int i = 0; i <= 4; ++i) len = (len << 8) + get();  // ReadHeader, runner-utils.cpp:29-36
  *out_len = len;
  if (len > cap) return -1;
  gmo_predictor* p = new gmo_predictor();  // analysis stays off in decompress
  u32 x1 = 0, x2 = 0xffffffff, x = 0;
  for (int i = 0; i < 4; ++i) x = (x << 8) + (get() & 0xff);
  for (u64 pos = 0; pos < len; ++pos) {
    int byte = 1;
    while (byte < 256) {
      const u32 p16 = Discretize(p->Predict());
      const u32 xmid = x1 + ((x2 - x1) >> 16) * p16 + (((x2 - x1) & 0xffff) * p16 >> 16);
      int bit = 0;
      if (x <= xmid) { bit = 1; x2 = xmid; } else x1 = xmid + 1;
      p->new_bit = bit;
      p->Learn();
      while (((x1 ^ x2) & 0xff000000) == 0) { x1 <<= 8; x2 = (x2 << 8) + 255; x = (x << 8) + get(); }
      byte += byte + bit;
    }
    out[pos] = (u8)byte;
  }
  delete p;
  return 0;
}

}  // extern "C"

// Debug accessor (tests only): copies the weight vector of mixer `m` for gate-context index `idx`;
// returns the number of weights, or -1 if that set does not exist yet. *steps gets MixerData::steps.
extern "C" int gmo_debug_mixer_set(gmo_predictor* p, int m, uint32_t idx, float* out, uint64_t* steps) {
  MixerSet* d = p->mixers[m].table[idx % p->mixers[m].table_size];
  if (!d) return -1;
  memcpy(out, d->w.data(), d->w.size() * sizeof(float));
  if (steps) *steps = d->steps;
  return (int)d->w.size();
}
