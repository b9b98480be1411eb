// Fixed model graph of the gmix predictor (reference src/predictor.cpp:17-358, SURVEY.md
// appendix A): 90 predictions, 41 Indirect, 6 Match, 9 IndirectHash, 24+8+1 mixers.
// Everything here is compile-time data shared by the device kernels and the host layout code.
#ifndef GMIX_B200_SPEC_CUH_
#define GMIX_B200_SPEC_CUH_
#include <stdint.h>

namespace gmx {

enum : int {
  NPRED = 90, NL0 = 24, NL1 = 8, NMIX = 33, NIND = 41, NMATCH = 6, NIH = 9,
  // prediction indices (ShortTermMemory::AddPrediction order)
  P_PPMD = 0, P_LSTM = 1, P_IND0 = 2, P_MATCH0 = 66, P_DIND0 = 72,
};

// Context slots of the per-stream blackboard (ShortTermMemory, short-term-memory.h:60-120).
enum CtxId : int {
  C_ZERO = 0, C_LAST_BYTE, C_BIT_CONTEXT, C_H2, C_H3, C_H4, C_H5, C_H6, C_LBPR, C_SLPR,
  C_RB1, C_RB2, C_RB3, C_RB4, C_RB5, C_RB6, C_RB7, C_RB8, C_RB9, C_LSTM, C_LONGEST,
  C_IV0,                // 9 interval contexts (predictor.cpp:54-76)
  C_SK0 = C_IV0 + 9,    // 15 skip contexts (predictor.cpp:122-185)
  C_IH0 = C_SK0 + 15,   // 9 indirect-hash contexts (predictor.cpp:210-249)
  C_COUNT = C_IH0 + 9   // 54
};

struct IndirectSpec { uint8_t ctx; uint8_t log2; uint8_t slow_lr; uint8_t pred; };
// 41 Indirect models in construction order: 17 of AddIndirect, 15 of AddSkip, 9 of
// AddDoubleIndirect (learning rate 1/200 instead of 0.02). `pred` = first prediction index.
#define GMX_INDIRECT_SPECS                                                                          \
  {C_LAST_BYTE, 8, 0, 2}, {C_H2, 16, 0, 4}, {C_H3, 15, 0, 6}, {C_H3, 16, 0, 8}, {C_H4, 15, 0, 10},   \
  {C_H5, 15, 0, 12}, {C_H6, 15, 0, 14}, {C_RB1, 8, 0, 16}, {C_RB2, 8, 0, 18}, {C_RB3, 8, 0, 20},     \
  {C_RB4, 8, 0, 22}, {C_RB5, 8, 0, 24}, {C_RB6, 8, 0, 26}, {C_RB7, 8, 0, 28}, {C_RB8, 8, 0, 30},     \
  {C_RB9, 8, 0, 32}, {C_LSTM, 8, 0, 34},                                                            \
  {C_SK0 + 0, 16, 0, 36}, {C_SK0 + 1, 16, 0, 38}, {C_SK0 + 2, 16, 0, 40}, {C_SK0 + 3, 16, 0, 42},    \
  {C_SK0 + 4, 16, 0, 44}, {C_SK0 + 5, 16, 0, 46}, {C_SK0 + 6, 16, 0, 48}, {C_SK0 + 7, 16, 0, 50},    \
  {C_SK0 + 8, 16, 0, 52}, {C_SK0 + 9, 16, 0, 54}, {C_SK0 + 10, 16, 0, 56}, {C_SK0 + 11, 16, 0, 58},  \
  {C_SK0 + 12, 16, 0, 60}, {C_SK0 + 13, 16, 0, 62}, {C_SK0 + 14, 16, 0, 64},                         \
  {C_IH0 + 0, 8, 1, 72}, {C_IH0 + 1, 16, 1, 74}, {C_IH0 + 2, 15, 1, 76}, {C_IH0 + 3, 8, 1, 78},      \
  {C_IH0 + 4, 16, 1, 80}, {C_IH0 + 5, 15, 1, 82}, {C_IH0 + 6, 8, 1, 84}, {C_IH0 + 7, 16, 1, 86},     \
  {C_IH0 + 8, 15, 1, 88}

// Hashed byte contexts: which ring bytes (0 = last byte) feed MurmurHash3_x86_32 (skip-context.cpp:9-19).
// First 5 = last 2..6 bytes (-> C_H2..C_H6), then the 15 skip contexts (-> C_SK0..).
struct SkipSpec { uint8_t n; uint8_t b[6]; };
#define GMX_SKIP_SPECS                                                                              \
  {2, {0, 1}}, {3, {0, 1, 2}}, {4, {0, 1, 2, 3}}, {5, {0, 1, 2, 3, 4}}, {6, {0, 1, 2, 3, 4, 5}},     \
  {2, {1, 2}}, {3, {1, 2, 3}}, {2, {0, 2}}, {3, {0, 2, 3}}, {4, {1, 2, 3, 4}}, {2, {0, 3}},          \
  {2, {0, 4}}, {2, {0, 5}}, {4, {0, 2, 3, 4}}, {3, {0, 3, 4}}, {2, {0, 6}}, {2, {0, 7}},             \
  {4, {0, 1, 3, 4}}, {3, {0, 4, 5}}, {4, {0, 1, 2, 4}}

// Interval contexts: ctx = mask & ((ctx << shift) + last_byte / div) (interval-context.cpp:3-23).
struct IntervalSpec { uint8_t div_log2; uint8_t shift; uint16_t mask; };
#define GMX_INTERVAL_SPECS                                                                          \
  {4, 4, 0xF}, {4, 4, 0xFF}, {4, 4, 0xFFF}, {5, 3, 0x7}, {5, 3, 0x3F}, {5, 3, 0xFFF},                \
  {6, 2, 0xF}, {6, 2, 0xFF}, {6, 2, 0xFFF}

// IndirectHash(outer_order, table 2^log2, inner_order) (predictor.cpp:213-245, indirect-hash.cpp:7-14).
struct IHSpec { uint8_t outer_order; uint8_t log2; uint8_t inner_order; };
#define GMX_IH_SPECS {1, 8, 1}, {1, 8, 2}, {1, 8, 3}, {2, 16, 1}, {2, 16, 2}, {2, 16, 3}, {3, 24, 1}, {4, 24, 2}, {4, 24, 3}

// Match(table 2^log2, byte context) (predictor.cpp:187-208).
struct MatchSpec { uint8_t ctx; uint8_t log2; };
#define GMX_MATCH_SPECS {C_LAST_BYTE, 8}, {C_H2, 16}, {C_H3, 24}, {C_H4, 21}, {C_H5, 21}, {C_H6, 21}

// Mixers (predictor.cpp:251-358): gate context, learning rate (double literal rounded to float
// exactly as `float learning_rate` does), table 2^log2. Order: 24 layer-0, 8 layer-1, 1 final.
struct MixerSpec { uint8_t ctx; uint8_t log2; float lr; };
#define GMX_MIXER_SPECS                                                                             \
  {C_LAST_BYTE, 8, (float)0.005}, {C_RB3, 8, (float)0.0055}, {C_SLPR, 16, (float)0.003}, {C_H4, 15, (float)0.0045},          \
  {C_IH0 + 6, 8, (float)0.006}, {C_RB1, 8, (float)0.004}, {C_LONGEST, 3, (float)0.0005}, {C_H2, 16, (float)0.0035},          \
  {C_RB2, 8, (float)0.0065}, {C_H3, 15, (float)0.0025}, {C_LAST_BYTE, 8, (float)0.001}, {C_LBPR, 16, (float)0.002},          \
  {C_IV0 + 0, 4, (float)0.005}, {C_IV0 + 1, 8, (float)0.0045}, {C_IV0 + 2, 12, (float)0.0055}, {C_IV0 + 3, 3, (float)0.004}, \
  {C_IV0 + 4, 6, (float)0.0035}, {C_SK0 + 2, 16, (float)0.006}, {C_IV0 + 5, 12, (float)0.003}, {C_IV0 + 6, 4, (float)0.0065},\
  {C_IV0 + 7, 8, (float)0.003}, {C_IV0 + 8, 12, (float)0.0025}, {C_LSTM, 8, (float)0.002}, {C_ZERO, 0, (float)0.0005},       \
  {C_RB1, 8, (float)0.0045}, {C_ZERO, 0, (float)0.0035}, {C_BIT_CONTEXT, 8, (float)0.003}, {C_RB2, 8, (float)0.002},         \
  {C_LAST_BYTE, 8, (float)0.0025}, {C_BIT_CONTEXT, 8, (float)0.00001}, {C_LONGEST, 3, (float)0.0008},                  \
  {C_ZERO, 0, (float)0.0004}, {C_ZERO, 0, (float)0.0005}

// Number of weights of mixer m (mixer.cpp:17-26; one skip connection = the LSTM prediction).
constexpr int MixerWeights(int m) { return m < NL0 ? NPRED + m : m < NL0 + NL1 ? NL0 + (m - NL0) + 1 : NL0 + NL1 + 1; }

// LSTM dimensions (lstm-model.cpp:7).
enum : int { L_CELLS = 50, L_HORIZON = 100, L_NIN = 307, L_ROW = 563, L_NOUT = 256, L_HID = 51, L_UPDATE_LIMIT = 3000 };
// Gate weights and their Adam moments are stored as [3 gates][L_ROWQ quads of input columns][L_CELLS][4]: the four
// consecutive input columns of one cell are one float4 and cells are adjacent, so the forward pass (one gate row per
// thread, lanes = cells) moves 512 contiguous bytes per warp load and keeps 4x the bytes in flight of scalar loads.
// Column 563 is padding (always 0).
enum : int { L_ROWQ = (L_ROW + 3) / 4, L_WSIZE = 3 * L_ROWQ * L_CELLS * 4 };
#if defined(__CUDACC__)
__host__ __device__
#endif
constexpr inline unsigned LstmW(int g, int j, int i) { return (unsigned)(((g * L_ROWQ + (j >> 2)) * L_CELLS + i) * 4 + (j & 3)); }

}  // namespace gmx
#endif
