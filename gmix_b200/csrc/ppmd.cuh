// Order-20 PPMd byte model for one stream, run by ONE WARP of the stream's CTA: scalar control on
// lane 0, the scans over a context's symbol statistics one state per lane.
//
// Behaviour follows the reference's ModPPMD (src/models/mod_ppmd.cpp; line cites below refer to
// that file): same 12-byte units, same free lists, same SEE / binary-context estimators, same
// 32-bit *virtual* heap offsets (0 .. 2000 MiB) because offset comparisons against `units_start`
// are semantic. Only three parts of that virtual heap are ever touched:
//   text area   [0, text_cap)                    grows up
//   low units   [units_start, units_start + x)   grows up
//   high units  [heap_end - y, heap_end)         grows down
// They are backed by ONE window of P = 2^k bytes addressed as heap[v & (P - 1)]: the text sits at the bottom, the
// low units start at units_start mod P and grow up, the high units end at heap_end mod P (the top of the window
// while P divides 2000 MiB, i.e. k <= 24; 464 MiB at k = 29, 976 MiB at k = 30) and grow down. The three never
// meet as long as text_cap <= units_start mod P and x + y <= units_cap = (heap_end mod P) - (units_start mod P);
// layout.h (PpmdWindowOf) only picks k for which that gap is one piece above the text area.
// The reference's out-of-memory machinery (AllocUnitsRare/GlueFreeBlocks
// :158-228, RestoreModelRare/cutOff :568-755, Expand/PrepareTextArea :299-348) only runs when the
// 2000 MiB heap is exhausted (> ~90 MB of input); here exhausting the *backed* part raises
// GMX_ERR_PPMD_ARENA for the stream instead (the host re-runs it with a larger window).
#ifndef GMIX_B200_PPMD_CUH_
#define GMIX_B200_PPMD_CUH_
#include <stdint.h>

#include "dmath.cuh"

#if defined(__CUDACC__)
#define GMX_DEV __device__
#define GMX_HOSTDEV __host__ __device__   // (GMX_HD, with force-inline, belongs to dmath.cuh)
#define GMX_NOINLINE __noinline__
#else
#define GMX_DEV
#define GMX_HOSTDEV
#define GMX_NOINLINE
#endif

namespace gmx {

enum : uint32_t {
  PPMD_N_INDEXES = 38, PPMD_MAX_FREQ = 124, PPMD_MAX_ORDER = 20, PPMD_UNIT = 12,
  PPMD_HEAP_END = 2000u << 20,                                       // Init(20, 2000, 1, 0) :1646
  PPMD_UNITS_START = PPMD_HEAP_END - (PPMD_HEAP_END / 8 / 12 * 7 * 12),  // InitSubAllocator :134-140
  PPMD_PERIOD_BITS = 7, PPMD_INTERVAL = 1 << 7, PPMD_BIN_SCALE = 1 << 14, PPMD_SCALE = 1 << 15,
};

struct PpmdSee2 { uint16_t summ; uint8_t shift; uint8_t count; };

// All PPMd state that is not in the heap; lives in the stream arena (global memory).
struct PpmdState {
  uint32_t bl_stamp[PPMD_N_INDEXES + 1], bl_next[PPMD_N_INDEXES + 1];
  uint32_t text_ptr, units_start, lo_unit, hi_unit;
  int32_t order_fall, bsumm, run_length, init_rl, num_masked, prev_success;
  uint32_t found_state, max_context, esc_count;
  uint32_t error;
  uint32_t char_mask[256];
  uint16_t bin_summ[25][64];
  PpmdSee2 see2[23][32];
  PpmdSee2 dummy_see2;
  uint8_t indx2units[PPMD_N_INDEXES], units2indx[128], ns2bs[256], qtable[260];
};

// Constant lookup tables of the model (PPMD_STARTUP :375-400); also filled by the host when a checkpoint
// is turned into an arena image (checkpoint.h).
GMX_HOSTDEV inline void PpmdFillTables(PpmdState* S) {
  int i, k, m, step;
  for (i = 0, k = 1; i < 4; i++, k += 1) S->indx2units[i] = (uint8_t)k;
  for (k++; i < 8; i++, k += 2) S->indx2units[i] = (uint8_t)k;
  for (k++; i < 12; i++, k += 3) S->indx2units[i] = (uint8_t)k;
  for (k++; i < (int)PPMD_N_INDEXES; i++, k += 4) S->indx2units[i] = (uint8_t)k;
  for (k = 0, i = 0; k < 128; k++) { i += S->indx2units[i] < k + 1; S->units2indx[k] = (uint8_t)i; }
  S->ns2bs[0] = 0; S->ns2bs[1] = 2; S->ns2bs[2] = 2;
  for (i = 3; i < 29; i++) S->ns2bs[i] = 4;
  for (i = 29; i < 256; i++) S->ns2bs[i] = 6;
  for (i = 0; i < 5; i++) S->qtable[i] = (uint8_t)i;
  for (m = i = 5, k = step = 1; i < 260; i++) { S->qtable[i] = (uint8_t)m; if (!--k) { k = ++step; m++; } }
}

struct Ppmd {
  PpmdState* S;
  uint8_t* heap;               // backing window of 2^k bytes: byte v of the virtual heap lives at heap[v & mask]
  uint32_t mask;
  uint32_t text_cap, units_cap;
  uint32_t* sqp;  // out: 256 symbol pseudo-probabilities (:1187)
  int lane;       // lane of the calling thread in the PPMd warp (UpdateByte / PrepareByte are warp-collective)
  // Shadow of the only question ever asked of CharMask (:449): "CharMask[sym] == EscCount?" as 256 bits in shared
  // memory, cleared whenever EscCount changes. The array in the arena is still written (it is checkpoint state) but
  // no longer read on the per-byte path.
  uint32_t* masked;
  // Segmented backing (mask == 0; overlay arenas of batched generation, layout.h MakeOverlayLayout): the three live
  // areas sit back to back in private memory instead of in a power-of-two window: text [0, text_cap) at heap + 0, low
  // units [units_start, units_start + lo_cap) at heap + seg_lo, high units [heap_end - hi_cap, heap_end) at heap +
  // seg_lo + lo_cap. A model trained on 1 MB would otherwise drag a 512 MiB window into every stream.
  uint32_t seg_lo = 0, lo_cap = 0, hi_cap = 0;
#if defined(GMX_NO_MASK_SHADOW)
  GMX_DEV bool Masked(uint32_t sy) const { return S->char_mask[sy] == S->esc_count; }
#else
  GMX_DEV bool Masked(uint32_t sy) const { return (masked[sy >> 5] >> (sy & 31)) & 1u; }
#endif
  GMX_DEV void Mask(uint32_t sy, uint32_t ec) const { S->char_mask[sy] = ec; atomicOr(&masked[sy >> 5], 1u << (sy & 31)); }
  GMX_DEV void NextEscCount() const { S->esc_count++; for (int i = 0; i < 8; ++i) masked[i] = 0u; }   // called by one lane

  // ---- virtual heap -> backed memory -------------------------------------------------------
  GMX_DEV uint8_t* At(uint32_t v) const {
#if !defined(GMX_OVERLAY) || !GMX_OVERLAY
    return heap + (v & mask);   // (kernels without overlay mode: no second translation, see stream_kernel.cuh GMX_OVERLAY)
#endif
    if (mask) return heap + (v & mask);
    const uint32_t hi_base = PPMD_HEAP_END - hi_cap;
    return heap + (v < PPMD_UNITS_START ? v : v < hi_base ? seg_lo + (v - PPMD_UNITS_START) : seg_lo + lo_cap + (v - hi_base));
  }
  GMX_DEV uint32_t R8(uint32_t v) const { return *At(v); }
  GMX_DEV void W8(uint32_t v, uint32_t x) const { *At(v) = (uint8_t)x; }
  GMX_DEV uint32_t R16(uint32_t v) const { return *(const uint16_t*)At(v); }
  GMX_DEV void W16(uint32_t v, uint32_t x) const { *(uint16_t*)At(v) = (uint16_t)x; }
  // 32-bit fields inside STATEs are only 2-byte aligned (6-byte packed records)
  GMX_DEV uint32_t R32(uint32_t v) const { const uint16_t* p = (const uint16_t*)At(v); return p[0] | ((uint32_t)p[1] << 16); }
  GMX_DEV void W32(uint32_t v, uint32_t x) const { uint16_t* p = (uint16_t*)At(v); p[0] = (uint16_t)x; p[1] = (uint16_t)(x >> 16); }
  // PPM_CONTEXT {u8 NumStats; u8 Flags; u16 SummFreq; u32 iStats; u32 iSuffix} :425-433
  GMX_DEV uint32_t NumStats(uint32_t c) const { return R8(c); }
  GMX_DEV void SetNumStats(uint32_t c, uint32_t x) const { W8(c, x); }
  GMX_DEV uint32_t Flags(uint32_t c) const { return R8(c + 1); }
  GMX_DEV void SetFlags(uint32_t c, uint32_t x) const { W8(c + 1, x); }
  GMX_DEV uint32_t SummFreq(uint32_t c) const { return R16(c + 2); }
  GMX_DEV void SetSummFreq(uint32_t c, uint32_t x) const { W16(c + 2, x); }
  GMX_DEV uint32_t Stats(uint32_t c) const { return R32(c + 4); }
  GMX_DEV void SetStats(uint32_t c, uint32_t x) const { W32(c + 4, x); }
  GMX_DEV uint32_t Suffix(uint32_t c) const { return R32(c + 8); }
  GMX_DEV void SetSuffix(uint32_t c, uint32_t x) const { W32(c + 8, x); }
  GMX_DEV static uint32_t OneState(uint32_t c) { return c + 2; }  // binary context overlay :432
  // STATE {u8 Symbol; u8 Freq; u32 iSuccessor} :406-410
  GMX_DEV uint32_t Sym(uint32_t s) const { return R8(s); }
  GMX_DEV uint32_t Freq(uint32_t s) const { return R8(s + 1); }
  GMX_DEV void SetSym(uint32_t s, uint32_t x) const { W8(s, x); }
  GMX_DEV void SetFreq(uint32_t s, uint32_t x) const { W8(s + 1, x); }
  GMX_DEV uint32_t Succ(uint32_t s) const { return R32(s + 2); }
  GMX_DEV void SetSucc(uint32_t s, uint32_t x) const { W32(s + 2, x); }
  GMX_DEV void CopyState(uint32_t dst, uint32_t src) const {
    const uint32_t a = R16(src), b = R16(src + 2), c = R16(src + 4);
    W16(dst, a); W16(dst + 2, b); W16(dst + 4, c);
  }
  GMX_DEV void SwapState(uint32_t x, uint32_t y) const {
    const uint32_t a = R16(x), b = R16(x + 2), c = R16(x + 4);
    CopyState(x, y);
    W16(y, a); W16(y + 2, b); W16(y + 4, c);
  }
  GMX_DEV void CopyUnits(uint32_t dst, uint32_t src, uint32_t nu) const {  // UnitsCpy :255
    const uint32_t* s = (const uint32_t*)At(src);
    uint32_t* d = (uint32_t*)At(dst);
    for (uint32_t i = 0; i < 3 * nu; ++i) d[i] = s[i];
  }

  // ---- sub-allocator -----------------------------------------------------------------------
  GMX_DEV void ListInsert(uint32_t i, uint32_t blk, uint32_t nu) const {  // insert :97-103
    uint32_t* b = (uint32_t*)At(blk);
    b[1] = S->bl_next[i]; S->bl_next[i] = blk;
    b[0] = 0xFFFFFFFFu; b[2] = nu;
    S->bl_stamp[i]++;
  }
  GMX_DEV uint32_t ListRemove(uint32_t i) const {  // remove :90-95
    const uint32_t blk = S->bl_next[i];
    S->bl_next[i] = ((const uint32_t*)At(blk))[1];
    S->bl_stamp[i]--;
    return blk;
  }
  GMX_DEV bool Backed() const {  // do the low and high unit areas still fit the backed memory?
#if defined(GMX_OVERLAY) && GMX_OVERLAY
    if (!mask) return S->lo_unit - PPMD_UNITS_START <= lo_cap && PPMD_HEAP_END - S->hi_unit <= hi_cap;
#endif
    return (uint64_t)(S->lo_unit - PPMD_UNITS_START) + (PPMD_HEAP_END - S->hi_unit) <= units_cap;
  }
  GMX_DEV void SplitBlock(uint32_t blk, uint32_t old_i, uint32_t new_i) const {  // :197-208
    uint32_t udiff = S->indx2units[old_i] - S->indx2units[new_i];
    uint32_t p = blk + PPMD_UNIT * S->indx2units[new_i];
    uint32_t i = S->units2indx[udiff - 1];
    if (S->indx2units[i] != udiff) {
      const uint32_t k = S->indx2units[--i];
      ListInsert(i, p, k);
      p += PPMD_UNIT * k;
      udiff -= k;
    }
    ListInsert(S->units2indx[udiff - 1], p, udiff);
  }
  GMX_DEV uint32_t AllocUnits(uint32_t nu) const {  // :230-238
    const uint32_t i = S->units2indx[nu - 1];
    if (S->bl_next[i]) return ListRemove(i);
    const uint32_t ret = S->lo_unit;
    S->lo_unit += PPMD_UNIT * S->indx2units[i];
    if (!Backed()) { S->error = 1; S->lo_unit = ret; return 0; }
    return ret;
  }
  GMX_DEV uint32_t AllocContext() const {  // :240-243
    S->hi_unit -= PPMD_UNIT;
    if (!Backed()) { S->error = 1; S->hi_unit += PPMD_UNIT; return 0; }
    return S->hi_unit;
  }
  GMX_DEV void FreeUnits(uint32_t p, uint32_t nu) const {  // :245-248
    const uint32_t i = S->units2indx[nu - 1];
    ListInsert(i, p, S->indx2units[i]);
  }
  GMX_DEV uint32_t ExpandUnits(uint32_t old, uint32_t old_nu) const {  // :257-267
    const uint32_t i0 = S->units2indx[old_nu - 1], i1 = S->units2indx[old_nu];
    if (i0 == i1) return old;
    const uint32_t p = AllocUnits(old_nu + 1);
    if (!p) return 0;
    CopyUnits(p, old, old_nu);
    ListInsert(i0, old, old_nu);
    return p;
  }
  GMX_DEV uint32_t ShrinkUnits(uint32_t old, uint32_t old_nu, uint32_t new_nu) const {  // :269-282
    const uint32_t i0 = S->units2indx[old_nu - 1], i1 = S->units2indx[new_nu - 1];
    if (i0 == i1) return old;
    if (S->bl_next[i1]) {
      const uint32_t p = ListRemove(i1);
      CopyUnits(p, old, new_nu);
      ListInsert(i0, old, S->indx2units[i0]);
      return p;
    }
    SplitBlock(old, i0, i1);
    return old;
  }

  // ---- model start: PPMD_STARTUP :375-400 + StartModelRare :659-713 ---------------------------
  GMX_DEV GMX_NOINLINE void Init() const {
    int i, k;
    PpmdFillTables(S);
#pragma unroll 1
    for (i = 0; i < 256; i++) S->char_mask[i] = 0;
    S->esc_count = 1;
    for (i = 0; i < 8; i++) masked[i] = 0u;
    S->order_fall = PPMD_MAX_ORDER;
    for (i = 0; i <= (int)PPMD_N_INDEXES; i++) { S->bl_stamp[i] = 0; S->bl_next[i] = 0; }
    S->text_ptr = 0;
    S->hi_unit = PPMD_HEAP_END;
    S->lo_unit = S->units_start = PPMD_UNITS_START;
    S->init_rl = -13; S->run_length = -13;
    S->error = 0;
    const uint32_t mc = AllocContext();
    S->max_context = mc;
    SetNumStats(mc, 255);
    SetSummFreq(mc, 257);
    const uint32_t st = AllocUnits(128);
    SetStats(mc, st);
    SetFlags(mc, 0);
    SetSuffix(mc, 0);
    S->prev_success = 0;
#pragma unroll 1
    for (i = 0; i < 256; i++) { SetSym(st + 6 * i, i); SetFreq(st + 6 * i, 1); SetSucc(st + 6 * i, 0); }
    const int esc_coef[12] = {16, -10, 1, 51, 14, 89, 23, 35, 64, 26, -42, 43};  // :35-36
    uint8_t i2f[25];
    for (k = i = 0; i < 25; i2f[i++] = (uint8_t)(k + 1)) while (S->qtable[k] == i) k++;
#pragma unroll 1
    for (k = 0; k < 64; k++) {
      int s = 0;
      for (i = 0; i < 6; i++) s += esc_coef[2 * i + ((k >> i) & 1)];
      s = s < 32 ? 32 : s > 224 ? 224 : s;
      s *= 128;
      for (i = 0; i < 25; i++) S->bin_summ[i][k] = (uint16_t)(PPMD_BIN_SCALE - s / i2f[i]);
    }
#pragma unroll 1
    for (i = 0; i < 23; i++)
#pragma unroll 1
      for (k = 0; k < 32; k++) {
      S->see2[i][k].shift = PPMD_PERIOD_BITS - 4;
      S->see2[i][k].summ = (uint16_t)((8 * i + 5) << (PPMD_PERIOD_BITS - 4));
      S->see2[i][k].count = 7;
    }
    S->dummy_see2.summ = 0; S->dummy_see2.shift = 0; S->dummy_see2.count = 0;
    S->found_state = 0; S->bsumm = 0; S->num_masked = 0;
  }

  GMX_DEV void See2Update(PpmdSee2* s) const {  // :478-493
    if (--s->count == 0) {
      uint32_t i = s->summ >> s->shift;
      i = PPMD_PERIOD_BITS - (i > 40) - (i > 280) - (i > 1020);
      if (i < s->shift) { s->summ >>= 1; s->shift--; }
      else if (i > s->shift) { s->summ <<= 1; s->shift++; }
      s->count = (uint8_t)(5 << s->shift);
    }
  }

  // rescale :498-566
  GMX_DEV GMX_NOINLINE uint32_t Rescale(uint32_t q, int order_fall, uint32_t fs) const {
    SetFlags(q, Flags(q) & 0x14);
    const uint32_t p1 = Stats(q);
    uint32_t t0 = R16(fs), t1 = R16(fs + 2), t2 = R16(fs + 4);
    uint32_t p;
    for (p = fs; p != p1; p -= 6) CopyState(p, p - 6);
    W16(p1, t0); W16(p1 + 2, t1); W16(p1 + 4, t2);
    const int of = (order_fall != 0);
    int a, i;
    const int f0 = (int)Freq(p);
    int sf = (int)SummFreq(q);
    int esc = sf - f0;
    SetFreq(p, (f0 + of) >> 1);
    SetSummFreq(q, Freq(p));
    const int ns = (int)NumStats(q);
    for (i = 0; i < ns; i++) {
      p += 6;
      a = (int)Freq(p);
      esc -= a;
      a = (a + of) >> 1;
      SetFreq(p, a);
      SetSummFreq(q, SummFreq(q) + a);
      if (a) SetFlags(q, Flags(q) | (0x08 * (Sym(p) >= 0x40)));
      if (a > (int)Freq(p - 6)) {
        t0 = R16(p); t1 = R16(p + 2); t2 = R16(p + 4);
        const uint32_t tf = t0 >> 8;  // Freq byte of the saved state
        uint32_t pp;
        for (pp = p; tf > Freq(pp - 6); pp -= 6) CopyState(pp, pp - 6);
        W16(pp, t0); W16(pp + 2, t1); W16(pp + 4, t2);
      }
    }
    if (Freq(p) == 0) {
      for (i = 0; Freq(p) == 0; i++, p -= 6) {}
      esc += i;
      a = (ns + 2) >> 1;
      const int ns_new = ns - i;
      SetNumStats(q, ns_new);
      if (ns_new == 0) {
        const uint32_t st = Stats(q);
        t0 = R16(st); t1 = R16(st + 2); t2 = R16(st + 4);
        int nf = (2 * (int)(t0 >> 8) + esc - 1) / esc;
        if (nf > (int)PPMD_MAX_FREQ / 3) nf = PPMD_MAX_FREQ / 3;
        t0 = (t0 & 0xff) | ((uint32_t)nf << 8);
        SetFlags(q, Flags(q) & 0x18);
        FreeUnits(st, a);
        const uint32_t os = OneState(q);
        W16(os, t0); W16(os + 2, t1); W16(os + 4, t2);
        return os;
      }
      SetStats(q, ShrinkUnits(Stats(q), a, (ns_new + 2) >> 1));
    }
    SetSummFreq(q, SummFreq(q) + ((esc + 1) >> 1));
    if (order_fall || (Flags(q) & 0x04) == 0) {
      a = (sf -= esc) - f0;
      const uint32_t v = (uint32_t)((f0 * (int)SummFreq(q) - sf * (int)Freq(Stats(q)) + a - 1) / a);
      a = (int)(v < 2u ? 2u : v > (PPMD_MAX_FREQ / 2u - 18u) ? (PPMD_MAX_FREQ / 2u - 18u) : v);
    } else {
      a = 2;
    }
    const uint32_t st = Stats(q);
    SetFreq(st, Freq(st) + a);
    SetSummFreq(q, SummFreq(q) + a);
    SetFlags(q, Flags(q) | 0x04);
    return st;
  }

  // CreateSuccessors :888-969 (`p` = state of the coded symbol in suffix(pc), or 0). Returns 0 on
  // arena exhaustion.
  GMX_DEV GMX_NOINLINE uint32_t CreateSuccessors(bool skip, uint32_t p, uint32_t pc) const {
    uint32_t ps[PPMD_MAX_ORDER + 4];
    int n = 0;
    uint32_t sym = Sym(S->found_state);
    const uint32_t up = Succ(S->found_state);
    bool no_loop = false;
    if (!skip) {
      ps[n++] = S->found_state;
      if (!Suffix(pc)) no_loop = true;
    }
    if (!no_loop) {
      bool first = p != 0;
      if (first) pc = Suffix(pc);
      for (;;) {
        if (!first) {
          pc = Suffix(pc);
          if (NumStats(pc)) {
            for (p = Stats(pc); Sym(p) != sym; p += 6) {}
            const uint32_t t = 2 * (Freq(p) < PPMD_MAX_FREQ - 1);
            SetFreq(p, Freq(p) + t);
            SetSummFreq(pc, SummFreq(pc) + t);
          } else {
            p = OneState(pc);
            SetFreq(p, Freq(p) + ((!NumStats(Suffix(pc))) & (Freq(p) < 16)));
          }
        }
        first = false;
        if (Succ(p) != up) { pc = Succ(p); break; }
        if (n < (int)PPMD_MAX_ORDER + 4) ps[n++] = p; else { S->error = 2; return 0; }
        if (!Suffix(pc)) break;
      }
    }
    if (n == 0) return pc;
    uint32_t flags = 0x10 * (sym >= 0x40);
    sym = R8(up);
    const uint32_t succ = up + 1;
    flags |= 0x08 * (sym >= 0x40);
    uint32_t freq;
    if (NumStats(pc)) {
      for (p = Stats(pc); Sym(p) != sym; p += 6) {}
      uint32_t cf = Freq(p) - 1;
      const uint32_t s0 = SummFreq(pc) - NumStats(pc) - cf;
      cf = 1 + ((2 * cf < s0) ? (uint32_t)(12 * cf > s0) : 2 + cf / s0);
      freq = cf < 7 ? cf : 7;
    } else {
      freq = Freq(OneState(pc));
    }
    do {
      const uint32_t pc1 = AllocContext();
      if (!pc1) return 0;
      W8(pc1, 0); W8(pc1 + 1, flags); W8(pc1 + 2, sym); W8(pc1 + 3, freq);
      W32(pc1 + 4, succ);
      SetSuffix(pc1, pc);
      pc = pc1;
      SetSucc(ps[--n], pc);
    } while (n);
    return pc;
  }

  GMX_DEV GMX_NOINLINE uint32_t ReduceOrder(uint32_t p, uint32_t pc) const {  // :971-1018
    const uint32_t pc1 = pc;
    SetSucc(S->found_state, S->text_ptr);
    const uint32_t sym = Sym(S->found_state);
    const uint32_t up = Succ(S->found_state);
    S->order_fall++;
    bool first = p != 0;
    if (first) pc = Suffix(pc);
    for (;;) {
      if (!first) {
        if (!Suffix(pc)) return pc;
        pc = Suffix(pc);
        if (NumStats(pc)) {
          for (p = Stats(pc); Sym(p) != sym; p += 6) {}
          const uint32_t t = 2 * (Freq(p) < PPMD_MAX_FREQ - 3);
          SetFreq(p, Freq(p) + t);
          SetSummFreq(pc, SummFreq(pc) + t);
        } else {
          p = OneState(pc);
          SetFreq(p, Freq(p) + (Freq(p) < 11));
        }
      }
      first = false;
      if (Succ(p)) break;
      SetSucc(p, up);
      S->order_fall++;
    }
    if (Succ(p) <= up) {
      const uint32_t saved = S->found_state;
      S->found_state = p;
      const uint32_t ns = CreateSuccessors(false, 0, pc);
      S->found_state = saved;
      if (!ns) return 0;
      SetSucc(p, ns);
    }
    if (S->order_fall == 1 && pc1 == S->max_context) {
      SetSucc(S->found_state, Succ(p));
      S->text_ptr--;
    }
    return Succ(p);
  }

  GMX_DEV GMX_NOINLINE void UpdateModel(uint32_t minc) const {  // :759-886
    const uint32_t fs = S->found_state;
    const uint32_t fsym = Sym(fs);
    const uint32_t ffreq = Freq(fs);
    uint32_t fsucc = Succ(fs);
    uint32_t p = 0, pc;
    if (Suffix(minc)) {
      pc = Suffix(minc);
      if (NumStats(pc)) {
        p = Stats(pc);
        if (Sym(p) != fsym) {
          for (p += 6; Sym(p) != fsym; p += 6) {}
          if (Freq(p) >= Freq(p - 6)) { SwapState(p, p - 6); p -= 6; }
        }
        if (Freq(p) < PPMD_MAX_FREQ - 3) {
          const uint32_t cf = 2 + (ffreq < 28);
          SetFreq(p, Freq(p) + cf);
          SetSummFreq(pc, SummFreq(pc) + cf);
        }
      } else {
        p = OneState(pc);
        SetFreq(p, Freq(p) + (Freq(p) < 14));
      }
    }
    if (!S->order_fall && fsucc) {
      const uint32_t ns = CreateSuccessors(true, p, minc);
      if (!ns) { S->error |= 1; return; }
      SetSucc(fs, ns);
      S->max_context = ns;
      return;
    }
    if (S->text_ptr + 1 >= text_cap) { S->error |= 4; return; }
    W8(S->text_ptr++, fsym);
    uint32_t succ = S->text_ptr;
    if (fsucc) {
      if (fsucc < S->units_start) fsucc = CreateSuccessors(false, p, minc);
    } else {
      fsucc = ReduceOrder(p, minc);
    }
    if (!fsucc) { S->error |= 1; return; }
    if (!--S->order_fall) {
      succ = fsucc;
      S->text_ptr -= (S->max_context != minc);
    }
    const uint32_t s0 = SummFreq(minc) - ffreq;
    const uint32_t ns = NumStats(minc);
    const uint32_t flag = 0x08 * (fsym >= 0x40);
    for (pc = S->max_context; pc != minc; pc = Suffix(pc)) {
      const uint32_t ns1 = NumStats(pc);
      if (ns1) {
        if (ns1 & 1) {
          const uint32_t np = ExpandUnits(Stats(pc), (ns1 + 1) >> 1);
          if (!np) { S->error |= 1; return; }
          SetStats(pc, np);
        }
        SetSummFreq(pc, SummFreq(pc) + (S->qtable[ns + 4] >> 3));
      } else {
        p = AllocUnits(1);
        if (!p) { S->error |= 1; return; }
        CopyState(p, OneState(pc));
        SetStats(pc, p);
        const uint32_t f = Freq(p);
        SetFreq(p, (f <= PPMD_MAX_FREQ / 3) ? (2 * f - 1) : (PPMD_MAX_FREQ - 15));
        const uint32_t exp_escape[16] = {51, 43, 18, 12, 11, 9, 8, 7, 6, 5, 4, 3, 3, 2, 2, 2};  // :39-40
        SetSummFreq(pc, Freq(p) + (ns > 1) + exp_escape[S->qtable[S->bsumm >> 8]]);
      }
      uint32_t cf = (ffreq - 1) * (5 + SummFreq(pc));
      const uint32_t sf = s0 + SummFreq(pc);
      if (cf <= 3 * sf) {
        cf = 1 + (2 * cf > sf) + (2 * cf > 3 * sf);
        SetSummFreq(pc, SummFreq(pc) + 4);
      } else {
        cf = 5 + (cf > 5 * sf) + (cf > 6 * sf) + (cf > 8 * sf) + (cf > 10 * sf) + (cf > 12 * sf);
        SetSummFreq(pc, SummFreq(pc) + cf);
      }
      const uint32_t nn = ns1 + 1;
      SetNumStats(pc, nn);
      p = Stats(pc) + 6 * nn;
      SetSucc(p, succ);
      SetSym(p, fsym);
      SetFreq(p, cf);
      SetFlags(pc, Flags(pc) | flag);
    }
    S->max_context = fsucc;
  }

  GMX_DEV uint16_t* BinSummFor(uint32_t q) const {  // :1026-1028 / :1224-1226
    const uint32_t rs = OneState(q);
    const int i = S->ns2bs[NumStats(Suffix(q))] + S->prev_success + (int)Flags(q) + ((S->run_length >> 26) & 0x20);
    return &S->bin_summ[S->qtable[Freq(rs) - 1]][i];
  }
  GMX_DEV PpmdSee2* See2For(uint32_t q, int cnum, int num_masked, int* see_freq) const {  // :1116-1124 / :1270-1278
    if (cnum != 0xFF) {
      PpmdSee2* s = S->see2[S->qtable[cnum + 3] - 4];
      s += ((int)SummFreq(q) > 10 * (cnum + 1));
      s += 2 * (2 * cnum < (int)NumStats(Suffix(q)) + num_masked) + (int)Flags(q);
      *see_freq = (s->summ >> s->shift) + 1;
      return s;
    }
    *see_freq = 1;
    return &S->dummy_see2;
  }

  // ---- warp-collective entry points --------------------------------------------------------
  // UpdateByte / PrepareByte are called by ALL 32 lanes of one warp with identical arguments. Control
  // flow is uniform (every lane reads the same scalars), the per-symbol scans over a context's
  // statistics run one state per lane, and everything that writes scalar state runs on lane 0
  // between two __syncwarp()s (GMX_L0). Only order-independent reductions cross lanes (integer
  // sums, first-hit ballots), so the result is identical to the serial code of the reference.
#define GMX_L0(...) do { __syncwarp(); if (lane == 0) { __VA_ARGS__; } __syncwarp(); } while (0)

  GMX_DEV int WarpSum(int v) const {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  }
  // index of the first of the n states at p whose symbol is sym, or -1
  GMX_DEV int FindSym(uint32_t p, int n, uint32_t sym) const {
    for (int base = 0; base < n; base += 32) {
      const int k = base + lane;
      const bool hit = k < n && Sym(p + 6 * k) == sym;
      const unsigned m = __ballot_sync(0xffffffffu, hit);
      if (m) return base + __ffs((int)m) - 1;
    }
    return -1;
  }

  // ppmd_UpdateByte :1351-1382 — code byte `c`, update the model.
  GMX_DEV void UpdateByte(uint32_t c) const {
    uint32_t minc = S->max_context;
    if (NumStats(minc)) {  // processSymbol1<0> :1049-1098
      const uint32_t p0 = Stats(minc);
      const int cnum = (int)NumStats(minc);
      const int i = FindSym(p0, cnum + 1, c);
      if (i >= 0) {
        GMX_L0(
          S->prev_success = 0;
          uint32_t p = p0 + 6 * i;
          SetFreq(p, Freq(p) + 4); SetSummFreq(minc, SummFreq(minc) + 4);
          if (i > 0 && Freq(p) > Freq(p - 6)) { SwapState(p, p - 6); p -= 6; }
          S->found_state = p;
          if (Freq(p) > PPMD_MAX_FREQ) S->found_state = Rescale(minc, S->order_fall, p));
      } else {
        const uint32_t ec = S->esc_count;
        for (int k = lane; k <= cnum; k += 32) Mask(Sym(p0 + 6 * k), ec);
        GMX_L0(S->prev_success = 0; S->num_masked = cnum; S->found_state = 0);
      }
    } else {  // processBinSymbol<0> :1023-1046
      GMX_L0(
        const uint32_t rs = OneState(minc);
        uint16_t* bs = BinSummFor(minc);
        S->bsumm = *bs;
        *bs = (uint16_t)(*bs - ((S->bsumm + 64) >> PPMD_PERIOD_BITS));
        if (Sym(rs) != c) {
          Mask(Sym(rs), S->esc_count); S->num_masked = 0; S->prev_success = 0; S->found_state = 0;
        } else {
          *bs = (uint16_t)(*bs + PPMD_INTERVAL);
          SetFreq(rs, Freq(rs) + (Freq(rs) < 196));
          S->run_length++; S->prev_success = 1; S->found_state = rs;
        });
    }
    while (!S->found_state) {
      int climbed = 0;
      const int nm = S->num_masked;
      do { climbed++; minc = Suffix(minc); } while ((int)NumStats(minc) == nm);
      // processSymbol2<0> :1104-1170
      const uint32_t p = Stats(minc);
      const int cnum = (int)NumStats(minc);
      const uint32_t ec = S->esc_count;
      int low = 0, hit_i = -1;
      for (int base = 0; base <= cnum; base += 32) {
        const int k = base + lane;
        bool hit = false;
        if (k <= cnum) {
          const uint32_t sy = Sym(p + 6 * k);
          if (!Masked(sy)) {
            Mask(sy, ec);
            low += (int)Freq(p + 6 * k);
            hit = sy == c;
          }
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (m) hit_i = base + __ffs((int)m) - 1;
      }
      low = WarpSum(low);
      GMX_L0(
        S->order_fall += climbed;
        int see_freq;
        PpmdSee2* see = See2For(minc, cnum, S->num_masked, &see_freq);
        const int total = see_freq + low;
        if (hit_i >= 0) {
          const uint32_t ph = p + 6 * hit_i;
          if (see_freq > 2) see->summ = (uint16_t)(see->summ - see_freq);
          See2Update(see);
          S->found_state = ph;
          SetFreq(ph, Freq(ph) + 4); SetSummFreq(minc, SummFreq(minc) + 4);
          if (Freq(ph) > PPMD_MAX_FREQ) S->found_state = Rescale(minc, S->order_fall, ph);
          S->run_length = S->init_rl;
          NextEscCount();
        } else {
          S->num_masked = cnum;
          see->summ = (uint16_t)(see->summ + (total - see_freq));
        });
    }
    GMX_L0(
      if (S->order_fall != 0 || Succ(S->found_state) < S->units_start) UpdateModel(minc);
      else S->max_context = Succ(S->found_state));
  }

  // ConvertSQ :1192-1209 for one symbol: cum * freq / total (+1 for a symbol, the new cum for an escape)
  // freq, total < 2^16 and freq <= total, so the 48-bit product divides in two 32-bit steps (schoolbook, base 2^16)
  // instead of the software 64-bit division: hi < total always, every partial dividend stays below 2^32.
  GMX_DEV static uint32_t Scale(uint32_t cum, uint32_t freq, uint32_t total) {
    const uint64_t prod = (uint64_t)cum * freq;
    const uint32_t hi = (uint32_t)(prod >> 32), lo = (uint32_t)prod;
    if (total > 0xffffu || hi >= total) return (uint32_t)(prod / total);   // never taken (16-bit statistics); keeps the function total
    uint32_t r = (hi << 16) | (lo >> 16);
    const uint32_t q1 = r / total;
    r = ((r - q1 * total) << 16) | (lo & 0xffffu);
    return (q1 << 16) + r / total;
  }

  // ppmd_PrepareByte :1322-1349 with the *_T walkers :1222-1297: full next-byte distribution.
  // (OrderFall is incremented and restored by the reference and read by nothing in between.)
  GMX_DEV void PrepareByte() const {
    uint32_t cum = 0xFFFFFF00u;
    for (int i = lane; i < 256; i += 32) sqp[i] = 0;
    uint32_t minc = S->max_context;
    const uint32_t ec = S->esc_count;
    int nm;
    __syncwarp();
    if (NumStats(minc)) {  // processSymbol1_T
      const uint32_t p = Stats(minc);
      const int cnum = (int)NumStats(minc);
      const uint32_t total = SummFreq(minc);
      int low = 0;
      for (int k = lane; k <= cnum; k += 32) {
        const uint32_t f = Freq(p + 6 * k), sy = Sym(p + 6 * k);
        sqp[sy] = Scale(cum, f, total) + 1;
        low += (int)f;
        Mask(sy, ec);
      }
      low = WarpSum(low);
      nm = cnum;
      cum = Scale(cum, (total - (uint32_t)low) & 0xffff, total);
    } else {  // processBinSymbol_T
      const uint32_t rs = OneState(minc);
      const int bsv = *BinSummFor(minc);
      const uint32_t sy = Sym(rs);
      GMX_L0(S->bsumm = bsv; sqp[sy] = Scale(cum, (uint32_t)(bsv + bsv) & 0xffff, PPMD_SCALE) + 1; Mask(sy, ec));
      cum = Scale(cum, (uint32_t)(PPMD_SCALE - bsv - bsv) & 0xffff, PPMD_SCALE);
      nm = 0;
    }
    __syncwarp();
    for (;;) {
      bool root = false;
      do {
        if (!Suffix(minc)) { root = true; break; }
        minc = Suffix(minc);
      } while ((int)NumStats(minc) == nm);
      if (root) break;
      // processSymbol2_T
      const uint32_t p = Stats(minc);
      const int cnum = (int)NumStats(minc);
      int see_freq;
      See2For(minc, cnum, nm, &see_freq);
      int low = 0;
      for (int k = lane; k <= cnum; k += 32)
        if (!Masked(Sym(p + 6 * k))) low += (int)Freq(p + 6 * k);
      low = WarpSum(low);
      const uint32_t total = ((uint32_t)see_freq + (uint32_t)low) & 0xffff;
      for (int k = lane; k <= cnum; k += 32) {
        const uint32_t sy = Sym(p + 6 * k);
        if (!Masked(sy)) { sqp[sy] = Scale(cum, Freq(p + 6 * k), total) + 1; Mask(sy, ec); }
      }
      cum = Scale(cum, (uint32_t)see_freq & 0xffff, total);
      nm = cnum;
      __syncwarp();
    }
    GMX_L0(NextEscCount(); S->num_masked = 0);
  }
#undef GMX_L0
};

}  // namespace gmx
#endif
