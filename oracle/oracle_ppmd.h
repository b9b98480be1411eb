// TEST INFRASTRUCTURE ONLY — see gmix_oracle.h.
// Fresh restatement of the reference's order-20 PPMd byte model (models/mod_ppmd.cpp; cites below
// are line numbers in that file). Integer-only. The model lives in a 2000 MB byte heap addressed by
// 32-bit offsets; comparisons between offsets and `units_start` are semantic,
// so the same virtual offsets are kept here (the heap is calloc'ed: untouched pages cost nothing).
//
// Not restated (unreachable for from-scratch streams below ~90 MB):
// the memory-exhaustion paths AllocUnitsRare/GlueFreeBlocks (:158-228), RestoreModelRare/cutOff
// (:568-755), ExpandTextArea/PrepareTextArea/MoveUnitsUp (:284-348). Reaching one aborts loudly.
#ifndef ORACLE_PPMD_H_
#define ORACLE_PPMD_H_
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace oracle_ppmd {

typedef uint8_t u8;
typedef uint16_t u16;
typedef uint32_t u32;
typedef uint64_t u64;

enum { N_INDEXES = 38, MAX_FREQ = 124, MAX_ORDER = 20, UNIT = 12 };
enum { INT_BITS = 7, PERIOD_BITS = 7, INTERVAL = 1 << INT_BITS, BIN_SCALE = 1 << (INT_BITS + PERIOD_BITS), SCALE = 1 << 15 };

struct See2 { u16 summ; u8 shift; u8 count; };  // :465-494
struct SqEntry { u16 sym, freq, total; };         // :1172-1182

struct Model {
  u8* heap = nullptr;
  u64 heap_size = 0;
  // sub-allocator (:113-122)
  u32 bl_stamp[N_INDEXES + 1], bl_next[N_INDEXES + 1];
  u32 text_ptr, units_start, lo_unit, hi_unit;
  // tables (:368-399)
  u8 indx2units[N_INDEXES], units2indx[128], ns2bs[256], qtable[260];
  // model state (:441-453, 496, 1020-1021, 1100-1101)
  int order_fall, bsumm, run_length, init_rl, num_masked, prev_success;
  u32 found_state;  // offset of the found STATE, 0 = none
  u32 max_context;
  u32 esc_count, char_mask[256];
  u16 bin_summ[25][64];
  See2 see2[23][32], dummy_see2;
  SqEntry sq[1024]; u32 sq_ptr;
  u32 sqp[256];

  ~Model() { free(heap); }

  // ---- raw heap access (structures are #pragma pack(1), :26) ----
  u8& B(u32 o) { return heap[o]; }
  u16 R16(u32 o) const { u16 v; memcpy(&v, heap + o, 2); return v; }
  u32 R32(u32 o) const { u32 v; memcpy(&v, heap + o, 4); return v; }
  void W16(u32 o, u16 v) { memcpy(heap + o, &v, 2); }
  void W32(u32 o, u32 v) { memcpy(heap + o, &v, 4); }
  // PPM_CONTEXT {u8 NumStats; u8 Flags; u16 SummFreq; u32 iStats; u32 iSuffix} (:425-433)
  u8& NumStats(u32 c) { return heap[c]; }
  u8& Flags(u32 c) { return heap[c + 1]; }
  u16 SummFreq(u32 c) const { return R16(c + 2); }
  void SetSummFreq(u32 c, u32 v) { W16(c + 2, (u16)v); }
  u32 Stats(u32 c) const { return R32(c + 4); }
  void SetStats(u32 c, u32 v) { W32(c + 4, v); }
  u32 Suffix(u32 c) const { return R32(c + 8); }
  void SetSuffix(u32 c, u32 v) { W32(c + 8, v); }
  static u32 OneState(u32 c) { return c + 2; }  // binary context: STATE overlays SummFreq/iStats
  // STATE {u8 Symbol; u8 Freq; u32 iSuccessor} (:406-410), 6 bytes
  u8& Sym(u32 s) { return heap[s]; }
  u8& Freq(u32 s) { return heap[s + 1]; }
  u32 Succ(u32 s) const { return R32(s + 2); }
  void SetSucc(u32 s, u32 v) { W32(s + 2, v); }
  void CopyState(u32 dst, u32 src) { memmove(heap + dst, heap + src, 6); }
  void SwapState(u32 a, u32 b) { u8 t[6]; memcpy(t, heap + a, 6); memcpy(heap + a, heap + b, 6); memcpy(heap + b, t, 6); }

  [[noreturn]] static void Exhausted(const char* where) {
    fprintf(stderr, "oracle_ppmd: heap exhausted in %s (path not restated)\n", where);
    abort();
  }

  // ---- free lists (:65-103) ----
  void ListInsert(int i, u32 blk, u32 nu) {
    W32(blk + 4, bl_next[i]); bl_next[i] = blk;
    W32(blk, 0xFFFFFFFFu); W32(blk + 8, nu);
    bl_stamp[i]++;
  }
  u32 ListRemove(int i) {
    u32 blk = bl_next[i];
    bl_next[i] = R32(blk + 4);
    bl_stamp[i]--;
    return blk;
  }
  void SplitBlock(u32 blk, int old_i, int new_i) {  // :197-208
    u32 udiff = indx2units[old_i] - indx2units[new_i];
    u32 p = blk + UNIT * indx2units[new_i];
    u32 i = units2indx[udiff - 1];
    if (indx2units[i] != udiff) {
      u32 k = indx2units[--i];
      ListInsert(i, p, k);
      p += UNIT * k;
      udiff -= k;
    }
    ListInsert(units2indx[udiff - 1], p, udiff);
  }
  u32 AllocUnits(u32 nu) {  // :230-238
    int i = units2indx[nu - 1];
    if (bl_next[i]) return ListRemove(i);
    u32 ret = lo_unit;
    lo_unit += UNIT * indx2units[i];
    if (lo_unit <= hi_unit) return ret;
    Exhausted("AllocUnits");
  }
  u32 AllocContext() {  // :240-243
    if (hi_unit != lo_unit) return hi_unit -= UNIT;
    Exhausted("AllocContext");
  }
  void FreeUnits(u32 p, u32 nu) { int i = units2indx[nu - 1]; ListInsert(i, p, indx2units[i]); }  // :245-248
  u32 ExpandUnits(u32 old, u32 old_nu) {  // :257-267
    int i0 = units2indx[old_nu - 1], i1 = units2indx[old_nu];
    if (i0 == i1) return old;
    u32 p = AllocUnits(old_nu + 1);
    memcpy(heap + p, heap + old, UNIT * old_nu);
    ListInsert(i0, old, old_nu);
    return p;
  }
  u32 ShrinkUnits(u32 old, u32 old_nu, u32 new_nu) {  // :269-282
    int i0 = units2indx[old_nu - 1], i1 = units2indx[new_nu - 1];
    if (i0 == i1) return old;
    if (bl_next[i1]) {
      u32 p = ListRemove(i1);
      memcpy(heap + p, heap + old, UNIT * new_nu);
      ListInsert(i0, old, indx2units[i0]);
      return p;
    }
    SplitBlock(old, i0, i1);
    return old;
  }

  void Init() {  // Init(20, 2000, 1, 0) :1302-1318 -> PPMD_STARTUP :375-400, StartModelRare :659-713
    heap_size = 2000ull << 20;
    heap = (u8*)calloc(heap_size, 1);
    if (!heap) { fprintf(stderr, "oracle_ppmd: cannot reserve heap\n"); abort(); }
    int i, k, m, step;
    for (i = 0, k = 1; i < 4; i++, k += 1) indx2units[i] = k;
    for (k++; i < 8; i++, k += 2) indx2units[i] = k;
    for (k++; i < 12; i++, k += 3) indx2units[i] = k;
    for (k++; i < N_INDEXES; i++, k += 4) indx2units[i] = k;
    for (k = 0, i = 0; k < 128; k++) { i += indx2units[i] < k + 1; units2indx[k] = i; }
    ns2bs[0] = 0; ns2bs[1] = 2; ns2bs[2] = 2;
    memset(ns2bs + 3, 4, 26); memset(ns2bs + 29, 6, 256 - 29);
    for (i = 0; i < 5; i++) qtable[i] = i;
    for (m = i = 5, k = step = 1; i < 260; i++) { qtable[i] = m; if (!--k) { k = ++step; m++; } }

    memset(char_mask, 0, sizeof(char_mask));
    esc_count = 1;
    order_fall = MAX_ORDER;
    memset(bl_stamp, 0, sizeof(bl_stamp)); memset(bl_next, 0, sizeof(bl_next));  // InitSubAllocator :134-140
    text_ptr = 0;
    hi_unit = (u32)heap_size;
    u64 diff = heap_size / 8 / UNIT * 7 * UNIT;
    lo_unit = units_start = hi_unit - (u32)diff;
    init_rl = -13; run_length = init_rl;
    max_context = AllocContext();
    NumStats(max_context) = 255;
    SetSummFreq(max_context, 257);
    SetStats(max_context, AllocUnits(128));
    Flags(max_context) = 0;
    SetSuffix(max_context, 0);
    prev_success = 0;
    u32 st = Stats(max_context);
    for (i = 0; i < 256; i++) { Sym(st + 6 * i) = i; Freq(st + 6 * i) = 1; SetSucc(st + 6 * i, 0); }
    static const signed char esc_coef[12] = {16, -10, 1, 51, 14, 89, 23, 35, 64, 26, -42, 43};  // :35-36
    u8 i2f[25];
    for (k = i = 0; i < 25; i2f[i++] = k + 1) while (qtable[k] == i) k++;
    for (k = 0; k < 64; k++) {
      int s = 0;
      for (i = 0; i < 6; i++) s += esc_coef[2 * i + ((k >> i) & 1)];
      s = s < 32 ? 32 : s > 224 ? 224 : s;
      s *= 128;
      for (i = 0; i < 25; i++) bin_summ[i][k] = BIN_SCALE - s / i2f[i];
    }
    for (i = 0; i < 23; i++) for (k = 0; k < 32; k++) {
      see2[i][k].shift = PERIOD_BITS - 4; see2[i][k].summ = (8 * i + 5) << (PERIOD_BITS - 4); see2[i][k].count = 7;
    }
    dummy_see2.summ = 0; dummy_see2.shift = 0; dummy_see2.count = 0;  // value-initialised object
    found_state = 0; bsumm = 0; num_masked = 0; sq_ptr = 0;
    memset(sqp, 0, sizeof(sqp));
  }

  void See2Update(See2& s) {  // :478-493
    if (--s.count == 0) {
      u32 i = s.summ >> s.shift;
      i = PERIOD_BITS - (i > 40) - (i > 280) - (i > 1020);
      if (i < s.shift) { s.summ >>= 1; s.shift--; }
      else if (i > s.shift) { s.summ <<= 1; s.shift++; }
      s.count = 5 << s.shift;
    }
  }

  u32 Rescale(u32 q, int of_in, u32 fs) {  // :498-566
    Flags(q) &= 0x14;
    u32 p1 = Stats(q);
    u8 tmp[6];
    memcpy(tmp, heap + fs, 6);
    u32 p;
    for (p = fs; p != p1; p -= 6) CopyState(p, p - 6);
    memcpy(heap + p1, tmp, 6);
    int of = (of_in != 0);
    int a, i;
    int f0 = Freq(p);
    int sf = SummFreq(q);
    int esc = sf - f0;
    Freq(p) = (f0 + of) >> 1;
    SetSummFreq(q, Freq(p));
    for (i = 0; i < NumStats(q); i++) {
      p += 6;
      a = Freq(p);
      esc -= a;
      a = (a + of) >> 1;
      Freq(p) = a;
      SetSummFreq(q, SummFreq(q) + a);
      if (a) Flags(q) |= 0x08 * (Sym(p) >= 0x40);
      if (a > Freq(p - 6)) {
        memcpy(tmp, heap + p, 6);
        u32 pp;
        for (pp = p; tmp[1] > Freq(pp - 6); pp -= 6) CopyState(pp, pp - 6);
        memcpy(heap + pp, tmp, 6);
      }
    }
    if (Freq(p) == 0) {
      for (i = 0; Freq(p) == 0; i++, p -= 6) {}
      esc += i;
      a = (NumStats(q) + 2) >> 1;
      NumStats(q) -= i;
      if (NumStats(q) == 0) {
        u32 st = Stats(q);
        memcpy(tmp, heap + st, 6);
        int nf = (2 * tmp[1] + esc - 1) / esc;
        tmp[1] = nf < MAX_FREQ / 3 ? nf : MAX_FREQ / 3;
        Flags(q) &= 0x18;
        FreeUnits(st, a);
        memcpy(heap + OneState(q), tmp, 6);
        return OneState(q);
      }
      SetStats(q, ShrinkUnits(Stats(q), a, (NumStats(q) + 2) >> 1));
    }
    SetSummFreq(q, SummFreq(q) + ((esc + 1) >> 1));
    if (of_in || (Flags(q) & 0x04) == 0) {
      a = (sf -= esc) - f0;
      u32 v = (u32)((f0 * (int)SummFreq(q) - sf * (int)Freq(Stats(q)) + a - 1) / a);
      a = v < 2u ? 2u : v > (MAX_FREQ / 2u - 18u) ? (MAX_FREQ / 2u - 18u) : v;
    } else {
      a = 2;
    }
    u32 st = Stats(q);
    Freq(st) += a;
    SetSummFreq(q, SummFreq(q) + a);
    Flags(q) |= 0x04;
    return st;
  }

  // :888-969. `p` = state of the coded symbol in suffix(pc) or 0.
  u32 CreateSuccessors(bool skip, u32 p, u32 pc) {
    u32 ps[64]; int n = 0;
    u8 sym = Sym(found_state);
    u32 up = Succ(found_state);
    bool no_loop = false;
    if (!skip) {
      ps[n++] = found_state;
      if (!Suffix(pc)) no_loop = true;
    }
    if (!no_loop) {
      bool first = true;
      if (p) { pc = Suffix(pc); } else first = false;
      for (;;) {
        if (!first) {
          pc = Suffix(pc);
          if (NumStats(pc)) {
            for (p = Stats(pc); Sym(p) != sym; p += 6) {}
            u8 t = 2 * (Freq(p) < MAX_FREQ - 1);
            Freq(p) += t;
            SetSummFreq(pc, SummFreq(pc) + t);
          } else {
            p = OneState(pc);
            Freq(p) += (!NumStats(Suffix(pc)) & (Freq(p) < 16));
          }
        }
        first = false;
        if (Succ(p) != up) { pc = Succ(p); break; }
        ps[n++] = p;
        if (!Suffix(pc)) break;
      }
    }
    if (n == 0) return pc;
    u8 ct[8];  // first 8 bytes of the new binary context: NumStats, Flags, Symbol, Freq, iSuccessor
    ct[0] = 0;
    ct[1] = 0x10 * (sym >= 0x40);
    sym = heap[up];
    u32 succ = up + 1;
    memcpy(ct + 4, &succ, 4);
    ct[2] = sym;
    ct[1] |= 0x08 * (sym >= 0x40);
    if (NumStats(pc)) {
      for (p = Stats(pc); Sym(p) != sym; p += 6) {}
      u32 cf = Freq(p) - 1;
      u32 s0 = SummFreq(pc) - NumStats(pc) - cf;
      cf = 1 + ((2 * cf < s0) ? (12 * cf > s0) : 2 + cf / s0);
      ct[3] = cf < 7 ? cf : 7;
    } else {
      ct[3] = Freq(OneState(pc));
    }
    do {
      u32 pc1 = AllocContext();
      memcpy(heap + pc1, ct, 8);
      SetSuffix(pc1, pc);
      pc = pc1;
      SetSucc(ps[--n], pc);
    } while (n);
    return pc;
  }

  u32 ReduceOrder(u32 p, u32 pc) {  // :971-1018
    u32 p1;
    u32 pc1 = pc;
    SetSucc(found_state, text_ptr);
    u8 sym = Sym(found_state);
    u32 up = Succ(found_state);
    order_fall++;
    bool first = p != 0;
    if (first) pc = Suffix(pc);
    for (;;) {
      if (!first) {
        if (!Suffix(pc)) return pc;
        pc = Suffix(pc);
        if (NumStats(pc)) {
          for (p = Stats(pc); Sym(p) != sym; p += 6) {}
          u8 t = 2 * (Freq(p) < MAX_FREQ - 3);
          Freq(p) += t;
          SetSummFreq(pc, SummFreq(pc) + t);
        } else {
          p = OneState(pc);
          Freq(p) += (Freq(p) < 11);
        }
      }
      first = false;
      if (Succ(p)) break;
      SetSucc(p, up);
      order_fall++;
    }
    if (Succ(p) <= up) {
      p1 = found_state;
      found_state = p;
      SetSucc(p, CreateSuccessors(false, 0, pc));
      found_state = p1;
    }
    if (order_fall == 1 && pc1 == max_context) {
      SetSucc(found_state, Succ(p));
      text_ptr--;
    }
    return Succ(p);
  }

  u32 UpdateModel(u32 minc) {  // :759-886 (returns new max context; never 0 here, see header)
    u8 fsym = Sym(found_state);
    u32 ffreq = Freq(found_state);
    u32 fsucc = Succ(found_state);
    u32 p = 0, pc;
    if (Suffix(minc)) {
      pc = Suffix(minc);
      if (NumStats(pc)) {
        p = Stats(pc);
        if (Sym(p) != fsym) {
          for (p += 6; Sym(p) != fsym; p += 6) {}
          if (Freq(p) >= Freq(p - 6)) { SwapState(p, p - 6); p -= 6; }
        }
        if (Freq(p) < MAX_FREQ - 3) {
          u32 cf = 2 + (ffreq < 28);
          Freq(p) += cf;
          SetSummFreq(pc, SummFreq(pc) + cf);
        }
      } else {
        p = OneState(pc);
        Freq(p) += (Freq(p) < 14);
      }
    }
    pc = max_context;
    if (!order_fall && fsucc) {
      SetSucc(found_state, CreateSuccessors(true, p, minc));
      max_context = Succ(found_state);
      return max_context;
    }
    heap[text_ptr++] = fsym;
    u32 succ = text_ptr;
    if (text_ptr >= units_start) Exhausted("UpdateModel(text)");
    if (fsucc) {
      if (fsucc < units_start) fsucc = CreateSuccessors(false, p, minc);
    } else {
      fsucc = ReduceOrder(p, minc);
    }
    if (!--order_fall) {
      succ = fsucc;
      text_ptr -= (max_context != minc);
    }
    u32 s0 = SummFreq(minc) - ffreq;
    u32 ns = NumStats(minc);
    u8 flag = 0x08 * (fsym >= 0x40);
    for (pc = max_context; pc != minc; pc = Suffix(pc)) {
      u32 ns1 = NumStats(pc);
      if (ns1) {
        if (ns1 & 1) SetStats(pc, ExpandUnits(Stats(pc), (ns1 + 1) >> 1));
        SetSummFreq(pc, SummFreq(pc) + (qtable[ns + 4] >> 3));
      } else {
        p = AllocUnits(1);
        CopyState(p, OneState(pc));
        SetStats(pc, p);
        Freq(p) = (Freq(p) <= MAX_FREQ / 3) ? (2 * Freq(p) - 1) : (MAX_FREQ - 15);
        static const u8 exp_escape[16] = {51, 43, 18, 12, 11, 9, 8, 7, 6, 5, 4, 3, 3, 2, 2, 2};  // :39-40
        SetSummFreq(pc, Freq(p) + (ns > 1) + exp_escape[qtable[bsumm >> 8]]);
      }
      u32 cf = (ffreq - 1) * (5 + SummFreq(pc));
      u32 sf = s0 + SummFreq(pc);
      if (cf <= 3 * sf) {
        cf = 1 + (2 * cf > sf) + (2 * cf > 3 * sf);
        SetSummFreq(pc, SummFreq(pc) + 4);
      } else {
        cf = 5 + (cf > 5 * sf) + (cf > 6 * sf) + (cf > 8 * sf) + (cf > 10 * sf) + (cf > 12 * sf);
        SetSummFreq(pc, SummFreq(pc) + cf);
      }
      p = Stats(pc) + 6 * (++NumStats(pc));
      SetSucc(p, succ);
      Sym(p) = fsym;
      Freq(p) = cf;
      Flags(pc) |= flag;
    }
    max_context = fsucc;
    return max_context;
  }

  u16& BinSummFor(u32 q) {  // shared by :1026-1028 and :1224-1226
    u32 rs = OneState(q);
    int i = ns2bs[NumStats(Suffix(q))] + prev_success + Flags(q) + ((run_length >> 26) & 0x20);
    return bin_summ[qtable[Freq(rs) - 1]][i];
  }
  See2* See2For(u32 q, int cnum, int& see_freq) {  // shared by :1116-1124 and :1270-1278
    if (cnum != 0xFF) {
      See2* s = see2[qtable[cnum + 3] - 4];
      s += (SummFreq(q) > 10 * (cnum + 1));
      s += 2 * (2 * cnum < NumStats(Suffix(q)) + num_masked) + Flags(q);
      see_freq = (s->summ >> s->shift) + 1;
      return s;
    }
    see_freq = 1;
    return &dummy_see2;
  }

  void UpdateByte(u32 c) {  // ppmd_UpdateByte :1351-1382
    u32 minc = max_context;
    if (NumStats(minc)) {  // processSymbol1<0> :1049-1098
      u32 p = Stats(minc);
      int cnum = NumStats(minc);
      prev_success = 0;
      if (Sym(p) == c) {
        Freq(p) += 4; SetSummFreq(minc, SummFreq(minc) + 4);
      } else {
        int i; bool hit = false;
        for (i = 1; i <= cnum; i++) if (Sym(p + 6 * i) == c) { hit = true; break; }
        if (hit) {
          Freq(p + 6 * i) += 4; SetSummFreq(minc, SummFreq(minc) + 4);
          if (Freq(p + 6 * i) > Freq(p + 6 * (i - 1))) { SwapState(p + 6 * i, p + 6 * (i - 1)); i--; }
          p = p + 6 * i;
        } else {
          num_masked = cnum;
          for (i = 0; i <= cnum; i++) char_mask[Sym(p + 6 * i)] = esc_count;
          p = 0;
        }
      }
      found_state = p;
      if (p && Freq(p) > MAX_FREQ) found_state = Rescale(minc, order_fall, found_state);
    } else {  // processBinSymbol<0> :1023-1046
      u32 rs = OneState(minc);
      u16& bs = BinSummFor(minc);
      bsumm = bs;
      bs -= (bsumm + 64) >> PERIOD_BITS;
      if (Sym(rs) != c) {
        char_mask[Sym(rs)] = esc_count; num_masked = 0; prev_success = 0; found_state = 0;
      } else {
        bs += INTERVAL; Freq(rs) += (Freq(rs) < 196); run_length++; prev_success = 1; found_state = rs;
      }
    }
    while (!found_state) {
      do { order_fall++; minc = Suffix(minc); } while (NumStats(minc) == num_masked);
      // processSymbol2<0> :1104-1170
      u32 p = Stats(minc);
      int cnum = NumStats(minc), see_freq;
      See2* see = See2For(minc, cnum, see_freq);
      int low = 0, hit_i = -1;
      for (int i = 0; i <= cnum; i++) {
        u32 s = Sym(p + 6 * i);
        if (char_mask[s] != esc_count) {
          char_mask[s] = esc_count;
          low += Freq(p + 6 * i);
          if (s == c) hit_i = i;
        }
      }
      int total = see_freq + low;
      if (hit_i >= 0) {
        p += 6 * hit_i;
        if (see_freq > 2) see->summ -= see_freq;
        See2Update(*see);
        found_state = p;
        Freq(p) += 4; SetSummFreq(minc, SummFreq(minc) + 4);
        if (Freq(p) > MAX_FREQ) found_state = Rescale(minc, order_fall, found_state);
        run_length = init_rl;
        esc_count++;
      } else {
        num_masked = cnum;
        see->summ += total - see_freq;
      }
    }
    if (order_fall != 0 || Succ(found_state) < units_start) UpdateModel(minc);
    else max_context = Succ(found_state);
  }

  void Store(u32 sym, u32 freq, u32 total) { sq[sq_ptr].sym = sym; sq[sq_ptr].freq = freq; sq[sq_ptr].total = total; sq_ptr++; }

  void PrepareByte() {  // ppmd_PrepareByte :1322-1349 with the *_T walkers :1222-1297
    sq_ptr = 0; num_masked = 0;
    int saved_order_fall = order_fall;
    u32 minc = max_context;
    if (NumStats(minc)) {  // processSymbol1_T
      u32 p = Stats(minc);
      int cnum = NumStats(minc), low = 0, total = SummFreq(minc);
      for (int i = 0; i <= cnum; i++) { int f = Freq(p + 6 * i); Store(Sym(p + 6 * i), f, total); low += f; }
      num_masked = cnum;
      for (int i = 0; i <= cnum; i++) char_mask[Sym(p + 6 * i)] = esc_count;
      Store(256, total - low, total);
    } else {  // processBinSymbol_T
      u32 rs = OneState(minc);
      bsumm = BinSummFor(minc);
      Store(Sym(rs), bsumm + bsumm, SCALE);
      Store(256, SCALE - bsumm - bsumm, SCALE);
      char_mask[Sym(rs)] = esc_count;
      num_masked = 0;
    }
    for (;;) {
      bool root = false;
      do {
        if (!Suffix(minc)) { root = true; break; }
        order_fall++;
        minc = Suffix(minc);
      } while (NumStats(minc) == num_masked);
      if (root) break;
      // processSymbol2_T
      u32 p = Stats(minc);
      int cnum = NumStats(minc), see_freq;
      See2For(minc, cnum, see_freq);
      int low = 0;
      for (int i = 0; i <= cnum; i++) if (char_mask[Sym(p + 6 * i)] != esc_count) low += Freq(p + 6 * i);
      int total = see_freq + low;
      for (int i = 0; i <= cnum; i++) {
        u32 s = Sym(p + 6 * i);
        if (char_mask[s] != esc_count) { Store(s, Freq(p + 6 * i), total); char_mask[s] = esc_count; }
      }
      Store(256, see_freq, total);
      num_masked = cnum;
    }
    esc_count++;
    num_masked = 0;
    order_fall = saved_order_fall;
    // ConvertSQ :1192-1209 (the trF/trT tree :1212-1219 is never read on the path)
    u32 cum = 0xFFFFFF00u;
    memset(sqp, 0, sizeof(sqp));
    for (u32 i = 0; i < sq_ptr; i++) {
      u32 prob = (u32)(((u64)cum * sq[i].freq) / sq[i].total);
      if (sq[i].sym < 256) sqp[sq[i].sym] = prob + 1; else cum = prob;
    }
  }
};

}  // namespace oracle_ppmd
#endif
