// TEST-ONLY: host check of the PPMd heap-window geometry MakeLayout picks (gmix_b200/csrc/layout.h).
// For every stream length / arena class: the text area [0, text_cap), the low unit area growing up from units_start and
// the high unit area growing down from heap_end must not alias inside the 2^k window for ANY split x + y <= units_cap.
#include "../emu/cuda_emu.h"

#include <stdio.h>

#include "../../gmix_b200/csrc/layout.h"

static bool Disjoint(uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1) { return a1 <= b0 || b1 <= a0; }

// window images of the virtual range [v0, v1) as at most two pieces
static int Pieces(uint64_t v0, uint64_t v1, uint64_t P, uint64_t out[2][2]) {
  if (v1 <= v0) return 0;
  if (v1 - v0 >= P) { out[0][0] = 0; out[0][1] = P; return 1; }
  const uint64_t a = v0 % P, b = a + (v1 - v0);
  if (b <= P) { out[0][0] = a; out[0][1] = b; return 1; }
  out[0][0] = a; out[0][1] = P; out[1][0] = 0; out[1][1] = b - P;
  return 2;
}
static bool RangesDisjoint(uint64_t a0, uint64_t a1, uint64_t b0, uint64_t b1, uint64_t P) {
  uint64_t pa[2][2], pb[2][2];
  const int na = Pieces(a0, a1, P, pa), nb = Pieces(b0, b1, P, pb);
  for (int i = 0; i < na; ++i) for (int j = 0; j < nb; ++j) if (!Disjoint(pa[i][0], pa[i][1], pb[j][0], pb[j][1])) return false;
  return true;
}

int main() {
  const uint64_t lens[] = {0, 1, 600, 4096, 15000, 16384, 65536, 66000, 70000, 100000, 262144, 411996, 560000, 1000000, 2500000,
                           3000000, 8000000, 16000000, 64000000, 100000000, 300000000};
  int bad = 0, windows_above_16m = 0;
  for (uint64_t len : lens)
    for (int roomy = 0; roomy < 2; ++roomy)
      for (int pre_units = 0; pre_units < 3; ++pre_units) {
        gmx::Preload pre;
        pre.ppmd_unit_bytes = pre_units == 0 ? 0 : pre_units == 1 ? (220ull << 20) : (730ull << 20);
        const gmx::ArenaLayout L = gmx::MakeLayout(len, roomy != 0, pre_units ? &pre : nullptr);
        const uint64_t P = (uint64_t)L.p_mask + 1;
        if (P > (16ull << 20)) ++windows_above_16m;
        const uint64_t cap = L.p_units_cap, text = L.p_text_cap;
        // extreme splits of the unit budget between the two areas, and the even one
        const uint64_t xs[3] = {cap, 0, cap / 2};
        for (uint64_t x : xs) {
          const uint64_t y = cap - x;
          const uint64_t lo0 = gmx::PPMD_UNITS_START, lo1 = lo0 + x, hi1 = gmx::PPMD_HEAP_END, hi0 = hi1 - y;
          const bool ok = RangesDisjoint(0, text, lo0, lo1, P) && RangesDisjoint(0, text, hi0, hi1, P) && RangesDisjoint(lo0, lo1, hi0, hi1, P);
          if (!ok) {
            printf("ALIAS len=%llu roomy=%d pre=%d P=%llu text_cap=%llu units_cap=%llu x=%llu\n", (unsigned long long)len, roomy, pre_units,
                   (unsigned long long)P, (unsigned long long)text, (unsigned long long)cap, (unsigned long long)x);
            ++bad;
          }
        }
        if (cap % 12 != 0) { printf("units_cap not a unit multiple\n"); ++bad; }
      }
  printf("windows above 16 MiB checked: %d\nbad %d\n", windows_above_16m, bad);
  return bad != 0;
}
