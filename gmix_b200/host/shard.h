// Static partition of independent streams over the GPUs of one box (the same rule as
// gmix_b200/shard.py, which the torch.distributed harness uses): rank r owns a contiguous stream range, ranges are
// balanced by input bytes (stream i goes to the rank whose byte interval contains the midpoint of stream i). Pure C++.
#ifndef GMIX_B200_HOST_SHARD_H_
#define GMIX_B200_HOST_SHARD_H_
#include <stdint.h>

#include <utility>
#include <vector>

namespace gmixb {

inline std::vector<std::pair<uint32_t, uint32_t>> ShardRanges(const std::vector<uint64_t>& lengths, int world) {
  const uint32_t n = (uint32_t)lengths.size();
  std::vector<std::pair<uint32_t, uint32_t>> out((size_t)world, {0u, 0u});
  if (n == 0) return out;
  uint64_t total = 0;
  for (uint64_t l : lengths) total += l;
  if (total == 0) {
    for (int r = 0; r < world; ++r) out[r] = {(uint32_t)((uint64_t)n * r / world), (uint32_t)((uint64_t)n * (r + 1) / world)};
    return out;
  }
  std::vector<uint32_t> count((size_t)world, 0);
  uint64_t cum = 0;
  for (uint32_t i = 0; i < n; ++i) {
    cum += lengths[i];
    const double mid = (double)cum - (double)lengths[i] / 2.0;
    int64_t owner = (int64_t)(mid * (double)world / (double)total);
    if (owner > world - 1) owner = world - 1;
    count[(size_t)owner]++;
  }
  uint32_t lo = 0;
  for (int r = 0; r < world; ++r) { out[r] = {lo, lo + count[r]}; lo += count[r]; }
  return out;
}

// FNV-1a 64, the function of the device ChecksumKernel (gmix_b200/csrc/host.cu).
inline uint64_t Fnv1a64(const uint8_t* p, uint64_t n) {
  uint64_t h = 0xcbf29ce484222325ull;
  for (uint64_t k = 0; k < n; ++k) h = (h ^ p[k]) * 0x100000001b3ull;
  return h;
}

}  // namespace gmixb
#endif
