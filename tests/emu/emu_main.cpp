// TEST-ONLY: runs the product's CUDA stream kernel under the CPU SIMT emulator (cuda_emu.h) on one
// stream, so kernel logic can be checked against the oracle without a GPU.
//   emu_main compress|decompress <in> <out> [bit_trace_out] [pred_trace_out]
#include "cuda_emu.h"

#include <stdio.h>
#include <vector>

#include "../../gmix_b200/csrc/layout.h"

#ifndef EMU_NT
#define EMU_NT 128
#endif

static std::vector<uint8_t> ReadAll(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path); exit(2); }
  std::vector<uint8_t> v; int c;
  while ((c = fgetc(f)) != EOF) v.push_back((uint8_t)c);
  fclose(f);
  return v;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: emu_main compress|decompress <in> <out> [bit_trace] [pred_trace]\n"); return 2; }
  const bool comp = argv[1][0] == 'c';
  std::vector<uint8_t> in = ReadAll(argv[2]);
  uint64_t raw_len = in.size();
  if (!comp) { raw_len = 0; for (int i = 0; i < 5 && i < (int)in.size(); ++i) raw_len = (raw_len << 8) + in[i]; }
  const bool force_roomy = getenv("EMU_ROOMY") != nullptr;
  bool roomy = force_roomy;
retry:
  gmx::ArenaLayout L = gmx::MakeLayout(raw_len, roomy);
  std::vector<uint8_t> arena(L.total + 256);
  std::vector<float> decay, adam, linit;
  gmx::FillDecayTable(decay, raw_len * 8 + 16);
  gmx::FillAdamTable(adam);
  gmx::FillLstmInit(linit);
  std::vector<uint8_t> out(comp ? raw_len + raw_len / 8 + 64 : raw_len + 8);
  uint64_t in_off[2] = {0, in.size()}, out_off[2] = {0, out.size()}, out_len[1] = {0};
  uint32_t status[1] = {0};
  static uint32_t queue;
  queue = 0;
  std::vector<uint64_t> bit_trace(argc > 4 ? raw_len * 8 : 0);
  std::vector<float> pred_trace(argc > 5 ? raw_len * 8 * 126 : 0);
  gmx::StreamParams P;
  memset(&P, 0, sizeof(P));
  P.in = in.data(); P.in_off = in_off; P.out = out.data(); P.out_off = out_off; P.out_len = out_len; P.status = status;
  P.n_streams = 1; P.queue = &queue;
  P.arenas = (uint8_t*)(((uintptr_t)arena.data() + 255) & ~(uintptr_t)255); P.arena_stride = L.total; P.layout = &L;
  P.lstm_init = linit.data(); P.decay = decay.data(); P.decay_len = (uint32_t)decay.size(); P.adam = adam.data();
  uint32_t usage[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  P.usage = usage;
  P.bit_trace = bit_trace.empty() ? nullptr : bit_trace.data();
  P.pred_trace = pred_trace.empty() ? nullptr : pred_trace.data();
  cuda_emu::RunBlock(EMU_NT, 0, 1, [&] {
    if (comp) gmx::StreamKernel<EMU_NT, gmx::MODE_COMPRESS, 1, false>(P);
    else gmx::StreamKernel<EMU_NT, gmx::MODE_DECOMPRESS, 1, false>(P);
  });
  if (!roomy && (status[0] == gmx::GMX_ERR_PPMD_ARENA || status[0] == gmx::GMX_ERR_MIXER_POOL || status[0] == gmx::GMX_ERR_SPARSE_FULL)) {
    fprintf(stderr, "status %u: retrying in a roomy arena (as the host library does)\n", status[0]);
    roomy = true;
    goto retry;
  }
  if (status[0]) { fprintf(stderr, "stream status %u\n", status[0]); return 1; }
  fprintf(stderr, "arena %llu KiB; sparse %u of %u (cap %u); mixer sets %u of %u; ppmd units %u of %u B; history %u\n",
          (unsigned long long)(L.total >> 10), usage[0], L.sparse_limit, L.sparse_mask ? L.sparse_mask + 1 : 0, usage[1], L.mix_pool_sets,
          usage[2], L.p_units_cap, usage[3]);
  FILE* f = fopen(argv[3], "wb"); fwrite(out.data(), 1, out_len[0], f); fclose(f);
  if (argc > 4) {  // same record format as ref_driver trace level 1/2
    FILE* t = fopen(argv[4], "wb");
    for (size_t i = 0; i < bit_trace.size(); ++i) {
      fwrite(&bit_trace[i], 8, 1, t);
      if (argc > 5) fwrite(&pred_trace[i * 126], 4, 126, t);
    }
    fclose(t);
  }
  return 0;
}
