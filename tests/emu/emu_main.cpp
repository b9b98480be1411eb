// TEST-ONLY: runs the product's CUDA stream kernel under the CPU SIMT emulator (cuda_emu.h) on one
// stream, so kernel logic can be checked against the oracle without a GPU.
//   emu_main compress|decompress <in> <out> [bit_trace_out] [pred_trace_out]
//   emu_main train    <in> <ckpt_prefix> [from_ckpt_prefix]   Predict/Perceive/Learn over <in> (analysis off), then the
//                                                             stream is written as <ckpt_prefix>.short/.long (checkpoint.h)
//   emu_main resume   <ckpt_prefix> <in> <out> [bit_trace_out]   `gmix -c <ckpt> <in> <out>`: compress starting from a checkpoint
//   emu_main expand   <ckpt_prefix> <in> <out>                   `gmix -d <ckpt> <in> <out>`
//   emu_main generate <ckpt_prefix> <prompt> <out> <size> <temperature>   `gmix -g ...` (sampling draws from this host's rand())
//   emu_main recode   <ckpt_prefix> <out_prefix>                 Parse then Serialize (must reproduce the files byte for byte)
//   emu_main parts    <in> <out> <split>                         compress <in> as ONE stream in two parts: bytes [0, split) with the
//                                                                header and no flush, a predictor + coder checkpoint (through the
//                                                                reference's file format), then the rest from that checkpoint
//   emu_main unparts  <in.gmix> <out> <split>                    decompress likewise: `split` bytes, checkpoint, the rest
//   emu_main steps    <in> <out>                                 compress through the Predictor facade's StepKernel (one launch per
//                                                                Predict / Learn, host coder), analysis as `gmix -c` sets it
#include "cuda_emu.h"

#include <stdio.h>
#include <algorithm>
#include <string>
#include <vector>

#include "../../gmix_b200/csrc/checkpoint.h"
#include "../../gmix_b200/csrc/gate_gemm.cuh"
#include "../../gmix_b200/host/coder.h"

// role split of the emulated CTA (warps of bit role / LSTM role; + one PPMd warp)
#ifndef EMU_WB
#define EMU_WB 2
#endif
#ifndef EMU_WL
#define EMU_WL 1
#endif
#ifndef EMU_SERIAL
#define EMU_SERIAL 0      // 1: compress without the role pipeline (kernels.h configuration 0)
#endif
#ifndef EMU_MINB
#define EMU_MINB 8        // resident CTAs per SM the configuration is built for; 1 selects the latency code paths (LAT)
#endif
#ifndef EMU_WS
#define EMU_WS 0          // 1: dense gate weights resident in (emulated) shared memory, refreshed by bulk copies after Adam
#endif
#define EMU_NT (32 * (EMU_WB + EMU_WL + 1))
namespace gmx { constexpr int kStepWB = 2, kStepWL = 1; }   // as kernels.h (which needs the CUDA runtime)

static std::vector<uint8_t> ReadAll(const std::string& path) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { fprintf(stderr, "cannot open %s\n", path.c_str()); exit(2); }
  std::vector<uint8_t> v;
  uint8_t buf[1 << 16];
  size_t n;
  while ((n = fread(buf, 1, sizeof(buf), f)) > 0) v.insert(v.end(), buf, buf + n);
  fclose(f);
  return v;
}
static void WriteAll(const std::string& path, const void* p, size_t n) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) { fprintf(stderr, "cannot write %s\n", path.c_str()); exit(2); }
  fwrite(p, 1, n, f);
  fclose(f);
}

struct Run {
  gmx::ArenaLayout L;
  std::vector<uint8_t> arena, tmpl_arena;
  std::vector<uint32_t> tmpl_state, final_state;
  std::vector<float> decay, adam, linit;
  uint32_t usage[8];
  uint32_t status;
};

// mode: gmx::MODE_*; ckpt: optional parsed checkpoint to start from; new_bytes: bytes the stream will add.
// overlay_learn > 0: the stream runs in overlay mode on top of the checkpoint (what gmx_generate_batch does): the checkpoint's
// arena stays read-only, the stream arena is a MakeOverlayLayout sized for overlay_learn learned bytes.
static int Execute(Run& R, int mode, const gmx::ckpt::Image* ckpt, uint64_t new_bytes, gmx::StreamParams& P, bool want_final, uint64_t overlay_learn = 0) {
  const bool force_roomy = getenv("EMU_ROOMY") != nullptr;
  bool roomy = force_roomy;
  gmx::Preload pre;
  if (ckpt) pre = gmx::ckpt::Count(*ckpt);
retry:
  R.L = gmx::MakeLayout(new_bytes, roomy, ckpt ? &pre : nullptr, getenv("EMU_DENSE") != nullptr);
  R.arena.assign(R.L.total + 256, 0);
  gmx::FillDecayTable(R.decay, pre.steps + new_bytes * 8 + 16);
  gmx::FillAdamTable(R.adam);
  gmx::FillLstmInit(R.linit);
  static uint32_t queue;
  queue = 0;
  P.n_streams = 1; P.queue = &queue;
  P.arenas = (uint8_t*)(((uintptr_t)R.arena.data() + 255) & ~(uintptr_t)255); P.arena_stride = R.L.total; P.layout = &R.L;
  P.lstm_init = R.linit.data(); P.decay = R.decay.data(); P.decay_len = (uint32_t)R.decay.size(); P.adam = R.adam.data();
  memset(R.usage, 0, sizeof(R.usage));
  P.usage = R.usage;
  R.status = 0;
  P.status = &R.status;
  if (ckpt) {
    R.tmpl_arena.assign(R.L.total + 256, 0);
    R.tmpl_state.assign(sizeof(gmx::StreamSmem) / 4 + 4, 0);
    uint8_t* ta = (uint8_t*)(((uintptr_t)R.tmpl_arena.data() + 255) & ~(uintptr_t)255);
    std::string err;
    if (!gmx::ckpt::ToArena(*ckpt, R.L, ta, (gmx::StreamSmem*)R.tmpl_state.data(), &err)) {
      if (!roomy) { fprintf(stderr, "%s: retrying with a roomy layout\n", err.c_str()); roomy = true; goto retry; }
      fprintf(stderr, "ToArena: %s\n", err.c_str());
      return 1;
    }
    P.tmpl_arena = ta; P.tmpl_state = R.tmpl_state.data();
    if (overlay_learn) {
      static gmx::ArenaLayout base;
      base = R.L;
      R.L = gmx::MakeOverlayLayout(base, pre, overlay_learn, new_bytes, getenv("EMU_SEGMENTED") != nullptr);
      R.arena.assign(R.L.total + 256, 0xCD);   // nothing in the overlay arena may be assumed zero
      P.arenas = (uint8_t*)(((uintptr_t)R.arena.data() + 255) & ~(uintptr_t)255); P.arena_stride = R.L.total; P.layout = &R.L;
      P.tmpl_layout = &base;
      fprintf(stderr, "overlay arena %llu KiB on top of a %llu KiB model\n", (unsigned long long)(R.L.total >> 10), (unsigned long long)(base.total >> 10));
    }
  }
  if (want_final) { R.final_state.assign(sizeof(gmx::StreamSmem) / 4 + 4, 0); P.final_state = R.final_state.data(); }
  cuda_emu::RunBlock(EMU_NT, 0, 1, [&] {
    if (mode == gmx::MODE_COMPRESS) gmx::StreamKernel<EMU_WB, EMU_WL, gmx::MODE_COMPRESS, EMU_MINB, false, EMU_SERIAL != 0, EMU_WS != 0>(P);
    else if (mode == gmx::MODE_DECOMPRESS) gmx::StreamKernel<EMU_WB, EMU_WL, gmx::MODE_DECOMPRESS, EMU_MINB, false, false, EMU_WS != 0>(P);
    else gmx::StreamKernel<EMU_WB, EMU_WL, gmx::MODE_GENERATE, EMU_MINB, false, false, EMU_WS != 0>(P);
  });
  if (!roomy && (R.status == gmx::GMX_ERR_PPMD_ARENA || R.status == gmx::GMX_ERR_MIXER_POOL || R.status == gmx::GMX_ERR_SPARSE_FULL)) {
    fprintf(stderr, "status %u: retrying in a roomy arena (as the host library does)\n", R.status);
    roomy = true;
    goto retry;
  }
  if (R.status) { fprintf(stderr, "stream status %u\n", R.status); return 1; }
  fprintf(stderr, "arena %llu KiB; sparse %u of %u (cap %u); mixer sets %u of %u; ppmd units %u of %u B; history %u\n",
          (unsigned long long)(R.L.total >> 10), R.usage[0], R.L.sparse_limit, R.L.sparse_mask ? R.L.sparse_mask + 1 : 0, R.usage[1], R.L.mix_pool_sets,
          R.usage[2], R.L.p_units_cap, R.usage[3]);
  return 0;
}

static bool LoadCkpt(const std::string& prefix, gmx::ckpt::Image* im) {
  std::vector<uint8_t> sh = ReadAll(prefix + ".short"), lo = ReadAll(prefix + ".long");
  std::string err;
  if (!gmx::ckpt::Parse(sh.data(), sh.size(), lo.data(), lo.size(), im, &err)) { fprintf(stderr, "Parse: %s\n", err.c_str()); return false; }
  return true;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: see the header of tests/emu/emu_main.cpp\n"); return 2; }
  const std::string mode = argv[1];
  Run R;
  gmx::StreamParams P;
  memset(&P, 0, sizeof(P));
  P.analysis = -1;
  uint64_t in_off[2], out_off[2], out_len[1] = {0};
  P.in_off = in_off; P.out_off = out_off; P.out_len = out_len;

  if (mode == "recode") {
    gmx::ckpt::Image im;
    if (!LoadCkpt(argv[2], &im)) return 1;
    std::vector<uint8_t> sh, lo;
    gmx::ckpt::Serialize(im, &sh, &lo);
    WriteAll(std::string(argv[3]) + ".short", sh.data(), sh.size());
    WriteAll(std::string(argv[3]) + ".long", lo.data(), lo.size());
    return 0;
  }
  if (mode == "train") {
    std::vector<uint8_t> in = ReadAll(argv[2]);
    gmx::ckpt::Image from;
    const bool has_from = argc > 4;
    if (has_from && !LoadCkpt(argv[4], &from)) return 1;
    std::vector<uint8_t> out(in.size() + in.size() / 8 + 64);
    in_off[0] = 0; in_off[1] = in.size(); out_off[0] = 0; out_off[1] = out.size();
    P.in = in.data(); P.out = out.data(); P.analysis = 0;
    if (Execute(R, gmx::MODE_COMPRESS, has_from ? &from : nullptr, in.size(), P, true)) return 1;
    gmx::ckpt::Image im;
    std::string err;
    if (!gmx::ckpt::FromArena(R.L, P.arenas, *(const gmx::StreamSmem*)R.final_state.data(), &im, &err)) { fprintf(stderr, "FromArena: %s\n", err.c_str()); return 1; }
    std::vector<uint8_t> sh, lo;
    gmx::ckpt::Serialize(im, &sh, &lo);
    WriteAll(std::string(argv[3]) + ".short", sh.data(), sh.size());
    WriteAll(std::string(argv[3]) + ".long", lo.data(), lo.size());
    return 0;
  }
  if (mode == "generate") {
    if (argc < 7) return 2;
    gmx::ckpt::Image im;
    if (!LoadCkpt(argv[2], &im)) return 1;
    std::vector<uint8_t> prompt = ReadAll(argv[3]);
    if (prompt.empty()) { fprintf(stderr, "empty prompt\n"); return 2; }
    const uint32_t size = (uint32_t)atoi(argv[5]);
    float temperature = (float)atof(argv[6]);
    if (temperature < 0.001) temperature = 0.001;   // runner-utils.cpp:170
    // what `gmix -g` draws: srand(0xDEADBEEF) by the Predictor constructor, 3*50*563 draws by the LSTM init, then sampling
    std::vector<float> unused;
    gmx::FillLstmInit(unused);
    std::vector<float> ru((size_t)size * 8);
    for (auto& r : ru) r = static_cast<float>(rand()) / static_cast<float>(RAND_MAX);
    std::vector<uint8_t> out(size + 8);
    in_off[0] = 0; in_off[1] = prompt.size(); out_off[0] = 0; out_off[1] = out.size();
    P.in = prompt.data(); P.out = out.data(); P.gen_bytes = size; P.temperature = temperature; P.rand_u = ru.data(); P.rand_stride = 0;
    if (getenv("EMU_LOCKSTEP")) {
      // what host.cu RunLockstepGenerate does: the prompt launch stops in front of the first gate product, then one exact batched
      // gate product + one GenStepKernel launch per sampled byte (one stream = one slot here)
      std::vector<uint32_t> park(sizeof(gmx::StreamSmem) / 4 + 4, 0), gsym(64, 0);
      std::vector<float> gx(gmx::GateXFloats(1), 0.0f), gg(gmx::GG_N * 2, 0.0f);
      P.lockstep = 1; P.stream_base = 0;
      P.gate_x = gx.data(); P.gate_sym = gsym.data(); P.gate_g = gg.data(); P.park = park.data();
      if (Execute(R, gmx::MODE_GENERATE, &im, prompt.size() + size, P, true, getenv("EMU_OVERLAY") ? prompt.size() : 0)) return 1;
      memcpy(park.data(), R.final_state.data(), sizeof(gmx::StreamSmem));
      const float* W = (const float*)(P.tmpl_arena + (P.tmpl_layout ? P.tmpl_layout->l_w : R.L.l_w));
      gmx::GenStepParams Q;
      Q.P = P; Q.n_slots = 1;
      for (uint32_t i = 0; i < size; ++i) {
        cuda_emu::RunBlock(gmx::GX_THREADS, 0, 1, [&] { gmx::GateDotsExactKernel(W, gx.data(), gsym.data(), gg.data(), 1); });
        Q.byte_index = i; Q.last = i + 1 == size;
        cuda_emu::RunBlock(EMU_NT, 0, 1, [&] { gmx::GenStepKernel<EMU_WB, EMU_WL, EMU_MINB>(Q); });
        if (R.status) { fprintf(stderr, "stream status %u at byte %u\n", R.status, i); return 1; }
      }
      WriteAll(argv[4], out.data(), size);
      return 0;
    }
    if (Execute(R, gmx::MODE_GENERATE, &im, prompt.size() + size, P, false, getenv("EMU_OVERLAY") ? prompt.size() : 0)) return 1;
    WriteAll(argv[4], out.data(), size);
    return 0;
  }

  if (mode == "parts" || mode == "unparts") {
    const bool comp = mode == "parts";
    std::vector<uint8_t> in = ReadAll(argv[2]);
    uint64_t total = in.size();
    if (!comp) { total = 0; for (int i = 0; i < 5; ++i) total = (total << 8) + in[i]; }
    const uint64_t split = strtoull(argv[4], nullptr, 10);
    if (split > total) return 2;
    std::vector<uint8_t> out(total + total / 8 + 64), result;
    uint32_t coder[8] = {0};
    // part 1 (from scratch)
    in_off[0] = 0; in_off[1] = comp ? split : in.size(); out_off[0] = 0; out_off[1] = out.size();
    P.in = in.data(); P.out = out.data(); P.analysis = 0;
    P.part = 1; P.part_header = 1; P.part_last = 0; P.part_total = comp ? total : split; P.coder_in = nullptr; P.coder_out = coder;
    if (Execute(R, comp ? gmx::MODE_COMPRESS : gmx::MODE_DECOMPRESS, nullptr, total, P, true)) return 1;
    result.assign(out.begin(), out.begin() + out_len[0]);
    const uint64_t consumed = coder[3];
    gmx::ckpt::Image im, im2;
    std::string err;
    if (!gmx::ckpt::FromArena(R.L, P.arenas, *(const gmx::StreamSmem*)R.final_state.data(), &im, &err)) { fprintf(stderr, "FromArena: %s\n", err.c_str()); return 1; }
    std::vector<uint8_t> sh, lo;
    gmx::ckpt::Serialize(im, &sh, &lo);
    if (!gmx::ckpt::Parse(sh.data(), sh.size(), lo.data(), lo.size(), &im2, &err)) { fprintf(stderr, "Parse: %s\n", err.c_str()); return 1; }
    // part 2 (from the checkpoint, coder state from part 1)
    Run R2;
    gmx::StreamParams Q;
    memset(&Q, 0, sizeof(Q));
    uint64_t io2[2], oo2[2] = {0, out.size()}, ol2[1] = {0};
    uint32_t coder_in[4] = {coder[0], coder[1], coder[2], 0}, coder2[8] = {0};
    Q.in_off = io2; Q.out_off = oo2; Q.out_len = ol2; Q.out = out.data(); Q.analysis = 0;
    if (comp) { Q.in = in.data() + split; io2[0] = 0; io2[1] = total - split; }
    else { Q.in = in.data() + consumed; io2[0] = 0; io2[1] = in.size() - consumed; }
    Q.part = 1; Q.part_header = 0; Q.part_last = 1; Q.part_total = comp ? total : total - split; Q.coder_in = coder_in; Q.coder_out = coder2;
    if (Execute(R2, comp ? gmx::MODE_COMPRESS : gmx::MODE_DECOMPRESS, &im2, total - split, Q, false)) return 1;
    result.insert(result.end(), out.begin(), out.begin() + ol2[0]);
    WriteAll(argv[3], result.data(), result.size());
    return 0;
  }
  if (mode == "steps") {   // what gmix_b200/host/predictor.h does over gmx_pred_*: STEP_INIT, then Predict / (Perceive) Learn per bit
    std::vector<uint8_t> in = ReadAll(argv[2]);
    Run R2;
    R2.L = gmx::MakeLayout(in.size() + 1, true);
    R2.arena.assign(R2.L.total + 256, 0);
    gmx::FillDecayTable(R2.decay, in.size() * 8 + 16);
    gmx::FillAdamTable(R2.adam);
    gmx::FillLstmInit(R2.linit);
    std::vector<uint32_t> state(sizeof(gmx::StreamSmem) / 4 + 4, 0);
    float prob = 0; uint32_t status = 0;
    gmx::StepParams Q;
    memset(&Q, 0, sizeof(Q));
    Q.P.arenas = (uint8_t*)(((uintptr_t)R2.arena.data() + 255) & ~(uintptr_t)255); Q.P.arena_stride = R2.L.total; Q.P.layout = &R2.L;
    Q.P.lstm_init = R2.linit.data(); Q.P.decay = R2.decay.data(); Q.P.decay_len = (uint32_t)R2.decay.size(); Q.P.adam = R2.adam.data();
    Q.state = state.data(); Q.prob_out = &prob; Q.status_out = &status;
    const int analysis = (8 * in.size() / 1000) > 0;
    constexpr int NTS = 32 * (gmx::kStepWB + gmx::kStepWL + 1);
    auto step = [&](int op, int has_bit, int bit) {
      Q.op = op; Q.has_bit = has_bit; Q.bit = bit; Q.analysis = analysis;
      cuda_emu::RunBlock(NTS, 0, 1, [&] { gmx::StepKernel<gmx::kStepWB, gmx::kStepWL>(Q); });
      if (status) { fprintf(stderr, "step status %u\n", status); exit(1); }
    };
    step(gmx::STEP_INIT, 0, 0);
    std::vector<uint8_t> out;
    for (int i = 4; i >= 0; --i) out.push_back((uint8_t)((uint64_t)in.size() >> (8 * i)));
    gmixb::Encoder enc(&out);
    for (size_t pos = 0; pos < in.size(); ++pos)
      for (int j = 7; j >= 0; --j) {
        const int bit = (in[pos] >> j) & 1;
        step(gmx::STEP_PREDICT, 0, 0);
        enc.Encode(bit, prob);
        step(gmx::STEP_LEARN, 1, bit);
      }
    enc.Flush();
    WriteAll(argv[3], out.data(), out.size());
    return 0;
  }
  if (mode == "compress2") {   // two streams through ONE persistent CTA and one arena: <in1> <in2> <out2> (what a batch larger than the arena count does)
    std::vector<uint8_t> a = ReadAll(argv[2]), b = ReadAll(argv[3]);
    std::vector<uint8_t> in(a); in.insert(in.end(), b.begin(), b.end());
    const uint64_t cap = std::max(a.size(), b.size()) * 9 / 8 + 64;
    std::vector<uint8_t> out(2 * cap);
    uint64_t io[3] = {0, a.size(), a.size() + b.size()}, oo[3] = {0, cap, 2 * cap}, ol[2] = {0, 0};
    uint32_t st2[2] = {0, 0};
    P.in = in.data(); P.out = out.data(); P.in_off = io; P.out_off = oo; P.out_len = ol;
    Run R2;
    gmx::StreamParams Q = P;
    // Execute() sets n_streams = 1 and its own status word: run it by hand for two streams
    const uint64_t layout_len = getenv("EMU_LAYOUT_LEN") ? strtoull(getenv("EMU_LAYOUT_LEN"), nullptr, 10) : std::max(a.size(), b.size());
    R2.L = gmx::MakeLayout(layout_len, getenv("EMU_ROOMY") != nullptr);   // EMU_LAYOUT_LEN: arenas configured for longer streams than these
    R2.arena.assign(R2.L.total + 256, 0);
    gmx::FillDecayTable(R2.decay, std::max(a.size(), b.size()) * 8 + 16);
    gmx::FillAdamTable(R2.adam);
    gmx::FillLstmInit(R2.linit);
    static uint32_t queue2;
    queue2 = 0;
    Q.n_streams = 2; Q.queue = &queue2; Q.status = st2;
    Q.arenas = (uint8_t*)(((uintptr_t)R2.arena.data() + 255) & ~(uintptr_t)255); Q.arena_stride = R2.L.total; Q.layout = &R2.L;
    Q.lstm_init = R2.linit.data(); Q.decay = R2.decay.data(); Q.decay_len = (uint32_t)R2.decay.size(); Q.adam = R2.adam.data();
    cuda_emu::RunBlock(EMU_NT, 0, 1, [&] { gmx::StreamKernel<EMU_WB, EMU_WL, gmx::MODE_COMPRESS, EMU_MINB, false, EMU_SERIAL != 0, EMU_WS != 0>(Q); });
    if (st2[0] || st2[1]) { fprintf(stderr, "status %u %u\n", st2[0], st2[1]); return 1; }
    WriteAll(argv[4], out.data() + cap, ol[1]);
    return 0;
  }
  const bool resume = mode == "resume" || mode == "expand";
  gmx::ckpt::Image im;
  if (resume) { if (!LoadCkpt(argv[2], &im)) return 1; argv += 1; argc -= 1; }
  const bool comp = mode == "compress" || mode == "resume";
  std::vector<uint8_t> in = ReadAll(argv[2]);
  uint64_t raw_len = in.size();
  if (!comp) { raw_len = 0; for (int i = 0; i < 5 && i < (int)in.size(); ++i) raw_len = (raw_len << 8) + in[i]; }
  std::vector<uint8_t> out(comp ? raw_len + raw_len / 8 + 64 : raw_len + 8);
  in_off[0] = 0; in_off[1] = in.size(); out_off[0] = 0; out_off[1] = out.size();
  std::vector<uint64_t> bit_trace(argc > 4 ? raw_len * 8 : 0);
  std::vector<float> pred_trace(argc > 5 ? raw_len * 8 * 126 : 0);
  P.in = in.data(); P.out = out.data();
  P.bit_trace = bit_trace.empty() ? nullptr : bit_trace.data();
  P.pred_trace = pred_trace.empty() ? nullptr : pred_trace.data();
  if (Execute(R, comp ? gmx::MODE_COMPRESS : gmx::MODE_DECOMPRESS, resume ? &im : nullptr, raw_len, P, false)) return 1;
  WriteAll(argv[3], out.data(), out_len[0]);
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
