// Batched LSTM gate product of lock-step generation: G[stream][gate row] = W_onehot[row][byte in front] + sum_k x[stream][k] * W[row][k]
// for all streams of a wave at once (reference: LstmLayer::ForwardPass lstm-layer.cpp:198-204 computes it per stream).
// While no stream has reached a BPTT pass the 3 x 50 x 563 gate matrix is the loaded model's for every stream (lstm.cpp:57-79),
// so the per-stream dot products are one dense contraction [streams x 307] . [307 x 150].
//
//   GateDotsExactKernel   the reference's arithmetic: one thread per gate row accumulates in the reference's column order with
//                         separately rounded fp32 multiplies and adds - bit-identical to LstmGateDots (stream_kernel.cuh). The
//                         dense weights (184.8 KB) are loaded into shared memory ONCE per CTA instead of once per stream and byte.
//   GateGemmTcKernel      opt-in: the same contraction on the 5th-generation tensor cores. tcgen05.mma kind::tf32 with M = 128
//                         streams, N = 160, K = 8 per instruction, accumulators in TMEM; both operands arrive as TMA bulk copies
//                         (cp.async.bulk + mbarrier) of pre-tiled K-major core-matrix images, two stages deep. fp32 accuracy is
//                         recovered with the 3xTF32 split (x = hi + lo, w = hi + lo; hi.hi + hi.lo + lo.hi accumulate in fp32):
//                         relative error ~2^-21 per product, but NOT the reference's summation order - sampled bytes can differ
//                         from `gmix -g`, which is why it is opt-in and its divergence is measured (bench.py, tests).
//   GateWeightPrepKernel  the model's dense gate columns -> the tiled hi / lo planes GateGemmTcKernel reads.
// Layouts: GateXIndex / GateWIndex (stream_kernel.cuh).
#ifndef GMIX_B200_GATE_GEMM_CUH_
#define GMIX_B200_GATE_GEMM_CUH_
#include "stream_kernel.cuh"

namespace gmx {

enum : int { GX_SLOTS = 8, GX_THREADS = 160 };   // exact kernel: slots per CTA pass, threads (150 gate rows + 10 helpers)
enum : int { GX_SMEM_BYTES = W_DENSE_BYTES + GX_SLOTS * GG_K * 4 };

// ---- exact ---------------------------------------------------------------------------------------------------------------
// grid: any; CTA c handles slot groups c, c + gridDim.x, ... of GX_SLOTS streams. W = the model's gate weights (arena layout).
__global__ void __launch_bounds__(GX_THREADS) GateDotsExactKernel(const float* W, const float* X, const uint32_t* sym, float* G, uint32_t n_slots) {
#if defined(__CUDACC__)
  extern __shared__ __align__(128) unsigned char gx_smem[];
#else
  static unsigned char gx_smem[GX_SMEM_BYTES] __attribute__((aligned(128)));
#endif
  float4* wd = (float4*)gx_smem;                          // [3][W_DENSE_Q][L_CELLS] float4, as in the arena
  float* xs = (float*)(gx_smem + W_DENSE_BYTES);          // [GX_SLOTS][GG_K]
  const int tid = (int)threadIdx.x;
  for (int g = 0; g < 3; ++g) {
    const float4* src = (const float4*)(W + LstmW(g, L_NOUT, 0));
    for (int i = tid; i < W_DENSE_Q * L_CELLS; i += GX_THREADS) wd[g * W_DENSE_Q * L_CELLS + i] = src[i];
  }
  const int g = tid / L_CELLS, i = tid - g * L_CELLS;
  for (uint32_t base = blockIdx.x * GX_SLOTS; base < n_slots; base += gridDim.x * GX_SLOTS) {
    __syncthreads();
    for (int t = tid; t < GX_SLOTS * GG_K; t += GX_THREADS) {
      const uint32_t slot = base + t / GG_K;
      const int k = t % GG_K;
      xs[t] = slot < n_slots ? f_add(X[GateXIndex(slot, k, 0)], X[GateXIndex(slot, k, 1)]) : 0.0f;   // hi + residual == the value
    }
    __syncthreads();
    if (tid < GG_ROWS) {
      float f[GX_SLOTS];
#pragma unroll
      for (int s = 0; s < GX_SLOTS; ++s) {
        const uint32_t slot = base + s < n_slots ? base + s : base;
        f[s] = W[LstmW(g, (int)(sym[slot] & 0xffu), i)];
      }
      const float4* wr = wd + g * W_DENSE_Q * L_CELLS + i;
#pragma unroll 2
      for (int q = 0; q < W_DENSE_Q - 1; ++q) {
        const float4 c = wr[q * L_CELLS];
#pragma unroll
        for (int s = 0; s < GX_SLOTS; ++s) {
          const float4 x = ((const float4*)(xs + s * GG_K))[q];
          f[s] = f_add(f[s], f_mul(x.x, c.x)); f[s] = f_add(f[s], f_mul(x.y, c.y));
          f[s] = f_add(f[s], f_mul(x.z, c.z)); f[s] = f_add(f[s], f_mul(x.w, c.w));
        }
      }
      {   // hidden 48, 49 and the bias input; the quad's fourth column is padding
        const float4 c = wr[(W_DENSE_Q - 1) * L_CELLS];
#pragma unroll
        for (int s = 0; s < GX_SLOTS; ++s) {
          const float4 x = ((const float4*)(xs + s * GG_K))[W_DENSE_Q - 1];
          f[s] = f_add(f[s], f_mul(x.x, c.x)); f[s] = f_add(f[s], f_mul(x.y, c.y)); f[s] = f_add(f[s], f_mul(x.z, c.z));
        }
      }
#pragma unroll
      for (int s = 0; s < GX_SLOTS; ++s)
        if (base + s < n_slots) G[(size_t)(base + s) * GG_N + tid] = f[s];
    }
  }
}

// ---- operand planes of the tensor-core variant ---------------------------------------------------------------------------
__global__ void GateWeightPrepKernel(const float* W, float* Wt) {
  const int t = (int)(blockIdx.x * blockDim.x + threadIdx.x);
  if (t >= GG_N * GG_K) return;
  const int row = t / GG_K, k = t % GG_K;
  float v = 0.0f;
  if (row < GG_ROWS && k < GG_KUSED) { const int g = row / L_CELLS; v = W[LstmW(g, L_NOUT + k, row - g * L_CELLS)]; }
  const float hi = Tf32Hi(v);
  Wt[GateWIndex(row, k, 0)] = hi;
  Wt[GateWIndex(row, k, 1)] = f_sub(v, hi);
}

#if defined(__CUDACC__)
// ---- tensor cores --------------------------------------------------------------------------------------------------------
// two operand stages + the tile's one-hot terms W[row][byte in front] ([GG_ROWS][GG_M] floats, gathered by warps 1-3 while warp 0
// runs the TMA / MMA loop, so that the epilogue adds them from shared memory instead of waiting on 150 scattered loads per stream)
enum : int { GT_STAGES = 2, GT_STAGE_BYTES = 2 * GG_A_FLOATS * 4 + 2 * GG_B_FLOATS * 4, GT_ONEHOT_BYTES = GG_ROWS * GG_M * 4,
             GT_SMEM_BYTES = GT_STAGES * GT_STAGE_BYTES + GT_ONEHOT_BYTES, GT_TMEM_COLS = 256 };

// shared-memory matrix descriptor, K-major, no swizzle (layout type 0), descriptor version 1 (sm_100):
// bits 0-13 start address >> 4, 16-29 leading byte offset >> 4 (between the two 16-byte k-slices of one MMA),
// 32-45 stride byte offset >> 4 (between 8-row groups)
GMX_DEV inline uint64_t UmmaDesc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor of kind::tf32: D fp32 (bits 4-5 = 1), A and B tf32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 in bits 17-22, M >> 4 in bits 24-28
constexpr uint32_t kGateIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(GG_N >> 3) << 17) | ((uint32_t)(GG_M >> 4) << 24);

GMX_DEV inline void UmmaTf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(da), "l"(db), "r"(kGateIdesc), "r"(accumulate) : "memory");
}
GMX_DEV inline void UmmaCommit(uint64_t* mbar) {   // arrives on mbar when all MMAs issued so far have completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)) : "memory");
}

// grid = tiles of 128 slots; 128 threads. sym / Wfull: the one-hot column is added in the epilogue.
__global__ void __launch_bounds__(128) GateGemmTcKernel(const float* Xt, const float* Wt, const float* Wfull, const uint32_t* sym, float* G,
                                                        uint32_t n_slots) {
  extern __shared__ __align__(128) unsigned char gt_smem[];
  __shared__ uint64_t full[GT_STAGES], freed[GT_STAGES], accum;
  __shared__ uint32_t tmem_base;
  __shared__ uint32_t tile_sym[GG_M];
  float* onehot = (float*)(gt_smem + GT_STAGES * GT_STAGE_BYTES);   // [row][stream of the tile]
  const int tid = (int)threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t tile = blockIdx.x;
  tile_sym[tid] = tile * GG_M + (uint32_t)tid < n_slots ? (sym[tile * GG_M + (uint32_t)tid] & 0xffu) : 0u;
  if (tid == 0) {
    for (int i = 0; i < GT_STAGES; ++i) { MbarInit(&full[i], 1); MbarInit(&freed[i], 1); }
    MbarInit(&accum, 1);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_base)), "r"((uint32_t)GT_TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base;

  if (warp == 0) {
    const float* xa = Xt + (size_t)tile * GG_NCHUNK * 2 * GG_A_FLOATS;
    auto load = [&](int chunk) {   // lane 0: one stage = {A hi, A lo, B hi, B lo} of this k-chunk, four bulk copies on one mbarrier
      const int st = chunk % GT_STAGES;
      unsigned char* dst = gt_smem + (size_t)st * GT_STAGE_BYTES;
      MbarExpectTx(&full[st], (uint32_t)GT_STAGE_BYTES);
      BulkG2S(dst, xa + (size_t)chunk * 2 * GG_A_FLOATS, 2 * GG_A_FLOATS * 4, &full[st]);
      BulkG2S(dst + 2 * GG_A_FLOATS * 4, Wt + (size_t)chunk * 2 * GG_B_FLOATS, 2 * GG_B_FLOATS * 4, &full[st]);
    };
    if (lane == 0) for (int c = 0; c < GT_STAGES - 1; ++c) load(c);
#pragma unroll 1
    for (int c = 0; c < GG_NCHUNK; ++c) {
      const int st = c % GT_STAGES;
      const int nx = c + GT_STAGES - 1;   // the chunk that goes into the stage chunk c - 1 used
      if (nx < GG_NCHUNK) {
        if (nx >= GT_STAGES) MbarWait(&freed[nx % GT_STAGES], (uint32_t)((nx / GT_STAGES - 1) & 1));   // that stage's previous MMAs are done
        if (lane == 0) load(nx);
      }
      MbarWait(&full[st], (uint32_t)((c / GT_STAGES) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      if (lane == 0) {
        const uint32_t a_hi = (uint32_t)__cvta_generic_to_shared(gt_smem + (size_t)st * GT_STAGE_BYTES);
        const uint32_t a_lo = a_hi + GG_A_FLOATS * 4, b_hi = a_hi + 2 * GG_A_FLOATS * 4, b_lo = b_hi + GG_B_FLOATS * 4;
        // leading offset = distance of the two 16-byte k-slices of one MMA, stride offset = distance of 8-row groups (GateXIndex)
        const uint32_t a_lbo = (GG_M / 8) * 128u, a_sbo = 128u, b_lbo = (GG_N / 8) * 128u, b_sbo = 128u;
#pragma unroll
        for (int ks = 0; ks < GG_KC / 8; ++ks) {   // one MMA covers 8 k = two 4-float slices
          const uint32_t ao = (uint32_t)ks * 2u * (GG_M / 8) * 128u, bo = (uint32_t)ks * 2u * (GG_N / 8) * 128u;
          UmmaTf32(tmem, UmmaDesc(a_hi + ao, a_lbo, a_sbo), UmmaDesc(b_hi + bo, b_lbo, b_sbo), (c | ks) != 0);
          UmmaTf32(tmem, UmmaDesc(a_hi + ao, a_lbo, a_sbo), UmmaDesc(b_lo + bo, b_lbo, b_sbo), 1u);
          UmmaTf32(tmem, UmmaDesc(a_lo + ao, a_lbo, a_sbo), UmmaDesc(b_hi + bo, b_lbo, b_sbo), 1u);
        }
        UmmaCommit(&freed[st]);
        if (c + 1 == GG_NCHUNK) UmmaCommit(&accum);
      }
      __syncwarp();
    }
  }
  else {
#pragma unroll 8
    for (int idx = tid - 32; idx < GG_ROWS * GG_M; idx += 96) {
      const int row = idx / GG_M, r = idx % GG_M, g = row / L_CELLS;
      onehot[idx] = Wfull[LstmW(g, (int)tile_sym[r], row - g * L_CELLS)];
    }
  }
  __syncthreads();
  MbarWait(&accum, 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: TMEM lane = stream row of the tile; warp w reads lanes 32 w .. 32 w + 31, eight columns per load
  const uint32_t slot = tile * GG_M + (uint32_t)tid;
  float* grow = G + (size_t)slot * GG_N;
#pragma unroll 2
  for (int col = 0; col < GG_N; col += 8) {
    uint32_t v[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)col));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (slot < n_slots) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int row = col + j;
        float r = __uint_as_float(v[j]);
        if (row < GG_ROWS) r = f_add(r, onehot[row * GG_M + tid]);
        grow[row] = r;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)GT_TMEM_COLS) : "memory");
}
#endif  // __CUDACC__

}  // namespace gmx
#endif
