// TEST-ONLY: a tiny SIMT emulator so that the CUDA kernels in gmix_b200/csrc/*.cuh can be compiled by
// g++ and executed on the CPU (this container has no GPU). One CUDA thread = one fiber (hand-rolled
// x86-64 context switch); __syncthreads / __syncwarp / __shfl*_sync are real barriers between
// fibers, so registers live across barriers exactly as on the device. One block runs at a time.
// It exists to debug kernel *logic* against the oracle; it is never part of the shipped library.
#ifndef GMIX_TESTS_CUDA_EMU_H_
#define GMIX_TESTS_CUDA_EMU_H_
#define GMX_EMU 1
#define GMX_OVERLAY 1   // the emulator build carries the overlay mode of the generation kernels
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>
#include <vector>

namespace cuda_emu {

struct Dim { unsigned x = 1, y = 1, z = 1; };

struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  bool done = false;
  unsigned tid = 0;
  unsigned shfl_parity = 0;
};

struct Block {
  std::vector<Fiber> fibers;
  void* sched_sp = nullptr;
  Fiber* cur = nullptr;
  unsigned nthreads = 0;
  unsigned bar_count = 0, bar_gen = 0;
  unsigned nbar_count[16] = {0}, nbar_gen[16] = {0};   // named barriers (bar.sync id, count)
  std::vector<unsigned> wbar_count, wbar_gen;
  std::vector<uint64_t> shfl_slots;  // [warp][parity][lane]
  std::function<void()> body;
  unsigned block_idx = 0, grid_dim = 1;
};

inline Block*& B() { static Block* b = nullptr; return b; }

extern "C" void cuda_emu_switch(void** save_sp, void* next_sp);
asm(R"(
.text
.globl cuda_emu_switch
.type cuda_emu_switch,@function
cuda_emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size cuda_emu_switch,.-cuda_emu_switch
)");

inline void Yield() { Block* b = B(); cuda_emu_switch(&b->cur->sp, b->sched_sp); }

inline void FiberEntry() {
  Block* b = B();
  b->body();
  b->cur->done = true;
  Yield();
  abort();  // never resumed
}

inline void RunBlock(unsigned nthreads, unsigned block_idx, unsigned grid_dim, std::function<void()> body) {
  Block blk;
  blk.nthreads = nthreads; blk.block_idx = block_idx; blk.grid_dim = grid_dim;
  blk.body = body;
  blk.fibers.resize(nthreads);
  const unsigned nwarps = (nthreads + 31) / 32;
  blk.wbar_count.assign(nwarps, 0); blk.wbar_gen.assign(nwarps, 0);
  blk.shfl_slots.assign((size_t)nwarps * 2 * 32, 0);
  const size_t kStack = 256 << 10;
  for (unsigned t = 0; t < nthreads; ++t) {
    Fiber& f = blk.fibers[t];
    f.tid = t;
    f.stack = (char*)aligned_alloc(64, kStack);
    // initial frame: 6 callee-saved registers + return address (FiberEntry); keep the SysV alignment
    // (rsp % 16 == 8 at function entry).
    uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)63;
    void** sp = (void**)(top - 8);  // entry rsp after `ret` pops the address => (top-8)+... see below
    *--sp = (void*)&FiberEntry;     // return address consumed by `ret`
    for (int i = 0; i < 6; ++i) *--sp = nullptr;
    f.sp = sp;
  }
  Block* saved = B();
  B() = &blk;
  // EMU_ORDER=reverse / EMU_ORDER=shuffle<seed>: other legal interleavings of the threads between two barriers
  // (default: ascending thread id), to flush out races inside a phase.
  const char* order = getenv("EMU_ORDER");
  const bool reverse = order && order[0] == 'r';
  uint64_t rng = order && order[0] == 's' ? 0x9E3779B97F4A7C15ull * (uint64_t)(atoi(order + 7) + 1) : 0;
  std::vector<unsigned> perm(nthreads);
  for (unsigned t = 0; t < nthreads; ++t) perm[t] = reverse ? nthreads - 1 - t : t;
  for (;;) {
    bool any = false;
    if (rng) for (unsigned t = nthreads - 1; t > 0; --t) { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; std::swap(perm[t], perm[rng % (t + 1)]); }
    for (unsigned k = 0; k < nthreads; ++k) {
      const unsigned t = perm[k];
      Fiber& f = blk.fibers[t];
      if (f.done) continue;
      any = true;
      blk.cur = &f;
      cuda_emu_switch(&blk.sched_sp, f.sp);
    }
    if (!any) break;
  }
  for (auto& f : blk.fibers) free(f.stack);
  B() = saved;
}

inline void BlockBarrier() {
  Block* b = B();
  const unsigned gen = b->bar_gen;
  if (++b->bar_count == b->nthreads) { b->bar_count = 0; b->bar_gen++; }
  else while (b->bar_gen == gen) Yield();
}
inline void NamedBarrier(int id, unsigned count) {   // bar.sync id, count: `count` threads of the block meet here
  Block* b = B();
  const unsigned gen = b->nbar_gen[id];
  if (++b->nbar_count[id] == count) { b->nbar_count[id] = 0; b->nbar_gen[id]++; }
  else while (b->nbar_gen[id] == gen) Yield();
}
inline void WarpBarrier() {
  Block* b = B();
  const unsigned w = b->cur->tid / 32;
  unsigned lanes = b->nthreads - w * 32; if (lanes > 32) lanes = 32;
  const unsigned gen = b->wbar_gen[w];
  if (++b->wbar_count[w] == lanes) { b->wbar_count[w] = 0; b->wbar_gen[w]++; }
  else while (b->wbar_gen[w] == gen) Yield();
}
template <typename T>
inline T Shfl(T v, unsigned src_lane) {
  Block* b = B();
  Fiber* f = b->cur;
  const unsigned w = f->tid / 32, lane = f->tid % 32;
  const unsigned par = f->shfl_parity; f->shfl_parity ^= 1;
  uint64_t* slots = &b->shfl_slots[((size_t)w * 2 + par) * 32];
  uint64_t raw = 0; memcpy(&raw, &v, sizeof(T));
  slots[lane] = raw;
  WarpBarrier();
  T out; memcpy(&out, &slots[src_lane % 32], sizeof(T));
  return out;
}

// cp.async emulation. Default: the copy happens at issue (earliest legal moment). With GMX_EMU_DEFER_CP the copy
// happens when the issuing thread WAITS for it (latest legal moment) and reads global memory as it is then: code
// that is only correct for one of the two timings has an ordering bug the synchronous emulation would hide.
struct PendingCopy { void* dst; const void* src; unsigned bytes; unsigned group; };
struct CpState { std::vector<PendingCopy> pending; unsigned open_group = 0; };
inline CpState& Cp() { static std::vector<CpState> v(2048); return v[B()->cur->tid]; }
inline void CpAsyncIssue(void* dst, const void* src, unsigned bytes) { CpState& c = Cp(); c.pending.push_back(PendingCopy{dst, src, bytes, c.open_group}); }
inline void CpAsyncCommitGroup() { Cp().open_group++; }
inline void CpAsyncComplete(unsigned first_pending_group) {   // runs every copy of a group < first_pending_group
  CpState& c = Cp();
  size_t keep = 0;
  for (size_t i = 0; i < c.pending.size(); ++i) {
    if (c.pending[i].group < first_pending_group) memcpy(c.pending[i].dst, c.pending[i].src, c.pending[i].bytes);
    else c.pending[keep++] = c.pending[i];
  }
  c.pending.resize(keep);
}
inline void CpAsyncWaitGroupN(unsigned n) { CpState& c = Cp(); CpAsyncComplete(c.open_group > n ? c.open_group - n : 0); }
inline void CpAsyncWaitEverything() { CpAsyncComplete(0xffffffffu); }

struct ThreadIdxProxy { unsigned y = 0, z = 0; struct X { operator unsigned() const { return B()->cur->tid; } } x; };
struct BlockIdxProxy { unsigned y = 0, z = 0; struct X { operator unsigned() const { return B()->block_idx; } } x; };
struct BlockDimProxy { unsigned y = 1, z = 1; struct X { operator unsigned() const { return B()->nthreads; } } x; };
struct GridDimProxy { unsigned y = 1, z = 1; struct X { operator unsigned() const { return B()->grid_dim; } } x; };

}  // namespace cuda_emu

static cuda_emu::ThreadIdxProxy threadIdx;
static cuda_emu::BlockIdxProxy blockIdx;
static cuda_emu::BlockDimProxy blockDim;
static cuda_emu::GridDimProxy gridDim;

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __launch_bounds__(...)
#define __restrict__

inline void __syncthreads() { cuda_emu::BlockBarrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { cuda_emu::WarpBarrier(); }
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return cuda_emu::Shfl(v, (unsigned)src); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int mask) {
  return cuda_emu::Shfl(v, (cuda_emu::B()->cur->tid % 32) ^ (unsigned)mask);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (cuda_emu::Shfl(pred ? 1u : 0u, (unsigned)l) & 1u) << l;
  return m;
}
inline void __threadfence_block() {}
inline void __nanosleep(unsigned) { cuda_emu::Yield(); }
inline int __ffs(int x) { return x == 0 ? 0 : __builtin_ctz((unsigned)x) + 1; }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
template <typename T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> inline T atomicCAS(T* p, T cmp, T v) { T o = *p; if (o == cmp) *p = v; return o; }
struct uint4 { unsigned x, y, z, w; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

#endif  // GMIX_TESTS_CUDA_EMU_H_
