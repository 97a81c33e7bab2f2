// fm_sm100.cuh -- hand-written sm_100a primitives: mbarrier, TMA, tcgen05 (MMA / TMEM ld,st /
// alloc / commit) and the UMMA shared-memory / instruction descriptors.  Inline PTX only.
//
// Descriptor bit layouts follow the PTX ISA "tcgen05 matrix descriptor" / "instruction
// descriptor" tables (the same fields CUTLASS names in cute/arch/mma_sm100_desc.hpp).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace fm {
namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier ----------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// make generic-proxy smem writes visible to the async proxy (TMA / tcgen05 operand reads)
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Developer cycle counters (-DFM_TIMING): FM_T(slot, cycles) adds into a per-file global table that
// fastmax_debug_timing_fetch_<file>() returns; compiled out otherwise.  Put FM_TIMING_DEFINE(<file>)
// once at namespace scope of the .cu that uses FM_T.
#ifdef FM_TIMING
#define FM_TIMING_DEFINE(name)                                                                        \
  static __device__ unsigned long long fm_timing_tab[64];                                             \
  extern "C" int fastmax_debug_timing_fetch_##name(unsigned long long* out, int reset) {              \
    cudaDeviceSynchronize();                                                                          \
    if (cudaMemcpyFromSymbol(out, fm_timing_tab, sizeof(unsigned long long) * 64) != cudaSuccess) return -4; \
    if (reset) {                                                                                      \
      unsigned long long z[64] = {};                                                                  \
      cudaMemcpyToSymbol(fm_timing_tab, z, sizeof(z));                                                \
    }                                                                                                 \
    return 0;                                                                                         \
  }
#define FM_T(slot, val) atomicAdd(&fm_timing_tab[slot], (unsigned long long)(val))
#define FM_NOW() clock64()
#else
#define FM_TIMING_DEFINE(name)
#define FM_T(slot, val) ((void)0)
#define FM_NOW() 0LL
#endif

// Wait with a watchdog: a protocol bug must trap (kernel error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = 0) {
  if (mbar_try_wait(bar, parity)) return;
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s
      printf("fastmax: mbarrier wait timed out (block %d,%d thread %d bar %u parity %u tag %d)\n", blockIdx.x,
             blockIdx.y, threadIdx.x, smem_u32(bar), parity, tag);
      __trap();
    }
  }
}

// wait without the time-out / diagnostic path (no printf arguments, no stack frame): for inner loops of kernels whose
// protocol is covered by the tests, where the registers the diagnostic path keeps live matter
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}

// Warp-collective wait: one lane polls, the warp reconverges.  Every polling thread costs an
// operation on the SM's (serial) mbarrier unit, so hundreds of threads spinning on the same
// barrier throttle every arrive/wait of the CTA; one poller per warp is enough.
__device__ __forceinline__ void mbar_wait_warp(uint64_t* bar, uint32_t parity) {
  if ((threadIdx.x & 31) == 0) mbar_wait(bar, parity);
  __syncwarp();
}

// ---------------------------------------------------------------- packed fp32 arithmetic ---
// sm_100 executes fp32 add / mul / fma on register PAIRS (.f32x2): half the instructions of an element-wise loop.
// Same rounding as the scalar forms.  The pack / unpack moves are register renames when the pair is adjacent.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ f32x2 pack2u(uint32_t lo, uint32_t hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint32_t pack_bf16x2_of(f32x2 v) {      // {lo, hi} -> bf16x2 (lo in the low half)
  float lo, hi;
  unpack2(v, lo, hi);
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// ---------------------------------------------------------------- TMA ---------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(m), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// bulk tensor store shared -> global (the shared tile is in the tensor map's swizzled layout)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(m), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// plain (1-D) bulk copy shared -> global; size a multiple of 16 bytes, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all bulk groups of this thread have finished READING their shared-memory source
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// ... and have completed (required before the CTA exits)
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// barrier among `nthreads` threads of the CTA (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_barrier_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void sts128(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(saddr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---------------------------------------------------------------- TMEM --------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait::ld that names the destination registers of the load it completes, so the compiler cannot
// move their uses above the wait (needed when other loads are kept in flight around it)
__device__ __forceinline__ void tmem_wait_ld_dep16(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait_ld_dep32(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// tcgen05.commit: mbarrier arrives when all prior tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32-bit, N consecutive columns per thread: thread t of warp w <-> lane 32*(w%4)+t
#define FM_R8(a, o) a[o], a[o + 1], a[o + 2], a[o + 3], a[o + 4], a[o + 5], a[o + 6], a[o + 7]
#define FM_OUT8(a, o)                                                                              \
  "=r"(a[o]), "=r"(a[o + 1]), "=r"(a[o + 2]), "=r"(a[o + 3]), "=r"(a[o + 4]), "=r"(a[o + 5]), "=r"(a[o + 6]), \
      "=r"(a[o + 7])
#define FM_IN8(a, o)                                                                            \
  "r"(a[o]), "r"(a[o + 1]), "r"(a[o + 2]), "r"(a[o + 3]), "r"(a[o + 4]), "r"(a[o + 5]), "r"(a[o + 6]), \
      "r"(a[o + 7])

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : FM_OUT8(r, 0)
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : FM_OUT8(r, 0), FM_OUT8(r, 8)
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : FM_OUT8(r, 0), FM_OUT8(r, 8), FM_OUT8(r, 16), FM_OUT8(r, 24)
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::FM_IN8(r, 0), "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
      ::FM_IN8(r, 0), FM_IN8(r, 8), "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31};"
      ::FM_IN8(r, 0), FM_IN8(r, 8), FM_IN8(r, 16), FM_IN8(r, 24), "r"(taddr)
      : "memory");
}

// ---------------------------------------------------------------- UMMA --------------------
// Shared-memory matrix descriptor (64 bit).  Fields (PTX ISA, tcgen05 matrix descriptor):
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4   [46,48) version = 1 on sm_100
//   [61,64) swizzle: 0 none, 2 = 128B, 4 = 64B, 6 = 32B
constexpr uint32_t SWZ_128B = 2;
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                   uint32_t swizzle) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;
  d |= static_cast<uint64_t>(swizzle) << 61;
  return d;
}
// Canonical 128B-swizzled tiles of bf16 with 128-byte rows (64 elements), rows packed at 128 B
// pitch, 8-row groups of 1024 B (what a TMA box of 64 x R elements with SWIZZLE_128B writes):
//  * K-major operand  (rows = M or N index, the 64 contiguous elements run along K):
//        SBO = 1024 (next 8-row group), LBO unused (=16); advance K by 16 elements: +32 B.
//  * MN-major operand (rows = K index, the 64 contiguous elements run along M/N):
//        SBO = 1024 (next 8 K-rows), LBO = byte offset of the next 64-wide MN block;
//        advance K by 16 rows: +2048 B.
__device__ __forceinline__ uint64_t desc_kmajor_sw128(uint32_t saddr) { return make_smem_desc(saddr, 16, 1024, SWZ_128B); }
__device__ __forceinline__ uint64_t desc_mnmajor_sw128(uint32_t saddr, uint32_t mn_block_bytes) {
  return make_smem_desc(saddr, mn_block_bytes, 1024, SWZ_128B);
}

// Instruction descriptor for kind::f16 with bf16 A/B, fp32 accumulate.
//   [4,6) D fmt (1=f32)  [7,10) A fmt (1=bf16)  [10,13) B fmt (1=bf16)
//   [15] A major (1=MN)  [16] B major (1=MN)  [17,23) N>>3  [24,29) M>>4
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]   -- one thread issues
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]   -- A: lane = row, 2 bf16 per 32-bit column (even k in the low half)
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- host-side descriptor construction ----------------------------------------------------
// tcgen05.mma operands are uniform-datapath registers.  When a descriptor is computed in ordinary
// registers nvcc 12.9 wraps every MMA in an ELECT / R2UR x4-7 / UTCHMMA sequence that costs ~110
// cycles per instruction (tools/mmarate.py); when it comes from the constant bank (kernel parameters)
// or is an immediate, and the MMAs sit together inside one elect.sync region of a converged warp,
// they issue back to back at the tensor pipe's rate (32 cycles for 128x64x16, tools/mmarate_const.py).
// So the kernels take their shared-memory descriptors as kernel parameters, built on the host for the
// known dynamic-shared-memory base (FM_SMEM_BASE, asserted on the device), and address tensor memory
// with immediates (allocation base asserted to be 0).
constexpr uint32_t FM_SMEM_BASE = 1024;   // dynamic smem of a kernel without static __shared__ (1 KiB reserved)
__host__ __device__ constexpr uint64_t make_smem_desc_const(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                                            uint32_t swizzle) {
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (static_cast<uint64_t>(swizzle) << 61);
}
__host__ __device__ constexpr uint64_t desc_kmajor_sw128_const(uint32_t saddr) {
  return make_smem_desc_const(saddr, 16, 1024, 2);
}
__host__ __device__ constexpr uint64_t desc_mnmajor_sw128_const(uint32_t saddr, uint32_t mn_block_bytes) {
  return make_smem_desc_const(saddr, mn_block_bytes, 1024, 2);
}
// tcgen05.commit to an mbarrier given by its shared-memory address (an immediate in the callers)
__device__ __forceinline__ void umma_commit_addr(uint32_t bar_saddr) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_saddr)
               : "memory");
}
// kernels that rely on the immediates above check them once
__device__ __forceinline__ void assert_bases(uint32_t tmem_base, const void* smem_aligned) {
  if (tmem_base != 0 || smem_u32(smem_aligned) != FM_SMEM_BASE) {
    if (threadIdx.x == 0)
      printf("fastmax: unexpected TMEM base %u or dynamic smem base %u (expected 0 / %u)\n", tmem_base,
             smem_u32(smem_aligned), FM_SMEM_BASE);
    __trap();
  }
}

// pack two floats to bf16x2: lo -> bits [15:0], hi -> bits [31:16]
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t mul_bf16x2(uint32_t a, uint32_t b) {
  uint32_t r;
  asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// broadcast the low (sel=0) or high (sel=1) bf16 of x to both halves
__device__ __forceinline__ uint32_t bcast_bf16(uint32_t x, int sel) {
  return sel ? __byte_perm(x, x, 0x3232) : __byte_perm(x, x, 0x1010);
}

}  // namespace sm100

// ---------------------------------------------------------------- host: tensor maps -------
// 3-D bf16 tensor map over a (batch, rows, cols) view; cols contiguous, box = (box_cols x box_rows x 1),
// 128B swizzle, zero fill out of bounds.  Returns FASTMAX_OK / FASTMAX_ERR_CUDA.
int make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                      uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_cols,
                      uint32_t box_rows);
int make_tmap_state_4d(CUtensorMap* map, const void* base, uint64_t nslab, uint64_t nrec, uint64_t rec_bytes,
                       uint32_t box_rows, uint32_t box_slabs);
// same for an fp32 tensor (box_cols * 4 bytes <= 128: one swizzle row)
int make_tmap_f32_3d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                      uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_cols,
                      uint32_t box_rows);

}  // namespace fm
