// fm_selftest.cu -- on-device known-answer test of the sm_100a primitives in fm_sm100.cuh:
// TMA 128B-swizzled tile loads, UMMA descriptors for K-major and MN-major B operands, A from
// shared memory (SS) and from tensor memory (TS), TMEM ld/st lane mapping.  tests/ compare the
// four 128x64 products against torch.matmul.  (Test infrastructure inside the library so the
// primitives are validated on the same code path the production kernels use.)
#include "../fm_common.cuh"
#include "../fm_sm100.cuh"

namespace fm {
namespace {
using namespace sm100;

__global__ void __launch_bounds__(128, 1)
selftest_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBk,
                const __grid_constant__ CUtensorMap tmBn, const __nv_bfloat16* __restrict__ A,
                float* __restrict__ D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // 128 rows x 128 B
  uint8_t* sBk = smem + 16384;     // 64 rows (n) x 128 B (k)
  uint8_t* sBn = smem + 24576;     // 64 rows (k) x 128 B (n)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 32768 + 64);

  const int tid = threadIdx.x, warp = tid >> 5;
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    mbar_arrive_expect_tx(&bars[0], 16384 + 8192 + 8192);
    tma_load_3d(sA, &tmA, &bars[0], 0, 0, 0);
    tma_load_3d(sBk, &tmBk, &bars[0], 0, 0, 0);
    tma_load_3d(sBn, &tmBn, &bars[0], 0, 0, 0);
  }
  // A into tensor memory: thread = row, 32 packed columns (64 bf16)
  {
    uint32_t r[32];
    const uint4* src = reinterpret_cast<const uint4*>(A + (size_t)tid * 64);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint4 x = src[i];
      r[4 * i] = x.x, r[4 * i + 1] = x.y, r[4 * i + 2] = x.z, r[4 * i + 3] = x.w;
    }
    tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 256, r);
    tmem_wait_st();
  }
  mbar_wait(&bars[0], 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) {
    const uint32_t id_kk = make_idesc_bf16(128, 64, false, false);
    const uint32_t id_kn = make_idesc_bf16(128, 64, false, true);
    const uint64_t dA = desc_kmajor_sw128(smem_u32(sA));
    const uint64_t dBk = desc_kmajor_sw128(smem_u32(sBk));
    const uint64_t dBn = desc_mnmajor_sw128(smem_u32(sBn), 8192);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      umma_ss(tmem + 0, dA + kk * 2, dBk + kk * 2, id_kk, kk > 0);
      umma_ss(tmem + 64, dA + kk * 2, dBn + kk * 128, id_kn, kk > 0);
      umma_ts(tmem + 128, tmem + 256 + kk * 8, dBk + kk * 2, id_kk, kk > 0);
      umma_ts(tmem + 192, tmem + 256 + kk * 8, dBn + kk * 128, id_kn, kk > 0);
    }
    umma_commit(&bars[1]);
  }
  mbar_wait(&bars[1], 0);
  tc_fence_after();
#pragma unroll 1
  for (int t = 0; t < 4; ++t) {
    uint32_t r[32];
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + t * 64 + half * 32, r);
      tmem_wait_ld();
#pragma unroll
      for (int c = 0; c < 32; ++c) D[((size_t)t * 128 + tid) * 64 + half * 32 + c] = __uint_as_float(r[c]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace
}  // namespace fm

extern "C" int fastmax_selftest_umma(const void* A, const void* Bk, const void* Bn, float* D, void* stream) {
  using namespace fm;
  if (!A || !Bk || !Bn || !D) return FASTMAX_ERR_BAD_ARG;
  CUtensorMap tmA, tmBk, tmBn;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tmA, A, 64, 128, 1, 128, 128 * 128, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmBk, Bk, 64, 64, 1, 128, 64 * 128, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmBn, Bn, 64, 64, 1, 128, 64 * 128, 64, 64))) return rc;
  const int smem = 32768 + 1024 + 1024;
  FM_CUDA(cudaFuncSetAttribute(selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  FM_LAUNCH("selftest_umma", (cudaStream_t)stream,
            selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(tmA, tmBk, tmBn, (const __nv_bfloat16*)A, D));
  return FASTMAX_OK;
}

// ---------------------------------------------------------------------------------------------
// Tensor-pipe rate probe (developer tool): cycles per tcgen05.mma for a few operand configurations.
// variant bit0: B MN-major, bit1: A from TMEM, bits 2-3: N = 64 << x, bit4: rotate over 4 accumulators
// ---------------------------------------------------------------------------------------------
namespace fm {
namespace {
using namespace sm100;
__global__ void __launch_bounds__(128, 1) mmarate_kernel(int variant, int reps, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 65536);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 65536 + 64);
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3C003C00u;  // small bf16 values
  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    fence_barrier_init();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const bool b_mn = variant & 1, a_tmem = variant & 2, rotate = variant & 16;
  const int N = 64 << ((variant >> 2) & 3);
  if (variant & 64) {
    // "clean" warp-uniform issue: the whole warp runs the loop, every operand is derived from
    // provably warp-uniform values (TMEM base asserted to be 0), one elected lane issues.
    const int nissuers = (variant & 128) ? 4 : 1;   // concurrent issuing warps (different accumulators)
    if (warp < nissuers) {
      const uint32_t tmu = __shfl_sync(0xffffffffu, tmem, 0);   // warp-uniform for the compiler
      const uint32_t idesc = make_idesc_bf16(128, N, false, b_mn);
      const uint32_t sb = __shfl_sync(0xffffffffu, smem_u32(smem), 0);
      const uint64_t dA = desc_kmajor_sw128(sb);
      const uint64_t dB = b_mn ? desc_mnmajor_sw128(sb + 16384, 8192) : desc_kmajor_sw128(sb + 16384);
      const uint32_t kstep = b_mn ? 128 : 2;
      long long t0 = clock64();
      for (int r = 0; r < reps; r += 4) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t d = tmu + (rotate ? kk * 64 : 0) + (nissuers > 1 ? __shfl_sync(0xffffffffu, warp, 0) * 64 : 0);
          if (elect_one()) {
            if (a_tmem) umma_ts(d, tmu + 448 + kk * 8, dB + kk * kstep, idesc, 1);
            else umma_ss(d, dA + kk * 2, dB + kk * kstep, idesc, 1);
          }
        }
      }
      long long t1 = clock64();
      if (tid == 0) {
        umma_commit(&bars[0]);
        mbar_wait(&bars[0], 0);
        long long t2 = clock64();
        out[0] = (unsigned long long)(t1 - t0);
        out[1] = (unsigned long long)(t2 - t0);
      }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
    return;
  }
  const bool uniform_issue = variant & 32;   // whole warp runs the loop, one elected lane issues
  if (uniform_issue ? warp == 0 : tid == 0) {
    const uint32_t idesc = make_idesc_bf16(128, N, false, b_mn);
    const uint64_t dA = desc_kmajor_sw128(smem_u32(smem));
    const uint64_t dB = b_mn ? desc_mnmajor_sw128(smem_u32(smem + 16384), 8192) : desc_kmajor_sw128(smem_u32(smem + 16384));
    const int nacc = rotate ? 4 : 1;
    const uint32_t tm = uniform_issue ? __shfl_sync(0xffffffffu, tmem, 0) : tmem;
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tm + (nacc > 1 ? (r & 3) * 64 : 0);
      const int kk = r & 3;
      if (!uniform_issue || elect_one()) {
        if (a_tmem) umma_ts(d, tm + 448 + kk * 8, dB + (b_mn ? kk * 128 : kk * 2), idesc, 1);
        else umma_ss(d, dA + kk * 2, dB + (b_mn ? kk * 128 : kk * 2), idesc, 1);
      }
      if (uniform_issue) __syncwarp();
    }
    long long t1 = clock64();
    if (tid == 0) {
      umma_commit(&bars[0]);
      mbar_wait(&bars[0], 0);
      long long t2 = clock64();
      out[0] = (unsigned long long)(t1 - t0);
      out[1] = (unsigned long long)(t2 - t0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}
}  // namespace
}  // namespace fm

// variant "const": every MMA operand comes from kernel parameters (constant bank) or immediates, so the
// compiler can keep them in uniform registers without R2UR moves.
// MODE 0: A in tensor memory, N=64   1: A in shared memory, N=64   2: A in smem, N=128   3: A in tmem, N=128
// Operands rotate over the four k-steps of a 128x64 A tile and of a B tile, like a real K=64 product.
namespace fm {
namespace {
struct ConstDescs {
  unsigned long long a[4];
  unsigned long long b[4];
  unsigned int idesc, smem_base_expected;
};
template <int MODE>
__global__ void __launch_bounds__(128, 1) mmarate_const_kernel(const __grid_constant__ ConstDescs cd, int reps,
                                                               unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3C003C00u;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    out[2] = smem_u32(smem_raw);
    out[3] = slot;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    long long t0 = clock64();
#pragma unroll 1
    for (int r = 0; r < reps; r += 4) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          if (MODE == 0) umma_ts(kk * 64, 448 + kk * 8, cd.b[kk], cd.idesc, 1);
          else if (MODE == 1) umma_ss(kk * 64, cd.a[kk], cd.b[kk], cd.idesc, 1);
          else if (MODE == 2) umma_ss((kk & 1) * 128, cd.a[kk], cd.b[kk], cd.idesc, 1);
          else if (MODE == 3) umma_ts((kk & 1) * 128, 448 + kk * 8, cd.b[kk], cd.idesc, 1);
          else if (MODE == 5) umma_ss(0, cd.a[kk], cd.b[kk], cd.idesc, 1);           // SS N=128, one accumulator
          else if (MODE == 6) umma_ts(0, 448 + kk * 8, cd.b[kk], cd.idesc, 1);       // TS N=64, one accumulator
          else if (MODE == 7) umma_ss(0, cd.a[kk], cd.b[kk], cd.idesc, 1);           // SS N=64, one accumulator
          else if (MODE == 8) umma_ts(0, 448 + kk * 8, cd.b[kk], cd.idesc, 1);       // TS N=128, one accumulator
        }
      }
      __syncwarp();
    }
    long long t1 = clock64();
    if (tid == 0) {
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      long long t2 = clock64();
      out[0] = (unsigned long long)(t1 - t0);
      out[1] = (unsigned long long)(t2 - t0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}
}  // namespace
}  // namespace fm

extern "C" int fastmax_selftest_mmarate_const(int reps, unsigned int smem_base, unsigned long long* out_dev, void* stream,
                                              int mode) {
  using namespace fm;
  ConstDescs cd;
  const int n = (mode == 2 || mode == 3 || mode == 5 || mode == 8) ? 128 : 64;
  for (int kk = 0; kk < 4; ++kk) {
    cd.a[kk] = sm100::desc_kmajor_sw128_const(smem_base) + kk * 2;                      // 128x64 K-major tile
    cd.b[kk] = sm100::desc_mnmajor_sw128_const(smem_base + 16384, 8192) + kk * 128;     // 64xN MN-major tile(s)
  }
  cd.idesc = sm100::make_idesc_bf16(128, n, false, true);
  cd.smem_base_expected = smem_base;
  const int smem = 65536 + 1024;
  cudaStream_t st = (cudaStream_t)stream;
#define FM_MRC(M)                                                                                            \
  case M:                                                                                                    \
    FM_CUDA(cudaFuncSetAttribute(mmarate_const_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    FM_LAUNCH("selftest_mmarate_const", st, mmarate_const_kernel<M><<<1, 128, smem, st>>>(cd, reps, out_dev));   \
    break;
  switch (mode) {
    FM_MRC(0) FM_MRC(1) FM_MRC(2) FM_MRC(3) FM_MRC(5) FM_MRC(6) FM_MRC(7) FM_MRC(8)
    default: return FASTMAX_ERR_BAD_ARG;
  }
#undef FM_MRC
  return FASTMAX_OK;
}

extern "C" int fastmax_selftest_mmarate(int variant, int reps, unsigned long long* out_dev, void* stream) {
  using namespace fm;
  const int smem = 65536 + 1024 + 1024;
  FM_CUDA(cudaFuncSetAttribute(mmarate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  FM_LAUNCH("selftest_mmarate", (cudaStream_t)stream,
            mmarate_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(variant, reps, out_dev));
  return FASTMAX_OK;
}

// ---------------------------------------------------------------------------------------------
// ALU rate probe (developer tool): cycles per warp-instruction of a few candidate element-wise ops,
// 8 independent chains per thread.  kind: 0 mul.bf16x2, 1 fma.bf16x2, 2 mul.f16x2, 3 mul.f32, 4 mul.f32x2
// ---------------------------------------------------------------------------------------------
namespace fm {
namespace {
template <int KIND>
__global__ void alurate_kernel(int iters, unsigned long long* out, uint32_t seed) {
  uint32_t r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = seed + i * 0x00010001u + threadIdx.x;
  const uint32_t m = 0x3F803F80u ^ (seed & 1);
  unsigned long long m2 = ((unsigned long long)0x3F800000u << 32) | 0x3F800000u;
  unsigned long long v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = ((unsigned long long)r[i] << 32) | r[i];
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (KIND == 0) asm volatile("mul.rn.bf16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(m));
      else if (KIND == 1) asm volatile("fma.rn.bf16x2 %0, %0, %1, %1;" : "+r"(r[i]) : "r"(m));
      else if (KIND == 2) asm volatile("mul.rn.f16x2 %0, %0, %1;" : "+r"(r[i]) : "r"(m));
      else if (KIND == 3) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(*reinterpret_cast<float*>(&r[i])) : "f"(1.0001f));
      else if (KIND == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(v[i]) : "l"(m2));
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= r[i] ^ (uint32_t)v[i];
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345678u) out[1] = acc;
}
}  // namespace
}  // namespace fm

extern "C" int fastmax_selftest_alurate(int kind, int iters, int threads, unsigned long long* out_dev, void* stream) {
  using namespace fm;
  cudaStream_t st = (cudaStream_t)stream;
  switch (kind) {
    case 0: FM_LAUNCH("selftest_alurate", st, alurate_kernel<0><<<1, threads, 0, st>>>(iters, out_dev, 12345u)); break;
    case 1: FM_LAUNCH("selftest_alurate", st, alurate_kernel<1><<<1, threads, 0, st>>>(iters, out_dev, 12345u)); break;
    case 2: FM_LAUNCH("selftest_alurate", st, alurate_kernel<2><<<1, threads, 0, st>>>(iters, out_dev, 12345u)); break;
    case 3: FM_LAUNCH("selftest_alurate", st, alurate_kernel<3><<<1, threads, 0, st>>>(iters, out_dev, 12345u)); break;
    default: FM_LAUNCH("selftest_alurate", st, alurate_kernel<4><<<1, threads, 0, st>>>(iters, out_dev, 12345u)); break;
  }
  return FASTMAX_OK;
}

// ---------------------------------------------------------------------------------------------
// TMEM load probe (developer tool): cycles per (tcgen05.ld xN + wait) with `inflight` loads per wait
// ---------------------------------------------------------------------------------------------
namespace fm {
namespace {
using namespace sm100;
template <int WIDTH, int INFLIGHT>
__global__ void ldrate_kernel(int iters, unsigned long long* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = slot + ((uint32_t)(32 * (warp & 3)) << 16);
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[INFLIGHT][WIDTH];
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j) {
      if (WIDTH == 16) tmem_ld16(tmem + ((it * INFLIGHT + j) * WIDTH) % 256, v[j]);
      else tmem_ld32(tmem + ((it * INFLIGHT + j) * WIDTH) % 256, v[j]);
    }
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j)
#pragma unroll
      for (int c = 0; c < WIDTH; ++c) acc += v[j][c];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (unsigned long long)(t1 - t0);
  if (acc == 0x12345u) out[1] = acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(slot, 512);
}
}  // namespace
}  // namespace fm

extern "C" int fastmax_selftest_ldrate(int width, int inflight, int iters, int threads, unsigned long long* out_dev,
                                       void* stream) {
  using namespace fm;
  cudaStream_t st = (cudaStream_t)stream;
  if (width == 16 && inflight == 1) { FM_LAUNCH("ldrate", st, ldrate_kernel<16, 1><<<1, threads, 0, st>>>(iters, out_dev)); }
  else if (width == 16 && inflight == 2) { FM_LAUNCH("ldrate", st, ldrate_kernel<16, 2><<<1, threads, 0, st>>>(iters, out_dev)); }
  else if (width == 16 && inflight == 4) { FM_LAUNCH("ldrate", st, ldrate_kernel<16, 4><<<1, threads, 0, st>>>(iters, out_dev)); }
  else if (width == 32 && inflight == 1) { FM_LAUNCH("ldrate", st, ldrate_kernel<32, 1><<<1, threads, 0, st>>>(iters, out_dev)); }
  else if (width == 32 && inflight == 2) { FM_LAUNCH("ldrate", st, ldrate_kernel<32, 2><<<1, threads, 0, st>>>(iters, out_dev)); }
  else return FASTMAX_ERR_BAD_ARG;
  return FASTMAX_OK;
}


// ---------------------------------------------------------------------------------------------
// Contention probe (developer tool): one converged warp issues tcgen05.mma back to back while
// `ldwarps` other warps stream tcgen05.ld (kind 0) or tcgen05.st (kind 1) from/to other columns of
// tensor memory.  out[0] = cycles of the MMA loop, out[1] = MMAs, out[2+2w] = cycles of warp w's
// loop, out[3+2w] = 16-column chunks it moved.
// MODE as in mmarate_const: 0 TS N=64, 1 SS N=64, 2 SS N=128, 3 TS N=128, 4 = no MMAs (spin).
// ---------------------------------------------------------------------------------------------
namespace fm {
namespace {
template <int MODE>
__global__ void __launch_bounds__(17 * 32, 1) contention_kernel(const __grid_constant__ ConstDescs cd, int reps,
                                                                 int ldwarps, int kind, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  __shared__ volatile int done;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 4; i += 17 * 32) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3C003C00u;
  if (warp == 16) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
    done = 0;
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 16) {
    long long t0 = clock64();
    while (clock64() - t0 < 20000) {}
    t0 = clock64();
    if (MODE < 4) {
#pragma unroll 1
      for (int r = 0; r < reps; r += 4) {
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (MODE == 0) umma_ts(kk * 64, 448 + kk * 8, cd.b[kk], cd.idesc, 1);
            else if (MODE == 1) umma_ss(kk * 64, cd.a[kk], cd.b[kk], cd.idesc, 1);
            else if (MODE == 2) umma_ss((kk & 1) * 128, cd.a[kk], cd.b[kk], cd.idesc, 1);
            else umma_ts((kk & 1) * 128, 448 + kk * 8, cd.b[kk], cd.idesc, 1);
          }
        }
        __syncwarp();
      }
      if (lane == 0) {
        umma_commit(&bar);
        mbar_wait(&bar, 0);
      }
    } else {
      while (clock64() - t0 < (long long)reps * 64) {}
    }
    __syncwarp();
    if (lane == 0) {
      out[0] = (unsigned long long)(clock64() - t0);
      out[1] = (unsigned long long)reps;
      done = 1;
    }
  } else if (warp < ldwarps) {
    const uint32_t tmem = slot + ((uint32_t)(32 * (warp & 3)) << 16) + 256;
    uint32_t acc = 0, chunks = 0;
    uint32_t v[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) v[c] = tid + c;
    const long long t0 = clock64();
    while (!done) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (kind == 0) {
          tmem_ld16(tmem + 16 * j + 128 * ((warp >> 2) & 1), v);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 16; ++c) acc ^= v[c];
        } else {
          tmem_st16(tmem + 16 * j + 128 * ((warp >> 2) & 1), v);
          tmem_wait_st();
        }
      }
      chunks += 8;
    }
    const long long t1 = clock64();
    if (lane == 0) {
      out[2 + 2 * warp] = (unsigned long long)(t1 - t0);
      out[3 + 2 * warp] = chunks + (acc == 0x12345u ? 1 : 0);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(slot, 512);
}
}  // namespace
}  // namespace fm

extern "C" int fastmax_selftest_contention(int mode, int reps, int ldwarps, int kind, unsigned int smem_base,
                                           unsigned long long* out_dev, void* stream) {
  using namespace fm;
  ConstDescs cd;
  const int n = (mode == 2 || mode == 3) ? 128 : 64;
  for (int kk = 0; kk < 4; ++kk) {
    cd.a[kk] = sm100::desc_kmajor_sw128_const(smem_base) + kk * 2;
    cd.b[kk] = sm100::desc_mnmajor_sw128_const(smem_base + 16384, 8192) + kk * 128;
  }
  cd.idesc = sm100::make_idesc_bf16(128, n, false, true);
  cd.smem_base_expected = smem_base;
  const int smem = 65536 + 1024;
  cudaStream_t st = (cudaStream_t)stream;
#define FM_CT(M)                                                                                          \
  case M:                                                                                                 \
    FM_CUDA(cudaFuncSetAttribute(contention_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    FM_LAUNCH("selftest_contention", st,                                                                  \
              contention_kernel<M><<<1, 17 * 32, smem, st>>>(cd, reps, ldwarps, kind, out_dev));           \
    break;
  switch (mode) {
    FM_CT(0) FM_CT(1) FM_CT(2) FM_CT(3) FM_CT(4)
    default: return FASTMAX_ERR_BAD_ARG;
  }
#undef FM_CT
  return FASTMAX_OK;
}

// ---------------------------------------------------------------------------------------------
// Multi-issuer probe (developer tool): `nissue` converged warps each issue groups of four TS N=64
// MMAs (issuer w: k-step w of four accumulators, like the build kernel) followed by one
// tcgen05.commit, back to back, no waits.  out[w] = cycles of issuer w's loop; `blocks` groups each.
// ---------------------------------------------------------------------------------------------
namespace fm {
namespace {
__global__ void __launch_bounds__(20 * 32, 1) multi_issue_kernel(const __grid_constant__ ConstDescs cd, int blocks,
                                                                  int nissue, int commits, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < 65536 / 4; i += 20 * 32) reinterpret_cast<uint32_t*>(smem_raw)[i] = 0x3C003C00u;
  if (warp == 0) {
    tmem_alloc(&slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], nissue);
    fence_barrier_init();
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp >= 16 && warp < 16 + nissue) {
    const int kk = warp - 16;
    const long long t0 = clock64();
#pragma unroll 1
    for (int b = 0; b < blocks; b += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (elect_one()) {
#define FM_MI(KK)                                                                   \
  _Pragma("unroll") for (int t = 0; t < 4; ++t)                                     \
      umma_ts(64 * t, 256 + 64 * t + 32 * (j & 1) + 8 * KK, cd.b[KK], cd.idesc, 1);
          if (kk == 0) { FM_MI(0) } else if (kk == 1) { FM_MI(1) } else if (kk == 2) { FM_MI(2) } else { FM_MI(3) }
#undef FM_MI
          for (int c = 0; c < commits; ++c) umma_commit(&bar[j]);
        }
        __syncwarp();
      }
    }
    if (lane == 0) out[kk] = (unsigned long long)(clock64() - t0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    if (lane == 0) {   // drain before freeing tensor memory
      umma_commit(&bar[0]);
    }
    long long t0 = clock64();
    while (clock64() - t0 < 200000) {}
    tmem_dealloc(slot, 512);
  }
}
}  // namespace
}  // namespace fm

extern "C" int fastmax_selftest_multi_issue(int blocks, int nissue, int commits, unsigned int smem_base,
                                            unsigned long long* out_dev, void* stream) {
  using namespace fm;
  ConstDescs cd;
  for (int kk = 0; kk < 4; ++kk) {
    cd.a[kk] = sm100::desc_kmajor_sw128_const(smem_base) + kk * 2;
    cd.b[kk] = sm100::desc_mnmajor_sw128_const(smem_base + 16384, 8192) + kk * 128;
  }
  cd.idesc = sm100::make_idesc_bf16(128, 64, false, true);
  cd.smem_base_expected = smem_base;
  const int smem = 65536 + 1024;
  cudaStream_t st = (cudaStream_t)stream;
  FM_CUDA(cudaFuncSetAttribute(multi_issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  FM_LAUNCH("selftest_multi_issue", st, multi_issue_kernel<<<1, 20 * 32, smem, st>>>(cd, blocks, nissue, commits, out_dev));
  return FASTMAX_OK;
}
