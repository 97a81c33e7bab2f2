// fm_selftest.cu -- on-device known-answer test of the sm_100a primitives in fm_sm100.cuh:
// TMA 128B-swizzled tile loads, UMMA descriptors for K-major and MN-major B operands, A from
// shared memory (SS) and from tensor memory (TS), TMEM ld/st lane mapping.  tests/ compare the
// four 128x64 products against torch.matmul.  (Test infrastructure inside the library so the
// primitives are validated on the same code path the production kernels use.)
#include "fm_common.cuh"
#include "fm_sm100.cuh"

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
