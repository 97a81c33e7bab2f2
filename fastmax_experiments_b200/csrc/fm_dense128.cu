// fm_dense128.cu -- FastMax for head size 128 (bf16) as dense tensor-core tiles on tcgen05.
//
// The moment kernels are built for D = 64 (a D = 128 state record would be 8.5 MB per boundary).  For the head size of
// Llama-2-7B (BASELINE.json config 4: 32 heads of 128, sequences of 4096) this file evaluates the polynomial attention
// directly, flash-attention style without max / exp, 128 x 128 score tiles at a time:
//
//   forward   per q tile i, for kv tiles j (<= i when causal):
//               S = Q_i K_j^T  ->  P = f(a S) o mask  (CUDA cores)  ->  acc += P [V_j | 1]      o = acc / acc[:,128], g = acc[:,128]
//   dq        per q tile i:  S = Q_i K_j^T, D = U_i V_j^T  ->  E = f'(a S) o (D - dlt_i) o mask  ->  dQ_i += E K_j        (x a)
//   dk, dv    per kv tile j, for q tiles i (>= j when causal):
//               S' = K_j Q_i^T, D' = V_j U_i^T  ->  P' = f(a S'), E' = f'(a S') o (D' - dlt_i)  ->  dV_j += P' U_i, dK_j += E' Q_i  (x a)
//
// with u = do / g, dlt = (do . o) / g (prep kernel below), f(x) = 1 + x (+ x^2/2), f'(x) = 1 (+ x): the same formulas as
// fm_exact.cu (reference: fastmax.py:219-322 forward, :433-691 backward), O(N^2 D) work but on the tensor pipe: every
// operand tile is two 64-column K-major panels (TMA, 128B swizzle), scores and accumulators live in tensor memory, the
// polynomial result goes back to tensor memory as the bf16 A operand of the second MMA.  One CTA per SM (512 tensor-
// memory columns): 16 compute warps (four per scheduler, so the polynomial's dependent instruction chains overlap;
// thread = one score row x 32 columns), a TMA producer warp, an MMA issuer warp.
#include "fm_common.cuh"
#include "fm_sm100.cuh"

namespace fm {
namespace d128 {
using namespace sm100;
typedef __nv_bfloat16 bf16;

constexpr int NCW = 16;                     // compute warps: quadrant = warp & 3, 32-column block = warp >> 2
constexpr int THREADS = (NCW + 2) * 32;     // + TMA producer warp + MMA issuer warp
constexpr int TILE = 32768;                 // one 128-token x 128-feature operand tile: two 16 KiB panels
constexpr int PANEL = 16384;

struct Params {
  int N, Npad, ntile, p, causal;
  float a;
  const float* dlt;          // (BH, Npad) fp32
  float* g;                  // (BH, N) (forward)
};

// Shared-memory descriptors of every MMA operand, built on the host for the fixed layouts below and passed in the
// constant bank: with immediate tensor-memory addresses (the CTA owns all 512 columns: base 0) one thread then issues
// at the tensor pipe's rate (fm_sm100.cuh).
struct Descs {
  unsigned long long resK[2][8];      // resident tile 0 / 1, K-major, 8 k-steps of 16 features
  unsigned long long strK[2][2][8];   // stage x streamed tile 0 / 1, K-major
  unsigned long long strMN[2][2][8];  // the same tiles read MN-major (tokens = K), 8 k-steps of 16 tokens
};

// S[tmem DT] = res[RA] . str[ST][OB]^T
template <uint32_t DT, int RA, int ST, int OB>
__device__ __forceinline__ void mma_scores(const Descs& D) {
  constexpr uint32_t ID = make_idesc_bf16(128, 128, false, false);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) umma_ss(DT, D.resK[RA][kk], D.strK[ST][OB][kk], ID, kk > 0);
}
// acc[tmem DT] += P[tmem PT, bf16 128 x 128] . str[ST][OB] read MN-major, N columns (128, or 144 with the ones panel)
template <int N, uint32_t DT, uint32_t PT, int ST, int OB>
__device__ __forceinline__ void mma_accum(const Descs& D, uint32_t acc_first) {
  constexpr uint32_t ID = make_idesc_bf16(128, N, false, true);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) umma_ts(DT, PT + 8 * kk, D.strMN[ST][OB][kk], ID, (acc_first || kk > 0) ? 1 : 0);
}

static void fill_descs(Descs& D, uint32_t res0, uint32_t res1, uint32_t st0, uint32_t st_stride, uint32_t op1_off) {
  const uint32_t base = sm100::FM_SMEM_BASE;
  for (int kk = 0; kk < 8; ++kk) {
    const uint32_t kp = (kk >> 2) * PANEL;
    D.resK[0][kk] = sm100::desc_kmajor_sw128_const(base + res0 + kp) + (kk & 3) * 2;
    D.resK[1][kk] = sm100::desc_kmajor_sw128_const(base + res1 + kp) + (kk & 3) * 2;
    for (int st = 0; st < 2; ++st)
      for (int op = 0; op < 2; ++op) {
        const uint32_t t = base + st0 + st * st_stride + op * op1_off;
        D.strK[st][op][kk] = sm100::desc_kmajor_sw128_const(t + kp) + (kk & 3) * 2;
        D.strMN[st][op][kk] = sm100::desc_mnmajor_sw128_const(t, PANEL) + kk * 128;
      }
  }
}

__device__ __forceinline__ void load_tile(uint8_t* dst, const CUtensorMap* tm, uint64_t* bar, int tok, int bh) {
  tma_load_3d(dst, tm, bar, 0, tok, bh);
  tma_load_3d(dst + PANEL, tm, bar, 64, tok, bh);
}

template <int P>
__device__ __forceinline__ float poly(float x) { return P == 2 ? fmaf(fmaf(0.5f, x, 1.f), x, 1.f) : 1.f + x; }

// ------------------------------------------------------------------------------------------------ forward
// shared memory: Q tile | 2 stages x [K tile | V tile | ones panel] | misc
constexpr int F_STAGE = 2 * TILE + PANEL;
constexpr int F_OFF_ST = TILE;
constexpr int F_OFF_MISC = F_OFF_ST + 2 * F_STAGE;
constexpr int F_SMEM = F_OFF_MISC + 1024 + 1024;

template <int P>
__global__ void __launch_bounds__(THREADS, 1)
fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
           const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const Params prm,
           const __grid_constant__ Descs D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + F_OFF_MISC);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* v_full = bars + 3;     // [2]
  uint64_t* k_free = bars + 5;     // [2] score MMA of the stage complete
  uint64_t* v_free = bars + 7;     // [2] accumulating MMA of the stage complete
  uint64_t* s_ready = bars + 9;    // [2] scores in tensor memory
  uint64_t* p_ready = bars + 11;   // [2] bf16 P written back
  uint64_t* acc_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y;
  const int it = prm.ntile - 1 - blockIdx.x;            // long rows first
  const int nj = prm.causal ? it + 1 : prm.ntile;
  const uint32_t sbase = smem_u32(smem);

  if (warp == NCW) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(q_full, 1), mbar_init(acc_done, 1);
    for (int s = 0; s < 2; ++s)
      mbar_init(&k_full[s], 1), mbar_init(&v_full[s], 1), mbar_init(&k_free[s], 1), mbar_init(&v_free[s], 1),
          mbar_init(&s_ready[s], 1), mbar_init(&p_ready[s], 1);
    fence_barrier_init();
  }
  // ones panels (third N panel of [V | 1]): element [tok][0] = 1
  for (int s = 0; s < 2; ++s) {
    uint8_t* ones = smem + F_OFF_ST + s * F_STAGE + 2 * TILE;
    for (int i = tid; i < PANEL / 16; i += THREADS) reinterpret_cast<uint4*>(ones)[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  if (tid < 128)
    for (int s = 0; s < 2; ++s)
      *reinterpret_cast<uint16_t*>(smem + F_OFF_ST + s * F_STAGE + 2 * TILE + tid * 128 + ((tid & 7) << 4)) = 0x3F80;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  assert_bases(tmem, smem);
  constexpr uint32_t C_ACC = 256;

  if (warp == NCW) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, TILE);
      load_tile(smem, &tmQ, q_full, it * 128, bh);
      for (int j = 0; j < nj; ++j) {
        const int s = j & 1;
        uint8_t* st = smem + F_OFF_ST + s * F_STAGE;
        if (j >= 2) mbar_wait(&k_free[s], ((j >> 1) - 1) & 1, 100 + j);
        mbar_arrive_expect_tx(&k_full[s], TILE);
        load_tile(st, &tmK, &k_full[s], j * 128, bh);
        if (j >= 2) mbar_wait(&v_free[s], ((j >> 1) - 1) & 1, 200 + j);
        mbar_arrive_expect_tx(&v_full[s], TILE);
        load_tile(st + TILE, &tmV, &v_full[s], j * 128, bh);
      }
    }
  } else if (warp == NCW + 1) {
    if (lane == 0) mbar_wait(q_full, 0, 300);
    __syncwarp();
    for (int j = 0; j <= nj; ++j) {
      const int s = j & 1;
      if (j < nj) {
        // scores of tile j into buffer s (its previous content, P of tile j-2, was consumed by an MMA issued earlier)
        if (lane == 0) mbar_wait(&k_full[s], (j >> 1) & 1, 400 + j);
        __syncwarp();
        tc_fence_after();
        if (elect_one()) {
          if (s == 0) mma_scores<0, 0, 0, 0>(D);
          else mma_scores<128, 0, 1, 0>(D);
          umma_commit(&k_free[s]);
          umma_commit(&s_ready[s]);
        }
        __syncwarp();
      }
      if (j > 0) {
        const int sp = (j - 1) & 1;
        if (lane == 0) {
          mbar_wait(&v_full[sp], ((j - 1) >> 1) & 1, 500 + j);
          mbar_wait(&p_ready[sp], ((j - 1) >> 1) & 1, 600 + j);
        }
        __syncwarp();
        tc_fence_after();
        if (elect_one()) {
          if (sp == 0) mma_accum<144, C_ACC, 0, 0, 1>(D, j > 1);
          else mma_accum<144, C_ACC, 128, 1, 1>(D, j > 1);
          umma_commit(&v_free[sp]);
          if (j == nj) umma_commit(acc_done);
        }
        __syncwarp();
      }
    }
  } else {
    const int q4 = warp & 3, cb = warp >> 2;
    const int r = 32 * q4 + lane;
    const int i_tok = it * 128 + r;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q4) << 16);
    const float a = prm.a;
    for (int j = 0; j < nj; ++j) {
      const int s = j & 1;
      mbar_wait_warp(&s_ready[s], (j >> 1) & 1);
      tc_fence_after();
      const bool diag = prm.causal && j == it;
      const bool tail = (j + 1) * 128 > prm.N;
      {
        const int c0 = 32 * cb;
        uint32_t v[32], pk[16];
        tmem_ld32(lane_addr + 128 * s + c0, v);
        tmem_wait_ld();
        // the P of columns [c0, c0+32) goes to tensor-memory columns [c0/2, c0/2+16), which another warp of this
        // quadrant may still be reading as scores: everyone has its scores in registers first
        named_barrier_sync(1, 32 * NCW);
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float x0 = poly<P>(a * __uint_as_float(v[c])), x1 = poly<P>(a * __uint_as_float(v[c + 1]));
          const int col = c0 + c;
          if (diag) x0 = col <= r ? x0 : 0.f, x1 = col + 1 <= r ? x1 : 0.f;
          if (tail) x0 = j * 128 + col < prm.N ? x0 : 0.f, x1 = j * 128 + col + 1 < prm.N ? x1 : 0.f;
          pk[c >> 1] = pack_bf16x2(x0, x1);
        }
        tmem_st16(lane_addr + 128 * s + (c0 >> 1), pk);
      }
      tmem_wait_st();
      tc_fence_before();
      named_barrier_sync(1, 32 * NCW);
      if (tid == 0) mbar_arrive(&p_ready[s]);
    }
    // epilogue: o = acc / den, g = den; rows leave through the Q tile's buffer (all score MMAs are complete)
    mbar_wait_warp(acc_done, 0);
    tc_fence_after();
    uint32_t w[8];
    tmem_ld8(lane_addr + C_ACC + 128, w);
    tmem_wait_ld();
    const float den = __uint_as_float(w[0]);
    const float rden = 1.f / den;
    if (cb == 0 && i_tok < prm.N) prm.g[(size_t)bh * prm.N + i_tok] = den;
    {
      const int hh = cb;
      uint32_t v[32];
      tmem_ld32(lane_addr + C_ACC + 32 * hh, v);
      tmem_wait_ld();
      const uint32_t dst = sbase + (hh >> 1) * PANEL + r * 128;
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        const int ch = (hh & 1) * 4 + j8;
        sts128(dst + ((uint32_t)(ch ^ (r & 7)) << 4),
               pack_bf16x2(__uint_as_float(v[8 * j8]) * rden, __uint_as_float(v[8 * j8 + 1]) * rden),
               pack_bf16x2(__uint_as_float(v[8 * j8 + 2]) * rden, __uint_as_float(v[8 * j8 + 3]) * rden),
               pack_bf16x2(__uint_as_float(v[8 * j8 + 4]) * rden, __uint_as_float(v[8 * j8 + 5]) * rden),
               pack_bf16x2(__uint_as_float(v[8 * j8 + 6]) * rden, __uint_as_float(v[8 * j8 + 7]) * rden));
      }
    }
    fence_proxy_async();
    named_barrier_sync(1, 32 * NCW);
    if (tid == 0) {
      tma_store_3d(&tmO, smem, 0, it * 128, bh);
      tma_store_3d(&tmO, smem + PANEL, 64, it * 128, bh);
      bulk_commit_group();
      bulk_wait_group_read0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCW) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ backward: prep
// u = do / g (bf16), dlt = (do . o) / g: one warp per row, four features per lane
__global__ void __launch_bounds__(256) prep_kernel(const bf16* __restrict__ o, const bf16* __restrict__ d_o,
                                                   const float* __restrict__ g, bf16* __restrict__ u,
                                                   float* __restrict__ dlt, size_t rows, int N, int Npad) {
  const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint2 ov = __ldg(reinterpret_cast<const uint2*>(o + row * 128) + lane);
  const uint2 dv = __ldg(reinterpret_cast<const uint2*>(d_o + row * 128) + lane);
  const float rg = 1.f / __ldg(g + row);
  const float d0 = __uint_as_float(dv.x << 16), d1 = __uint_as_float(dv.x & 0xffff0000u);
  const float d2 = __uint_as_float(dv.y << 16), d3 = __uint_as_float(dv.y & 0xffff0000u);
  float dot = d0 * __uint_as_float(ov.x << 16) + d1 * __uint_as_float(ov.x & 0xffff0000u) +
              d2 * __uint_as_float(ov.y << 16) + d3 * __uint_as_float(ov.y & 0xffff0000u);
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, m);
  reinterpret_cast<uint2*>(u + row * 128)[lane] = make_uint2(pack_bf16x2(d0 * rg, d1 * rg), pack_bf16x2(d2 * rg, d3 * rg));
  if (lane == 0) dlt[(row / N) * Npad + row % N] = dot * rg;
}

// ------------------------------------------------------------------------------------------------ backward: dq
// shared memory: Q tile | U tile | 2 stages x [K tile | V tile] | misc
constexpr int B_OFF_ST = 2 * TILE;
constexpr int B_STAGE = 2 * TILE;
constexpr int B_OFF_MISC = B_OFF_ST + 2 * B_STAGE;
constexpr int B_SMEM = B_OFF_MISC + 2048 + 1024;

template <int P>
__global__ void __launch_bounds__(THREADS, 1)
dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmU,
          const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
          const __grid_constant__ CUtensorMap tmO, const Params prm,
           const __grid_constant__ Descs D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_MISC);
  uint64_t* res_full = bars;
  uint64_t* st_full = bars + 1;    // [2]
  uint64_t* st_free = bars + 3;    // [2]
  uint64_t* sd_ready = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* acc_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y;
  const int it = prm.ntile - 1 - blockIdx.x;
  const int nj = prm.causal ? it + 1 : prm.ntile;
  const uint32_t sbase = smem_u32(smem);

  if (warp == NCW) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(res_full, 1), mbar_init(sd_ready, 1), mbar_init(p_ready, 1), mbar_init(acc_done, 1);
    for (int s = 0; s < 2; ++s) mbar_init(&st_full[s], 1), mbar_init(&st_free[s], 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  assert_bases(tmem, smem);
  constexpr uint32_t C_D = 128, C_ACC = 256;

  if (warp == NCW) {
    if (lane == 0) {
      mbar_arrive_expect_tx(res_full, 2 * TILE);
      load_tile(smem, &tmQ, res_full, it * 128, bh);
      load_tile(smem + TILE, &tmU, res_full, it * 128, bh);
      for (int j = 0; j < nj; ++j) {
        const int s = j & 1;
        uint8_t* st = smem + B_OFF_ST + s * B_STAGE;
        if (j >= 2) mbar_wait(&st_free[s], ((j >> 1) - 1) & 1, 100 + j);
        mbar_arrive_expect_tx(&st_full[s], 2 * TILE);
        load_tile(st, &tmK, &st_full[s], j * 128, bh);
        load_tile(st + TILE, &tmV, &st_full[s], j * 128, bh);
      }
    }
  } else if (warp == NCW + 1) {
    if (lane == 0) mbar_wait(res_full, 0, 300);
    __syncwarp();
    for (int j = 0; j < nj; ++j) {
      const int s = j & 1;
      if (lane == 0) mbar_wait(&st_full[s], (j >> 1) & 1, 400 + j);
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
        if (s == 0) mma_scores<0, 0, 0, 0>(D), mma_scores<C_D, 1, 0, 1>(D);     // S = Q K^T, D = U V^T
        else mma_scores<0, 0, 1, 0>(D), mma_scores<C_D, 1, 1, 1>(D);
        umma_commit(sd_ready);
      }
      __syncwarp();
      if (lane == 0) mbar_wait(p_ready, j & 1, 500 + j);
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
        if (s == 0) mma_accum<128, C_ACC, 0, 0, 0>(D, j > 0);                    // dQ += E K
        else mma_accum<128, C_ACC, 0, 1, 0>(D, j > 0);
        umma_commit(&st_free[s]);
        if (j == nj - 1) umma_commit(acc_done);
      }
      __syncwarp();
    }
  } else {
    const int q4 = warp & 3, cb = warp >> 2;
    const int r = 32 * q4 + lane;
    const int i_tok = it * 128 + r;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q4) << 16);
    const float a = prm.a;
    const float dl = i_tok < prm.N ? prm.dlt[(size_t)bh * prm.Npad + i_tok] : 0.f;
    for (int j = 0; j < nj; ++j) {
      mbar_wait_warp(sd_ready, j & 1);
      tc_fence_after();
      const bool diag = prm.causal && j == it;
      {
        const int c0 = 32 * cb;
        uint32_t sv[32], dv[32], pk[16];
        tmem_ld32(lane_addr + c0, sv);
        tmem_ld32(lane_addr + C_D + c0, dv);
        tmem_wait_ld();
        named_barrier_sync(1, 32 * NCW);       // (E overwrites score columns other warps of the quadrant read)
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float e0 = __uint_as_float(dv[c]) - dl, e1 = __uint_as_float(dv[c + 1]) - dl;
          if (P == 2) e0 *= fmaf(a, __uint_as_float(sv[c]), 1.f), e1 *= fmaf(a, __uint_as_float(sv[c + 1]), 1.f);
          const int col = c0 + c;
          if (diag) e0 = col <= r ? e0 : 0.f, e1 = col + 1 <= r ? e1 : 0.f;
          pk[c >> 1] = pack_bf16x2(e0, e1);
        }
        tmem_st16(lane_addr + (c0 >> 1), pk);
      }
      tmem_wait_st();
      tc_fence_before();
      named_barrier_sync(1, 32 * NCW);
      if (tid == 0) mbar_arrive(p_ready);
    }
    mbar_wait_warp(acc_done, 0);
    tc_fence_after();
    {
      const int hh = cb;
      uint32_t v[32];
      tmem_ld32(lane_addr + C_ACC + 32 * hh, v);
      tmem_wait_ld();
      const uint32_t dst = sbase + (hh >> 1) * PANEL + r * 128;
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        const int ch = (hh & 1) * 4 + j8;
        sts128(dst + ((uint32_t)(ch ^ (r & 7)) << 4),
               pack_bf16x2(__uint_as_float(v[8 * j8]) * a, __uint_as_float(v[8 * j8 + 1]) * a),
               pack_bf16x2(__uint_as_float(v[8 * j8 + 2]) * a, __uint_as_float(v[8 * j8 + 3]) * a),
               pack_bf16x2(__uint_as_float(v[8 * j8 + 4]) * a, __uint_as_float(v[8 * j8 + 5]) * a),
               pack_bf16x2(__uint_as_float(v[8 * j8 + 6]) * a, __uint_as_float(v[8 * j8 + 7]) * a));
      }
    }
    fence_proxy_async();
    named_barrier_sync(1, 32 * NCW);
    if (tid == 0) {
      tma_store_3d(&tmO, smem, 0, it * 128, bh);
      tma_store_3d(&tmO, smem + PANEL, 64, it * 128, bh);
      bulk_commit_group();
      bulk_wait_group_read0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCW) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------ backward: dk, dv
// shared memory: K tile | V tile | 2 stages x [Q tile | U tile] | misc (dlt of the stage's q tokens)
template <int P>
__global__ void __launch_bounds__(THREADS, 1)
dkdv_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
            const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmU,
            const __grid_constant__ CUtensorMap tmDK, const __grid_constant__ CUtensorMap tmDV, const Params prm,
            const __grid_constant__ Descs D) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_MISC);
  uint64_t* res_full = bars;
  uint64_t* st_full = bars + 1;    // [2]
  uint64_t* st_free = bars + 3;    // [2]
  uint64_t* sd_ready = bars + 5;
  uint64_t* p_ready = bars + 6;
  uint64_t* acc_done = bars + 7;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  float* dcol = reinterpret_cast<float*>(smem + B_OFF_MISC + 128);    // [2][128]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y;
  const int jt = blockIdx.x;                                // (causal: the kv tiles with the most q tiles come first)
  const int i0 = prm.causal ? jt : 0;
  const int ni = prm.ntile - i0;
  const uint32_t sbase = smem_u32(smem);

  if (warp == NCW) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(res_full, 1), mbar_init(sd_ready, 1), mbar_init(p_ready, 1), mbar_init(acc_done, 1);
    for (int s = 0; s < 2; ++s) mbar_init(&st_full[s], 1), mbar_init(&st_free[s], 1);
    fence_barrier_init();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  assert_bases(tmem, smem);
  constexpr uint32_t C_D = 128, C_DV = 256, C_DK = 384;

  if (warp == NCW) {
    if (lane == 0) {
      mbar_arrive_expect_tx(res_full, 2 * TILE);
      load_tile(smem, &tmK, res_full, jt * 128, bh);
      load_tile(smem + TILE, &tmV, res_full, jt * 128, bh);
      for (int n = 0; n < ni; ++n) {
        const int s = n & 1;
        uint8_t* st = smem + B_OFF_ST + s * B_STAGE;
        if (n >= 2) mbar_wait(&st_free[s], ((n >> 1) - 1) & 1, 100 + n);
        mbar_arrive_expect_tx(&st_full[s], 2 * TILE);
        load_tile(st, &tmQ, &st_full[s], (i0 + n) * 128, bh);
        load_tile(st + TILE, &tmU, &st_full[s], (i0 + n) * 128, bh);
      }
    }
  } else if (warp == NCW + 1) {
    if (lane == 0) mbar_wait(res_full, 0, 300);
    __syncwarp();
    for (int n = 0; n < ni; ++n) {
      const int s = n & 1;
      if (lane == 0) mbar_wait(&st_full[s], (n >> 1) & 1, 400 + n);
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
        if (s == 0) mma_scores<0, 0, 0, 0>(D), mma_scores<C_D, 1, 0, 1>(D);     // S' = K Q^T, D' = V U^T
        else mma_scores<0, 0, 1, 0>(D), mma_scores<C_D, 1, 1, 1>(D);
        umma_commit(sd_ready);
      }
      __syncwarp();
      if (lane == 0) mbar_wait(p_ready, n & 1, 500 + n);
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
        if (s == 0) mma_accum<128, C_DV, 0, 0, 1>(D, n > 0), mma_accum<128, C_DK, C_D, 0, 0>(D, n > 0);   // dV += P' U, dK += E' Q
        else mma_accum<128, C_DV, 0, 1, 1>(D, n > 0), mma_accum<128, C_DK, C_D, 1, 0>(D, n > 0);
        umma_commit(&st_free[s]);
        if (n == ni - 1) umma_commit(acc_done);
      }
      __syncwarp();
    }
  } else {
    const int q4 = warp & 3, cb = warp >> 2;
    const int r = 32 * q4 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q4) << 16);
    const float a = prm.a;
    for (int n = 0; n < ni; ++n) {
      const int i = i0 + n, s = n & 1;
      const int q_tok = i * 128 + r;
      // (the buffer of two iterations ago was last read before that iteration's P barrier)
      if (cb == 0) dcol[128 * s + r] = q_tok < prm.N ? prm.dlt[(size_t)bh * prm.Npad + q_tok] : 0.f;
      mbar_wait_warp(sd_ready, n & 1);
      tc_fence_after();
      const bool diag = prm.causal && i == jt;
      const float* dc = dcol + 128 * s;
      {
        const int c0 = 32 * cb;
        uint32_t sv[32], dv[32], pp[16], pe[16];
        tmem_ld32(lane_addr + c0, sv);
        tmem_ld32(lane_addr + C_D + c0, dv);
        tmem_wait_ld();
        named_barrier_sync(1, 32 * NCW);       // scores in registers everywhere, dlt of the q tokens visible
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          const int col = c0 + c;
          const float x0 = a * __uint_as_float(sv[c]), x1 = a * __uint_as_float(sv[c + 1]);
          float p0 = poly<P>(x0), p1 = poly<P>(x1);
          float e0 = __uint_as_float(dv[c]) - dc[col], e1 = __uint_as_float(dv[c + 1]) - dc[col + 1];
          if (P == 2) e0 *= 1.f + x0, e1 *= 1.f + x1;
          if (diag) {
            const bool k0 = col >= r, k1 = col + 1 >= r;
            p0 = k0 ? p0 : 0.f, p1 = k1 ? p1 : 0.f, e0 = k0 ? e0 : 0.f, e1 = k1 ? e1 : 0.f;
          }
          pp[c >> 1] = pack_bf16x2(p0, p1);
          pe[c >> 1] = pack_bf16x2(e0, e1);
        }
        tmem_st16(lane_addr + (c0 >> 1), pp);
        tmem_st16(lane_addr + C_D + (c0 >> 1), pe);
      }
      tmem_wait_st();
      tc_fence_before();
      named_barrier_sync(1, 32 * NCW);
      if (tid == 0) mbar_arrive(p_ready);
    }
    mbar_wait_warp(acc_done, 0);
    tc_fence_after();
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {             // dV into the K tile's buffer, dK into the V tile's
      const float sc = which == 0 ? 1.f : a;
      {
        const int hh = cb;
        uint32_t v[32];
        tmem_ld32(lane_addr + (which == 0 ? C_DV : C_DK) + 32 * hh, v);
        tmem_wait_ld();
        const uint32_t dst = sbase + which * TILE + (hh >> 1) * PANEL + r * 128;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          const int ch = (hh & 1) * 4 + j8;
          sts128(dst + ((uint32_t)(ch ^ (r & 7)) << 4),
                 pack_bf16x2(__uint_as_float(v[8 * j8]) * sc, __uint_as_float(v[8 * j8 + 1]) * sc),
                 pack_bf16x2(__uint_as_float(v[8 * j8 + 2]) * sc, __uint_as_float(v[8 * j8 + 3]) * sc),
                 pack_bf16x2(__uint_as_float(v[8 * j8 + 4]) * sc, __uint_as_float(v[8 * j8 + 5]) * sc),
                 pack_bf16x2(__uint_as_float(v[8 * j8 + 6]) * sc, __uint_as_float(v[8 * j8 + 7]) * sc));
        }
      }
    }
    fence_proxy_async();
    named_barrier_sync(1, 128);
    if (tid == 0) {
      tma_store_3d(&tmDV, smem, 0, jt * 128, bh);
      tma_store_3d(&tmDV, smem + PANEL, 64, jt * 128, bh);
      tma_store_3d(&tmDK, smem + TILE, 0, jt * 128, bh);
      tma_store_3d(&tmDK, smem + TILE + PANEL, 64, jt * 128, bh);
      bulk_commit_group();
      bulk_wait_group_read0();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCW) tmem_dealloc(tmem, 512);
}

}  // namespace d128

using namespace d128;

static int tmap128(CUtensorMap* tm, const void* p, int N, int BH) {
  return make_tmap_bf16_3d(tm, p, 128, N, BH, 256, (uint64_t)N * 256, 64, 128);
}

bool dense128_usable(const Problem& pb) {
  return pb.D == 128 && pb.dtype == FASTMAX_DTYPE_BF16 && pb.N >= 128 && !pb.sp_init_fwd && !pb.sp_init_rev;
}

int dense128_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g) {
  const int BH = pb.B * pb.H;
  CUtensorMap tQ, tK, tV, tO;
  int rc;
  if ((rc = tmap128(&tQ, q, pb.N, BH)) || (rc = tmap128(&tK, k, pb.N, BH)) || (rc = tmap128(&tV, v, pb.N, BH)) ||
      (rc = tmap128(&tO, o, pb.N, BH)))
    return rc;
  Params P{};
  P.N = pb.N, P.ntile = (pb.N + 127) / 128, P.p = pb.p, P.causal = pb.causal, P.a = pb.inv_c, P.g = g;
  dim3 grid(P.ntile, BH);
  Descs D{};
  fill_descs(D, 0, 0, F_OFF_ST, F_STAGE, TILE);          // resident Q; stages [K | V | ones]
  if (pb.p == 2) {
    FM_CUDA(cudaFuncSetAttribute(fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    FM_LAUNCH("dense128_fwd", pb.stream, fwd_kernel<2><<<grid, THREADS, F_SMEM, pb.stream>>>(tQ, tK, tV, tO, P, D));
  } else {
    FM_CUDA(cudaFuncSetAttribute(fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    FM_LAUNCH("dense128_fwd", pb.stream, fwd_kernel<1><<<grid, THREADS, F_SMEM, pb.stream>>>(tQ, tK, tV, tO, P, D));
  }
  return FASTMAX_OK;
}

size_t dense128_bwd_workspace_bytes(int B, int H, int N) {
  return align_up((size_t)B * H * N * 256, 1024) + align_up((size_t)B * H * align_up(N, 64) * 4, 1024);
}

int dense128_bwd(const Problem& pb, const void* q, const void* k, const void* v, const void* o, const float* g,
                 const void* d_o, void* dq, void* dk, void* dv, void* ws) {
  const int BH = pb.B * pb.H, Npad = (int)align_up(pb.N, 64);
  bf16* u = (bf16*)ws;
  float* dlt = (float*)((uint8_t*)ws + align_up((size_t)BH * pb.N * 256, 1024));
  const size_t rows = (size_t)BH * pb.N;
  FM_CUDA(cudaMemsetAsync(dlt, 0, (size_t)BH * Npad * 4, pb.stream));
  FM_LAUNCH("dense128_prep", pb.stream,
            prep_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, pb.stream>>>((const bf16*)o, (const bf16*)d_o, g, u, dlt, rows,
                                                                           pb.N, Npad));
  CUtensorMap tQ, tK, tV, tU, tDQ, tDK, tDV;
  int rc;
  if ((rc = tmap128(&tQ, q, pb.N, BH)) || (rc = tmap128(&tK, k, pb.N, BH)) || (rc = tmap128(&tV, v, pb.N, BH)) ||
      (rc = tmap128(&tU, u, pb.N, BH)) || (rc = tmap128(&tDQ, dq, pb.N, BH)) || (rc = tmap128(&tDK, dk, pb.N, BH)) ||
      (rc = tmap128(&tDV, dv, pb.N, BH)))
    return rc;
  Params P{};
  P.N = pb.N, P.Npad = Npad, P.ntile = (pb.N + 127) / 128, P.p = pb.p, P.causal = pb.causal, P.a = pb.inv_c, P.dlt = dlt;
  dim3 grid(P.ntile, BH);
  Descs D{};
  fill_descs(D, 0, TILE, B_OFF_ST, B_STAGE, TILE);       // resident [Q | U] resp. [K | V]; stages [K | V] resp. [Q | U]
  if (pb.p == 2) {
    FM_CUDA(cudaFuncSetAttribute(dq_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    FM_CUDA(cudaFuncSetAttribute(dkdv_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    FM_LAUNCH("dense128_dq", pb.stream, dq_kernel<2><<<grid, THREADS, B_SMEM, pb.stream>>>(tQ, tU, tK, tV, tDQ, P, D));
    FM_LAUNCH("dense128_dkdv", pb.stream, dkdv_kernel<2><<<grid, THREADS, B_SMEM, pb.stream>>>(tK, tV, tQ, tU, tDK, tDV, P, D));
  } else {
    FM_CUDA(cudaFuncSetAttribute(dq_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    FM_CUDA(cudaFuncSetAttribute(dkdv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    FM_LAUNCH("dense128_dq", pb.stream, dq_kernel<1><<<grid, THREADS, B_SMEM, pb.stream>>>(tQ, tU, tK, tV, tDQ, P, D));
    FM_LAUNCH("dense128_dkdv", pb.stream, dkdv_kernel<1><<<grid, THREADS, B_SMEM, pb.stream>>>(tK, tV, tQ, tU, tDK, tDV, P, D));
  }
  return FASTMAX_OK;
}

}  // namespace fm
