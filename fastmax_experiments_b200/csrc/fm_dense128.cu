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
//
// SPLIT = true runs fp32 problems of head size <= 64 through the same kernels at fp32-level accuracy: every fp32 operand x
// is stored as the bf16 pair [hi | lo] (hi = bf16(x), lo = bf16(x - hi): 16 mantissa bits) -- exactly the two-panel tile of
// a 128-wide head -- and each product A.B is taken as hi.hi + hi.lo + lo.hi on the tensor pipe with fp32 accumulation
// (three MMA groups instead of two).  The polynomial results P / E are split the same way when they go back to tensor
// memory, the normaliser is summed in fp32 on CUDA cores, outputs are written as fp32.
#include "fm_common.cuh"
#include "fm_sm100.cuh"
#include <type_traits>

#ifdef FM_TRACE
// developer build: event times of CTA (0,0) of the forward kernel, [event][tile]
static __device__ long long fm_trace_d128[8][64];
extern "C" int fastmax_debug_trace_fetch_d128(long long* out) {
  return cudaMemcpyFromSymbol(out, fm_trace_d128, sizeof(fm_trace_d128)) == cudaSuccess ? 0 : -4;
}
#define FM_TRD(ev, j) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (j) < 64) fm_trace_d128[ev][j] = clock64(); } while (0)
#else
#define FM_TRD(ev, j) ((void)0)
#endif

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
  float *out0, *out1;        // SPLIT: fp32 outputs (BH, N, Dout): o | dq | dv, dk
  int Dout;
};

// Shared-memory descriptors of every MMA operand, built on the host for the fixed layouts below and passed in the
// constant bank: with immediate tensor-memory addresses (the CTA owns all 512 columns: base 0) one thread then issues
// at the tensor pipe's rate (fm_sm100.cuh).
struct Descs {
  unsigned long long resK[2][8];      // resident tile 0 / 1, K-major, 8 k-steps of 16 features
  unsigned long long strK[2][2][8];   // stage x streamed tile 0 / 1, K-major
  unsigned long long strMN[2][2][8];  // the same tiles read MN-major (tokens = K), 8 k-steps of 16 tokens
};

// S[tmem DT] = res[RA] . str[ST][OB]^T.  SPLIT: panels are [hi | lo] of a 64-wide operand: hi.hi + hi.lo + lo.hi
template <bool SPLIT, uint32_t DT, int RA, int ST, int OB>
__device__ __forceinline__ void mma_scores(const Descs& D) {
  constexpr uint32_t ID = make_idesc_bf16(128, 128, false, false);
  if constexpr (!SPLIT) {
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) umma_ss(DT, D.resK[RA][kk], D.strK[ST][OB][kk], ID, kk > 0);
  } else {
#pragma unroll
    for (int t = 0; t < 12; ++t) {
      const int ka = (t < 8 ? 0 : 4) + (t & 3), kb = (t >= 4 && t < 8 ? 4 : 0) + (t & 3);
      umma_ss(DT, D.resK[RA][ka], D.strK[ST][OB][kb], ID, t > 0);
    }
  }
}
// acc[tmem DT] += P[tmem PT, bf16 128 x 128] . str[ST][OB] read MN-major, N columns (128, or 144 with the ones panel).
// SPLIT: acc[:, 0:128) += P_hi . [B_hi | B_lo], acc[:, 0:64) += P_lo . B_hi (P_lo in tensor-memory columns PT + 64 ...)
template <bool SPLIT, int N, uint32_t DT, uint32_t PT, int ST, int OB>
__device__ __forceinline__ void mma_accum(const Descs& D, uint32_t acc_first) {
  constexpr uint32_t ID = make_idesc_bf16(128, SPLIT ? 128 : N, false, true);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) umma_ts(DT, PT + 8 * kk, D.strMN[ST][OB][kk], ID, (acc_first || kk > 0) ? 1 : 0);
  if constexpr (SPLIT) {
    constexpr uint32_t ID64 = make_idesc_bf16(128, 64, false, true);
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) umma_ts(DT, PT + 64 + 8 * kk, D.strMN[ST][OB][kk], ID64, 1);
  }
}

// x -> bf16 pair (hi, lo)
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
  hi = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
  lo = pack_bf16x2(x0 - __bfloat162float(h0), x1 - __bfloat162float(h1));
}

// fp32 rows of a SPLIT epilogue: columns [32 cb2, +32) of (accA + accB) * scale -> out[(bh N + tok) Dout + ...]
__device__ __forceinline__ void store_rows_f32(float* out, size_t row, int Dout, int cb2, const uint32_t* va,
                                               const uint32_t* vb, float scale) {
  float* dst = out + row * Dout + 32 * cb2;
  if (Dout == 64) {
#pragma unroll
    for (int c = 0; c < 32; c += 4)
      *reinterpret_cast<float4*>(dst + c) =
          make_float4((__uint_as_float(va[c]) + __uint_as_float(vb[c])) * scale,
                      (__uint_as_float(va[c + 1]) + __uint_as_float(vb[c + 1])) * scale,
                      (__uint_as_float(va[c + 2]) + __uint_as_float(vb[c + 2])) * scale,
                      (__uint_as_float(va[c + 3]) + __uint_as_float(vb[c + 3])) * scale);
  } else {
#pragma unroll
    for (int c = 0; c < 32; ++c)
      if (32 * cb2 + c < Dout) dst[c] = (__uint_as_float(va[c]) + __uint_as_float(vb[c])) * scale;
  }
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

// f(a s) straight from the raw score: 1 + a s (+ (a^2 / 2) s^2), one or two FMAs (c1 = a, c2 = a^2 / 2)
template <int P>
__device__ __forceinline__ float poly(float s, float c1, float c2) {
  return P == 2 ? fmaf(fmaf(c2, s, c1), s, 1.f) : fmaf(c1, s, 1.f);
}

// ------------------------------------------------------------------------------------------------ forward
// shared memory: Q tile | 2 stages x [K tile | V tile | ones panel] | misc
constexpr int F_STAGE = 2 * TILE + PANEL;
constexpr int F_OFF_ST = TILE;
constexpr int F_OFF_MISC = F_OFF_ST + 2 * F_STAGE;
constexpr int F_SMEM = F_OFF_MISC + 4096 + 1024;

template <int P, bool SPLIT>
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
  float* rs_sm = reinterpret_cast<float*>(smem + F_OFF_MISC + 128);   // [4][128] row sums per column block (SPLIT)
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
        if (lane == 0) {
          mbar_wait(&k_full[s], (j >> 1) & 1, 400 + j);
          FM_TRD(0, j);
        }
        __syncwarp();
        tc_fence_after();
        if (elect_one()) {
          if (s == 0) mma_scores<SPLIT, 0, 0, 0, 0>(D);
          else mma_scores<SPLIT, 128, 0, 1, 0>(D);
          umma_commit(&k_free[s]);
          umma_commit(&s_ready[s]);
          FM_TRD(1, j);
        }
        __syncwarp();
      }
      if (j > 0) {
        const int sp = (j - 1) & 1;
        if (lane == 0) {
          mbar_wait(&v_full[sp], ((j - 1) >> 1) & 1, 500 + j);
          mbar_wait(&p_ready[sp], ((j - 1) >> 1) & 1, 600 + j);
          FM_TRD(2, j - 1);
        }
        __syncwarp();
        tc_fence_after();
        if (elect_one()) {
          if (sp == 0) mma_accum<SPLIT, 144, C_ACC, 0, 0, 1>(D, j > 1);
          else mma_accum<SPLIT, 144, C_ACC, 128, 1, 1>(D, j > 1);
          umma_commit(&v_free[sp]);
          if (j == nj) umma_commit(acc_done);
          FM_TRD(3, j - 1);
        }
        __syncwarp();
      }
    }
  } else {
    const int q4 = warp & 3, cb = warp >> 2;
    const int r = 32 * q4 + lane;
    const int i_tok = it * 128 + r;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q4) << 16);
    const float a = prm.a, c2 = 0.5f * a * a;
    float rsum = 0.f;
    for (int j = 0; j < nj; ++j) {
      const int s = j & 1;
      mbar_wait_warp(&s_ready[s], (j >> 1) & 1);
      tc_fence_after();
      if (tid == 0) FM_TRD(4, j);
      const bool diag = prm.causal && j == it;
      const bool tail = (j + 1) * 128 > prm.N;
      {
        const int c0 = 32 * cb;
        uint32_t v[32], pk[16], pl[16];
        tmem_ld32(lane_addr + 128 * s + c0, v);
        tmem_wait_ld();
        // the P of columns [c0, c0+32) goes to tensor-memory columns [c0/2, c0/2+16), which another warp of this
        // quadrant may still be reading as scores: everyone has its scores in registers first
        named_barrier_sync(1, 32 * NCW);
        if (tid == 0) FM_TRD(5, j);
        // (interior tiles take the loop without the mask selects: the polynomial is ~1300 issue cycles per tile pair
        //  and scheduler otherwise, more than the pair's MMAs)
        auto body = [&](auto masked) {
          constexpr bool M = decltype(masked)::value;
          const int lim = M ? min(diag ? r : 127, prm.N - 1 - j * 128) : 0;     // last column this row keeps
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float x0 = poly<P>(__uint_as_float(v[c]), a, c2), x1 = poly<P>(__uint_as_float(v[c + 1]), a, c2);
            if (M) x0 = c0 + c <= lim ? x0 : 0.f, x1 = c0 + c + 1 <= lim ? x1 : 0.f;
            if constexpr (SPLIT) {
              rsum += x0 + x1;
              split2(x0, x1, pk[c >> 1], pl[c >> 1]);
            } else {
              pk[c >> 1] = pack_bf16x2(x0, x1);
            }
          }
        };
        if (diag || tail) body(std::true_type{});
        else body(std::false_type{});
        tmem_st16(lane_addr + 128 * s + (c0 >> 1), pk);
        if constexpr (SPLIT) tmem_st16(lane_addr + 128 * s + 64 + (c0 >> 1), pl);
      }
      tmem_wait_st();
      tc_fence_before();
      named_barrier_sync(1, 32 * NCW);
      if (tid == 0) mbar_arrive(&p_ready[s]);
      if (tid == 0) FM_TRD(6, j);
    }
    // epilogue: o = acc / den, g = den; rows leave through the Q tile's buffer (all score MMAs are complete)
    mbar_wait_warp(acc_done, 0);
    tc_fence_after();
    float den;
    if constexpr (SPLIT) {
      rs_sm[128 * cb + r] = rsum;
      named_barrier_sync(1, 32 * NCW);
      den = rs_sm[r] + rs_sm[128 + r] + rs_sm[256 + r] + rs_sm[384 + r];
    } else {
      uint32_t w[8];
      tmem_ld8(lane_addr + C_ACC + 128, w);
      tmem_wait_ld();
      den = __uint_as_float(w[0]);
    }
    const float rden = 1.f / den;
    if (cb == 0 && i_tok < prm.N) prm.g[(size_t)bh * prm.N + i_tok] = den;
    if constexpr (SPLIT) {
      if (cb < 2) {
        uint32_t va[32], vb[32];
        tmem_ld32(lane_addr + C_ACC + 32 * cb, va);
        tmem_ld32(lane_addr + C_ACC + 64 + 32 * cb, vb);
        tmem_wait_ld();
        if (i_tok < prm.N) store_rows_f32(prm.out0, (size_t)bh * prm.N + i_tok, prm.Dout, cb, va, vb, rden);
      }
    } else {
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
    if (!SPLIT && tid == 0) {
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

// ------------------------------------------------------------------------------------------------ fp32 -> [hi | lo]
// (rows, D <= 64) fp32 -> (rows, 128) bf16: columns [0, 64) = hi, [64, 128) = lo, zero beyond D; one warp per row
__global__ void __launch_bounds__(256) split_kernel(const float* __restrict__ x, bf16* __restrict__ out, size_t rows,
                                                    int D, const float* __restrict__ g) {
  const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float rg = g ? 1.f / __ldg(g + row) : 1.f;
  const float x0 = 2 * lane < D ? __ldg(x + row * D + 2 * lane) * rg : 0.f;
  const float x1 = 2 * lane + 1 < D ? __ldg(x + row * D + 2 * lane + 1) * rg : 0.f;
  uint32_t hi, lo;
  split2(x0, x1, hi, lo);
  reinterpret_cast<uint32_t*>(out + row * 128)[lane] = hi;
  reinterpret_cast<uint32_t*>(out + row * 128 + 64)[lane] = lo;
}
// dlt = (do . o) / g for fp32 rows
__global__ void __launch_bounds__(256) dlt_f32_kernel(const float* __restrict__ o, const float* __restrict__ d_o,
                                                      const float* __restrict__ g, float* __restrict__ dlt, size_t rows,
                                                      int D, int N, int Npad) {
  const size_t row = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  float dot = 0.f;
  for (int c = lane; c < D; c += 32) dot += __ldg(o + row * D + c) * __ldg(d_o + row * D + c);
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, m);
  if (lane == 0) dlt[(row / N) * Npad + row % N] = dot / __ldg(g + row);
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

template <int P, bool SPLIT>
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
        if (s == 0) mma_scores<SPLIT, 0, 0, 0, 0>(D), mma_scores<SPLIT, C_D, 1, 0, 1>(D);     // S = Q K^T, D = U V^T
        else mma_scores<SPLIT, 0, 0, 1, 0>(D), mma_scores<SPLIT, C_D, 1, 1, 1>(D);
        umma_commit(sd_ready);
      }
      __syncwarp();
      if (lane == 0) mbar_wait(p_ready, j & 1, 500 + j);
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
        if (s == 0) mma_accum<SPLIT, 128, C_ACC, 0, 0, 0>(D, j > 0);             // dQ += E K
        else mma_accum<SPLIT, 128, C_ACC, 0, 1, 0>(D, j > 0);
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
        uint32_t sv[32], dv[32], pk[16], pl[SPLIT ? 16 : 1];
        tmem_ld32(lane_addr + c0, sv);
        tmem_ld32(lane_addr + C_D + c0, dv);
        tmem_wait_ld();
        named_barrier_sync(1, 32 * NCW);       // (E overwrites score columns other warps of the quadrant read)
        auto body = [&](auto masked) {
          constexpr bool M = decltype(masked)::value;
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float e0 = __uint_as_float(dv[c]) - dl, e1 = __uint_as_float(dv[c + 1]) - dl;
            if (P == 2) e0 *= fmaf(a, __uint_as_float(sv[c]), 1.f), e1 *= fmaf(a, __uint_as_float(sv[c + 1]), 1.f);
            if (M) e0 = c0 + c <= r ? e0 : 0.f, e1 = c0 + c + 1 <= r ? e1 : 0.f;
            if constexpr (SPLIT) split2(e0, e1, pk[c >> 1], pl[c >> 1]);
            else pk[c >> 1] = pack_bf16x2(e0, e1);
          }
        };
        if (diag) body(std::true_type{});
        else body(std::false_type{});
        tmem_st16(lane_addr + (c0 >> 1), pk);
        if constexpr (SPLIT) tmem_st16(lane_addr + 64 + (c0 >> 1), pl);
      }
      tmem_wait_st();
      tc_fence_before();
      named_barrier_sync(1, 32 * NCW);
      if (tid == 0) mbar_arrive(p_ready);
    }
    mbar_wait_warp(acc_done, 0);
    tc_fence_after();
    if constexpr (SPLIT) {
      if (cb < 2) {
        uint32_t va[32], vb[32];
        tmem_ld32(lane_addr + C_ACC + 32 * cb, va);
        tmem_ld32(lane_addr + C_ACC + 64 + 32 * cb, vb);
        tmem_wait_ld();
        if (i_tok < prm.N) store_rows_f32(prm.out0, (size_t)bh * prm.N + i_tok, prm.Dout, cb, va, vb, a);
      }
    } else {
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
    if (!SPLIT && tid == 0) {
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
template <int P, bool SPLIT>
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
        if (s == 0) mma_scores<SPLIT, 0, 0, 0, 0>(D), mma_scores<SPLIT, C_D, 1, 0, 1>(D);     // S' = K Q^T, D' = V U^T
        else mma_scores<SPLIT, 0, 0, 1, 0>(D), mma_scores<SPLIT, C_D, 1, 1, 1>(D);
        umma_commit(sd_ready);
      }
      __syncwarp();
      if (lane == 0) mbar_wait(p_ready, n & 1, 500 + n);
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
        if (s == 0) mma_accum<SPLIT, 128, C_DV, 0, 0, 1>(D, n > 0), mma_accum<SPLIT, 128, C_DK, C_D, 0, 0>(D, n > 0);   // dV += P' U, dK += E' Q
        else mma_accum<SPLIT, 128, C_DV, 0, 1, 1>(D, n > 0), mma_accum<SPLIT, 128, C_DK, C_D, 1, 0>(D, n > 0);
        umma_commit(&st_free[s]);
        if (n == ni - 1) umma_commit(acc_done);
      }
      __syncwarp();
    }
  } else {
    const int q4 = warp & 3, cb = warp >> 2;
    const int r = 32 * q4 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q4) << 16);
    const float a = prm.a, c2 = 0.5f * a * a;
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
        uint32_t lp[SPLIT ? 16 : 1], le[SPLIT ? 16 : 1];
        tmem_ld32(lane_addr + c0, sv);
        tmem_ld32(lane_addr + C_D + c0, dv);
        tmem_wait_ld();
        named_barrier_sync(1, 32 * NCW);       // scores in registers everywhere, dlt of the q tokens visible
        auto body = [&](auto masked) {
          constexpr bool M = decltype(masked)::value;
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            const int col = c0 + c;
            const float s0 = __uint_as_float(sv[c]), s1 = __uint_as_float(sv[c + 1]);
            float p0 = poly<P>(s0, a, c2), p1 = poly<P>(s1, a, c2);
            const float2 dd = *reinterpret_cast<const float2*>(dc + col);
            float e0 = __uint_as_float(dv[c]) - dd.x, e1 = __uint_as_float(dv[c + 1]) - dd.y;
            if (P == 2) e0 *= fmaf(a, s0, 1.f), e1 *= fmaf(a, s1, 1.f);
            if (M) {
              const bool k0 = col >= r, k1 = col + 1 >= r;
              p0 = k0 ? p0 : 0.f, p1 = k1 ? p1 : 0.f, e0 = k0 ? e0 : 0.f, e1 = k1 ? e1 : 0.f;
            }
            if constexpr (SPLIT) {
              split2(p0, p1, pp[c >> 1], lp[c >> 1]);
              split2(e0, e1, pe[c >> 1], le[c >> 1]);
            } else {
              pp[c >> 1] = pack_bf16x2(p0, p1);
              pe[c >> 1] = pack_bf16x2(e0, e1);
            }
          }
        };
        if (diag) body(std::true_type{});
        else body(std::false_type{});
        tmem_st16(lane_addr + (c0 >> 1), pp);
        tmem_st16(lane_addr + C_D + (c0 >> 1), pe);
        if constexpr (SPLIT) {
          tmem_st16(lane_addr + 64 + (c0 >> 1), lp);
          tmem_st16(lane_addr + C_D + 64 + (c0 >> 1), le);
        }
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
      if constexpr (SPLIT) {
        const int kv_tok = jt * 128 + r;
        if (cb < 2) {
          uint32_t va[32], vb[32];
          tmem_ld32(lane_addr + (which == 0 ? C_DV : C_DK) + 32 * cb, va);
          tmem_ld32(lane_addr + (which == 0 ? C_DV : C_DK) + 64 + 32 * cb, vb);
          tmem_wait_ld();
          if (kv_tok < prm.N)
            store_rows_f32(which == 0 ? prm.out0 : prm.out1, (size_t)bh * prm.N + kv_tok, prm.Dout, cb, va, vb, sc);
        }
      } else {
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
    named_barrier_sync(1, 32 * NCW);
    if (!SPLIT && tid == 0) {
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
// fp32, head size <= 64: the same kernels on [hi | lo] bf16 pairs
bool dense_split_usable(const Problem& pb) {
  return pb.D <= 64 && pb.dtype == FASTMAX_DTYPE_F32 && pb.N >= 128 && !pb.sp_init_fwd && !pb.sp_init_rev;
}

static size_t split_tensor_bytes(int B, int H, int N) { return align_up((size_t)B * H * N * 256, 1024); }
size_t dense_split_fwd_workspace_bytes(int B, int H, int N) { return 3 * split_tensor_bytes(B, H, N); }
size_t dense_split_bwd_workspace_bytes(int B, int H, int N) {
  return 4 * split_tensor_bytes(B, H, N) + align_up((size_t)B * H * align_up(N, 64) * 4, 1024);
}
size_t dense128_bwd_workspace_bytes(int B, int H, int N) {
  return align_up((size_t)B * H * N * 256, 1024) + align_up((size_t)B * H * align_up(N, 64) * 4, 1024);
}

template <bool SPLIT>
static int launch_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g) {
  const int BH = pb.B * pb.H;
  CUtensorMap tQ, tK, tV, tO;
  int rc;
  if ((rc = tmap128(&tQ, q, pb.N, BH)) || (rc = tmap128(&tK, k, pb.N, BH)) || (rc = tmap128(&tV, v, pb.N, BH))) return rc;
  if ((rc = tmap128(&tO, SPLIT ? q : o, pb.N, BH))) return rc;          // (SPLIT writes fp32 rows itself)
  Params P{};
  P.N = pb.N, P.ntile = (pb.N + 127) / 128, P.p = pb.p, P.causal = pb.causal, P.a = pb.inv_c, P.g = g;
  P.out0 = SPLIT ? (float*)o : nullptr, P.Dout = pb.D;
  dim3 grid(P.ntile, BH);
  Descs D{};
  fill_descs(D, 0, 0, F_OFF_ST, F_STAGE, TILE);          // resident Q; stages [K | V | ones]
  const char* name = SPLIT ? "dense_split_fwd" : "dense128_fwd";
  if (pb.p == 2) {
    FM_CUDA(cudaFuncSetAttribute(fwd_kernel<2, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    FM_LAUNCH(name, pb.stream, (fwd_kernel<2, SPLIT><<<grid, THREADS, F_SMEM, pb.stream>>>(tQ, tK, tV, tO, P, D)));
  } else {
    FM_CUDA(cudaFuncSetAttribute(fwd_kernel<1, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    FM_LAUNCH(name, pb.stream, (fwd_kernel<1, SPLIT><<<grid, THREADS, F_SMEM, pb.stream>>>(tQ, tK, tV, tO, P, D)));
  }
  return FASTMAX_OK;
}

int dense128_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g) {
  return launch_fwd<false>(pb, q, k, v, o, g);
}

static int launch_split(const Problem& pb, const void* x, void* out, const float* g, const char* name) {
  const size_t rows = (size_t)pb.B * pb.H * pb.N;
  FM_LAUNCH(name, pb.stream,
            split_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, pb.stream>>>((const float*)x, (bf16*)out, rows, pb.D, g));
  return FASTMAX_OK;
}

int dense_split_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g, void* ws) {
  const size_t tb = split_tensor_bytes(pb.B, pb.H, pb.N);
  uint8_t* w8 = (uint8_t*)ws;
  int rc;
  if ((rc = launch_split(pb, q, w8, nullptr, "dense_split_q")) || (rc = launch_split(pb, k, w8 + tb, nullptr, "dense_split_k")) ||
      (rc = launch_split(pb, v, w8 + 2 * tb, nullptr, "dense_split_v")))
    return rc;
  return launch_fwd<true>(pb, w8, w8 + tb, w8 + 2 * tb, o, g);
}

template <bool SPLIT>
static int launch_bwd(const Problem& pb, int Npad, const void* q, const void* k, const void* v, const void* u,
                      const float* dlt, void* dq, void* dk, void* dv) {
  const int BH = pb.B * pb.H;
  CUtensorMap tQ, tK, tV, tU, tDQ, tDK, tDV;
  int rc;
  if ((rc = tmap128(&tQ, q, pb.N, BH)) || (rc = tmap128(&tK, k, pb.N, BH)) || (rc = tmap128(&tV, v, pb.N, BH)) ||
      (rc = tmap128(&tU, u, pb.N, BH)) || (rc = tmap128(&tDQ, SPLIT ? q : dq, pb.N, BH)) ||
      (rc = tmap128(&tDK, SPLIT ? q : dk, pb.N, BH)) || (rc = tmap128(&tDV, SPLIT ? q : dv, pb.N, BH)))
    return rc;
  Params P{};
  P.N = pb.N, P.Npad = Npad, P.ntile = (pb.N + 127) / 128, P.p = pb.p, P.causal = pb.causal, P.a = pb.inv_c, P.dlt = dlt;
  P.Dout = pb.D;
  Params Pq = P, Pkv = P;
  if (SPLIT) Pq.out0 = (float*)dq, Pkv.out0 = (float*)dv, Pkv.out1 = (float*)dk;
  dim3 grid(P.ntile, BH);
  Descs D{};
  fill_descs(D, 0, TILE, B_OFF_ST, B_STAGE, TILE);       // resident [Q | U] resp. [K | V]; stages [K | V] resp. [Q | U]
  const char* n1 = SPLIT ? "dense_split_dq" : "dense128_dq";
  const char* n2 = SPLIT ? "dense_split_dkdv" : "dense128_dkdv";
  if (pb.p == 2) {
    FM_CUDA(cudaFuncSetAttribute(dq_kernel<2, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    FM_CUDA(cudaFuncSetAttribute(dkdv_kernel<2, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    FM_LAUNCH(n1, pb.stream, (dq_kernel<2, SPLIT><<<grid, THREADS, B_SMEM, pb.stream>>>(tQ, tU, tK, tV, tDQ, Pq, D)));
    FM_LAUNCH(n2, pb.stream, (dkdv_kernel<2, SPLIT><<<grid, THREADS, B_SMEM, pb.stream>>>(tK, tV, tQ, tU, tDK, tDV, Pkv, D)));
  } else {
    FM_CUDA(cudaFuncSetAttribute(dq_kernel<1, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    FM_CUDA(cudaFuncSetAttribute(dkdv_kernel<1, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    FM_LAUNCH(n1, pb.stream, (dq_kernel<1, SPLIT><<<grid, THREADS, B_SMEM, pb.stream>>>(tQ, tU, tK, tV, tDQ, Pq, D)));
    FM_LAUNCH(n2, pb.stream, (dkdv_kernel<1, SPLIT><<<grid, THREADS, B_SMEM, pb.stream>>>(tK, tV, tQ, tU, tDK, tDV, Pkv, D)));
  }
  return FASTMAX_OK;
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
  return launch_bwd<false>(pb, Npad, q, k, v, u, dlt, dq, dk, dv);
}

int dense_split_bwd(const Problem& pb, const void* q, const void* k, const void* v, const void* o, const float* g,
                    const void* d_o, void* dq, void* dk, void* dv, void* ws) {
  const int BH = pb.B * pb.H, Npad = (int)align_up(pb.N, 64);
  const size_t tb = split_tensor_bytes(pb.B, pb.H, pb.N);
  uint8_t* w8 = (uint8_t*)ws;
  float* dlt = (float*)(w8 + 4 * tb);
  const size_t rows = (size_t)BH * pb.N;
  int rc;
  FM_CUDA(cudaMemsetAsync(dlt, 0, (size_t)BH * Npad * 4, pb.stream));
  FM_LAUNCH("dense_split_dlt", pb.stream,
            dlt_f32_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, pb.stream>>>((const float*)o, (const float*)d_o, g, dlt, rows,
                                                                              pb.D, pb.N, Npad));
  if ((rc = launch_split(pb, q, w8, nullptr, "dense_split_q")) || (rc = launch_split(pb, k, w8 + tb, nullptr, "dense_split_k")) ||
      (rc = launch_split(pb, v, w8 + 2 * tb, nullptr, "dense_split_v")) ||
      (rc = launch_split(pb, d_o, w8 + 3 * tb, g, "dense_split_u")))
    return rc;
  return launch_bwd<true>(pb, Npad, w8, w8 + tb, w8 + 2 * tb, w8 + 3 * tb, dlt, dq, dk, dv);
}

}  // namespace fm
