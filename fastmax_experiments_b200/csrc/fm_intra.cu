// fm_intra.cu -- intra-segment part of the FastMax chunked scan on tcgen05 (bf16, D = 64).
//
// Inside one segment the causal polynomial scores are evaluated directly, flash-attention style
// but with no running max / exp: P = f(Q K^T / c) o mask needs only two FMAs per score.
//
//   MODE_FWD : O_i  += f(S) V_j ,  g_i += rowsum f(S)          S = Q_i K_j^T
//   MODE_DQ  : dQ_i += E K_j ,     E = f'(S) o (U_i V_j^T - dlt_i)           (unscaled by 1/c)
//   MODE_DKDV: dV_j += P'^T-form:  P' = f(S') , E' = f'(S') o (V_j U_i^T - dlt_i),  S' = K_j Q_i^T
//              dV_j += P' U_i ,  dK_j += E' Q_i                               (dK unscaled by 1/c)
//
// with U = dO/g, dlt = (dO.O)/g (fm_moments.cu prep kernel).  These are the j-in-segment terms of
// compute_F_masked / compute_g_masked (fastmax.py:219-322) and gradient_o_{q,k,v}_masked
// (fastmax.py:433-691); the out-of-segment terms come from the moment states.
//
// One CTA owns four 128-row output tiles (two for DKDV) of one (b,h) and streams 64-token tiles of
// the other operand pair through a TMA ring.  Two compute warpgroups each own half of the output
// tiles; per (output tile, stream tile) pair a warpgroup's elected thread issues the score MMAs
// (operands from shared memory) into tensor memory, the 128 threads read their score row, apply the
// polynomial + mask, write the bf16 result back to tensor memory as the A operand, and the elected
// thread issues the second MMAs that accumulate into the output tile's fp32 accumulator.
// Outputs are fp32 partials in the workspace (the apply kernel adds the moment part).
#include "fm_common.cuh"
#include "fm_sm100.cuh"
#include "fm_moments.cuh"
#include <type_traits>

namespace fm {
namespace mom {
using namespace sm100;
typedef __nv_bfloat16 bf16;

enum IntraMode { INTRA_FWD = 0, INTRA_DQ = 1, INTRA_DKDV = 2 };

struct IntraParams {
  int N, Npad, L, p;
  float a;                 // 1/c
  const float* dlt;        // (BH, Npad) fp32, zero padded
  float* out0;             // FWD: O partial; DQ: dQ partial; DKDV: dV partial   (BH,N,64) fp32
  float* out1;             // DKDV: dK partial
  float* g;                // FWD: normaliser (BH,N): intra rowsum + moment part of the segment boundary
  const uint8_t* state;    // boundary states whose normaliser column is folded in here (FWD, DQ: forward-
                           // direction states; DKDV: reverse-direction states), nullable
  int nseg;
  int fused;               // p = 1: the first-order state term (slab 64) is applied here as well and the final
                           // bf16 outputs are written (no apply kernels, no fp32 partials)
  int noncausal;           // no intra-segment part at all: only the normaliser column of the one state per (b,h)
                           // is applied (g; the -dlt(N1 + a N2 q) part of dq; the RN1 + a RN2 k part of dk)
  int edge_state;          // sequence parallel: the first (FWD, DQ) / last (DKDV) local segment has a state too
  float count_offset;      // tokens of earlier ranks (normaliser count term)
  int H, Hkv;              // grouped-query attention: query heads / kv heads per batch entry (fm_common.cuh kv_index)
};

template <int MODE> struct IntraCfg {
  static constexpr int OUT_TILES = MODE == INTRA_DKDV ? 2 : 4;
  static constexpr int RES_OPS = MODE == INTRA_FWD ? 1 : 2;      // resident operand tensors
  static constexpr int STAGES = MODE == INTRA_DQ ? 4 : 6;
  static constexpr int STAGE_BYTES = 2 * 8192 + (MODE == INTRA_DKDV ? 256 : 0);
  static constexpr int STAGE_STRIDE = 2 * 8192 + 1024;            // keeps every tile 1024-aligned
  static constexpr int RES_BYTES = OUT_TILES * RES_OPS * 16384;
  static constexpr int NORM_BYTES = 8192 + 1024;   // slab 65 of the boundary state (bf16)
  static constexpr int SMEM = RES_BYTES + STAGES * STAGE_STRIDE + NORM_BYTES + 1024 + 1024;
};

// Shared-memory descriptors of every MMA operand, built on the host (fm_sm100.cuh: operands that come
// from the constant bank let one thread issue at the tensor pipe's rate).
struct IntraDescs {
  unsigned long long resK[8][4];        // resident tile (R1 tiles, then R2 tiles) x k-step, K-major
  unsigned long long strK[6][2][4];     // ring stage x operand (T1, T2) x k-step, K-major
  unsigned long long strMN[6][2][4];    // same tiles read MN-major (B of the accumulating MMAs)
  unsigned long long normK[4];          // normaliser slab (p = 1 fused: the first-order slab), K-major
  unsigned long long normMN[4];         // same tile read MN-major
};

template <int MODE> __host__ __device__ constexpr int intra_otile(int w, int x) {
  return MODE == INTRA_DKDV ? w : (w == 0 ? (x == 0 ? 0 : 3) : (x == 0 ? 1 : 2));
}
template <int MODE> __host__ __device__ constexpr uint32_t intra_bar0() {
  return FM_SMEM_BASE + IntraCfg<MODE>::RES_BYTES + IntraCfg<MODE>::STAGES * IntraCfg<MODE>::STAGE_STRIDE +
         IntraCfg<MODE>::NORM_BYTES;
}

// score MMAs of one (output tile, stream tile) pair: S = R1_o . T1^T  (and dP = R2_o . T2^T)
template <int MODE, int W, int X>
__device__ __forceinline__ void intra_issue_scores(const IntraDescs& D, int st) {
  constexpr int O = intra_otile<MODE>(W, X);
  constexpr uint32_t IDK = make_idesc_bf16(128, 64, false, false);
  constexpr uint32_t SB = 256 + 128 * W;
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ss(SB, D.resK[O][kk], D.strK[st][0][kk], IDK, kk > 0);
  if constexpr (MODE != INTRA_FWD) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
      umma_ss(SB + 64, D.resK[IntraCfg<MODE>::OUT_TILES + O][kk], D.strK[st][1][kk], IDK, kk > 0);
  }
  umma_commit_addr(intra_bar0<MODE>() + 8 * (17 + W));   // s_full[W]
}

// accumulating MMAs of the pair (A = bf16 scores in tensor memory)
template <int MODE, int W, int X>
__device__ __forceinline__ void intra_issue_acc(const IntraDescs& D, int st) {
  constexpr int O = intra_otile<MODE>(W, X);
  constexpr uint32_t IDN = make_idesc_bf16(128, 64, false, true);
  constexpr uint32_t SB = 256 + 128 * W;
  if constexpr (MODE == INTRA_FWD) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) umma_ts(64 * O, SB + 8 * kk, D.strMN[st][1][kk], IDN, 1);
  } else if constexpr (MODE == INTRA_DQ) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) umma_ts(64 * O, SB + 8 * kk, D.strMN[st][0][kk], IDN, 1);
  } else {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) umma_ts(128 * O, SB + 8 * kk, D.strMN[st][1][kk], IDN, 1);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) umma_ts(128 * O + 64, SB + 32 + 8 * kk, D.strMN[st][0][kk], IDN, 1);
  }
  umma_commit_addr(intra_bar0<MODE>() + 8 * (19 + W));   // s_free[W]
}

// normaliser column: T_o = R1_o . N2 into the two halves of the warpgroup's score buffer
template <int MODE, int W, int X>
__device__ __forceinline__ void intra_issue_norm(const IntraDescs& D) {
  constexpr int O = intra_otile<MODE>(W, X);
  constexpr uint32_t IDK = make_idesc_bf16(128, 64, false, false);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ss(256 + 128 * W + 64 * X, D.resK[O][kk], D.normK[kk], IDK, kk > 0);
}

#define FM_INTRA_CALL_WX(FN, ...)                          \
  do {                                                     \
    if (w == 0) {                                          \
      if (x == 0) FN<MODE, 0, 0>(__VA_ARGS__);             \
      else FN<MODE, 0, 1>(__VA_ARGS__);                    \
    } else {                                               \
      if (x == 0) FN<MODE, 1, 0>(__VA_ARGS__);             \
      else FN<MODE, 1, 1>(__VA_ARGS__);                    \
    }                                                      \
  } while (0)

template <int MODE>
__global__ void __launch_bounds__(384, 1)
intra_kernel(const __grid_constant__ CUtensorMap tmR1, const __grid_constant__ CUtensorMap tmR2,
             const __grid_constant__ CUtensorMap tmT1, const __grid_constant__ CUtensorMap tmT2,
             const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmB0,
             const __grid_constant__ CUtensorMap tmB1, const __grid_constant__ IntraDescs DS, const IntraParams P) {
  using C = IntraCfg<MODE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* res = smem;                         // [RES_OPS][OUT_TILES][16 KiB]
  uint8_t* ring = smem + C::RES_BYTES;         // STAGES x (T1 8 KiB | T2 8 KiB | dlt 256 B)
  uint8_t* norm_s = ring + C::STAGES * C::STAGE_STRIDE;   // [8 KiB normaliser matrix | 256 B vector]
  uint64_t* bars = reinterpret_cast<uint64_t*>(norm_s + C::NORM_BYTES);
  uint64_t* full = bars;              // [STAGES]
  uint64_t* empty = bars + 8;         // [STAGES]   count 2 (one arrival per warpgroup)
  uint64_t* res_full = bars + 16;     // [1]
  uint64_t* s_full = bars + 17;       // [2] score tile of warpgroup w is in tensor memory
  uint64_t* s_free = bars + 19;       // [2] second MMAs of the previous pair have drained the buffer
  uint64_t* acc_done = bars + 21;     // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y;
  constexpr int CT = C::OUT_TILES * 128;       // tokens of the output chunk
  const int c0 = blockIdx.x * CT;              // first output row of the CTA
  const int seg_start = (c0 / P.L) * P.L;
  const int seg_end = min(seg_start + P.L, P.N);
  // stream range (absolute token index of the first stream tile, number of 64-token tiles)
  const int st0 = MODE == INTRA_DKDV ? c0 : seg_start;
  const int st1 = MODE == INTRA_DKDV ? seg_end : min(c0 + CT, P.N);
  const int nstream = P.noncausal ? 0 : (st1 - st0 + 63) / 64;
  // moment part of the normaliser column (FWD: g, DQ: the -dlt (N1 + a N2 q) term): only segments
  // that have a boundary state
  const int seg = c0 / P.L;
  const bool has_norm =
      P.state != nullptr &&
      (P.noncausal || (MODE == INTRA_DKDV ? (seg < P.nseg - 1 || P.edge_state) : (seg > 0 || P.edge_state)));
  // grouped queries: k, v and the forward states belong to the (b, kv head); q, u, dlt, the reverse states and every
  // output to the query head.  FWD / DQ: resident = q (, u), streamed = k, v; DKDV: resident = k, v, streamed = q, u
  const int kvh = kv_index(bh, P.H, P.Hkv);
  const int bh_res = MODE == INTRA_DKDV ? kvh : bh, bh_str = MODE == INTRA_DKDV ? bh : kvh;
  const int bh_st = MODE == INTRA_DKDV ? bh : kvh;
  const int rec_idx = P.noncausal ? bh_st : bh_st * P.nseg + seg;

  if (warp == 0) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    for (int i = 0; i < C::STAGES; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], 2);
    mbar_init(res_full, 1);
    for (int i = 0; i < 2; ++i) mbar_init(&s_full[i], 1), mbar_init(&s_free[i], 1), mbar_init(&acc_done[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmR1);
    tma_prefetch_desc(&tmT1);
    tma_prefetch_desc(&tmT2);
  }
  // fp32 vector part of the boundary state ([0,64): S0 / R0, [64,128): N1 / RN1), read by every epilogue row
  float* vec_s = reinterpret_cast<float*>(norm_s + 8192);
  if (tid < 128)
    vec_s[tid] = has_norm ? reinterpret_cast<const float*>(P.state + (size_t)rec_idx * REC_BYTES + VEC_OFF)[tid] : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  assert_bases(tmem, smem);
  // the output accumulators start at zero; every accumulating MMA then has accumulate = 1.  Only the
  // compute warpgroups wait for that (each zeroes half of the columns, so they wait for each other); the
  // producer warp starts its loads right away.
  if (warp >= 4) {
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = 0u;
    const uint32_t la = ((uint32_t)(32 * (warp & 3)) << 16) + 128 * ((warp - 4) >> 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) tmem_st32(la + 32 * i, z);
    tmem_wait_st();
    tc_fence_before();
    asm volatile("bar.sync 4, 256;" ::: "memory");
    tc_fence_after();
  }

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(res_full, C::RES_BYTES + (has_norm ? 8192 : 0));
      if (has_norm) tma_load_3d(norm_s, &tmS, res_full, 0, (P.fused ? 64 : 65) * 64, rec_idx);
      for (int o = 0; o < C::OUT_TILES; ++o)
        for (int h = 0; h < 2; ++h) {
          tma_load_3d(res + o * 16384 + h * 8192, &tmR1, res_full, 0, c0 + o * 128 + h * 64, bh_res);
          if (C::RES_OPS == 2)
            tma_load_3d(res + (C::OUT_TILES + o) * 16384 + h * 8192, &tmR2, res_full, 0, c0 + o * 128 + h * 64, bh_res);
        }
      for (int jt = 0; jt < nstream; ++jt) {
        const int st = jt % C::STAGES, ph = (jt / C::STAGES) & 1;
        mbar_wait(&empty[st], ph ^ 1);
        uint8_t* sp = ring + st * C::STAGE_STRIDE;
        mbar_arrive_expect_tx(&full[st], C::STAGE_BYTES);
        tma_load_3d(sp, &tmT1, &full[st], 0, st0 + jt * 64, bh_str);
        tma_load_3d(sp + 8192, &tmT2, &full[st], 0, st0 + jt * 64, bh_str);
        if (MODE == INTRA_DKDV) {
          const float* src = P.dlt + (size_t)bh * P.Npad + st0 + jt * 64;
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                  smem_u32(sp + 16384)),
              "l"(src), "r"(256), "r"(smem_u32(&full[st]))
              : "memory");
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- compute warpgroups
    const int w = (warp - 4) >> 2, q = warp & 3;
    const bool issuer = (warp & 3) == 0 && lane == 0;
    const int rloc = 32 * q + lane;                        // row inside a 128-row output tile
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
    const uint32_t sbuf = lane_addr + 256 + 128 * w;      // score buffer of this warpgroup (128 cols)
    // output tiles of this warpgroup
    // (an expression, not an array: a dynamically indexed local array lives in local memory)
    auto otile_of = [&](int x) { return intra_otile<MODE>(w, x); };
    int n_ot;
    n_ot = MODE == INTRA_DKDV ? 1 : 2;

    const float a = P.a, c2 = P.p == 2 ? 0.5f * P.a * P.a : 0.f, a1 = P.p == 2 ? P.a : 0.f;
    float gsum[2] = {0.f, 0.f};
    float dlt_row[2] = {0.f, 0.f};
    if constexpr (MODE == INTRA_DQ) {
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        const int r = c0 + otile_of(x) * 128 + rloc;
        dlt_row[x] = r < P.N ? P.dlt[(size_t)bh * P.Npad + r] : 0.f;
      }
    }
    const uint32_t id_kk = make_idesc_bf16(128, 64, false, false);
    const uint32_t id_kn = make_idesc_bf16(128, 64, false, true);

    mbar_wait(res_full, 0);
    int npair = 0;
    for (int jt = 0; jt < nstream; ++jt) {
      const int st = jt % C::STAGES, ph = (jt / C::STAGES) & 1;
      const int ts = st0 + jt * 64;                       // absolute index of the first stream token
      uint8_t* sp = ring + st * C::STAGE_STRIDE;
      bool used = false;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (x >= n_ot) break;
        const int o = otile_of(x);
        const int ob = c0 + o * 128;                      // first row of the output tile
        // does output tile o interact with this stream tile?
        bool need;
        if (MODE == INTRA_DKDV) need = ts + 63 >= ob;     // query tile reaches the key tile
        else need = ts <= ob + 127;                       // key tile starts at or before the last row
        if (ob >= P.N) need = false;
        if (!need) continue;
        if (!used) {
          mbar_wait(&full[st], ph);
          used = true;
        }
        // ---- score MMAs (issued by the warpgroup's first warp, converged, one elected lane)
        if (q == 0) {
          if (lane == 0 && npair > 0) mbar_wait(&s_free[w], (npair - 1) & 1);
          __syncwarp();
          tc_fence_after();
          if (elect_one()) FM_INTRA_CALL_WX(intra_issue_scores, DS, st);
          __syncwarp();
        }
        mbar_wait(&s_full[w], npair & 1);
        tc_fence_after();
        // ---- polynomial on CUDA cores: thread = output row, 64 stream columns
        const int r_abs = ob + rloc;
        const bool diag = MODE == INTRA_DKDV ? (ts < ob + 128) : (ts + 63 > ob);   // tile needs the mask
        uint32_t pk[32];
        uint32_t pe[MODE == INTRA_DKDV ? 32 : 1];
        // two copies of the element loop: interior tiles skip the causal mask (compare + select per score)
        auto chunk = [&](auto diag_c, int h) {
          constexpr bool DIAG = decltype(diag_c)::value;
          uint32_t sv[32];
          tmem_ld32(sbuf + 32 * h, sv);
          uint32_t dv[MODE == INTRA_FWD ? 1 : 32];
          if constexpr (MODE != INTRA_FWD) tmem_ld32(sbuf + 64 + 32 * h, dv);
          tmem_wait_ld();
          float dl[MODE == INTRA_DKDV ? 32 : 1];
          if constexpr (MODE == INTRA_DKDV) {
            const float4* dsrc = reinterpret_cast<const float4*>(sp + 16384) + 8 * h;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              float4 t4 = dsrc[c4];
              dl[4 * c4] = t4.x, dl[4 * c4 + 1] = t4.y, dl[4 * c4 + 2] = t4.z, dl[4 * c4 + 3] = t4.w;
            }
          }
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float r0, r1, e0 = 0.f, e1 = 0.f;
            const float s0 = __uint_as_float(sv[c]), s1 = __uint_as_float(sv[c + 1]);
            // f(s) = 1 + a s + c2 s^2 (c2 = 0 for p = 1);  f'(s)/a = 1 + a1 s (a1 = 0 for p = 1)
            if constexpr (MODE == INTRA_FWD) {
              r0 = fmaf(s0, fmaf(c2, s0, a), 1.f);
              r1 = fmaf(s1, fmaf(c2, s1, a), 1.f);
            } else if constexpr (MODE == INTRA_DQ) {
              const float d0 = __uint_as_float(dv[c]) - dlt_row[x], d1 = __uint_as_float(dv[c + 1]) - dlt_row[x];
              r0 = fmaf(a1, s0, 1.f) * d0;
              r1 = fmaf(a1, s1, 1.f) * d1;
            } else {
              const float d0 = __uint_as_float(dv[c]) - dl[c], d1 = __uint_as_float(dv[c + 1]) - dl[c + 1];
              r0 = fmaf(s0, fmaf(c2, s0, a), 1.f);
              r1 = fmaf(s1, fmaf(c2, s1, a), 1.f);
              e0 = fmaf(a1, s0, 1.f) * d0;
              e1 = fmaf(a1, s1, 1.f) * d1;
            }
            if constexpr (DIAG) {
              const int col = ts + 32 * h + c;
              bool v0, v1;
              if constexpr (MODE == INTRA_DKDV) v0 = col >= r_abs, v1 = col + 1 >= r_abs;   // query >= key
              else v0 = col <= r_abs, v1 = col + 1 <= r_abs;                                // key <= query
              r0 = v0 ? r0 : 0.f, r1 = v1 ? r1 : 0.f;
              e0 = v0 ? e0 : 0.f, e1 = v1 ? e1 : 0.f;
            }
            if constexpr (MODE == INTRA_FWD) gsum[x] += r0 + r1;
            pk[16 * h + (c >> 1)] = pack_bf16x2(r0, r1);
            if constexpr (MODE == INTRA_DKDV) pe[16 * h + (c >> 1)] = pack_bf16x2(e0, e1);
          }
        };
        if (diag) {
          chunk(std::true_type{}, 0);
          chunk(std::true_type{}, 1);
        } else {
          chunk(std::false_type{}, 0);
          chunk(std::false_type{}, 1);
        }
        // ---- result back to tensor memory as the A operand of the accumulating MMAs
        tmem_st32(sbuf, pk);
        if constexpr (MODE == INTRA_DKDV) tmem_st32(sbuf + 32, pe);
        tmem_wait_st();
        tc_fence_before();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
        if (q == 0) {
          tc_fence_after();
          if (elect_one()) FM_INTRA_CALL_WX(intra_issue_acc, DS, st);
          __syncwarp();
        }
        ++npair;
      }
      // release the ring stage: both warpgroups arrive, with or without work on it.  A warpgroup
      // that skips the tile still waits for it to land first, so that its arrival cannot be counted
      // towards the previous use of the stage.
      if (issuer) {
        if (used) umma_commit(&empty[st]);
        else {
          mbar_wait(&full[st], ph);
          mbar_arrive(&empty[st]);
        }
      }
    }
    // ---- moment part of the normaliser column: T_o = R1_o . N2 (one more score-like MMA per output
    // tile, into the two halves of the warpgroup's score buffer)
    if (has_norm && q == 0) {
      if (lane == 0 && npair > 0) mbar_wait(&s_free[w], (npair - 1) & 1);
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
        if (!P.fused) {
#pragma unroll
          for (int x = 0; x < 2; ++x) {
            if (x >= n_ot) break;
            FM_INTRA_CALL_WX(intra_issue_norm, DS);
          }
        } else {
          // p = 1: first-order state term of each output tile into the (drained) score buffer.
          //   FWD : T  = Q . S1        (contract the slab's rows: MN-major B)        -> columns 64x
          //   DQ  : T' = U . S1^T      (contract its columns: K-major B, A = R2 = U)  -> columns 64x
          //   DKDV: T  = K . R1 -> columns [0,64)  and  T' = V . R1^T -> columns [64,128)
          constexpr uint32_t IDK64 = make_idesc_bf16(128, 64, false, false);
          constexpr uint32_t IDN64 = make_idesc_bf16(128, 64, false, true);
          const uint32_t sb = 256 + 128 * w;
          if constexpr (MODE == INTRA_DKDV) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_ss(sb, DS.resK[w][kk], DS.normMN[kk], IDN64, kk > 0);
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_ss(sb + 64, DS.resK[C::OUT_TILES + w][kk], DS.normK[kk], IDK64, kk > 0);
          } else {
            for (int x = 0; x < 2; ++x) {
              const int o = otile_of(x);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                if constexpr (MODE == INTRA_FWD) umma_ss(sb + 64 * x, DS.resK[o][kk], DS.normMN[kk], IDN64, kk > 0);
                else umma_ss(sb + 64 * x, DS.resK[C::OUT_TILES + o][kk], DS.normK[kk], IDK64, kk > 0);
              }
            }
          }
        }
      }
      __syncwarp();
    }
    if (issuer) umma_commit(&acc_done[w]);
    mbar_wait(&acc_done[w], 0);
    tc_fence_after();
    if (P.fused) {
      // ---------------------------------------------------------------- epilogue, p = 1: final outputs
      //   FWD : o  = (O + a T + S0) / g,   g = rowsum + count + a q.N1
      //   DQ  : dq = a (O + T' - dlt N1)
      //   DKDV: dv = O_dv + a T + R0 ;  dk = a (O_dk + T' + RN1)
      // (state terms only where the segment has a boundary state), written as bf16 through the free stream
      // ring with bulk tensor stores.
      asm volatile("bar.sync 3, 256;" ::: "memory");
      uint8_t* stg = ring + w * 32768;
      const bool storer = q == 0 && lane == 0;
      int nstored = 0;
#pragma unroll
      for (int x = 0; x < 2; ++x) {
        if (x >= n_ot) break;
        const int o = otile_of(x);
        const int ob = c0 + o * 128;
        const int r_abs = ob + rloc;
        if (ob >= P.N) continue;                               // uniform over the warpgroup
        const size_t grow = (size_t)bh * P.N + r_abs;
        float sc0 = 1.f;
        if constexpr (MODE == INTRA_FWD) {
          float gmom = 0.f;
          if (has_norm) {
            const uint8_t* qrow = res + o * 16384 + rloc * 128;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
              const uint4 qv = *reinterpret_cast<const uint4*>(qrow + ((c8 ^ (rloc & 7)) << 4));
              const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                gmom = fmaf(__uint_as_float(qw[j] << 16), vec_s[64 + 8 * c8 + 2 * j], gmom);
                gmom = fmaf(__uint_as_float(qw[j] & 0xffff0000u), vec_s[64 + 8 * c8 + 2 * j + 1], gmom);
              }
            }
          }
          const float count = P.noncausal ? (float)P.N : (float)(seg * P.L) + P.count_offset;
          const float gv = gsum[x] + (has_norm ? count + a * gmom : 0.f);
          if (r_abs < P.N) P.g[grow] = gv;
          sc0 = 1.f / gv;
        } else if constexpr (MODE == INTRA_DQ) {
          sc0 = a;
        }
        const int nacc = MODE == INTRA_DKDV ? 2 : 1;
        for (int a2 = 0; a2 < nacc; ++a2) {
          const uint32_t col = MODE == INTRA_DKDV ? 128 * o + 64 * a2 : 64 * o;
          const uint32_t tcol = MODE == INTRA_DKDV ? 64 * a2 : 64 * x;
          const float sc = (MODE == INTRA_DKDV && a2 == 1) ? a : sc0;
          if (nstored > 0) {       // the previous tile's stores have read the staging area
            if (storer) bulk_wait_group_read0();
            asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
          }
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32], tn[32];
            tmem_ld32(lane_addr + col + 32 * h, v);
            if (has_norm) tmem_ld32(sbuf + tcol + 32 * h, tn);
            tmem_wait_ld();
            uint32_t pk[16];
#pragma unroll
            for (int c = 0; c < 32; c += 2) {
              float r[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const int idx = 32 * h + c + e;
                float val = __uint_as_float(v[c + e]);
                if (has_norm) {
                  const float t = __uint_as_float(tn[c + e]);
                  if constexpr (MODE == INTRA_FWD) val += fmaf(a, t, vec_s[idx]);
                  else if constexpr (MODE == INTRA_DQ) val += t - dlt_row[x] * vec_s[64 + idx];
                  else val += a2 == 0 ? fmaf(a, t, vec_s[idx]) : t + vec_s[64 + idx];
                }
                r[e] = val * sc;
              }
              pk[c >> 1] = pack_bf16x2(r[0], r[1]);
            }
            const uint32_t dst = smem_u32(stg) + rloc * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              sts128(dst + ((uint32_t)((4 * h + j) ^ (rloc & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
          }
          fence_proxy_async();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
          if (storer) {
            const CUtensorMap* tm = a2 == 0 ? &tmB0 : &tmB1;
            tma_store_3d(tm, stg, 0, ob, bh);                         // rows past N are clipped by the tensor map
            if (ob + 64 < P.N) tma_store_3d(tm, stg + 8192, 0, ob + 64, bh);
            bulk_commit_group();
          }
          ++nstored;
        }
      }
      if (storer) bulk_wait_group_read0();   // shared memory must outlive the bulk stores' reads (not their global writes)
    } else {
    // ---------------------------------------------------------------- epilogue: fp32 partials
    // Written through shared memory with bulk tensor stores: a thread owns a 256-byte output row, so
    // storing from registers means 16-byte pieces at a 256-byte stride (measured: a third of the CTA's
    // time).  Once both warpgroups' MMAs are complete the stream ring is free; each warpgroup stages one
    // 128x64 fp32 tile there in the blocked partial layout and one thread sends it as a 32 KiB bulk copy.
    asm volatile("bar.sync 3, 256;" ::: "memory");
    uint8_t* stg = ring + w * 32768;
    const bool storer = q == 0 && lane == 0;
    int nstored = 0;
#pragma unroll
    for (int x = 0; x < 2; ++x) {
      if (x >= n_ot) break;
      const int o = otile_of(x);
      const int r_abs = c0 + o * 128 + rloc;
      const bool tile_live = c0 + o * 128 < P.N;          // uniform over the warpgroup
      if (!tile_live) continue;
      const size_t grow = (size_t)bh * P.N + r_abs;
      const int nacc = MODE == INTRA_DKDV ? 2 : 1;
      float gmom = 0.f;   // FWD: a q.N1 + (a^2/2) q.T
      for (int a2 = 0; a2 < nacc; ++a2) {
        const uint32_t col = MODE == INTRA_DKDV ? 128 * o + 64 * a2 : 64 * o;
        const bool store = (a2 == 0 ? P.out0 : P.out1) != nullptr;   // non-causal: some partials do not exist
        if (!store && MODE != INTRA_FWD) continue;                   // (FWD still needs the pass for g)
        if (store && nstored > 0) {       // the previous tile's stores have read the staging area
          if (storer) bulk_wait_group_read0();
          asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
        }
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld32(lane_addr + col + 32 * h, v);
          if constexpr (MODE == INTRA_DKDV) {
            if (has_norm && a2 == 1) {   // dk_part[l] += RN1[l] + a (RN2 k)[l]
              uint32_t tn[32];
              tmem_ld32(sbuf + 32 * h, tn);
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < 32; ++c)
                v[c] = __float_as_uint(__uint_as_float(v[c]) +
                                       fmaf(a1, __uint_as_float(tn[c]), vec_s[64 + 32 * h + c]));
            }
          } else {
            if (has_norm) {   // CTA-uniform
              uint32_t tn[32];
              tmem_ld32(sbuf + 64 * x + 32 * h, tn);
              tmem_wait_ld();
              if constexpr (MODE == INTRA_DQ) {
                // dq_part[l] -= dlt (N1[l] + a T[l])
#pragma unroll
                for (int c = 0; c < 32; ++c)
                  v[c] = __float_as_uint(__uint_as_float(v[c]) -
                                         dlt_row[x] * fmaf(a1, __uint_as_float(tn[c]), vec_s[64 + 32 * h + c]));
              } else {
                // g += a q.N1 + (a^2/2) q.T  with the thread's own q row read back from the resident tile
                const uint8_t* qrow = res + o * 16384 + rloc * 128;
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                  const uint4 qv = *reinterpret_cast<const uint4*>(qrow + (((4 * h + c8) ^ (rloc & 7)) << 4));
                  const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const int c = 8 * c8 + 2 * j;
                    const float q0 = __uint_as_float(qw[j] << 16), q1 = __uint_as_float(qw[j] & 0xffff0000u);
                    const float c2n = c2;
                    gmom += q0 * fmaf(c2n, __uint_as_float(tn[c]), a * vec_s[64 + 32 * h + c]);
                    gmom += q1 * fmaf(c2n, __uint_as_float(tn[c + 1]), a * vec_s[64 + 32 * h + c + 1]);
                  }
                }
              }
            }
          }
          tmem_wait_ld();
          if (store) {   // staging in the partial's own layout [c4][row] (fm_moments.cuh): conflict-free 512-byte rows
            const uint32_t dst = smem_u32(stg) + (uint32_t)(8 * h) * 2048 + rloc * 16;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4)
              sts128(dst + c4 * 2048, v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
          }
        }
        if (!store) continue;
        fence_proxy_async();
        asm volatile("bar.sync %0, 128;" ::"r"(1 + w) : "memory");
        if (storer) {   // one contiguous 32 KiB block per tile (the buffer is padded to whole tiles)
          float* outp = a2 == 0 ? P.out0 : P.out1;
          const size_t tile = (size_t)bh * ((P.N + 127) / 128) + (size_t)(c0 / 128 + o);
          bulk_store_1d(outp + tile * PART_TILE_FLOATS, stg, 32768);
          bulk_commit_group();
        }
        ++nstored;
      }
      if (MODE == INTRA_FWD && r_abs < P.N) {
        const float count = P.noncausal ? (float)P.N : (float)(seg * P.L) + P.count_offset;
        P.g[grow] = gsum[x] + (has_norm ? count + gmom : 0.f);
      }
    }
    if (storer) bulk_wait_group_read0();   // shared memory must outlive the bulk stores' reads (not their global writes)
    }   // !fused
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

}  // namespace mom

using namespace mom;

// R1/R2: resident operands (rows = output rows), T1/T2: streamed operands.  All (BH, N, 64) bf16.
template <int MODE>
static int launch_intra_t(const Problem& pb, int L, int Npad, const void* r1, const void* r2, const void* t1,
                          const void* t2, const float* dlt, float* out0, float* out1, float* g,
                          const uint8_t* state, const char* name, void* final0 = nullptr, void* final1 = nullptr) {
  using C = IntraCfg<MODE>;
  const int BH = pb.B * pb.H, BHkv = pb.B * pb.kv_heads();
  const int BHres = MODE == INTRA_DKDV ? BHkv : BH, BHstr = MODE == INTRA_DKDV ? BH : BHkv;
  CUtensorMap m[4];
  const void* ptr[4] = {r1, r2 ? r2 : r1, t1, t2};
  for (int i = 0; i < 4; ++i) {
    int rc = make_tmap_bf16_3d(&m[i], ptr[i], 64, pb.N, i < 2 ? BHres : BHstr, 128, (uint64_t)pb.N * 128, 64, 64);
    if (rc) return rc;
  }
  IntraParams P{};
  P.N = pb.N, P.Npad = Npad, P.L = L, P.p = pb.p, P.a = pb.inv_c, P.dlt = dlt, P.out0 = out0, P.out1 = out1, P.g = g;
  P.noncausal = !pb.causal;
  P.fused = pb.p == 1 && final0 != nullptr;
  P.nseg = pb.causal ? (pb.N + L - 1) / L : 1;
  P.edge_state = MODE == INTRA_DKDV ? pb.sp_init_rev != nullptr : pb.sp_init_fwd != nullptr;
  P.count_offset = (float)pb.sp_offset;
  P.H = pb.H, P.Hkv = pb.kv_heads();
  P.state = (P.nseg > 1 || P.edge_state || P.noncausal) ? state : nullptr;
  CUtensorMap tmS = m[0];
  if (P.state) {
    int rc = make_tmap_bf16_3d(&tmS, state, 64, NSLAB * 64, (uint64_t)(MODE == INTRA_DKDV ? BH : BHkv) * P.nseg, 128,
                               REC_BYTES, 64, 64);
    if (rc) return rc;
  }
  CUtensorMap tmB0 = m[0], tmB1 = m[0];   // bf16 outputs of the fused p = 1 mode (fp32 partials are plain bulk copies)
  if (P.fused) {
    int rc = make_tmap_bf16_3d(&tmB0, final0, 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 64);
    if (rc) return rc;
    tmB1 = tmB0;
    if (final1 && (rc = make_tmap_bf16_3d(&tmB1, final1, 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 64))) return rc;
  }
  static_assert(C::STAGES * C::STAGE_STRIDE >= 2 * 32768, "the epilogue stages one fp32 tile per warpgroup in the ring");
  auto kern = intra_kernel<MODE>;
  FM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM));
  constexpr int CT = C::OUT_TILES * 128;
  dim3 grid((pb.N + CT - 1) / CT, BH);
  IntraDescs DS{};
  {
    const uint32_t base = sm100::FM_SMEM_BASE;
    for (int i = 0; i < C::OUT_TILES * C::RES_OPS; ++i)
      for (int kk = 0; kk < 4; ++kk) DS.resK[i][kk] = sm100::desc_kmajor_sw128_const(base + i * 16384) + kk * 2;
    for (int st = 0; st < C::STAGES; ++st)
      for (int op = 0; op < 2; ++op)
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t a = base + C::RES_BYTES + st * C::STAGE_STRIDE + op * 8192;
          DS.strK[st][op][kk] = sm100::desc_kmajor_sw128_const(a) + kk * 2;
          DS.strMN[st][op][kk] = sm100::desc_mnmajor_sw128_const(a, 8192) + kk * 128;
        }
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t na = base + C::RES_BYTES + C::STAGES * C::STAGE_STRIDE;
      DS.normK[kk] = sm100::desc_kmajor_sw128_const(na) + kk * 2;
      DS.normMN[kk] = sm100::desc_mnmajor_sw128_const(na, 8192) + kk * 128;
    }
  }
  FM_LAUNCH(name, pb.stream, kern<<<grid, 384, C::SMEM, pb.stream>>>(m[0], m[1], m[2], m[3], tmS, tmB0, tmB1, DS, P));
  return FASTMAX_OK;
}

int intra_fwd(const Problem& pb, int L, int Npad, const void* q, const void* k, const void* v, float* o_part,
              float* g, const uint8_t* state, void* o_final) {
  return launch_intra_t<INTRA_FWD>(pb, L, Npad, q, nullptr, k, v, nullptr, o_part, nullptr, g, state, "mom_intra_fwd",
                                   o_final);
}
int intra_dq(const Problem& pb, int L, int Npad, const void* q, const void* u, const void* k, const void* v,
             const float* dlt, float* dq_part, const uint8_t* state, void* dq_final) {
  return launch_intra_t<INTRA_DQ>(pb, L, Npad, q, u, k, v, dlt, dq_part, nullptr, nullptr, state, "mom_intra_dq",
                                  dq_final);
}
int intra_dkdv(const Problem& pb, int L, int Npad, const void* k, const void* v, const void* q, const void* u,
               const float* dlt, float* dk_part, float* dv_part, const uint8_t* rstate, void* dk_final, void* dv_final) {
  return launch_intra_t<INTRA_DKDV>(pb, L, Npad, k, v, q, u, dlt, dv_part, dk_part, nullptr, rstate, "mom_intra_dkdv",
                                    dv_final, dk_final);
}

}  // namespace fm
