// fm_seg.cu -- intra-segment forward part of the p = 2 chunked scan, second generation (bf16, D = 64).
//
// Same contract as intra_kernel<INTRA_FWD> of fm_intra.cu (fp32 partial numerators in the blocked layout of
// fm_moments.cuh, the normaliser g including the moment part of the segment's boundary state) but organised like the
// dense tiles of fm_dense128.cu, which measured 2-3x fewer cycles per 128 x 128 score tile:
//
//   * ONE chain per CTA instead of one per warpgroup: all 16 compute warps (four per scheduler) work on the same
//     128 x 128 score tile, thread = one row x 32 columns, so the polynomial's dependent chains overlap;
//   * scores double-buffered in tensor memory and issued one tile ahead by a dedicated issuer warp, so the score MMA,
//     its completion and the wake-up run under the previous tile's polynomial;
//   * 128-token stream tiles (half as many hand-overs per score as the 64-token tiles of fm_intra.cu), and a stream
//     tile is used for all the CTA's q tiles that need it before the next one is touched.
//
// A CTA owns four 128-row q tiles (512 tokens: one segment at the usual L = 512) with their fp32 accumulators in tensor
// memory (4 x 64 columns) next to the two score buffers (2 x 128).
#include "fm_common.cuh"
#include "fm_sm100.cuh"
#include "fm_moments.cuh"

#ifdef FM_TRACE
// developer build: event times of one CTA, [event][index]
static __device__ long long fm_trace_seg[16][16];
extern "C" int fastmax_debug_trace_fetch_seg(long long* out) {
  return cudaMemcpyFromSymbol(out, fm_trace_seg, sizeof(fm_trace_seg)) == cudaSuccess ? 0 : -4;
}
static __device__ long long fm_trace_seg2[2][16][16];
extern "C" int fastmax_debug_trace_fetch_seg2(long long* out) {
  return cudaMemcpyFromSymbol(out, fm_trace_seg2, sizeof(fm_trace_seg2)) == cudaSuccess ? 0 : -4;
}
#define FM_TRW(ev, w, j) do { if (blockIdx.x == 5 && blockIdx.y == 100 && (j) < 16 && (threadIdx.x & 31) == 0) fm_trace_seg2[ev][w][j] = clock64(); } while (0)
#define FM_TRS(ev, j) do { if (blockIdx.x == 5 && blockIdx.y == 100 && (j) < 16) fm_trace_seg[ev][j] = clock64(); } while (0)
#else
#define FM_TRW(ev, w, j) ((void)0)
#define FM_TRS(ev, j) ((void)0)
#endif

namespace fm {
namespace seg {
using namespace sm100;
using namespace mom;
typedef __nv_bfloat16 bf16;

constexpr int NCW = 16;
constexpr int THREADS = (NCW + 2) * 32;
constexpr int QT = 4;
constexpr int NST = 3;
constexpr int OFF_Q = 0;
constexpr int OFF_RING = QT * 16384;
constexpr int STAGE = 32768;                       // K tile | V tile (128 tokens each)
constexpr int OFF_NORM = OFF_RING + NST * STAGE;   // slab 65 of the boundary state
constexpr int OFF_MISC = OFF_NORM + 8192;
constexpr int SMEM = OFF_MISC + 12288 + 1024;

struct Descs {
  unsigned long long qK[QT][4];       // q tile, K-major
  unsigned long long kK[NST][4];      // stage K tile, K-major (128 rows)
  unsigned long long vMN[NST][8];     // stage V tile read MN-major, 8 k-steps of 16 tokens
  unsigned long long normK[4];        // normaliser slab, K-major
};

struct Params {
  int N, L, nseg, edge_state;
  int H, Hkv;                // grouped-query attention (fm_common.cuh kv_index): k, v and the states per (b, kv head)
  float a, count_offset;
  float* out;                // blocked fp32 partial
  float* g;
  const uint8_t* state;      // nullable
};

template <int T, int ST, int B>
__device__ __forceinline__ void issue_scores(const Descs& D) {
  constexpr uint32_t ID = make_idesc_bf16(128, 128, false, false);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ss(128 * B, D.qK[T][kk], D.kK[ST][kk], ID, kk > 0);
}
template <int T, int ST, int B>
__device__ __forceinline__ void issue_pv(const Descs& D, uint32_t acc) {
  constexpr uint32_t ID = make_idesc_bf16(128, 64, false, true);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) umma_ts(256 + 64 * T, 128 * B + 8 * kk, D.vMN[ST][kk], ID, (acc || kk > 0) ? 1 : 0);
}
#define FM_SEG_DISPATCH(FN, t, st, b, ...)                                                          \
  do {                                                                                              \
    switch ((t) * 6 + (st) * 2 + (b)) {                                                             \
      case 0: FN<0, 0, 0>(__VA_ARGS__); break;   case 1: FN<0, 0, 1>(__VA_ARGS__); break;           \
      case 2: FN<0, 1, 0>(__VA_ARGS__); break;   case 3: FN<0, 1, 1>(__VA_ARGS__); break;           \
      case 4: FN<0, 2, 0>(__VA_ARGS__); break;   case 5: FN<0, 2, 1>(__VA_ARGS__); break;           \
      case 6: FN<1, 0, 0>(__VA_ARGS__); break;   case 7: FN<1, 0, 1>(__VA_ARGS__); break;           \
      case 8: FN<1, 1, 0>(__VA_ARGS__); break;   case 9: FN<1, 1, 1>(__VA_ARGS__); break;           \
      case 10: FN<1, 2, 0>(__VA_ARGS__); break;  case 11: FN<1, 2, 1>(__VA_ARGS__); break;          \
      case 12: FN<2, 0, 0>(__VA_ARGS__); break;  case 13: FN<2, 0, 1>(__VA_ARGS__); break;          \
      case 14: FN<2, 1, 0>(__VA_ARGS__); break;  case 15: FN<2, 1, 1>(__VA_ARGS__); break;          \
      case 16: FN<2, 2, 0>(__VA_ARGS__); break;  case 17: FN<2, 2, 1>(__VA_ARGS__); break;          \
      case 18: FN<3, 0, 0>(__VA_ARGS__); break;  case 19: FN<3, 0, 1>(__VA_ARGS__); break;          \
      case 20: FN<3, 1, 0>(__VA_ARGS__); break;  case 21: FN<3, 1, 1>(__VA_ARGS__); break;          \
      case 22: FN<3, 2, 0>(__VA_ARGS__); break;  default: FN<3, 2, 1>(__VA_ARGS__); break;          \
    }                                                                                               \
  } while (0)

// Pair order of a CTA that owns a whole 512-token segment (tq = 4, joff = 0): stream tile j ascending, q tiles
// t = j .. 3.  The issuer warp runs this schedule fully unrolled (no bookkeeping between two MMA groups).
__host__ __device__ constexpr int full_j(int np) { return np < 4 ? 0 : np < 7 ? 1 : np < 9 ? 2 : 3; }
__host__ __device__ constexpr int full_t(int np) { return np < 4 ? np : np < 7 ? np - 3 : np < 9 ? np - 5 : 3; }
constexpr int FULL_PAIRS = 10;

// compact (run-time indexed) issue routines of the whole-segment schedule: the descriptors still come from the
// constant bank, but through indexed loads, so the issuer's code is one small loop that stays in the instruction cache
__device__ __forceinline__ void issue_scores_rt(const Descs& D, int t, int st, int b) {
  constexpr uint32_t ID = make_idesc_bf16(128, 128, false, false);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ss(128 * b, D.qK[t][kk], D.kK[st][kk], ID, kk > 0);
}
__device__ __forceinline__ void issue_pv_rt(const Descs& D, int t, int st, int b, uint32_t acc) {
  constexpr uint32_t ID = make_idesc_bf16(128, 64, false, true);
#pragma unroll
  for (int kk = 0; kk < 8; ++kk) umma_ts(256 + 64 * t, 128 * b + 8 * kk, D.vMN[st][kk], ID, (acc || kk > 0) ? 1 : 0);
}
constexpr unsigned long long FULL_T_TAB = 0x3323213210ULL, FULL_J_TAB = 0x3221110000ULL;   // nibble np

__global__ void __launch_bounds__(THREADS, 1)
fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
           const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmS,
           const __grid_constant__ Descs D, const Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_MISC);
  uint64_t* q_full = bars;         // q tiles 1.. , normaliser slab, fp32 vectors
  uint64_t* q0_full = bars + 13;   // q tile 0 alone: the first scores do not wait for the rest
  uint64_t* full = bars + 1;       // [3]
  uint64_t* empty = bars + 4;      // [3]
  uint64_t* s_ready = bars + 7;    // [2]
  uint64_t* p_ready = bars + 9;    // [2] count NCW: one arrival per compute warp
  uint64_t* acc_done = bars + 11;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* vec_s = reinterpret_cast<float*>(smem + OFF_MISC + 128);    // [128] S0 | N1 of the boundary state
  float* red = vec_s + 128;                                          // [4 tiles][4][128] per-column-block partial sums
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) FM_TRS(0, 0);
  const int bh = blockIdx.y;
  const int c0 = blockIdx.x * (QT * 128);
  const int seg = c0 / P.L;
  const int st0 = seg * P.L;                                          // first stream token (start of the segment)
  const int joff = (c0 - st0) / 128;                                  // stream tiles before the CTA's first q tile
  const int tq = min(QT, (P.N - c0 + 127) / 128);                     // live q tiles
  const int nstream = joff + tq;                                      // stream tiles [st0, c0 + 128 tq)
  const bool has_norm = P.state != nullptr && (seg > 0 || P.edge_state);
  const int kvh = kv_index(bh, P.H, P.Hkv);
  const int rec_idx = kvh * P.nseg + seg;
  const bool whole = tq == QT && joff == 0;
  // pairs (j, t), j ascending, t = max(0, j - joff) .. tq-1
  auto tmin = [&](int j) { return max(0, j - joff); };
  int npairs = 0;
  for (int j = 0; j < nstream; ++j) npairs += tq - tmin(j);

  // stream tiles: K rows in token order; V rows in the order of the P operand's k index (see the compute warps):
  // k rows [32 cb, 32 cb + 16) <- tokens [16 cb, +16), k rows [32 cb + 16, +16) <- tokens [64 + 16 cb, +16)
  auto load_stream = [&](int j) {
    const int s = j % NST;
    mbar_arrive_expect_tx(&full[s], STAGE);
    uint8_t* dst = smem + OFF_RING + s * STAGE;
    tma_load_3d(dst, &tmK, &full[s], 0, st0 + j * 128, kvh);
#pragma unroll
    for (int m = 0; m < 8; ++m)
      tma_load_3d(dst + 16384 + m * 2048, &tmV, &full[s], 0, st0 + j * 128 + 16 * (m >> 1) + 64 * (m & 1), kvh);
  };
  if (warp == NCW) {
    // ---------------------------------------------------------------- TMA producer, first part
    // (its own barriers are set up here so the first loads leave before the CTA-wide set-up barrier)
    if (lane == 0) {
      mbar_init(q_full, 1), mbar_init(q0_full, 1);
      for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
      fence_barrier_init();
    }
    __syncwarp();
    // the producer only ARRIVES on the set-up barrier: its ~40 bulk-tensor instructions run under the others' set-up
    asm volatile("bar.arrive 2, %0;" ::"r"(THREADS) : "memory");
    if (lane == 0) {
      mbar_arrive_expect_tx(q0_full, 16384);
      tma_load_3d(smem + OFF_Q, &tmQ, q0_full, 0, c0, bh);
      mbar_arrive_expect_tx(q_full, (tq - 1) * 16384 + (has_norm ? 8192 + 512 : 0));
      load_stream(0);
      for (int t = 1; t < tq; ++t) tma_load_3d(smem + OFF_Q + t * 16384, &tmQ, q_full, 0, c0 + t * 128, bh);
      if (has_norm) {
        tma_load_3d(smem + OFF_NORM, &tmS, q_full, 0, 65 * 64, rec_idx);
        // S0 | N1 of the boundary state (fp32), used by the epilogue only
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(vec_s)), "l"(P.state + (size_t)rec_idx * REC_BYTES + VEC_OFF), "r"(512), "r"(smem_u32(q_full))
                     : "memory");
      }
      for (int j = 1; j < min(nstream, NST); ++j) load_stream(j);
    }
  } else if (warp == NCW + 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(acc_done, 1);
    for (int s = 0; s < 2; ++s) mbar_init(&s_ready[s], 1), mbar_init(&p_ready[s], NCW);
    fence_barrier_init();
  }
  tc_fence_before();
  if (warp != NCW) named_barrier_sync(2, THREADS);
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  assert_bases(tmem, smem);
  if (tid == 0) FM_TRS(0, 1);

  if (warp == NCW) {
    // ---------------------------------------------------------------- TMA producer: stream tiles that reuse a stage
    if (lane == 0) {
      for (int j = NST; j < nstream; ++j) {
        mbar_wait(&empty[j % NST], ((j / NST) - 1) & 1, 100 + j);
        load_stream(j);
      }
    }
  } else if (warp == NCW + 1) {
    // ---------------------------------------------------------------- MMA issuer
    if (lane == 0) {
      mbar_wait(q0_full, 0, 299);
      if (!whole) mbar_wait(q_full, 0, 300);
    }
    __syncwarp();
    if (lane == 0) FM_TRS(0, 2);
    if (whole) {
      // scores two pairs ahead of the polynomial: S(0), S(1) up front, then P.V(np) + S(np + 2) per step
      if (lane == 0) mbar_wait(&full[0], 0, 400);
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
        issue_scores<0, 0, 0>(D);
        umma_commit(&s_ready[0]);
      }
      __syncwarp();
      if (lane == 0) mbar_wait(q_full, 0, 300);
      __syncwarp();
      if (elect_one()) {
        issue_scores<1, 0, 1>(D);
        umma_commit(&s_ready[1]);
      }
      __syncwarp();
#pragma unroll 1
      for (int np = 0; np < FULL_PAIRS; ++np) {
        // P.V of pair np as soon as its P is stored, then the scores of pair np + 2 into the score buffer that P.V
        // has just consumed (the tensor pipe executes in issue order)
        const int j = (int)(FULL_J_TAB >> (4 * np)) & 15, t = (int)(FULL_T_TAB >> (4 * np)) & 15, b = np & 1;
        const bool more = np + 2 < FULL_PAIRS;
        const int j2 = (int)(FULL_J_TAB >> (4 * np + 8)) & 15, t2 = (int)(FULL_T_TAB >> (4 * np + 8)) & 15;
        mbar_wait_spin(&p_ready[b], (np >> 1) & 1);                                   // (every lane polls)
        if (more && t2 == j2) mbar_wait_spin(&full[j2 % NST], (j2 / NST) & 1);        // first use of stream tile j2
        __syncwarp();
        tc_fence_after();
        if (lane == 0) FM_TRS(8, np);
        if (elect_one()) {
          issue_pv_rt(D, t, j % NST, b, j > 0 ? 1u : 0u);
          if (t == QT - 1) umma_commit(&empty[j % NST]);
          if (more) {
            issue_scores_rt(D, t2, j2 % NST, b);
            umma_commit(&s_ready[b]);
          }
          FM_TRS(12, np);
        }
        __syncwarp();
      }
    } else {
      // ragged tail / CTA inside a longer segment: same pairs, scores one pair ahead
      int js = 0, ts = tmin(0), ns = 0;        // next score group to issue
      int jp = 0, tp = tmin(0);                // next P.V group to issue
      auto issue_s = [&]() {
        const int st = js % NST, b = ns & 1;
        if (ts == tmin(js)) {
          if (lane == 0) mbar_wait(&full[st], (js / NST) & 1, 400 + js);
          __syncwarp();
        }
        tc_fence_after();
        if (elect_one()) {
          FM_SEG_DISPATCH(issue_scores, ts, st, b, D);
          umma_commit(&s_ready[b]);
        }
        __syncwarp();
        ++ns;
        if (++ts >= tq) ++js, ts = tmin(js);
      };
      if (npairs > 0) issue_s();
      for (int np = 0; np < npairs; ++np) {
        if (ns < npairs) issue_s();
        const int st = jp % NST, b = np & 1;
        if (lane == 0) mbar_wait(&p_ready[b], (np >> 1) & 1, 500 + np);
        __syncwarp();
        tc_fence_after();
        if (elect_one()) {
          FM_SEG_DISPATCH(issue_pv, tp, st, b, D, jp > 0 ? 1u : 0u);
          if (tp == tq - 1) umma_commit(&empty[st]);
        }
        __syncwarp();
        if (++tp >= tq) ++jp, tp = tmin(jp);
      }
    }
    // moment part of the normaliser: T_t = Q_t . N2 into the (drained) score columns [64 t, 64 t + 64)
    tc_fence_after();
    if (elect_one()) {
      if (has_norm) {
        constexpr uint32_t IDK = make_idesc_bf16(128, 64, false, false);
        for (int t = 0; t < tq; ++t) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss(64 * t, D.qK[t][kk], D.normK[kk], IDK, kk > 0);
        }
      }
      umma_commit(acc_done);
    }
    __syncwarp();
  } else if (warp < NCW) {
    // ---------------------------------------------------------------- compute warps: row x 32 score columns
    // Warp (q4, cb) owns lanes [32 q4, +32) and the score columns [16 cb, +16) and [64 + 16 cb, +16).  Its bf16 P
    // values go to the 32-bit columns [16 cb, +16) of the same buffer -- inside the columns it has just read, so no
    // warp waits for another -- which makes the k index of P:  k in [32 cb, +16) = tokens [16 cb, +16),
    // k in [32 cb + 16, +16) = tokens [64 + 16 cb, +16); the producer lays the V rows out in that order.
    const int q4 = warp & 3, cb = warp >> 2;
    const int r = 32 * q4 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q4) << 16);
    const float a = P.a, c2 = 0.5f * a * a;
    float gs0 = 0.f, gs1 = 0.f, gs2 = 0.f, gs3 = 0.f;
    int j = 0, t = tmin(0);
    for (int np = 0; np < npairs; ++np) {
      const int b = np & 1;
      if (tid == 0) FM_TRS(1, np);
      mbar_wait_spin(&s_ready[b], (np >> 1) & 1);      // every lane polls: fastest wake-up (measured)
      tc_fence_after();
      if (tid == 0) FM_TRS(2, np);
      FM_TRW(0, warp, np);
      const bool diag = j - joff == t;
      uint32_t v0[16], v1[16], pk[16];
      tmem_ld16(lane_addr + 128 * b + 16 * cb, v0);
      tmem_ld16(lane_addr + 128 * b + 64 + 16 * cb, v1);
      tmem_wait_ld();
      if (tid == 0) FM_TRS(3, np);
      float ps = 0.f;
      if (diag) {
        const f32x2 C2 = pack2(c2, c2), A2 = pack2(a, a), ONE2 = pack2(1.f, 1.f);
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          const f32x2 s01 = pack2u(v0[c], v0[c + 1]), s23 = pack2u(v1[c], v1[c + 1]);
          float x0, x1, x2, x3;
          unpack2(fma2(fma2(C2, s01, A2), s01, ONE2), x0, x1);
          unpack2(fma2(fma2(C2, s23, A2), s23, ONE2), x2, x3);
          x0 = 16 * cb + c <= r ? x0 : 0.f, x1 = 16 * cb + c + 1 <= r ? x1 : 0.f;
          x2 = 64 + 16 * cb + c <= r ? x2 : 0.f, x3 = 64 + 16 * cb + c + 1 <= r ? x3 : 0.f;
          ps += (x0 + x1) + (x2 + x3);
          pk[c >> 1] = pack_bf16x2(x0, x1);
          pk[8 + (c >> 1)] = pack_bf16x2(x2, x3);
        }
      } else {
        // interior tile: packed fp32 arithmetic (fm_sm100.cuh), half the instructions of the scalar loop
        const f32x2 C2 = pack2(c2, c2), A2 = pack2(a, a), ONE2 = pack2(1.f, 1.f);
        f32x2 ps2 = pack2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          const f32x2 s01 = pack2u(v0[c], v0[c + 1]), s23 = pack2u(v1[c], v1[c + 1]);
          const f32x2 x01 = fma2(fma2(C2, s01, A2), s01, ONE2), x23 = fma2(fma2(C2, s23, A2), s23, ONE2);
          ps2 = add2(ps2, add2(x01, x23));
          pk[c >> 1] = pack_bf16x2_of(x01);
          pk[8 + (c >> 1)] = pack_bf16x2_of(x23);
        }
        float pa, pb;
        unpack2(ps2, pa, pb);
        ps = pa + pb;
      }
      gs0 += t == 0 ? ps : 0.f, gs1 += t == 1 ? ps : 0.f, gs2 += t == 2 ? ps : 0.f, gs3 += t == 3 ? ps : 0.f;
      if (tid == 0) FM_TRS(4, np);
      tmem_st16(lane_addr + 128 * b + 16 * cb, pk);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[b]);
      FM_TRW(1, warp, np);
      if (tid == 0) FM_TRS(5, np);
      if (++t >= tq) ++j, t = tmin(j);
    }
    // ---------------------------------------------------------------- epilogue: fp32 partials + g
    // The blocked partial layout ([tile][c4][row] float4, fm_moments.cuh) makes a warp's 16-byte stores 512
    // contiguous bytes, so the accumulators go from registers straight to global memory: no staging, no bulk
    // stores to wait for.  Only the row sums of g cross warps (four column blocks per row), through `red`.
    if (tid == 0) FM_TRS(0, 3);
    mbar_wait_warp(acc_done, 0);
    mbar_wait_warp(q_full, 0);        // (long complete: makes the bulk-copied vec_s visible to this warp)
    tc_fence_after();
    if (tid == 0) FM_TRS(0, 4);
#pragma unroll
    for (int tt = 0; tt < QT; ++tt) {
      float gpart = tt == 0 ? gs0 : tt == 1 ? gs1 : tt == 2 ? gs2 : gs3;
      if (has_norm && tt < tq) {
        // g += a q.N1 + (a^2/2) q.T over this thread's 16 features, q row read back from the resident tile
        uint32_t tn[16];
        tmem_ld16(lane_addr + 64 * tt + 16 * cb, tn);
        tmem_wait_ld();
        const uint8_t* qrow = smem + OFF_Q + tt * 16384 + r * 128;
#pragma unroll
        for (int c8 = 0; c8 < 2; ++c8) {
          const uint4 qv = *reinterpret_cast<const uint4*>(qrow + (((2 * cb + c8) ^ (r & 7)) << 4));
          const uint32_t qw[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = 8 * c8 + 2 * e;
            const float q0 = __uint_as_float(qw[e] << 16), q1 = __uint_as_float(qw[e] & 0xffff0000u);
            gpart += q0 * fmaf(c2, __uint_as_float(tn[c]), a * vec_s[64 + 16 * cb + c]);
            gpart += q1 * fmaf(c2, __uint_as_float(tn[c + 1]), a * vec_s[64 + 16 * cb + c + 1]);
          }
        }
      }
      red[(4 * tt + cb) * 128 + r] = gpart;
    }
    if (tid == 0) FM_TRS(6, 0);
    named_barrier_sync(1, 32 * NCW);
    if (tid == 0) FM_TRS(6, 1);
#pragma unroll
    for (int tt = 0; tt < QT; ++tt) {
      if (tt >= tq) break;
      uint32_t o[16];
      tmem_ld16(lane_addr + 256 + 64 * tt + 16 * cb, o);
      tmem_wait_ld();
      const size_t tile = (size_t)bh * ((P.N + 127) / 128) + (size_t)(c0 / 128 + tt);
      uint4* dst = reinterpret_cast<uint4*>(P.out + tile * PART_TILE_FLOATS) + (4 * cb) * 128 + r;
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i * 128] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      const int r_abs = c0 + tt * 128 + r;
      if (cb == tt && r_abs < P.N) {
        const float count = (float)(seg * P.L) + P.count_offset;
        const float* rd = red + 512 * tt + r;
        P.g[(size_t)bh * P.N + r_abs] = rd[0] + rd[128] + rd[256] + rd[384] + (has_norm ? count : 0.f);
      }
    }
    if (tid == 0) FM_TRS(0, 5);
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0) FM_TRS(0, 6);
  if (warp == NCW + 1) tmem_dealloc(tmem, 512);
}

}  // namespace seg

using namespace seg;

bool seg_fwd_usable(const Problem& pb, int L) {
  return pb.p == 2 && pb.causal && pb.D == 64 && pb.dtype == FASTMAX_DTYPE_BF16 && L % 512 == 0;
}

// drop-in for intra_fwd (p = 2, causal): o_part fp32 blocked partial, g, boundary states (nullable)
int seg_fwd(const Problem& pb, int L, const void* q, const void* k, const void* v, float* o_part, float* g,
            const uint8_t* state) {
  const int BH = pb.B * pb.H, BHkv = pb.B * pb.kv_heads();
  CUtensorMap tQ, tK, tV, tS;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tQ, q, 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tK, k, 64, pb.N, BHkv, 128, (uint64_t)pb.N * 128, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tV, v, 64, pb.N, BHkv, 128, (uint64_t)pb.N * 128, 64, 16))) return rc;
  Params P{};
  P.N = pb.N, P.L = L, P.nseg = (pb.N + L - 1) / L, P.edge_state = pb.sp_init_fwd != nullptr;
  P.a = pb.inv_c, P.count_offset = (float)pb.sp_offset, P.out = o_part, P.g = g;
  P.H = pb.H, P.Hkv = pb.kv_heads();
  P.state = (P.nseg > 1 || P.edge_state) ? state : nullptr;
  tS = tQ;
  if (P.state && (rc = make_tmap_bf16_3d(&tS, state, 64, NSLAB * 64, (uint64_t)BHkv * P.nseg, 128, REC_BYTES, 64, 64))) return rc;
  Descs D{};
  const uint32_t base = sm100::FM_SMEM_BASE;
  for (int kk = 0; kk < 4; ++kk) {
    for (int t = 0; t < QT; ++t) D.qK[t][kk] = sm100::desc_kmajor_sw128_const(base + OFF_Q + t * 16384) + kk * 2;
    for (int s = 0; s < NST; ++s) D.kK[s][kk] = sm100::desc_kmajor_sw128_const(base + OFF_RING + s * STAGE) + kk * 2;
    D.normK[kk] = sm100::desc_kmajor_sw128_const(base + OFF_NORM) + kk * 2;
  }
  for (int kk = 0; kk < 8; ++kk)
    for (int s = 0; s < NST; ++s)
      D.vMN[s][kk] = sm100::desc_mnmajor_sw128_const(base + OFF_RING + s * STAGE + 16384, 8192) + kk * 128;
  FM_CUDA(cudaFuncSetAttribute(fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  dim3 grid((pb.N + QT * 128 - 1) / (QT * 128), BH);
  FM_LAUNCH("mom_seg_fwd", pb.stream, fwd_kernel<<<grid, THREADS, SMEM, pb.stream>>>(tQ, tK, tV, tS, D, P));
  return FASTMAX_OK;
}

}  // namespace fm
