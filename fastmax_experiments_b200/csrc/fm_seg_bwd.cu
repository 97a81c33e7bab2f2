// fm_seg_bwd.cu -- intra-segment dk / dv part of the p = 2 chunked scan, second generation (bf16, D = 64).
//
// Same contract as intra_kernel<INTRA_DKDV> of fm_intra.cu: fp32 partials of dv and dk in the blocked layout of
// fm_moments.cuh, dk including the normaliser column of the segment's reverse boundary state (RN1 + a RN2 k):
//
//   S' = K_j Q_i^T ,  dP' = V_j U_i^T          (rows = keys j of the CTA's tile, columns = queries i of the stream tile)
//   P' = f(S') o mask ,  E' = f'(S')/a o (dP' - dlt_i) o mask ,  mask: query >= key
//   dV_j += P' U_i ,  dK_j += E' Q_i           (dK unscaled by 1/c: the apply kernel scales)
//
// organised like fm_seg.cu (what its event trace taught): all 16 compute warps work on ONE 128 x 64 score pair at a time
// (thread = key row x 16 query columns of S' and of dP'), the scores of pair np + 2 are issued behind the accumulating
// MMAs of pair np by a dedicated issuer warp whose code is one compact loop with run-time indexed constant-bank
// descriptors, and a warp writes its bf16 P' | E' values into the score columns it has just read, so no warp waits
// for another (P' -> columns [16 cb, +8), E' -> [16 cb + 8, +8): k-step cb of each A operand starts at column 16 cb
// resp. 16 cb + 8, in token order).
//
// A CTA owns two 128-row key tiles (accumulators dV | dK: 2 x 128 columns) and streams 64-token q / u tiles (+ their
// dlt) from its first key to the end of the segment; two score buffers (S' | dP': 2 x 128 columns).
#include "fm_common.cuh"
#include "fm_sm100.cuh"
#include "fm_moments.cuh"

#ifdef FM_TRACE
static __device__ long long fm_trace_segb[16][16];
extern "C" int fastmax_debug_trace_fetch_segb(long long* out) {
  return cudaMemcpyFromSymbol(out, fm_trace_segb, sizeof(fm_trace_segb)) == cudaSuccess ? 0 : -4;
}
#define FM_TRB(ev, j) do { if (blockIdx.x == 4 && blockIdx.y == 100 && (j) < 16) fm_trace_segb[ev][j] = clock64(); } while (0)
#else
#define FM_TRB(ev, j) ((void)0)
#endif

namespace fm {
namespace segb {
using namespace sm100;
using namespace mom;
typedef __nv_bfloat16 bf16;

constexpr int NCW = 16;
constexpr int THREADS = (NCW + 2) * 32;
constexpr int KT = 2;                              // key tiles per CTA
constexpr int NST = 4;                             // ring stages
constexpr int STAGE = 2 * 8192 + 1024;             // Q tile | U tile (64 tokens each) | dlt (256 B)
constexpr int OFF_K = 0;
constexpr int OFF_V = KT * 16384;
constexpr int OFF_RING = 2 * KT * 16384;
constexpr int OFF_NORM = OFF_RING + NST * STAGE;   // slab 65 of the reverse boundary state
constexpr int OFF_MISC = OFF_NORM + 8192;
constexpr int SMEM = OFF_MISC + 2048 + 1024;
static_assert(OFF_NORM % 1024 == 0, "tile alignment");
static_assert((NST & (NST - 1)) == 0, "ring stages: power of two");

struct Descs {
  unsigned long long kK[KT][4], vK[KT][4];        // resident key / value tiles, K-major (A of the score MMAs)
  unsigned long long qK[NST][4], uK[NST][4];      // stream tiles, K-major (B of the score MMAs, 64 rows)
  unsigned long long qMN[NST][4], uMN[NST][4];    // stream tiles read MN-major (B of the accumulating MMAs)
  unsigned long long normK[4];                    // normaliser slab, K-major
};

struct Params {
  int N, Npad, L, nseg, edge_state;
  int H, Hkv;                // grouped-query attention: k, v per (b, kv head); q, u, dlt, states, outputs per query head
  float a;
  const float* dlt;          // (BH, Npad) fp32, zero padded
  float* out_dv;             // blocked fp32 partials
  float* out_dk;
  const uint8_t* state;      // reverse boundary states, nullable
};

// pair np -> (stream tile m, key tile t): m ascending, t = 0 .. min(tk - 1, m / 2)
struct PairIt {
  int m = 0, t = 0;
  __device__ __forceinline__ void next(int tk) {
    if (t + 1 < tk && m >= 2 * (t + 1)) ++t;
    else ++m, t = 0;
  }
};

__device__ __forceinline__ void issue_scores(const Descs& D, int t, int st, int b) {
  constexpr uint32_t ID = make_idesc_bf16(128, 64, false, false);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ss(256 + 128 * b, D.kK[t][kk], D.qK[st][kk], ID, kk > 0);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ss(256 + 128 * b + 64, D.vK[t][kk], D.uK[st][kk], ID, kk > 0);
}
__device__ __forceinline__ void issue_acc(const Descs& D, int t, int st, int b, uint32_t acc) {
  constexpr uint32_t ID = make_idesc_bf16(128, 64, false, true);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ts(128 * t, 256 + 128 * b + 16 * kk, D.uMN[st][kk], ID, (acc || kk > 0) ? 1 : 0);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ts(128 * t + 64, 256 + 128 * b + 16 * kk + 8, D.qMN[st][kk], ID, (acc || kk > 0) ? 1 : 0);
}

__global__ void __launch_bounds__(THREADS, 1)
dkdv_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
            const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmU,
            const __grid_constant__ CUtensorMap tmS, const __grid_constant__ Descs D, const Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_MISC);
  uint64_t* res_full = bars;
  uint64_t* full = bars + 1;        // [4]
  uint64_t* empty = bars + 5;       // [4]
  uint64_t* s_ready = bars + 9;     // [2]
  uint64_t* p_ready = bars + 11;    // [2] count NCW
  uint64_t* acc_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* vec_s = reinterpret_cast<float*>(smem + OFF_MISC + 256);    // [128] R0 | RN1 of the boundary state
  // pair schedule, one word per pair: m | t << 8 | (first pair of key tile t) << 16 | (last pair on stream tile m) << 17
  uint32_t* sched = reinterpret_cast<uint32_t*>(smem + OFF_MISC + 1024);   // [256]: 14 pairs at L = 512, 254 at L = 8192
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) FM_TRB(0, 0);
  const int bh = blockIdx.y;
  const int kvh = kv_index(bh, P.H, P.Hkv);
  const int c0 = blockIdx.x * (KT * 128);                  // first key of the CTA
  const int seg = c0 / P.L;
  const int seg_end = min((seg + 1) * P.L, P.N);
  const int tk = min(KT, (P.N - c0 + 127) / 128);          // live key tiles
  const int M = (seg_end - c0 + 63) / 64;                  // stream tiles [c0, seg_end)
  const bool has_norm = P.state != nullptr && (seg < P.nseg - 1 || P.edge_state);
  const int rec_idx = bh * P.nseg + seg;
  int npairs = 0;
  for (int m = 0; m < M; ++m) npairs += min(tk, m / 2 + 1);

  auto load_stream = [&](int m) {
    const int s = m % NST;
    uint8_t* dst = smem + OFF_RING + s * STAGE;
    mbar_arrive_expect_tx(&full[s], 2 * 8192 + 256);
    tma_load_3d(dst, &tmQ, &full[s], 0, c0 + m * 64, bh);
    tma_load_3d(dst + 8192, &tmU, &full[s], 0, c0 + m * 64, bh);
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst + 16384)), "l"(P.dlt + (size_t)bh * P.Npad + c0 + m * 64), "r"(256), "r"(smem_u32(&full[s]))
                 : "memory");
  };
  if (warp == NCW) {
    // ---------------------------------------------------------------- TMA producer, first part (before the set-up barrier)
    if (lane == 0) {
      mbar_init(res_full, 1);
      for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("bar.arrive 2, %0;" ::"r"(THREADS) : "memory");
    if (lane == 0) {
      mbar_arrive_expect_tx(res_full, 2 * tk * 16384 + (has_norm ? 8192 + 512 : 0));
      tma_load_3d(smem + OFF_K, &tmK, res_full, 0, c0, kvh);
      tma_load_3d(smem + OFF_V, &tmV, res_full, 0, c0, kvh);
      if (M > 0) load_stream(0);
      if (tk > 1) {
        tma_load_3d(smem + OFF_K + 16384, &tmK, res_full, 0, c0 + 128, kvh);
        tma_load_3d(smem + OFF_V + 16384, &tmV, res_full, 0, c0 + 128, kvh);
      }
      if (has_norm) {
        tma_load_3d(smem + OFF_NORM, &tmS, res_full, 0, 65 * 64, rec_idx);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(vec_s)), "l"(P.state + (size_t)rec_idx * REC_BYTES + VEC_OFF), "r"(512), "r"(smem_u32(res_full))
                     : "memory");
      }
      for (int m = 1; m < min(M, NST); ++m) load_stream(m);
    }
  } else if (warp == NCW + 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 32) {       // (a compute warp's lane: the issuer warp's own scalar work is what bounds this kernel)
    PairIt it;
    for (int np = 0; np < npairs; ++np) {
      sched[np] = (uint32_t)it.m | ((uint32_t)it.t << 8) | ((it.m == 2 * it.t ? 1u : 0u) << 16) |
                  ((it.t == min(tk - 1, it.m / 2) ? 1u : 0u) << 17);
      it.next(tk);
    }
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

  if (warp == NCW) {
    // ---------------------------------------------------------------- TMA producer: stream tiles that reuse a stage
    if (lane == 0) {
      for (int m = NST; m < M; ++m) {
        mbar_wait(&empty[m % NST], ((m / NST) - 1) & 1, 100 + m);
        load_stream(m);
      }
    }
  } else if (warp == NCW + 1) {
    // ---------------------------------------------------------------- MMA issuer
    // One warp, one elected region per pair: the accumulating MMAs of pair np, then the scores of pair np + 2 into the
    // buffer they have just consumed (the pipe executes in issue order).  Measured alternative: separate warps for
    // the scores and the accumulating MMAs (hand-over through an "issued" barrier): 1.95 -> 2.26 ms, each warp's
    // issue blocks behind the other's queued MMAs.
    if (lane == 0) {
      mbar_wait(res_full, 0, 300);
      if (npairs > 0) mbar_wait(&full[0], 0, 400);
      if (npairs > 1) mbar_wait(&full[1 % NST], 0, 401);
    }
    __syncwarp();
    tc_fence_after();
    if (elect_one()) {
      if (npairs > 0) {
        issue_scores(D, 0, 0, 0);                 // pair 0 = (m 0, t 0)
        umma_commit(&s_ready[0]);
      }
      if (npairs > 1) {
        issue_scores(D, 0, 1 % NST, 1);           // pair 1 = (m 1, t 0)
        umma_commit(&s_ready[1]);
      }
    }
    __syncwarp();
#pragma unroll 1
    for (int np = 0; np < npairs; ++np) {
      const int b = np & 1;
      const bool more = np + 2 < npairs;
      const uint32_t wa = sched[np], ws = sched[more ? np + 2 : np];
      const int ma = wa & 255, ta = (wa >> 8) & 255, ms = ws & 255, ts = (ws >> 8) & 255;
      mbar_wait_spin(&p_ready[b], (np >> 1) & 1);                                          // (every lane polls)
      if (more && ts == 0) mbar_wait_spin(&full[ms & (NST - 1)], (ms / NST) & 1);          // first use of stream tile ms
      __syncwarp();
      tc_fence_after();
      if (lane == 0) FM_TRB(5, np);
      if (elect_one()) {
        issue_acc(D, ta, ma & (NST - 1), b, (wa >> 16) & 1 ? 0u : 1u);
        if ((wa >> 17) & 1) umma_commit(&empty[ma & (NST - 1)]);      // last pair on this stream tile
        FM_TRB(6, np);
        if (more) {
          issue_scores(D, ts, ms & (NST - 1), b);
          umma_commit(&s_ready[b]);
        }
        FM_TRB(7, np);
      }
      __syncwarp();
    }
    // normaliser column of the boundary state: T_t = K_t . RN2 into the (drained) score columns [256 + 64 t, +64)
    tc_fence_after();
    if (elect_one()) {
      if (has_norm) {
        constexpr uint32_t IDK = make_idesc_bf16(128, 64, false, false);
        for (int t = 0; t < tk; ++t) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss(256 + 64 * t, D.kK[t][kk], D.normK[kk], IDK, kk > 0);
        }
      }
      umma_commit(acc_done);
    }
    __syncwarp();
  } else if (warp < NCW) {
    // ---------------------------------------------------------------- compute warps: key row x 16 query columns
    const int q4 = warp & 3, cb = warp >> 2;
    const int r = 32 * q4 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q4) << 16);
    const float a = P.a, c2 = 0.5f * a * a;
    PairIt pc;
    for (int np = 0; np < npairs; ++np) {
      const int b = np & 1, st = pc.m % NST;
      const uint32_t sbuf = lane_addr + 256 + 128 * b;
      if (tid == 0) FM_TRB(1, np);
      mbar_wait_spin(&s_ready[b], (np >> 1) & 1);      // every lane polls: fastest wake-up (measured)
      tc_fence_after();
      if (tid == 0) FM_TRB(2, np);
      uint32_t sv[16], dv[16], w[16];
      tmem_ld16(sbuf + 16 * cb, sv);
      tmem_ld16(sbuf + 64 + 16 * cb, dv);
      const float4* dsrc = reinterpret_cast<const float4*>(smem + OFF_RING + st * STAGE + 16384) + 4 * cb;
      float dl[16];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 t4 = dsrc[c4];
        dl[4 * c4] = t4.x, dl[4 * c4 + 1] = t4.y, dl[4 * c4 + 2] = t4.z, dl[4 * c4 + 3] = t4.w;
      }
      tmem_wait_ld();
      const int key = 128 * pc.t + r;                      // relative to c0
      const int col0 = 64 * pc.m + 16 * cb;                // query of this thread's first column, relative to c0
      const bool diag = pc.m < 2 * pc.t + 2;               // the tile reaches below the diagonal
      if (diag) {
        const f32x2 C2 = pack2(c2, c2), A2 = pack2(a, a), ONE2 = pack2(1.f, 1.f);
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          const f32x2 s01 = pack2u(sv[c], sv[c + 1]);
          float p0, p1, e0, e1;
          unpack2(fma2(s01, fma2(C2, s01, A2), ONE2), p0, p1);
          unpack2(mul2(fma2(A2, s01, ONE2), sub2(pack2u(dv[c], dv[c + 1]), pack2(dl[c], dl[c + 1]))), e0, e1);
          const bool v0 = col0 + c >= key, v1 = col0 + c + 1 >= key;      // query >= key
          p0 = v0 ? p0 : 0.f, p1 = v1 ? p1 : 0.f, e0 = v0 ? e0 : 0.f, e1 = v1 ? e1 : 0.f;
          w[c >> 1] = pack_bf16x2(p0, p1);
          w[8 + (c >> 1)] = pack_bf16x2(e0, e1);
        }
      } else {
        // interior tile: packed fp32 arithmetic (fm_sm100.cuh), half the instructions of the scalar loop
        const f32x2 C2 = pack2(c2, c2), A2 = pack2(a, a), ONE2 = pack2(1.f, 1.f);
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          const f32x2 s01 = pack2u(sv[c], sv[c + 1]);
          const f32x2 p01 = fma2(s01, fma2(C2, s01, A2), ONE2);
          const f32x2 e01 = mul2(fma2(A2, s01, ONE2), sub2(pack2u(dv[c], dv[c + 1]), pack2(dl[c], dl[c + 1])));
          w[c >> 1] = pack_bf16x2_of(p01);
          w[8 + (c >> 1)] = pack_bf16x2_of(e01);
        }
      }
      if (tid == 0) FM_TRB(3, np);
      tmem_st16(sbuf + 16 * cb, w);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[b]);
      if (tid == 0) FM_TRB(4, np);
      pc.next(tk);
    }
    // ---------------------------------------------------------------- epilogue: fp32 partials straight from registers
    if (tid == 0) FM_TRB(0, 1);
    mbar_wait_warp(acc_done, 0);
    mbar_wait_warp(res_full, 0);      // (long complete: makes the bulk-copied vec_s visible to this warp)
    tc_fence_after();
    if (tid == 0) FM_TRB(0, 2);
    for (int t = 0; t < tk; ++t) {
      const size_t tile = (size_t)bh * ((P.N + 127) / 128) + (size_t)(c0 / 128 + t);
#pragma unroll
      for (int a2 = 0; a2 < 2; ++a2) {
        uint32_t o[16];
        tmem_ld16(lane_addr + 128 * t + 64 * a2 + 16 * cb, o);
        if (a2 == 1 && has_norm) {      // dk_part[l] += RN1[l] + a (RN2 k)[l]
          uint32_t tn[16];
          tmem_ld16(lane_addr + 256 + 64 * t + 16 * cb, tn);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 16; ++c)
            o[c] = __float_as_uint(__uint_as_float(o[c]) + fmaf(a, __uint_as_float(tn[c]), vec_s[64 + 16 * cb + c]));
        } else {
          tmem_wait_ld();
        }
        uint4* dst = reinterpret_cast<uint4*>((a2 == 0 ? P.out_dv : P.out_dk) + tile * PART_TILE_FLOATS) + (4 * cb) * 128 + r;
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i * 128] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
      }
    }
  }
  if (tid == 0) FM_TRB(0, 3);
  tc_fence_before();
  __syncthreads();
  if (tid == 0) FM_TRB(0, 4);
  if (warp == NCW + 1) tmem_dealloc(tmem, 512);
}

}  // namespace segb


// =============================================================================================
// dq: same organisation with the roles swapped -- rows = queries of the CTA's four tiles (q, u resident), 64-token
// k / v tiles streamed from the start of the segment:
//   S = Q_i K_j^T ,  dP = U_i V_j^T ,  E = f'(S)/a o (dP - dlt_i) o mask (key <= query) ,  dQ_i += E K_j
// plus the normaliser column of the forward boundary state, dq_part -= dlt_i (N1 + a N2 q_i)   (intra_kernel<INTRA_DQ>)
// Four accumulators (4 x 64 columns) + two (S | dP) buffers (2 x 128 columns).  1.31 -> 1.21 ms at the headline shape
// (same box, same run); with 128-token stream tiles only ONE (S | dP) buffer fits next to the accumulators and the
// issuer and the compute warps strictly alternate: 1.59 ms.
// =============================================================================================
namespace segq {
using namespace sm100;
using namespace mom;

constexpr int NCW = 16;
constexpr int THREADS = (NCW + 2) * 32;
constexpr int QT = 4;                              // query tiles per CTA
constexpr int NST = 3;                             // (4 stages = 203 KiB of shared memory: same speed, but the kernel then
                                                   //  fails to launch under ncu metric replay, which needs some itself)
constexpr int STAGE = 2 * 8192;                    // K tile | V tile (64 tokens each)
constexpr int OFF_Q = 0;
constexpr int OFF_U = QT * 16384;
constexpr int OFF_RING = 2 * QT * 16384;
constexpr int OFF_NORM = OFF_RING + NST * STAGE;
constexpr int OFF_MISC = OFF_NORM + 8192;
constexpr int SMEM = OFF_MISC + 2048 + 1024;

struct Descs {
  unsigned long long qK[QT][4], uK[QT][4];        // resident query / u tiles, K-major (A of the score MMAs)
  unsigned long long kK[NST][4], vK[NST][4];      // stream tiles, K-major (B of the score MMAs, 64 rows)
  unsigned long long kMN[NST][4];                 // stream K tile read MN-major (B of the accumulating MMAs)
  unsigned long long normK[4];
};

struct Params {
  int N, Npad, L, nseg, edge_state;
  int H, Hkv;
  float a;
  const float* dlt;
  float* out;                // blocked fp32 partial of dq
  const uint8_t* state;      // forward boundary states, nullable
};

// pair np -> (stream tile m, query tile t): m ascending, t = max(0, m / 2 - J) .. tq - 1   (J: query tiles of the
// segment in front of the CTA)
struct PairIt {
  int m = 0, t = 0;
  __device__ __forceinline__ void next(int tq, int J) {
    if (t + 1 < tq) ++t;
    else ++m, t = max(0, (m >> 1) - J);
  }
};

__device__ __forceinline__ void issue_scores(const Descs& D, int t, int st, int b) {
  constexpr uint32_t ID = make_idesc_bf16(128, 64, false, false);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ss(256 + 128 * b, D.qK[t][kk], D.kK[st][kk], ID, kk > 0);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ss(256 + 128 * b + 64, D.uK[t][kk], D.vK[st][kk], ID, kk > 0);
}
// E (bf16) of warp cb sits in columns [16 cb, +8) of the score buffer: k-step cb starts there
__device__ __forceinline__ void issue_acc(const Descs& D, int t, int st, int b, uint32_t acc) {
  constexpr uint32_t ID = make_idesc_bf16(128, 64, false, true);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) umma_ts(64 * t, 256 + 128 * b + 16 * kk, D.kMN[st][kk], ID, (acc || kk > 0) ? 1 : 0);
}

__global__ void __launch_bounds__(THREADS, 1)
dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmU,
          const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
          const __grid_constant__ CUtensorMap tmS, const __grid_constant__ Descs D, const Params P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_MISC);
  uint64_t* res_full = bars;
  uint64_t* full = bars + 1;        // [NST]
  uint64_t* empty = bars + 5;       // [NST]
  uint64_t* s_ready = bars + 9;     // [2]
  uint64_t* p_ready = bars + 11;    // [2] count NCW
  uint64_t* acc_done = bars + 13;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  float* vec_s = reinterpret_cast<float*>(smem + OFF_MISC + 256);    // [128] S0 | N1 of the boundary state
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y;
  const int kvh = kv_index(bh, P.H, P.Hkv);
  const int c0 = blockIdx.x * (QT * 128);
  const int seg = c0 / P.L;
  const int st0 = seg * P.L;
  const int J = (c0 - st0) / 128;
  const int tq = min(QT, (P.N - c0 + 127) / 128);          // live query tiles
  const int M = 2 * (J + tq);                              // stream tiles [st0, c0 + 128 tq)
  const bool has_norm = P.state != nullptr && (seg > 0 || P.edge_state);
  const int rec_idx = kvh * P.nseg + seg;
  int npairs = 0;
  for (int m = 0; m < M; ++m) npairs += tq - max(0, (m >> 1) - J);

  auto load_stream = [&](int m) {
    const int s = m % NST;
    uint8_t* dst = smem + OFF_RING + s * STAGE;
    mbar_arrive_expect_tx(&full[s], STAGE);
    tma_load_3d(dst, &tmK, &full[s], 0, st0 + m * 64, kvh);
    tma_load_3d(dst + 8192, &tmV, &full[s], 0, st0 + m * 64, kvh);
  };
  if (warp == NCW) {
    if (lane == 0) {
      mbar_init(res_full, 1);
      for (int s = 0; s < NST; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], 1);
      fence_barrier_init();
    }
    __syncwarp();
    asm volatile("bar.arrive 2, %0;" ::"r"(THREADS) : "memory");
    if (lane == 0) {
      mbar_arrive_expect_tx(res_full, 2 * tq * 16384 + (has_norm ? 8192 + 512 : 0));
      tma_load_3d(smem + OFF_Q, &tmQ, res_full, 0, c0, bh);
      tma_load_3d(smem + OFF_U, &tmU, res_full, 0, c0, bh);
      load_stream(0);
      for (int t = 1; t < tq; ++t) {
        tma_load_3d(smem + OFF_Q + t * 16384, &tmQ, res_full, 0, c0 + t * 128, bh);
        tma_load_3d(smem + OFF_U + t * 16384, &tmU, res_full, 0, c0 + t * 128, bh);
      }
      if (has_norm) {
        tma_load_3d(smem + OFF_NORM, &tmS, res_full, 0, 65 * 64, rec_idx);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(vec_s)), "l"(P.state + (size_t)rec_idx * REC_BYTES + VEC_OFF), "r"(512), "r"(smem_u32(res_full))
                     : "memory");
      }
      for (int m = 1; m < min(M, NST); ++m) load_stream(m);
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

  if (warp == NCW) {
    if (lane == 0) {
      for (int m = NST; m < M; ++m) {
        mbar_wait(&empty[m % NST], ((m / NST) - 1) & 1, 100 + m);
        load_stream(m);
      }
    }
  } else if (warp == NCW + 1) {
    // ---------------------------------------------------------------- MMA issuer (see dkdv_kernel)
    PairIt ps, pa;                  // next pair to issue scores for / to accumulate
    mbar_wait_spin(res_full, 0);
    if (npairs > 0) mbar_wait_spin(&full[0], 0);
    __syncwarp();
    tc_fence_after();
    for (int ns = 0; ns < min(2, npairs); ++ns) {    // pairs 0 and 1: both on stream tile 0 unless the CTA has one query tile
      if (ps.m > 0) mbar_wait_spin(&full[ps.m % NST], 0);
      __syncwarp();
      if (elect_one()) {
        issue_scores(D, ps.t, ps.m % NST, ns);
        umma_commit(&s_ready[ns]);
      }
      __syncwarp();
      ps.next(tq, J);
    }
#pragma unroll 1
    for (int np = 0; np < npairs; ++np) {
      const int b = np & 1, st = pa.m % NST;
      const bool more = np + 2 < npairs;
      const bool fresh = more && ps.t == max(0, (ps.m >> 1) - J);      // first pair on stream tile ps.m
      mbar_wait_spin(&p_ready[b], (np >> 1) & 1);                       // (every lane polls)
      if (fresh) mbar_wait_spin(&full[ps.m % NST], (ps.m / NST) & 1);
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
        issue_acc(D, pa.t, st, b, pa.m > 0 ? 1u : 0u);
        if (pa.t == tq - 1) umma_commit(&empty[st]);                   // last pair on this stream tile
        if (more) {
          issue_scores(D, ps.t, ps.m % NST, b);
          umma_commit(&s_ready[b]);
        }
      }
      __syncwarp();
      pa.next(tq, J);
      ps.next(tq, J);
    }
    // normaliser column of the boundary state: T_t = Q_t . N2 into the (drained) score columns [256 + 64 t, +64)
    tc_fence_after();
    if (elect_one()) {
      if (has_norm) {
        constexpr uint32_t IDK = make_idesc_bf16(128, 64, false, false);
        for (int t = 0; t < tq; ++t) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss(256 + 64 * t, D.qK[t][kk], D.normK[kk], IDK, kk > 0);
        }
      }
      umma_commit(acc_done);
    }
    __syncwarp();
  } else if (warp < NCW) {
    // ---------------------------------------------------------------- compute warps: query row x 16 key columns
    const int q4 = warp & 3, cb = warp >> 2;
    const int r = 32 * q4 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q4) << 16);
    const float a = P.a;
    float dl0, dl1, dl2, dl3;       // dlt of this thread's row in each query tile
    {
      const float* dsrc = P.dlt + (size_t)bh * P.Npad + c0 + r;
      dl0 = c0 + r < P.N ? dsrc[0] : 0.f;
      dl1 = c0 + 128 + r < P.N ? dsrc[128] : 0.f;
      dl2 = c0 + 256 + r < P.N ? dsrc[256] : 0.f;
      dl3 = c0 + 384 + r < P.N ? dsrc[384] : 0.f;
    }
    PairIt pc;
    for (int np = 0; np < npairs; ++np) {
      const int b = np & 1;
      const uint32_t sbuf = lane_addr + 256 + 128 * b;
      mbar_wait_spin(&s_ready[b], (np >> 1) & 1);      // every lane polls: fastest wake-up (measured)
      tc_fence_after();
      uint32_t sv[16], dv[16], w[8];
      tmem_ld16(sbuf + 16 * cb, sv);
      tmem_ld16(sbuf + 64 + 16 * cb, dv);
      tmem_wait_ld();
      const float dl = pc.t == 0 ? dl0 : pc.t == 1 ? dl1 : pc.t == 2 ? dl2 : dl3;
      const int qrow = 128 * (pc.t + J) + r;               // relative to the start of the segment
      const int col0 = 64 * pc.m + 16 * cb;                // key of this thread's first column, same origin
      const bool diag = pc.m >= 2 * (pc.t + J);            // the tile reaches above the diagonal
      if (diag) {
        const f32x2 A2 = pack2(a, a), ONE2 = pack2(1.f, 1.f), DL2 = pack2(dl, dl);
#pragma unroll
        for (int c = 0; c < 16; c += 2) {
          float e0, e1;
          unpack2(mul2(fma2(A2, pack2u(sv[c], sv[c + 1]), ONE2), sub2(pack2u(dv[c], dv[c + 1]), DL2)), e0, e1);
          e0 = col0 + c <= qrow ? e0 : 0.f, e1 = col0 + c + 1 <= qrow ? e1 : 0.f;      // key <= query
          w[c >> 1] = pack_bf16x2(e0, e1);
        }
      } else {
        const f32x2 A2 = pack2(a, a), ONE2 = pack2(1.f, 1.f), DL2 = pack2(dl, dl);      // packed fp32 arithmetic
#pragma unroll
        for (int c = 0; c < 16; c += 2)
          w[c >> 1] = pack_bf16x2_of(mul2(fma2(A2, pack2u(sv[c], sv[c + 1]), ONE2), sub2(pack2u(dv[c], dv[c + 1]), DL2)));
      }
      tmem_st8(sbuf + 16 * cb, w);      // inside the S columns this warp has just read: no warp waits for another
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[b]);
      pc.next(tq, J);
    }
    // ---------------------------------------------------------------- epilogue: fp32 partial straight from registers
    mbar_wait_warp(acc_done, 0);
    mbar_wait_warp(res_full, 0);
    tc_fence_after();
#pragma unroll
    for (int t = 0; t < QT; ++t) {
      if (t >= tq) break;
      uint32_t o[16];
      tmem_ld16(lane_addr + 64 * t + 16 * cb, o);
      if (has_norm) {               // dq_part[l] -= dlt (N1[l] + a T[l])
        uint32_t tn[16];
        tmem_ld16(lane_addr + 256 + 64 * t + 16 * cb, tn);
        tmem_wait_ld();
        const float dl = t == 0 ? dl0 : t == 1 ? dl1 : t == 2 ? dl2 : dl3;
#pragma unroll
        for (int c = 0; c < 16; ++c)
          o[c] = __float_as_uint(__uint_as_float(o[c]) - dl * fmaf(a, __uint_as_float(tn[c]), vec_s[64 + 16 * cb + c]));
      } else {
        tmem_wait_ld();
      }
      const size_t tile = (size_t)bh * ((P.N + 127) / 128) + (size_t)(c0 / 128 + t);
      uint4* dst = reinterpret_cast<uint4*>(P.out + tile * PART_TILE_FLOATS) + (4 * cb) * 128 + r;
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i * 128] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == NCW + 1) tmem_dealloc(tmem, 512);
}

}  // namespace segq

bool seg_dq_usable(const Problem& pb, int L) {
  return pb.p == 2 && pb.causal && pb.D == 64 && pb.dtype == FASTMAX_DTYPE_BF16 && L % 512 == 0;
}

// drop-in for intra_dq (p = 2, causal): fp32 blocked partial of dq, forward boundary states (nullable)
int seg_dq(const Problem& pb, int L, int Npad, const void* q, const void* u, const void* k, const void* v,
           const float* dlt, float* dq_part, const uint8_t* state) {
  const int BH = pb.B * pb.H, BHkv = pb.B * pb.kv_heads();
  CUtensorMap tQ, tU, tK, tV, tS;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tQ, q, 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tU, u, 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tK, k, 64, pb.N, BHkv, 128, (uint64_t)pb.N * 128, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&tV, v, 64, pb.N, BHkv, 128, (uint64_t)pb.N * 128, 64, 64))) return rc;
  segq::Params P{};
  P.N = pb.N, P.Npad = Npad, P.L = L, P.nseg = (pb.N + L - 1) / L, P.edge_state = pb.sp_init_fwd != nullptr;
  P.H = pb.H, P.Hkv = pb.kv_heads();
  P.a = pb.inv_c, P.dlt = dlt, P.out = dq_part;
  P.state = (P.nseg > 1 || P.edge_state) ? state : nullptr;
  tS = tQ;
  if (P.state && (rc = make_tmap_bf16_3d(&tS, state, 64, mom::NSLAB * 64, (uint64_t)BHkv * P.nseg, 128, mom::REC_BYTES, 64, 64))) return rc;
  segq::Descs D{};
  const uint32_t base = sm100::FM_SMEM_BASE;
  for (int kk = 0; kk < 4; ++kk) {
    for (int t = 0; t < segq::QT; ++t) {
      D.qK[t][kk] = sm100::desc_kmajor_sw128_const(base + segq::OFF_Q + t * 16384) + kk * 2;
      D.uK[t][kk] = sm100::desc_kmajor_sw128_const(base + segq::OFF_U + t * 16384) + kk * 2;
    }
    for (int s = 0; s < segq::NST; ++s) {
      const uint32_t ka = base + segq::OFF_RING + s * segq::STAGE, va = ka + 8192;
      D.kK[s][kk] = sm100::desc_kmajor_sw128_const(ka) + kk * 2;
      D.vK[s][kk] = sm100::desc_kmajor_sw128_const(va) + kk * 2;
      D.kMN[s][kk] = sm100::desc_mnmajor_sw128_const(ka, 8192) + kk * 128;
    }
    D.normK[kk] = sm100::desc_kmajor_sw128_const(base + segq::OFF_NORM) + kk * 2;
  }
  FM_CUDA(cudaFuncSetAttribute(segq::dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, segq::SMEM));
  dim3 grid((pb.N + segq::QT * 128 - 1) / (segq::QT * 128), BH);
  FM_LAUNCH("mom_seg_dq", pb.stream,
            segq::dq_kernel<<<grid, segq::THREADS, segq::SMEM, pb.stream>>>(tQ, tU, tK, tV, tS, D, P));
  return FASTMAX_OK;
}

using namespace segb;

bool seg_dkdv_usable(const Problem& pb, int L) {
  // (L <= 8192: the pair schedule of a CTA -- at most 2 * L / 64 - 2 words -- has 256 words of shared memory)
  return pb.p == 2 && pb.causal && pb.D == 64 && pb.dtype == FASTMAX_DTYPE_BF16 && L % 512 == 0 && L <= 8192;
}

// drop-in for intra_dkdv (p = 2, causal): fp32 blocked partials of dk and dv, reverse boundary states (nullable)
int seg_dkdv(const Problem& pb, int L, int Npad, const void* k, const void* v, const void* q, const void* u,
             const float* dlt, float* dk_part, float* dv_part, const uint8_t* rstate) {
  const int BH = pb.B * pb.H, BHkv = pb.B * pb.kv_heads();
  CUtensorMap tK, tV, tQ, tU, tS;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tK, k, 64, pb.N, BHkv, 128, (uint64_t)pb.N * 128, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tV, v, 64, pb.N, BHkv, 128, (uint64_t)pb.N * 128, 64, 128))) return rc;
  if ((rc = make_tmap_bf16_3d(&tQ, q, 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&tU, u, 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 64))) return rc;
  Params P{};
  P.N = pb.N, P.Npad = Npad, P.L = L, P.nseg = (pb.N + L - 1) / L, P.edge_state = pb.sp_init_rev != nullptr;
  P.H = pb.H, P.Hkv = pb.kv_heads();
  P.a = pb.inv_c, P.dlt = dlt, P.out_dv = dv_part, P.out_dk = dk_part;
  P.state = (P.nseg > 1 || P.edge_state) ? rstate : nullptr;
  tS = tQ;
  if (P.state && (rc = make_tmap_bf16_3d(&tS, rstate, 64, NSLAB * 64, (uint64_t)BH * P.nseg, 128, REC_BYTES, 64, 64))) return rc;
  Descs D{};
  const uint32_t base = sm100::FM_SMEM_BASE;
  for (int kk = 0; kk < 4; ++kk) {
    for (int t = 0; t < KT; ++t) {
      D.kK[t][kk] = sm100::desc_kmajor_sw128_const(base + OFF_K + t * 16384) + kk * 2;
      D.vK[t][kk] = sm100::desc_kmajor_sw128_const(base + OFF_V + t * 16384) + kk * 2;
    }
    for (int s = 0; s < NST; ++s) {
      const uint32_t qa = base + OFF_RING + s * STAGE, ua = qa + 8192;
      D.qK[s][kk] = sm100::desc_kmajor_sw128_const(qa) + kk * 2;
      D.uK[s][kk] = sm100::desc_kmajor_sw128_const(ua) + kk * 2;
      D.qMN[s][kk] = sm100::desc_mnmajor_sw128_const(qa, 8192) + kk * 128;
      D.uMN[s][kk] = sm100::desc_mnmajor_sw128_const(ua, 8192) + kk * 128;
    }
    D.normK[kk] = sm100::desc_kmajor_sw128_const(base + OFF_NORM) + kk * 2;
  }
  FM_CUDA(cudaFuncSetAttribute(dkdv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  dim3 grid((pb.N + KT * 128 - 1) / (KT * 128), BH);
  FM_LAUNCH("mom_seg_dkdv", pb.stream, dkdv_kernel<<<grid, THREADS, SMEM, pb.stream>>>(tK, tV, tQ, tU, tS, D, P));
  return FASTMAX_OK;
}

}  // namespace fm
