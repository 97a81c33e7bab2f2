// fm_moments.cu -- FASTMAX_ALGO_MOMENTS: the chunked causal scan of FastMax on tcgen05/TMEM/TMA.
//
// A sequence is cut into segments of L tokens.  Within a segment the causal polynomial scores
// are evaluated directly (intra part); everything before the segment enters through the moment
// state (fm_moments.cuh) of the segment boundary:
//
//   build : one pass over the tokens keeps S1,S2 (fp32) resident in tensor memory, adds
//           Phi(k_j)^T [v_j] per 64-token step with tcgen05.mma (A = Khatri-Rao rows k[x]*k[y]
//           generated in registers and written to TMEM, B = the V tile brought by TMA) and
//           snapshots the running state as bf16 at every segment boundary.
//   apply : per 512 tokens, streams the 65 slabs of the boundary state through shared memory
//           with TMA; two-stage contraction: T_x = Y_tile . slab_x on the tensor pipe, then
//           O += X[x] * T_x on the CUDA cores while T is read out of tensor memory.
//   The same two kernels serve the backward pass with other operands (see moments_bwd); for
//   p = 1 the intra kernels (fm_intra.cu) apply the single-slab state term themselves.
//
// Reference functions covered: compute_F_masked / compute_g_masked (fastmax.py:219-322) and
// gradient_o_{q,k,v}_masked (fastmax.py:433-691); the non-causal variants use one whole-sequence
// state (fastmax.py:185-285, 384-645).
#include "fm_common.cuh"
#include "fm_sm100.cuh"
#include "fm_moments.cuh"
#include <stdlib.h>

FM_TIMING_DEFINE(moments)

#ifdef FM_TRACE
// developer build: event times of one apply CTA, [event][slab pair]
static __device__ long long fm_trace_apply2[8][40];
extern "C" int fastmax_debug_trace_fetch_apply2(long long* out) {
  return cudaMemcpyFromSymbol(out, fm_trace_apply2, sizeof(fm_trace_apply2)) == cudaSuccess ? 0 : -4;
}
#define FM_TRA(ev, j) do { if (blockIdx.x == 5 && blockIdx.y == 100 && (j) < 40) fm_trace_apply2[ev][j] = clock64(); } while (0)
#else
#define FM_TRA(ev, j) ((void)0)
#endif

namespace fm {
namespace mom {
using namespace sm100;

typedef __nv_bfloat16 bf16;

// =============================================================================================
// prep: blocked transpose (B*H, N, 64) -> (B*H, N/64, 64 features, 64 tokens) so that the build
// kernel can read token-major runs of one feature with 16-byte shared loads and every 64-token
// block is one contiguous 8 KiB TMA box.  HBM-bound (reads + writes the tensor once).
// =============================================================================================
__global__ void __launch_bounds__(256) transpose_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int N,
                                                        int Npad) {
  // tile[d][tok ^ 4*(d >> 3)], rows padded to 72 elements (144 B, 16-byte aligned for the vector
  // read-back).  The XOR spreads the 2-byte transposing stores of a warp (8 feature chunks x 4 tokens)
  // over the banks: with 16-byte aligned rows an 8-row stride alone always hits one bank.
  __shared__ __align__(16) bf16 tile[64][72];
  const int bh = blockIdx.y, n0 = blockIdx.x * 64;
  const bf16* src = in + (size_t)bh * N * 64;
  bf16* dst = out + ((size_t)bh * (Npad / 64) + blockIdx.x) * 4096;
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = threadIdx.x + 256 * it;       // 512 chunks of 8 features: token r, features [8c, 8c+8)
    const int r = idx >> 3, c = idx & 7;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (n0 + r < N) v = __ldg(reinterpret_cast<const uint4*>(src + (size_t)(n0 + r) * 64 + 8 * c));
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    const int col = r ^ (4 * c);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      tile[8 * c + 2 * j][col] = __ushort_as_bfloat16((unsigned short)(w[j] & 0xffffu));
      tile[8 * c + 2 * j + 1][col] = __ushort_as_bfloat16((unsigned short)(w[j] >> 16));
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 2; ++it) {
    const int idx = threadIdx.x + 256 * it;       // feature d, tokens [8m, 8m+8)
    const int d = idx >> 3, m = idx & 7;
    const int k = d >> 3;                          // tokens were stored at tok ^ 4k
    uint4 v = *reinterpret_cast<const uint4*>(&tile[d][8 * (m ^ (k >> 1))]);
    if (k & 1) v = make_uint4(v.z, v.w, v.x, v.y); // 4-token halves of the chunk are swapped
    *reinterpret_cast<uint4*>(dst + d * 64 + 8 * m) = v;
  }
}

// =============================================================================================
// build kernel
// =============================================================================================
struct BuildParams {
  int N, nblk, nproc, Lb, nseg, reverse, causal, p;
  int special_cx;   // blockIdx.x of the CTA that owns the special tile (S1 | normaliser): last CTA for p=2, 0 for p=1
  int packed;       // 1: symmetric-packed second-order tiles (5 regular CTAs per (b,h)), 0: all 64x64 pairs (8 CTAs)
  uint8_t* state;   // records: causal -> [bh][nseg], else [bh]
  const uint8_t* init;   // sequence parallel: record per (b,h) the scan starts from (nullable)
  unsigned long long dW[4][4];    // B-operand descriptors [ring stage][k-step] of the W tile (host-built)
  unsigned long long dW2[4][4];   // ... and of the W2 tile (special CTA)
};

// CTA layout of the build kernel (p = 2): 8 "regular" CTAs per (b,h), each owning 4 tiles
// (= 8 slabs x 64 y rows, four 64-column fp32 accumulators) plus one "special" CTA with the two
// accumulators of the special tile.  Tensor memory: accumulators in columns [0,256), A operand
// buffers (4 tiles x 2 stages x 32 columns) in [256,512).
//
// Roles (25 warps): generator group k = warps 4k..4k+3 produces MMA k-step k (16 of the 64 tokens of a
// block) of all four tiles; issuer warp 16+k issues that k-step's four MMAs (converged warp, operands
// from the constant bank / immediates, see fm_sm100.cuh), so every issuer waits on one barrier per
// block; warp 20 is the TMA producer; warps 21..24 write the bf16 snapshots at segment boundaries.
constexpr int B_STAGES = 4;
constexpr int B_STAGE_BYTES = 3 * 8192;
constexpr int B_SNAP_BYTES = 8 * SLAB_BYTES;             // bf16 staging of the CTA's eight slabs at a boundary
constexpr int B_MIRROR_BYTES = 8 * SLAB_BYTES;           // ... and of their mirror images (packed tiles)
constexpr int B_SMEM = B_STAGES * B_STAGE_BYTES + B_SNAP_BYTES + B_MIRROR_BYTES + 1024 /*barriers*/ + 1024 /*align slack*/;
constexpr int B_ACOL = 256;
constexpr int B_CTAS_P2 = 9;
constexpr int B_CTAS_P2_PACKED = 6;

// Symmetric packing.  S2[x][y][a] = S2[y][x][a], so only the 16x16 blocks (gx, by) of the (x, y) plane with
// by >= gx are accumulated: ten blocks of 256 rows = twenty 128-row tiles instead of thirty-two, four tiles
// (two blocks) per CTA.  Row r of a tile covers x = 16 gx + 8 half + (r >> 4), y = 16 by + (r & 15).  At a
// segment boundary the off-diagonal blocks are written twice, as rows y of slab x and as rows x of slab y,
// so the state record stays the full symmetric tensor that the K-major contractions (dq, dk) need.
__device__ __forceinline__ void packed_block(int b, int& gx, int& by) {
  gx = b < 4 ? 0 : (b < 7 ? 1 : (b < 9 ? 2 : 3));
  by = gx + b - (gx == 0 ? 0 : (gx == 1 ? 4 : (gx == 2 ? 7 : 9)));
}
// feature pair (x, y) accumulated in row `row` of tile `t` of regular CTA `cx`
__device__ __forceinline__ void tile_row_xy(int packed, int cx, int t, int row, int& x, int& y) {
  if (packed) {
    int gx, by;
    packed_block(2 * cx + (t >> 1), gx, by);
    x = 16 * gx + 8 * (t & 1) + (row >> 4), y = 16 * by + (row & 15);
  } else {
    x = 8 * cx + 2 * t + (row >> 6), y = row & 63;
  }
}
constexpr int B_THREADS = 25 * 32;

// Segment-boundary schedule of the scan without integer division in the hot loops.
struct BoundaryIter {
  int c, Lb, reverse, causal, last;
  __device__ BoundaryIter(const BuildParams& P) : Lb(P.Lb), reverse(P.reverse), causal(P.causal), last(P.nproc - 1) {
    c = P.reverse ? (P.nblk - 1) % P.Lb : 0;
  }
  // true if a snapshot is due after block number `it` of the scan
  __device__ bool step(int it) {
    if (!causal) return it == last;
    bool bnd;
    if (reverse) {
      bnd = c == 0;
      c = bnd ? Lb - 1 : c - 1;
    } else {
      bnd = c == Lb - 1;
      c = bnd ? 0 : c + 1;
    }
    return bnd;
  }
};

__device__ __forceinline__ void lds128(uint32_t saddr, uint32_t* r) {
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}

__global__ void __launch_bounds__(B_THREADS, 1)
build_kernel(const __grid_constant__ CUtensorMap tmXT, const __grid_constant__ CUtensorMap tmW,
             const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmS,
             const __grid_constant__ CUtensorMap tmS16, const BuildParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* snap = smem + B_STAGES * B_STAGE_BYTES;   // [8 slabs][8 KiB], 128B-swizzled rows (TMA store source)
  uint8_t* mir = snap + B_SNAP_BYTES;                 // packed tiles: the same rows transposed (rows x of slab y)
  uint64_t* bars = reinterpret_cast<uint64_t*>(mir + B_MIRROR_BYTES);
  uint64_t* full = bars;                 // [4] block operands landed (TMA)
  uint64_t* done = bars + 4;             // [4] every MMA of the block in this ring slot complete (count ntile)
  uint64_t* a_ready = bars + 8;          // [4 tiles][2] A buffer written by the 4 warps of the group (count 4)
  uint64_t* acc_full = bars + 16;        // accumulators final for this segment boundary (count 4: one commit per issuer)
  uint64_t* acc_free = bars + 20;        // ... and read out by the four snapshot warps (count 4)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y, cx = blockIdx.x;
  const bool special = cx == P.special_cx;
  const int ntile = special ? 1 : 4;
  [[maybe_unused]] const long long cta_t0 = FM_NOW();
  const uint32_t stage_tx = special ? 3 * 8192 : 2 * 8192;

  if (warp == 20) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    for (int i = 0; i < B_STAGES; ++i) mbar_init(&full[i], 1), mbar_init(&done[i], 4);
    for (int i = 0; i < 8; ++i) mbar_init(&a_ready[i], 4);
    mbar_init(acc_full, 4), mbar_init(acc_free, 4);
    fence_barrier_init();
    tma_prefetch_desc(&tmXT);
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmW2);
    tma_prefetch_desc(&tmS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  assert_bases(tmem, smem);
  // Accumulators start from the incoming state of the sequence-parallel exchange (the state of the
  // earlier -- reverse scan: later -- ranks, so every snapshot already contains it) or from zero;
  // every MMA then accumulates.
  const bool has_init = P.init != nullptr;
  if (warp < 16 && (warp >> 2) < ntile) {
    const int g = warp >> 2, q = warp & 3;
    const int row = 32 * q + lane;
    int xg, y;
    tile_row_xy(P.packed, cx, g, row, xg, y);
    if (special) y = row & 63;
    const uint8_t* rec = has_init ? P.init + (size_t)bh * REC_BYTES : nullptr;
    const int nacc = special ? 2 : 1;
    for (int a2 = 0; a2 < nacc; ++a2) {
      const int slab = special ? 64 + a2 : xg;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = 0u;
        if (has_init) {
          const float* vec = reinterpret_cast<const float*>(rec + VEC_OFF);
          if (!special || row < 64) {
            if (special && a2 == 1) {          // fp32 copy of the normaliser matrix
#pragma unroll
              for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(vec[128 + y * 64 + 32 * h + c]);
            } else {
              const uint4* src = reinterpret_cast<const uint4*>(rec + (size_t)slab * SLAB_BYTES + (size_t)y * 128 + 64 * h);
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                const uint4 w4 = src[c8];
                const uint32_t w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) v[8 * c8 + 2 * j] = w[j] << 16, v[8 * c8 + 2 * j + 1] = w[j] & 0xffff0000u;
              }
            }
          } else if (row == 64) {
#pragma unroll
            for (int c = 0; c < 32; ++c) v[c] = __float_as_uint(vec[a2 * 64 + 32 * h + c]);
          }
        }
        tmem_st32(((uint32_t)(32 * q) << 16) + 64 * (g + a2) + 32 * h, v);
      }
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (warp < 16) {
    // ------------------------------------------------------------------ generator groups ("k-split")
    // The 64 tokens of a block are four MMA k-steps of 16 tokens.  Group k (4 warps = the 4 TMEM lane
    // quadrants) produces k-step k of ALL tiles of the CTA: it reads its 16-token run of feature row y
    // once and multiplies it with the runs of the tiles' x rows.  (One group per tile would re-read
    // the y rows four times: 640 shared-memory wavefronts per block against 512 cycles of MMA.)
    const int k = warp >> 2, q = warp & 3;
    const int row = 32 * q + lane;                // TMEM lane == feature pair (x, y) of every tile
    // byte offsets inside the 128B-swizzled XT tile of the two 16-byte chunks holding tokens
    // [16k, 16k+16) of the y rows (one per pair of tiles) and of the x rows of the four tiles
    uint32_t yo[2][2], xo[4][2];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      int x, y;
      tile_row_xy(P.packed, cx, t, row, x, y);
      if (special) x = 0, y = row & 63;
      x &= 63;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        xo[t][c] = (uint32_t)x * 128 + ((uint32_t)((2 * k + c) ^ (x & 7)) << 4);
        if ((t & 1) == 0) yo[t >> 1][c] = (uint32_t)y * 128 + ((uint32_t)((2 * k + c) ^ (y & 7)) << 4);
      }
    }
    const uint32_t fill = row == 64 ? 0x3F803F80u : 0u;
    const uint32_t a_col0 = ((uint32_t)(32 * q) << 16) + B_ACOL + 8 * k;
#pragma unroll 1
    for (int it = 0; it < P.nproc; ++it) {
      const int s = it & 3, ph = (it >> 2) & 1, buf = it & 1;
      mbar_wait(&full[s], ph);
      const uint32_t xt = FM_SMEM_BASE + s * B_STAGE_BYTES;
      uint32_t a[4][8];
      if (!special) {
#pragma unroll
        for (int tp = 0; tp < 2; ++tp) {         // a pair of tiles shares its y rows
          uint32_t yr[8];
          lds128(xt + yo[tp][0], yr);
          lds128(xt + yo[tp][1], yr + 4);
#pragma unroll
          for (int th = 0; th < 2; ++th) {
            const int t = 2 * tp + th;
            uint32_t v[8];
            lds128(xt + xo[t][0], v);
            lds128(xt + xo[t][1], v + 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) a[t][i] = mul_bf16x2(v[i], yr[i]);
          }
        }
      } else {
        // special tile: rows 0..63 = Y (S1 / normaliser matrix), row 64 = ones (S0 / normaliser vector)
        uint32_t yr[8];
        lds128(xt + yo[0][0], yr);
        lds128(xt + yo[0][1], yr + 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[0][i] = row < 64 ? yr[i] : fill;
      }
      // A buffer `buf` was last read by the MMAs of block it-2
      if (it >= 2) mbar_wait(&done[(it - 2) & 3], ((it - 2) >> 2) & 1);
      tc_fence_after();
#pragma unroll
      for (int t = 0; t < 4; ++t)
        if (t < ntile) tmem_st8(a_col0 + 64 * t + 32 * buf, a[t]);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&a_ready[k * 2 + buf]);
    }
  } else if (warp < 20) {
    // ------------------------------------------------------------------ MMA issuers (converged warps)
    // issuer kk: k-step kk of every tile, as soon as generator group kk has written it
    const int kk = warp - 16;
    constexpr uint32_t IDESC = make_idesc_bf16(128, 64, false, true);
    constexpr uint32_t BAR0 = FM_SMEM_BASE + B_STAGES * B_STAGE_BYTES + B_SNAP_BYTES + B_MIRROR_BYTES;   // shared address of bars[0]
    BoundaryIter bi(P);
    int snaps = 0;
    bool pending = false;
#pragma unroll 1
    for (int it0 = 0; it0 < P.nproc; it0 += 4) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {          // j = ring stage (compile-time), j & 1 = A buffer
        const int it = it0 + j;
        if (it < P.nproc) {
          const bool bnd = bi.step(it);
          if (lane == 0) {
            [[maybe_unused]] const long long i0 = FM_NOW();
            if (pending) mbar_wait(acc_free, (snaps - 1) & 1);
            [[maybe_unused]] const long long i1 = FM_NOW();
            // group kk read the block's tiles itself, so their TMA has landed
            mbar_wait(&a_ready[kk * 2 + (j & 1)], (it >> 1) & 1);
#ifdef FM_TIMING
            if (kk == 0 && cx == 0) FM_T(20, i1 - i0), FM_T(21, FM_NOW() - i1), FM_T(22, 1);
            if (kk == 0 && special) FM_T(28, i1 - i0), FM_T(29, FM_NOW() - i1), FM_T(30, 1);
#endif
          }
          __syncwarp();
          tc_fence_after();
          if (elect_one()) {
            // (the k-step offsets of the A columns and of the descriptors are selected per warp)
#define FM_BUILD_ISSUE(KK)                                                                              \
  if (!special) {                                                                                       \
    _Pragma("unroll") for (int t = 0; t < 4; ++t)                                                       \
        umma_ts(64 * t, B_ACOL + 64 * t + 32 * (j & 1) + 8 * KK, P.dW[j][KK], IDESC, 1);                 \
  } else {                                                                                              \
    umma_ts(0, B_ACOL + 32 * (j & 1) + 8 * KK, P.dW[j][KK], IDESC, 1);                                   \
    umma_ts(64, B_ACOL + 32 * (j & 1) + 8 * KK, P.dW2[j][KK], IDESC, 1);                                 \
  }
            if (kk == 0) { FM_BUILD_ISSUE(0) } else if (kk == 1) { FM_BUILD_ISSUE(1) }
            else if (kk == 2) { FM_BUILD_ISSUE(2) } else { FM_BUILD_ISSUE(3) }
            umma_commit_addr(BAR0 + 8 * (4 + j));   // done[j]
            if (bnd) umma_commit(acc_full);
          }
          __syncwarp();
          if (bnd) ++snaps;
          pending = bnd;
        }
      }
    }
  } else if (warp == 20) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      for (int it = 0; it < P.nproc; ++it) {
        const int s = it & 3;
        if (it >= B_STAGES) mbar_wait(&done[s], ((it - B_STAGES) >> 2) & 1);   // MMAs of block it-4 complete
        const int b = P.reverse ? P.nblk - 1 - it : it;
        uint8_t* st = smem + s * B_STAGE_BYTES;
        mbar_arrive_expect_tx(&full[s], stage_tx);
        tma_load_3d(st, &tmXT, &full[s], 0, b * 64, bh);
        tma_load_3d(st + 8192, &tmW, &full[s], 0, b * 64, bh);
        if (special) tma_load_3d(st + 16384, &tmW2, &full[s], 0, b * 64, bh);
      }
    }
  } else {
    // ------------------------------------------------------------------ snapshot writers (warps 21..24)
    // The scan stalls while the accumulators are read, so the read-out is kept short: tcgen05.ld ->
    // bf16 -> shared memory (swizzled rows), release the accumulators, and only then let one thread
    // send the eight slabs to the record with bulk tensor stores.  (Storing from registers, 16 bytes
    // per lane at a 128-byte stride, held the accumulators for ~4.8k cycles per boundary.)
    const int q = warp & 3;
    const int row = 32 * q + lane;
    const int xsel = row >> 6, y = row & 63;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
    const bool storer = warp == 21 && lane == 0;
    // packed tiles: the CTA's two blocks (gx, by); off-diagonal ones are also stored mirrored
    int pgx[2] = {0, 0}, pby[2] = {0, 0};
    if (P.packed && !special) {
      packed_block(2 * cx, pgx[0], pby[0]);
      packed_block(2 * cx + 1, pgx[1], pby[1]);
    }
    BoundaryIter bi(P);
    int kq = 0;
    for (int it = 0; it < P.nproc; ++it) {
      if (!bi.step(it)) continue;
      const int b = P.reverse ? P.nblk - 1 - it : it;
      const int seg = !P.causal ? 0 : (P.reverse ? b / P.Lb - 1 : (b + 1) / P.Lb);
      const int rec_idx = P.causal ? bh * P.nseg + seg : bh;
      uint8_t* rec = P.state + (size_t)rec_idx * REC_BYTES;
      if (storer) bulk_wait_group_read0();          // previous boundary's stores have left the staging tiles
      [[maybe_unused]] const long long s0 = FM_NOW();
      mbar_wait(acc_full, kq & 1);
      [[maybe_unused]] const long long s1 = FM_NOW();
      tc_fence_after();
      if (!special) {
        named_barrier_sync(1, 128);
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
          uint32_t v0[32], v1[32];
          tmem_ld32(lane_addr + 64 * g, v0);
          tmem_ld32(lane_addr + 64 * g + 32, v1);
          tmem_wait_ld_dep32(v0);
          tmem_wait_ld_dep32(v1);
          // staging row of this thread: unpacked -> slab (2g + xsel) of the CTA, row y;
          // packed -> 16-row unit (8g + (row >> 4)) = rows [16 by, 16 by + 16) of slab x, row (row & 15)
          const int srow = P.packed ? (row & 15) : y;
          const uint32_t dst = smem_u32(snap) + (P.packed ? (uint32_t)(8 * g + (row >> 4)) * 2048u + (uint32_t)srow * 128
                                                          : (uint32_t)(2 * g + xsel) * SLAB_BYTES + (uint32_t)y * 128);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t* v = c < 4 ? v0 + 8 * c : v1 + 8 * (c - 4);
            sts128(dst + ((uint32_t)(c ^ (srow & 7)) << 4),
                   pack_bf16x2(__uint_as_float(v[0]), __uint_as_float(v[1])),
                   pack_bf16x2(__uint_as_float(v[2]), __uint_as_float(v[3])),
                   pack_bf16x2(__uint_as_float(v[4]), __uint_as_float(v[5])),
                   pack_bf16x2(__uint_as_float(v[6]), __uint_as_float(v[7])));
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_free);
        [[maybe_unused]] const long long s2 = FM_NOW();
        fence_proxy_async();
        named_barrier_sync(1, 128);
        if (!P.packed) {
          if (storer) {
#pragma unroll 1
            for (int sl = 0; sl < 8; ++sl) tma_store_3d(&tmS, snap + sl * SLAB_BYTES, 0, (8 * cx + sl) * 64, rec_idx);
            bulk_commit_group();
          }
        } else {
          if (storer) {
            // block blk = staging units [16 blk, 16 blk + 16): rows [16 by, +16) of the 16 slabs x = 16 gx + xr --
            // ONE bulk tensor store per block through the 4-D map (a store per 2 KiB unit, 64 per boundary, cost
            // ~100 cycles each: the snapshot warps then needed longer than the scan takes between boundaries)
            for (int blk = 0; blk < 2; ++blk) tma_store_4d(&tmS16, snap + blk * 32768, 0, 16 * pby[blk], 16 * pgx[blk], rec_idx);
          }
          // Mirror images of the off-diagonal blocks: atom (xr, yi, chunk c) of the staging moves from unit xr,
          // row yi to unit yi, row xr.  Lane l of an 8-lane phase takes xr = l (mod 8), yi = l + s (mod 8) with one
          // chunk c, so the eight 16-byte reads and the eight writes of a phase each cover all 32 banks once.
          const int l7 = lane & 7, gid = 4 * q + (lane >> 3);
          const int xr = 8 * (gid & 1) + l7, yh = (gid >> 1) & 1, s_lo = gid >> 2;
#pragma unroll 1
          for (int blk = 0; blk < 2; ++blk) {
            if (pgx[blk] == pby[blk]) continue;
#pragma unroll 4
            for (int i2 = 0; i2 < 16; ++i2) {
              const int sft = s_lo + 4 * (i2 >> 3), c = i2 & 7;
              const int yi = 8 * yh + ((l7 + sft) & 7);
              uint32_t w[4];
              lds128(smem_u32(snap) + (uint32_t)(16 * blk + xr) * 2048u + (uint32_t)yi * 128 + ((uint32_t)(c ^ (yi & 7)) << 4), w);
              sts128(smem_u32(mir) + (uint32_t)(16 * blk + yi) * 2048u + (uint32_t)xr * 128 + ((uint32_t)(c ^ (xr & 7)) << 4),
                     w[0], w[1], w[2], w[3]);
            }
          }
          fence_proxy_async();
          named_barrier_sync(1, 128);
          if (storer) {
            // mirror of block blk: rows [16 gx, +16) of the 16 slabs y = 16 by + yi
            for (int blk = 0; blk < 2; ++blk)
              if (pgx[blk] != pby[blk]) tma_store_4d(&tmS16, mir + blk * 32768, 0, 16 * pgx[blk], 16 * pby[blk], rec_idx);
            bulk_commit_group();
          }
        }
#ifdef FM_TIMING
        if (storer && cx == 0) FM_T(16, 1), FM_T(17, s1 - s0), FM_T(18, s2 - s1), FM_T(19, FM_NOW() - s2);
#endif
      } else {
        // special tile (S1 | normaliser matrix, row 64: S0 | normaliser vector): 1/33 of the state, stored directly
        for (int a2 = 0; a2 < 2; ++a2) {
          const int slab = 64 + a2;
#pragma unroll 1
          for (int h = 0; h < 2; ++h) {
            uint32_t v[32];
            tmem_ld32(lane_addr + 64 * a2 + 32 * h, v);
            tmem_wait_ld_dep32(v);
            if (a2 == 1 && h == 1) {
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(acc_free);
            }
            if (row < 64) {
              uint4* dst = reinterpret_cast<uint4*>(rec + (size_t)slab * SLAB_BYTES + (size_t)y * 128 + 64 * h);
#pragma unroll
              for (int c8 = 0; c8 < 4; ++c8) {
                uint4 o;
                o.x = pack_bf16x2(__uint_as_float(v[8 * c8]), __uint_as_float(v[8 * c8 + 1]));
                o.y = pack_bf16x2(__uint_as_float(v[8 * c8 + 2]), __uint_as_float(v[8 * c8 + 3]));
                o.z = pack_bf16x2(__uint_as_float(v[8 * c8 + 4]), __uint_as_float(v[8 * c8 + 5]));
                o.w = pack_bf16x2(__uint_as_float(v[8 * c8 + 6]), __uint_as_float(v[8 * c8 + 7]));
                dst[c8] = o;
              }
              if (slab == 65) {  // fp32 copy of the normaliser matrix
                uint4* d32 = reinterpret_cast<uint4*>(rec + VEC_OFF + (size_t)(128 + y * 64 + 32 * h) * 4);
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) d32[c4] = make_uint4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
              }
            } else if (row == 64) {
              uint4* d32 = reinterpret_cast<uint4*>(rec + VEC_OFF + (size_t)(a2 * 64 + 32 * h) * 4);
#pragma unroll
              for (int c4 = 0; c4 < 8; ++c4) d32[c4] = make_uint4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
            }
          }
        }
#ifdef FM_TIMING
        if (storer) FM_T(24, 1), FM_T(25, s1 - s0), FM_T(26, FM_NOW() - s1);
#endif
      }
      ++kq;
    }
    if (storer) bulk_wait_group_read0();            // shared memory must outlive the bulk stores' reads (not their global writes)
  }
  tc_fence_before();
  __syncthreads();
#ifdef FM_TIMING
  if (tid == 0) FM_T(special ? 33 : 32, 1), FM_T(special ? 35 : 34, FM_NOW() - cta_t0);
#endif
  if (warp == 20) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// apply kernel ("two-stage" contraction)
//
//   out[i][c] = scale_i ( sum_x coef X[i][x] * T_x[i][c]  +  coef64 * T_64[i][c]  + vec[c] + partial[i][c] )
//   T_x = Y_tile . B_x        (B_x = slab x of the boundary state, 64x64)
//
// i.e. the second-order contraction  sum_{x,y} X[x] Y[y] S[x][y][c]  is split into a plain GEMM of the
// token tile Y (128x64, resident in shared memory) against FOUR slabs at once (one 128x256x16
// tcgen05.mma per k-step: large-N MMAs are the only ones a single issuing thread can keep the
// tensor pipe busy with, see tools/mmarate.py) and a per-token scalar multiply-add over x that runs
// on the CUDA cores while the fp32 T tile is read out of tensor memory.  No operand generation, and
// the products with X stay in fp32.
//
//   forward o : X = Y = q, slabs read MN-major (contract y)          coef a^2/2, coef64 a
//   dq        : X = q, Y = do/g, slabs read K-major (contract a)      coef a,     coef64 1
//   dv        : X = Y = k on the reverse states, MN-major             coef a^2/2, coef64 a
//   dk        : X = k, Y = v on the reverse states, K-major           coef a,     coef64 1
// =============================================================================================
enum ApplyMode { APPLY_FWD = 0, APPLY_DQ = 1, APPLY_DV = 2, APPLY_DK = 3 };

struct ApplyParams {
  int N, L, nseg, causal, p, mode;
  int state_reverse;        // 1: records are exclusive suffixes (last segment has none)
  int b_kmajor;             // 1: contraction over the slab's contiguous index (dq, dk)
  float coef_reg, coef64, out_scale;
  int vec_sel;              // -1 none, 0: vec[0:64), 1: vec[64:128)
  int edge_state;           // sequence parallel: first (forward states) / last (reverse) local segment has a state
  int H, Hkv;               // grouped-query attention: heads per batch entry of the outputs / of k, v
  int state_kv, xy_kv;      // 1: the state records / the X and Y operands are indexed by (b, kv head)
  const bf16* X;            // per-token scalars
  const float* partial;     // (BH, N, 64) fp32 intra-segment partial (nullable)
  const float* gden;        // (BH, N) fp32: divide rows by it (forward) (nullable)
  bf16* out;
  const uint8_t* state;
};

// Shared-memory descriptors of the apply kernel's MMA operands, built on the host (constant bank: one thread then
// issues an MMA in ~50 cycles instead of ~110 with descriptors assembled in registers; event trace of one CTA: the
// four MMAs + commit of a (tile, slab pair) step took ~500 cycles of a ~1800-cycle step)
struct ApplyDescs {
  unsigned long long aK[4][4];        // resident Y tile t, K-major, k-step kk
  unsigned long long bD[6][4];        // ring stage, k-step kk: slabs read MN-major (o, dv) or K-major (dq, dk)
};

constexpr int A_STAGES = 6;
constexpr int A_STAGE_BYTES = 2 * 8192;                  // two slabs per MMA (N = 128)
constexpr int A_TOK = APPLY_TOK;                         // 4 token tiles per CTA
constexpr int A_RES_BYTES = 4 * 16384;                   // four resident 128x64 Y tiles
constexpr int A_SMEM = A_RES_BYTES + A_STAGES * A_STAGE_BYTES + 1024 + 1024;
#ifndef FM_APPLY_NAP
#define FM_APPLY_NAP 40
#endif
constexpr int A_ISSUERS = 2;   // (4, one per tile: 4.2 ms per launch instead of 2.2 -- a 21st warp costs every thread registers)
constexpr int A_PRODUCER = 16 + A_ISSUERS;               // warp index of the TMA producer
constexpr int A_THREADS = (17 + A_ISSUERS) * 32;         // 16 epilogue warps, MMA issuers, TMA producer

// Each token tile owns a private 128-column T buffer in tensor memory and its own barrier pair.
// In-kernel cycle counters (-DFM_TIMING, tools/timing.py) at the headline shape: per tile and slab pair
// the chain  MMA (4 x 128x128x16 = 256 cycles) -> commit -> read-out of 128 columns (~1300 cycles: 16
// warps share the tensor-memory read path at ~2.6 cycles per 32-lane column per quadrant) -> free  is
// serial, because 512 columns are exactly four single buffers; tensor pipe ~45 % and read-out warps ~75 %
// busy.  MMAs and tcgen05.ld do not slow each other (tools/contention.py), but a tcgen05.commit stalls
// its issuing thread ~90 cycles (tools/multi_issue.py), which is why two issuer warps beat one.
// Measured alternatives (ms per apply launch; this layout 2.3): one converged constant-bank issuer that
// also feeds the TMA ring 2.9; register re-balance (setmaxnreg) with two read-out loads in flight 2.9;
// two tiles per CTA, eight warps per tile, two N=128 buffers per tile, one issuer 3.5, two issuers 3.7
// (the state is then streamed twice); two 64-column buffers per tile (N=64 MMAs from shared memory cost
// 48 cycles instead of 32) 3.7; two shared N=256 T buffers with one issuer 2.6; two tiles per CTA with a
// column-split epilogue 2.9; issuing from the epilogue warpgroups 3.2; generated A operands in tensor
// memory 3.2-5.2.
__global__ void __launch_bounds__(A_THREADS, 1)
apply_kernel(const __grid_constant__ CUtensorMap tmS, const __grid_constant__ CUtensorMap tmY,
             const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ ApplyDescs DS, const ApplyParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* res = smem;                       // [4 tiles][16 KiB]
  uint8_t* ring = smem + A_RES_BYTES;        // [6 stages][16 KiB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + A_STAGES * A_STAGE_BYTES);
  uint64_t* full = bars;            // [6] slab pair landed
  uint64_t* empty = bars + 6;       // [6] slab pair consumed by the MMAs of all four tiles (count 2: one per issuer)
  uint64_t* t_full = bars + 12;     // [4] T tile of (slab pair, tile t) is in tensor memory
  uint64_t* t_free = bars + 16;     // [4] ... and has been read out by warpgroup t (count 4: one per warp)
  uint64_t* res_full = bars + 20;    // [4] resident Y tile t landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  float* vec_s = reinterpret_cast<float*>(bars + 26);  // [64]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y;
  const int t0 = blockIdx.x * A_TOK;
  const int seg = t0 / P.L;
  const bool has_state =
      !P.causal || (P.state_reverse ? (seg < P.nseg - 1 || P.edge_state) : (seg > 0 || P.edge_state));
  const int kvh = kv_index(bh, P.H, P.Hkv);
  const int bh_st = P.state_kv ? kvh : bh, bh_xy = P.xy_kv ? kvh : bh;
  const int rec_idx = P.causal ? bh_st * P.nseg + seg : bh_st;
  const int ngroups = (P.p == 2 ? 32 : 0) + 1;     // 32 pairs of second-order slabs + slab 64

  if (warp == A_PRODUCER) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    for (int i = 0; i < A_STAGES; ++i) mbar_init(&full[i], 1), mbar_init(&empty[i], A_ISSUERS);
    for (int i = 0; i < 4; ++i) mbar_init(&t_full[i], 1), mbar_init(&t_free[i], 4);
    for (int i = 0; i < 4; ++i) mbar_init(&res_full[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmS);
    tma_prefetch_desc(&tmY);
  }
  if (tid < 64) {
    float v = 0.f;
    if (has_state && P.vec_sel >= 0)
      v = reinterpret_cast<const float*>(P.state + (size_t)rec_idx * REC_BYTES + VEC_OFF)[P.vec_sel * 64 + tid];
    vec_s[tid] = v;
  }
  // The output rows start from the additive terms.  The intra-segment partial (fp32, blocked layout of
  // fm_moments.cuh: [tile][c4][row] float4, 512 contiguous bytes per warp and load) is requested first
  // thing, before the set-up barrier: it is the longest latency in front of the first read-out.
  float O[64];
  if (warp < 16) {
    const float4* part = nullptr;
    if (P.partial && t0 + 128 * (warp >> 2) < P.N)
      part = reinterpret_cast<const float4*>(P.partial + ((size_t)bh * ((P.N + 127) / 128) + (size_t)(t0 / 128 + (warp >> 2))) *
                                                             PART_TILE_FLOATS) + (32 * (warp & 3) + lane);
#pragma unroll
    for (int c4 = 0; c4 < 16; ++c4) {
      const float4 pv = part ? __ldg(part + c4 * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
      O[4 * c4 + 0] = pv.x, O[4 * c4 + 1] = pv.y, O[4 * c4 + 2] = pv.z, O[4 * c4 + 3] = pv.w;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  assert_bases(tmem, smem);

  if (warp == A_PRODUCER) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0 && has_state) {
      // operands of the very first MMAs first: Y tile 0, slab pair 0, then the other tiles
      auto load_tile = [&](int t) {
        mbar_arrive_expect_tx(&res_full[t], 16384);
        for (int h = 0; h < 2; ++h)
          tma_load_3d(res + t * 16384 + h * 8192, &tmY, &res_full[t], 0, t0 + t * 128 + h * 64, bh_xy);
      };
      load_tile(0);
      for (int gi = 0; gi < ngroups; ++gi) {
        const int st = gi % A_STAGES, ph = (gi / A_STAGES) & 1;
        const bool last = gi == ngroups - 1;
        mbar_wait(&empty[st], ph ^ 1, 1000 + gi);
        mbar_arrive_expect_tx(&full[st], last ? 8192 : A_STAGE_BYTES);
        uint8_t* dst = ring + st * A_STAGE_BYTES;
        if (last) {
          tma_load_3d(dst, &tmS, &full[st], 0, 64 * 64, rec_idx);
        } else {
          tma_load_3d(dst, &tmS, &full[st], 0, (2 * gi) * 64, rec_idx);
          tma_load_3d(dst + 8192, &tmS, &full[st], 0, (2 * gi + 1) * 64, rec_idx);
        }
        if (gi == 0)
          for (int t = 1; t < 4; ++t) load_tile(t);
      }
    }
  } else if (warp >= 16) {
    // ---------------------------------------------------------------- MMA issuers (tiles 0,1 | tiles 2,3)
    // converged warps, one elected lane issues; descriptors from the constant bank through indexed loads
    if (has_state) {
      constexpr int TPI = 4 / A_ISSUERS;            // tiles per issuer
      const int tb = (warp - 16) * TPI;
      const uint32_t id128 = make_idesc_bf16(128, 128, false, !P.b_kmajor);
      const uint32_t id64 = make_idesc_bf16(128, 64, false, !P.b_kmajor);
#pragma unroll 1
      for (int gi = 0; gi < ngroups; ++gi) {
        const int st = gi % A_STAGES;
        const bool last = gi == ngroups - 1;
        const uint32_t idesc = last ? id64 : id128;
        if (lane == 0) mbar_wait(&full[st], (gi / A_STAGES) & 1, 2000 + gi);
#pragma unroll
        for (int tt = 0; tt < TPI; ++tt) {
          const int t = tb + tt;
          if (t == 0 && lane == 0) FM_TRA(0, gi);
          if (gi > 0) mbar_wait_spin(&t_free[t], (gi - 1) & 1);                // (every lane polls)
          else mbar_wait_spin(&res_full[t], 0);
          __syncwarp();
          tc_fence_after();
          if (t == 0 && lane == 0) FM_TRA(1, gi);
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_ss(128 * t, DS.aK[t][kk], DS.bD[st][kk], idesc, kk > 0);
            umma_commit(&t_full[t]);
            if (tt == TPI - 1) umma_commit(&empty[st]);
            if (t == 0) FM_TRA(2, gi);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ---------------------------------------------------------------- epilogue warpgroups (warps 0..15)
    const int t = warp >> 2, q = warp & 3;
    const int i = t0 + 128 * t + 32 * q + lane;   // token row of this thread
    const bool valid = i < P.N;
    const size_t grow = (size_t)bh * P.N + i;
    const uint32_t tbuf = tmem + ((uint32_t)(32 * q) << 16) + 128 * t;
#pragma unroll
    for (int c = 0; c < 64; ++c) O[c] += vec_s[c];
    // (the normaliser of the row is fetched here, not in the tail where its latency would be exposed)
    const float inv_g = (valid && P.gden) ? 1.f / __ldg(P.gden + grow) : 1.f;
    if (has_state) {
      const uint32_t* xrow = reinterpret_cast<const uint32_t*>(P.X + (valid ? (size_t)bh_xy * P.N + i : 0) * 64);
      // the pair of X coefficients of the NEXT group is fetched one iteration ahead (ncu: 8 % of the
      // kernel's stall samples sat on this load when it was issued at the point of use)
      uint32_t xw_next = (valid && ngroups > 1) ? __ldg(xrow) : 0u;
#pragma unroll 1
      for (int gi = 0; gi < ngroups; ++gi) {
        const bool last = gi == ngroups - 1;
        float xs[2];
        if (!last) {
          const uint32_t xw = xw_next;
          xw_next = (valid && gi + 2 < ngroups) ? __ldg(xrow + gi + 1) : 0u;
          xs[0] = __uint_as_float(xw << 16) * P.coef_reg;
          xs[1] = __uint_as_float(xw & 0xffff0000u) * P.coef_reg;
        } else {
          xs[0] = P.coef64, xs[1] = 0.f;
        }
        if (warp == 0 && lane == 0) FM_TRA(3, gi);
        // a short nap between polls: the step runs at the power cap, and 512 threads re-polling for about half of the
        // kernel's time cost more than the ~40 ns of wake-up latency (sustained bench 18.8 -> 18.4 ms on the same box)
        while (!mbar_try_wait(&t_full[t], gi & 1)) __nanosleep(FM_APPLY_NAP);
        tc_fence_after();
        if (warp == 0 && lane == 0) FM_TRA(4, gi);
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (j > 0 && last) break;
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            uint32_t v[16];
            tmem_ld16(tbuf + 64 * j + 16 * h, v);
            tmem_wait_ld();
#pragma unroll
            // packed fp32 FMA (fma.rn.f32x2, same rounding as fmaf): two columns per instruction, half the FMA issue of
            // the read-out loop: apply -3 % relative to the build kernel of the same run
            for (int c = 0; c < 16; c += 2) {
              unsigned long long acc, val, xx;
              asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(O[16 * h + c]), "f"(O[16 * h + c + 1]));
              asm("mov.b64 %0, {%1, %2};" : "=l"(val) : "r"(v[c]), "r"(v[c + 1]));
              asm("mov.b64 %0, {%1, %1};" : "=l"(xx) : "f"(xs[j]));
              asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(xx), "l"(val));
              asm("mov.b64 {%0, %1}, %2;" : "=f"(O[16 * h + c]), "=f"(O[16 * h + c + 1]) : "l"(acc));
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&t_free[t]);
        if (warp == 0 && lane == 0) FM_TRA(5, gi);
        if (warp == 3 && lane == 0) FM_TRA(6, gi);
      }
    }
    // The bf16 rows leave through shared memory: the tile's own resident Y tile is free once its last T tile
    // has been read (all of its MMAs are complete), so the 128x64 output tile is staged there in the 128B-swizzled
    // layout and one thread sends it with two bulk tensor stores (rows past N are clipped by the tensor map).
    // (Storing from registers -- a 128-byte row per thread, 16 bytes per lane and instruction at a 128-byte
    // stride -- was most of the ~7k-cycle tail of the CTA.)
    const float scale = P.out_scale * inv_g;
    const int ob = t0 + 128 * t;
    if (ob < P.N) {                                   // uniform over the tile's warpgroup
      const int rloc = 32 * q + lane;
      const uint32_t dst = smem_u32(res) + t * 16384 + rloc * 128;
#pragma unroll
      for (int c8 = 0; c8 < 8; ++c8)
        sts128(dst + ((uint32_t)(c8 ^ (rloc & 7)) << 4),
               pack_bf16x2(O[8 * c8 + 0] * scale, O[8 * c8 + 1] * scale), pack_bf16x2(O[8 * c8 + 2] * scale, O[8 * c8 + 3] * scale),
               pack_bf16x2(O[8 * c8 + 4] * scale, O[8 * c8 + 5] * scale), pack_bf16x2(O[8 * c8 + 6] * scale, O[8 * c8 + 7] * scale));
      fence_proxy_async();
      named_barrier_sync(1 + t, 128);
      if (q == 0 && lane == 0) {
        tma_store_3d(&tmOut, res + t * 16384, 0, ob, bh);
        if (ob + 64 < P.N) tma_store_3d(&tmOut, res + t * 16384 + 8192, 0, ob + 64, bh);
        bulk_commit_group();
        bulk_wait_group_read0();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == A_PRODUCER) tmem_dealloc(tmem, 512);
}

// =============================================================================================
// backward prep: u = do/g (bf16), dlt = (do.o)/g (fp32), qd = -dlt*q (bf16).  HBM-bound pass.
// =============================================================================================
constexpr int PREP_ROWS = 128;   // rows per CTA
template <bool QD>   // QD = false: only u and dlt (the first-order scan kernels do not use qd and q is not read)
__global__ void __launch_bounds__(256) prep_bwd_kernel(const bf16* __restrict__ q, const bf16* __restrict__ o,
                                                        const bf16* __restrict__ d_o, const float* __restrict__ g,
                                                        bf16* __restrict__ u, float* __restrict__ dlt,
                                                        bf16* __restrict__ qd, size_t rows, int N, int Npad) {
  // eight threads per row, 16 bytes (8 features) each; a CTA covers 128 rows in four sweeps of 32 whose loads are all
  // issued before the first use (four times the bytes in flight per thread: the pass is latency-bound otherwise)
  constexpr int SW = PREP_ROWS / 32;
  const int c = threadIdx.x & 7;
  uint4 dv[SW], ov[SW], qv[SW];
  float gg[SW];
  size_t row[SW];
#pragma unroll
  for (int s = 0; s < SW; ++s) {
    row[s] = (size_t)blockIdx.x * PREP_ROWS + 32 * s + (threadIdx.x >> 3);
    const size_t off = (row[s] < rows ? row[s] : 0) * 64 + 8 * c;
    dv[s] = __ldg(reinterpret_cast<const uint4*>(d_o + off));
    ov[s] = __ldg(reinterpret_cast<const uint4*>(o + off));
    qv[s] = QD ? __ldg(reinterpret_cast<const uint4*>(q + off)) : make_uint4(0u, 0u, 0u, 0u);
    gg[s] = g[row[s] < rows ? row[s] : 0];
  }
#pragma unroll
  for (int s = 0; s < SW; ++s) {
    const bool live = row[s] < rows;
    const size_t off = (live ? row[s] : 0) * 64 + 8 * c;
    const float ig = 1.f / gg[s];
    const uint32_t dw[4] = {dv[s].x, dv[s].y, dv[s].z, dv[s].w}, ow[4] = {ov[s].x, ov[s].y, ov[s].z, ov[s].w};
    const uint32_t qw[4] = {qv[s].x, qv[s].y, qv[s].z, qv[s].w};
    float d[8], qf[8], sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      d[2 * j] = __uint_as_float(dw[j] << 16), d[2 * j + 1] = __uint_as_float(dw[j] & 0xffff0000u);
      qf[2 * j] = __uint_as_float(qw[j] << 16), qf[2 * j + 1] = __uint_as_float(qw[j] & 0xffff0000u);
      sum = fmaf(d[2 * j], __uint_as_float(ow[j] << 16), sum);
      sum = fmaf(d[2 * j + 1], __uint_as_float(ow[j] & 0xffff0000u), sum);
    }
#pragma unroll
    for (int x = 4; x; x >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, x);
    const float dl = sum * ig;
    if (!live) continue;
    uint4 uo, qo;
    uo.x = pack_bf16x2(d[0] * ig, d[1] * ig), uo.y = pack_bf16x2(d[2] * ig, d[3] * ig);
    uo.z = pack_bf16x2(d[4] * ig, d[5] * ig), uo.w = pack_bf16x2(d[6] * ig, d[7] * ig);
    qo.x = pack_bf16x2(-dl * qf[0], -dl * qf[1]), qo.y = pack_bf16x2(-dl * qf[2], -dl * qf[3]);
    qo.z = pack_bf16x2(-dl * qf[4], -dl * qf[5]), qo.w = pack_bf16x2(-dl * qf[6], -dl * qf[7]);
    *reinterpret_cast<uint4*>(u + off) = uo;
    if (QD) *reinterpret_cast<uint4*>(qd + off) = qo;
    if (c == 0) dlt[(row[s] / N) * Npad + row[s] % N] = dl;   // (BH, Npad) layout, pad stays zero
  }
}

// =============================================================================================
// grouped queries: dk, dv of a (b, kv head) = sum over the query heads of its group (fp32 sum of the bf16 rows).
// HBM-bound: reads 2 x (q_per_kv x N x 64) bf16 and writes 2 x (N x 64) per (b, kv head).
// =============================================================================================
__global__ void __launch_bounds__(256) group_sum_kernel(const bf16* __restrict__ dk_h, const bf16* __restrict__ dv_h,
                                                        bf16* __restrict__ dk, bf16* __restrict__ dv, int q_per_kv, int N,
                                                        size_t n8) {
  const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;       // chunk of 8 features of one (b, kv head, token)
  if (i >= n8) return;
  const size_t per_head = (size_t)N * 8;
  const size_t kvh = i / per_head, rem = i % per_head;
#pragma unroll
  for (int which = 0; which < 2; ++which) {
    const uint4* src = reinterpret_cast<const uint4*>(which ? dv_h : dk_h) + kvh * q_per_kv * per_head + rem;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int h = 0; h < q_per_kv; ++h) {
      const uint4 x = __ldg(src + (size_t)h * per_head);
      const uint32_t w[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[2 * j] += __uint_as_float(w[j] << 16), acc[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]), o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]), o.w = pack_bf16x2(acc[6], acc[7]);
    reinterpret_cast<uint4*>(which ? dv : dk)[i] = o;
  }
}

}  // namespace mom

// =============================================================================================
// host orchestration
// =============================================================================================
using namespace mom;

struct MomLayout {
  int BH, BHkv, N, Npad, L, nseg, nrec;   // BHkv: (b, kv head) pairs = owners of k, v and of the forward states
  size_t off_xt, off_xt2, off_state, off_state2, off_part, off_part2, off_part3, off_u, off_qd, off_dlt;
  size_t off_dkh, off_dvh;                // grouped queries: per-query-head dk, dv before the sum over the group
  size_t total;
  MomLayout kv() const {                  // the same layout seen by the kernels that run once per (b, kv head)
    MomLayout k = *this;
    k.BH = BHkv;
    return k;
  }
};

static MomLayout mom_layout(int B, int H, int N, int L, int causal, bool bwd, int Hkv = 0) {
  MomLayout m{};
  m.BH = B * H, m.BHkv = B * (Hkv > 0 ? Hkv : H), m.N = N, m.Npad = (int)align_up(N, 64), m.L = L;
  m.nseg = (N + L - 1) / L;
  m.nrec = causal ? m.nseg : 1;
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off += align_up(bytes, 1024);
    return o;
  };
  const size_t tens = (size_t)m.BH * N * 64;
  const size_t part_floats = (size_t)m.BH * ((N + 127) / 128) * PART_TILE_FLOATS;   // whole 128-token tiles
  m.off_xt = take((size_t)m.BHkv * 64 * m.Npad * 2);
  m.off_state = take((size_t)m.BHkv * m.nrec * REC_BYTES);
  m.off_part = take(part_floats * 4);
  if (bwd) {
    m.off_xt2 = take((size_t)m.BH * 64 * m.Npad * 2);
    m.off_state2 = take((size_t)m.BH * m.nrec * REC_BYTES);
    m.off_part2 = take(part_floats * 4);
    m.off_part3 = take(part_floats * 4);
    m.off_u = take(tens * 2);
    m.off_qd = take(tens * 2);
    m.off_dlt = take((size_t)m.BH * m.Npad * 4);
    if (m.BHkv != m.BH) m.off_dkh = take(tens * 2), m.off_dvh = take(tens * 2);
  }
  m.total = off;
  return m;
}

// Segment length of the chunked scan.  Cost per token: intra-segment tiles ~ L, moment kernels constant
// (and absent with a single segment).  Measured at D = 64: the moment kernels cost about as much as the
// intra tiles of a 2048-token segment, so sequences up to 2048 tokens run as ONE segment (pure
// tensor-core quadratic evaluation, no states); longer ones use 512-token segments, doubled only while
// the boundary states of one direction would exceed 24 GiB (B*H = 32, N = 1M: L = 1024).
int moments_seg_len(int B, int H, int N, int p) {
  // (p = 1: the moment kernels handle one slab instead of 65, so segments stay at 512 tokens)
  if (N <= (p == 2 ? 2048 : 512)) return (N + 511) / 512 * 512;
  int L = 512;
  const double budget = 24.0 * 1024 * 1024 * 1024;
  while ((double)B * H * ((N + L - 1) / L) * (double)REC_BYTES > budget) L *= 2;
  return L;
}

size_t moments_workspace_bytes(int B, int H, int N, int p, int causal, bool bwd, int Hkv) {
  return mom_layout(B, H, N, moments_seg_len(B, H, N, p), causal, bwd, Hkv).total;
}

// developer switch: FASTMAX_P1=chunked keeps first-order causal problems on the chunked scan (round-1 path)
static bool p1_scan_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FASTMAX_P1");
    v = (e && !strcmp(e, "chunked")) ? 0 : 1;
  }
  return v != 0;
}

// developer switch: FASTMAX_INTRA=v1 keeps the first-generation intra-segment kernels of fm_intra.cu for p = 2 causal
static bool seg_kernels_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FASTMAX_INTRA");
    v = (e && !strcmp(e, "v1")) ? 0 : 1;
  }
  return v != 0;
}

static int launch_transpose(const Problem& pb, const void* in, void* out, int Npad) {
  dim3 grid(Npad / 64, pb.B * pb.H);
  FM_LAUNCH("mom_transpose", pb.stream,
            transpose_kernel<<<grid, 256, 0, pb.stream>>>((const bf16*)in, (bf16*)out, pb.N, Npad));
  return FASTMAX_OK;
}

// developer switch: FASTMAX_BUILD=full accumulates all 64x64 feature pairs (round-1 layout, 9 CTAs per (b,h));
// default: symmetric-packed tiles (6 CTAs)
static int build_packed() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FASTMAX_BUILD");
    v = (e && !strcmp(e, "full")) ? 0 : 1;
  }
  return v;
}

static int launch_build(const Problem& pb, const MomLayout& m, const void* xt, const void* w, const void* w2,
                        uint8_t* state, int reverse, const void* init = nullptr) {
  CUtensorMap tmXT, tmW, tmW2, tmS, tmS16;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tmS, state, 64, NSLAB * 64, (uint64_t)m.BH * m.nrec, 128, REC_BYTES, 64, 64))) return rc;
  if ((rc = make_tmap_state_4d(&tmS16, state, NSLAB, (uint64_t)m.BH * m.nrec, REC_BYTES, 16, 16))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmXT, xt, 64, m.Npad, m.BH, 128, (uint64_t)m.Npad * 128, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmW, w, 64, m.N, m.BH, 128, (uint64_t)m.N * 128, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmW2, w2, 64, m.N, m.BH, 128, (uint64_t)m.N * 128, 64, 64))) return rc;
  BuildParams P{};
  P.N = m.N, P.nblk = (m.N + 63) / 64, P.Lb = m.L / 64, P.nseg = m.nseg, P.reverse = reverse, P.causal = pb.causal;
  P.p = pb.p, P.state = state, P.init = (const uint8_t*)init;
  P.packed = pb.p == 2 && build_packed();
  const int nctas = pb.p == 2 ? (P.packed ? B_CTAS_P2_PACKED : B_CTAS_P2) : 1;
  P.special_cx = nctas - 1;
  for (int st = 0; st < B_STAGES; ++st)
    for (int kk = 0; kk < 4; ++kk) {
      const uint32_t base = sm100::FM_SMEM_BASE + st * B_STAGE_BYTES;
      P.dW[st][kk] = sm100::desc_mnmajor_sw128_const(base + 8192, 8192) + kk * 128;
      P.dW2[st][kk] = sm100::desc_mnmajor_sw128_const(base + 16384, 8192) + kk * 128;
    }
  if (init && pb.causal) {
    // the boundary state of the first (reverse scan: last) local segment is the incoming state itself
    uint8_t* dst = state + (size_t)(reverse ? m.nseg - 1 : 0) * REC_BYTES;
    FM_CUDA(cudaMemcpy2DAsync(dst, (size_t)m.nseg * REC_BYTES, init, REC_BYTES, REC_BYTES, m.BH,
                              cudaMemcpyDeviceToDevice, pb.stream));
  }
  if (!pb.causal) P.nproc = P.nblk;
  else if (reverse) P.nproc = P.nblk - P.Lb;
  else P.nproc = (m.nseg - 1) * P.Lb;
  if (P.nproc <= 0) return FASTMAX_OK;   // single segment: no boundary state needed
  FM_CUDA(cudaFuncSetAttribute(build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
  dim3 grid(nctas, m.BH);
  FM_LAUNCH(reverse ? "mom_build_rev" : "mom_build", pb.stream,
            build_kernel<<<grid, B_THREADS, B_SMEM, pb.stream>>>(tmXT, tmW, tmW2, tmS, tmS16, P));
  return FASTMAX_OK;
}

// developer switch: FASTMAX_APPLY=gen selects the generated-operand kernel (fm_apply.cu: correct, still slower
// than the two-stage kernel below); default: two-stage
static int apply_variant() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("FASTMAX_APPLY");
    v = (e && !strcmp(e, "gen")) ? 0 : -1;
  }
  return v;
}

static int launch_apply(const Problem& pb, const MomLayout& m, const char* name, ApplyParams P, const void* Y) {
  if (apply_variant() >= 0 && pb.p == 2 && m.BHkv == m.BH) {
    ApplyDesc d{};
    d.L = m.L, d.nseg = m.nseg, d.state_reverse = P.state_reverse, d.b_kmajor = P.b_kmajor, d.edge_state = P.edge_state;
    d.vec_sel = P.vec_sel, d.coef_reg = P.coef_reg, d.coef64 = P.coef64, d.out_scale = P.out_scale;
    d.X = P.X, d.Y = Y, d.partial = P.partial, d.gden = P.gden, d.out = P.out, d.state = P.state;
    return launch_apply_gen(pb, m.BH, m.nrec, name, d, apply_variant());
  }
  CUtensorMap tmS, tmY, tmOut;
  int rc;
  // grouped queries: forward states (o, dq) and the k, v operands (dv, dk) live once per (b, kv head)
  P.H = pb.H, P.Hkv = pb.kv_heads();
  P.state_kv = P.mode == APPLY_FWD || P.mode == APPLY_DQ, P.xy_kv = !P.state_kv;
  const uint64_t bh_st = P.state_kv ? m.BHkv : m.BH, bh_xy = P.xy_kv ? m.BHkv : m.BH;
  if ((rc = make_tmap_bf16_3d(&tmS, P.state, 64, NSLAB * 64, bh_st * m.nrec, 128, REC_BYTES, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmY, Y, 64, m.N, bh_xy, 128, (uint64_t)m.N * 128, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&tmOut, P.out, 64, m.N, m.BH, 128, (uint64_t)m.N * 128, 64, 64))) return rc;
  P.N = m.N, P.L = m.L, P.nseg = m.nseg, P.causal = pb.causal, P.p = pb.p;
  FM_CUDA(cudaFuncSetAttribute(apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM));
  dim3 grid((m.N + A_TOK - 1) / A_TOK, m.BH);
  ApplyDescs DS{};
  for (int kk = 0; kk < 4; ++kk) {
    for (int t = 0; t < 4; ++t) DS.aK[t][kk] = sm100::desc_kmajor_sw128_const(sm100::FM_SMEM_BASE + t * 16384) + kk * 2;
    for (int st = 0; st < A_STAGES; ++st) {
      const uint32_t sb = sm100::FM_SMEM_BASE + A_RES_BYTES + st * A_STAGE_BYTES;
      DS.bD[st][kk] = P.b_kmajor ? sm100::desc_kmajor_sw128_const(sb) + kk * 2
                                 : sm100::desc_mnmajor_sw128_const(sb, 8192) + kk * 128;
    }
  }
  FM_LAUNCH(name, pb.stream, apply_kernel<<<grid, A_THREADS, A_SMEM, pb.stream>>>(tmS, tmY, tmOut, DS, P));
  return FASTMAX_OK;
}


int moments_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g, void* ws) {
  const MomLayout m = mom_layout(pb.B, pb.H, pb.N, moments_seg_len(pb.B, pb.H, pb.N, pb.p), pb.causal, false, pb.Hkv);
  Problem pk = pb;                 // the problem as the per-(b, kv head) kernels see it
  pk.H = pb.kv_heads(), pk.Hkv = 0;
  uint8_t* w8 = (uint8_t*)ws;
  int rc;
  const float a = pb.inv_c;
  uint8_t* state = pb.saved_states ? (uint8_t*)pb.saved_states : w8 + m.off_state;
  if (p1_scan_enabled() && p1scan_usable(pb)) return p1scan_fwd(pb, q, k, v, o, g, w8 + m.off_state);
  if (p1_scan_enabled() && p1full_usable(pb)) return p1full_fwd(pb, q, k, v, o, g, state);
  float* part = (float*)(w8 + m.off_part);
  // 1. K^T and the boundary states (forward direction)
  if ((rc = launch_transpose(pk, k, w8 + m.off_xt, m.Npad))) return rc;
  if ((rc = launch_build(pk, m.kv(), w8 + m.off_xt, v, k, state, 0, pb.sp_init_fwd))) return rc;
  // 2. intra-segment part (causal only) and the normaliser
  // intra-segment scores + the normaliser (its moment part rides along as one more score MMA); non-causal:
  // only that MMA runs (no intra-segment part, no partial)
  if (pb.p == 1)   // first order: the intra kernel applies the (single-slab) state term itself and writes o
    return intra_fwd(pb, m.L, m.Npad, q, k, v, nullptr, g, state, o);
  if (seg_kernels_enabled() && seg_fwd_usable(pb, m.L)) {
    if ((rc = seg_fwd(pb, m.L, q, k, v, part, g, state))) return rc;
  } else if ((rc = intra_fwd(pb, m.L, m.Npad, q, k, v, pb.causal ? part : nullptr, g, state))) {
    return rc;
  }
  // 3. inter-segment part + normalisation
  ApplyParams ap{};
  ap.mode = APPLY_FWD, ap.state_reverse = 0, ap.b_kmajor = 0, ap.edge_state = pb.sp_init_fwd != nullptr;
  ap.coef_reg = 0.5f * a * a, ap.coef64 = a, ap.out_scale = 1.f, ap.vec_sel = 0;
  ap.X = (const bf16*)q, ap.partial = pb.causal ? part : nullptr, ap.gden = g;
  ap.out = (bf16*)o, ap.state = state;
  return launch_apply(pb, m, "mom_apply_fwd", ap, q);
}

int moments_bwd(const Problem& pb, const void* q, const void* k, const void* v, const void* o, const float* g,
                const void* d_o, void* dq, void* dk, void* dv, void* ws) {
  const MomLayout m = mom_layout(pb.B, pb.H, pb.N, moments_seg_len(pb.B, pb.H, pb.N, pb.p), pb.causal, true, pb.Hkv);
  Problem pk = pb;                 // the problem as the per-(b, kv head) kernels see it
  pk.H = pb.kv_heads(), pk.Hkv = 0;
  const bool grouped = m.BHkv != m.BH;
  uint8_t* w8 = (uint8_t*)ws;
  int rc;
  const float a = pb.inv_c;
  uint8_t* S = pb.saved_states ? (uint8_t*)pb.saved_states : w8 + m.off_state;
  uint8_t* R = w8 + m.off_state2;
  float* dq_part = (float*)(w8 + m.off_part);
  float* dk_part = (float*)(w8 + m.off_part2);
  float* dv_part = (float*)(w8 + m.off_part3);
  bf16* u = (bf16*)(w8 + m.off_u);
  bf16* qd = (bf16*)(w8 + m.off_qd);
  float* dlt = (float*)(w8 + m.off_dlt);
  const size_t rows = (size_t)m.BH * m.N;
  // 1. u = do/g, dlt = (do.o)/g, qd = -dlt*q
  FM_CUDA(cudaMemsetAsync(dlt, 0, (size_t)m.BH * m.Npad * 4, pb.stream));
  if (p1_scan_enabled() && p1scan_usable(pb)) {
    FM_LAUNCH("mom_prep_bwd", pb.stream,
              prep_bwd_kernel<false><<<(unsigned)((rows + PREP_ROWS - 1) / PREP_ROWS), 256, 0, pb.stream>>>(
                  (const bf16*)q, (const bf16*)o, (const bf16*)d_o, g, u, dlt, qd, rows, m.N, m.Npad));
    return p1scan_bwd(pb, m.Npad, q, k, v, u, dlt, dq, dk, dv, w8 + m.off_state, R);
  }
  if (p1_scan_enabled() && p1full_usable(pb)) {
    FM_LAUNCH("mom_prep_bwd", pb.stream,
              prep_bwd_kernel<false><<<(unsigned)((rows + PREP_ROWS - 1) / PREP_ROWS), 256, 0, pb.stream>>>(
                  (const bf16*)q, (const bf16*)o, (const bf16*)d_o, g, u, dlt, qd, rows, m.N, m.Npad));
    return p1full_bwd(pb, m.Npad, q, k, v, u, dlt, dq, dk, dv, S, pb.saved_states != nullptr, R);
  }
  FM_LAUNCH("mom_prep_bwd", pb.stream,
            prep_bwd_kernel<true><<<(unsigned)((rows + PREP_ROWS - 1) / PREP_ROWS), 256, 0, pb.stream>>>(
                (const bf16*)q, (const bf16*)o, (const bf16*)d_o, g, u, dlt, qd, rows, m.N, m.Npad));
  // 2. forward-direction states of (k, v) (recomputed, not saved) and reverse-direction states of (q, u)
  if (!pb.saved_states) {
    if ((rc = launch_transpose(pk, k, w8 + m.off_xt, m.Npad))) return rc;
    if ((rc = launch_build(pk, m.kv(), w8 + m.off_xt, v, k, S, 0, pb.sp_init_fwd))) return rc;
  }
  if ((rc = launch_transpose(pb, q, w8 + m.off_xt2, m.Npad))) return rc;
  if ((rc = launch_build(pb, m, w8 + m.off_xt2, u, qd, R, 1, pb.sp_init_rev))) return rc;
  // 3. intra-segment partials
  // (non-causal: the same kernels only apply the normaliser column of the single state: dq_part, dk_part)
  if (pb.p == 1) {   // first order: no apply kernels, no fp32 partials
    if ((rc = intra_dq(pb, m.L, m.Npad, q, u, k, v, dlt, nullptr, S, dq))) return rc;
    return intra_dkdv(pb, m.L, m.Npad, k, v, q, u, dlt, nullptr, nullptr, R, dk, dv);
  }
  if (seg_kernels_enabled() && seg_dq_usable(pb, m.L)) {
    if ((rc = seg_dq(pb, m.L, m.Npad, q, u, k, v, dlt, dq_part, S))) return rc;
  } else if ((rc = intra_dq(pb, m.L, m.Npad, q, u, k, v, dlt, dq_part, S))) {
    return rc;
  }
  if (seg_kernels_enabled() && seg_dkdv_usable(pb, m.L)) {
    if ((rc = seg_dkdv(pb, m.L, m.Npad, k, v, q, u, dlt, dk_part, dv_part, R))) return rc;
  } else if ((rc = intra_dkdv(pb, m.L, m.Npad, k, v, q, u, dlt, dk_part, pb.causal ? dv_part : nullptr, R))) {
    return rc;
  }
  // 4. inter-segment parts
  ApplyParams ap{};
  ap.mode = APPLY_DQ, ap.state_reverse = 0, ap.b_kmajor = 1, ap.edge_state = pb.sp_init_fwd != nullptr;
  ap.coef_reg = a, ap.coef64 = 1.f, ap.out_scale = a, ap.vec_sel = -1;
  ap.X = (const bf16*)q, ap.partial = dq_part, ap.gden = nullptr, ap.out = (bf16*)dq, ap.state = S;
  if ((rc = launch_apply(pb, m, "mom_apply_dq", ap, u))) return rc;

  ap = ApplyParams{};
  ap.mode = APPLY_DV, ap.state_reverse = 1, ap.b_kmajor = 0, ap.edge_state = pb.sp_init_rev != nullptr;
  ap.coef_reg = 0.5f * a * a, ap.coef64 = a, ap.out_scale = 1.f, ap.vec_sel = 0;
  // grouped queries: dk, dv of every query head first (its own reverse states against the shared k, v), then the
  // sum over the group -- what autograd does with the expand() of model.py:408-416, without the expanded copies
  bf16* dk_h = grouped ? (bf16*)(w8 + m.off_dkh) : (bf16*)dk;
  bf16* dv_h = grouped ? (bf16*)(w8 + m.off_dvh) : (bf16*)dv;
  ap.X = (const bf16*)k, ap.partial = pb.causal ? dv_part : nullptr, ap.gden = nullptr;
  ap.out = dv_h, ap.state = R;
  if ((rc = launch_apply(pb, m, "mom_apply_dv", ap, k))) return rc;

  ap = ApplyParams{};
  // the normaliser column of the reverse state (RN1 + a RN2 k) is already inside dk_part
  ap.mode = APPLY_DK, ap.state_reverse = 1, ap.b_kmajor = 1, ap.edge_state = pb.sp_init_rev != nullptr;
  ap.coef_reg = a, ap.coef64 = 1.f, ap.out_scale = a, ap.vec_sel = -1;
  ap.X = (const bf16*)k, ap.partial = dk_part, ap.gden = nullptr;
  ap.out = dk_h, ap.state = R;
  if ((rc = launch_apply(pb, m, "mom_apply_dk", ap, v))) return rc;
  if (grouped) {
    const size_t n8 = (size_t)m.BHkv * m.N * 8;       // 16-byte chunks of one output tensor
    FM_LAUNCH("mom_group_sum", pb.stream,
              group_sum_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, pb.stream>>>(dk_h, dv_h, (bf16*)dk, (bf16*)dv,
                                                                                   pb.H / pb.kv_heads(), m.N, n8));
  }
  return FASTMAX_OK;
}

// =============================================================================================
// sequence parallelism: per-rank total states and their exclusive prefix / suffix over the ranks
// =============================================================================================
size_t moments_state_bytes(int B, int H) { return (size_t)B * H * REC_BYTES; }

size_t moments_saved_state_bytes(int B, int H, int N, int p, int causal, int Hkv) {
  const MomLayout m = mom_layout(B, H, N, moments_seg_len(B, H, N, p), causal, false, Hkv);
  return (size_t)m.BHkv * m.nrec * REC_BYTES;
}

size_t moments_local_state_workspace_bytes(int B, int H, int N) {
  return align_up((size_t)B * H * 64 * align_up(N, 64) * 2, 1024);
}

// total forward-direction state of the local tokens: x = k, w = v, w2 = k  (one record per (b,h))
int moments_local_state(const Problem& pb_in, const void* x, const void* w, const void* w2, void* state_out, void* ws) {
  Problem pb = pb_in;
  pb.causal = 0;
  const MomLayout m = mom_layout(pb.B, pb.H, pb.N, moments_seg_len(pb.B, pb.H, pb.N, pb.p), 0, false);
  int rc;
  if ((rc = launch_transpose(pb, x, ws, m.Npad))) return rc;
  return launch_build(pb, m, ws, w, w2, (uint8_t*)state_out, 0);
}

size_t moments_local_state_rev_workspace_bytes(int B, int H, int N) {
  const size_t tens = (size_t)B * H * N * 64, npad = align_up(N, 64);
  return align_up((size_t)B * H * 64 * npad * 2, 1024) + 2 * align_up(tens * 2, 1024) + align_up((size_t)B * H * npad * 4, 1024);
}

// total reverse-direction state of the local tokens: x = q, w = do/g, w2 = -dlt*q
int moments_local_state_rev(const Problem& pb_in, const void* q, const void* o, const float* g, const void* d_o,
                            void* state_out, void* ws) {
  Problem pb = pb_in;
  pb.causal = 0;
  const MomLayout m = mom_layout(pb.B, pb.H, pb.N, moments_seg_len(pb.B, pb.H, pb.N, pb.p), 0, false);
  uint8_t* w8 = (uint8_t*)ws;
  const size_t tens = (size_t)m.BH * m.N * 64;
  size_t off = align_up((size_t)m.BH * 64 * m.Npad * 2, 1024);
  bf16* u = (bf16*)(w8 + off);
  off += align_up(tens * 2, 1024);
  bf16* qd = (bf16*)(w8 + off);
  off += align_up(tens * 2, 1024);
  float* dlt = (float*)(w8 + off);
  const size_t rows = (size_t)m.BH * m.N;
  FM_CUDA(cudaMemsetAsync(dlt, 0, (size_t)m.BH * m.Npad * 4, pb.stream));
  FM_LAUNCH("mom_prep_bwd", pb.stream,
            prep_bwd_kernel<true><<<(unsigned)((rows + PREP_ROWS - 1) / PREP_ROWS), 256, 0, pb.stream>>>(
                (const bf16*)q, (const bf16*)o, (const bf16*)d_o, g, u, dlt, qd, rows, m.N, m.Npad));
  int rc;
  if ((rc = launch_transpose(pb, q, w8, m.Npad))) return rc;
  return launch_build(pb, m, w8, u, qd, (uint8_t*)state_out, 1);
}

// out[bh] = sum over ranks r' < rank (dir 0) or r' > rank (dir 1) of gathered[r'][bh]; bf16 slabs are
// summed in fp32.  HBM-bound, G records in, one out.
__global__ void __launch_bounds__(256) reduce_states_kernel(const uint8_t* __restrict__ gathered, int G, int rank,
                                                            int dir, uint8_t* __restrict__ out, size_t bh_count) {
  const size_t words_per_rec = REC_BYTES / 4;
  const size_t total = bh_count * words_per_rec;
  const int r0 = dir == 0 ? 0 : rank + 1, r1 = dir == 0 ? rank : G;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const size_t bh = idx / words_per_rec, w = idx % words_per_rec;
    const bool is_vec = w * 4 >= VEC_OFF;
    float a0 = 0.f, a1 = 0.f;
    for (int r = r0; r < r1; ++r) {
      const uint32_t x = reinterpret_cast<const uint32_t*>(gathered + ((size_t)r * bh_count + bh) * REC_BYTES)[w];
      if (is_vec) a0 += __uint_as_float(x);
      else a0 += __uint_as_float(x << 16), a1 += __uint_as_float(x & 0xffff0000u);
    }
    reinterpret_cast<uint32_t*>(out + bh * REC_BYTES)[w] = is_vec ? __float_as_uint(a0) : sm100::pack_bf16x2(a0, a1);
  }
}

int moments_reduce_states(const void* gathered, int G, int rank, int dir, void* out, int B, int H, cudaStream_t stream) {
  const size_t bh = (size_t)B * H;
  FM_LAUNCH("mom_reduce_states", stream,
            reduce_states_kernel<<<148 * 4, 256, 0, stream>>>((const uint8_t*)gathered, G, rank, dir, (uint8_t*)out, bh));
  return FASTMAX_OK;
}

// executed tensor-core flops of the plan (for the utilisation report; never the algorithmic count)
void moments_flops(int B, int H, int N, int p, int causal, double* fwd, double* bwd, double* rebuild) {
  const int L = moments_seg_len(B, H, N, p);
  const MomLayout m = mom_layout(B, H, N, L, causal, false);
  const double mma = 2.0 * 128 * 64 * 16;
  const int nblk = (N + 63) / 64, Lb = L / 64;
  const double acc = p == 2 ? (build_packed() ? 22 : 34) : 2;   // 128x64 accumulators per (b,h): second-order tiles + 2 special
  const double nproc_f = causal ? (double)(m.nseg - 1) * Lb : nblk;
  const double nproc_r = causal ? (double)(nblk - Lb) : nblk;
  const double build_f = m.BH * (nproc_f > 0 ? nproc_f : 0) * acc * 4 * mma;
  const double build_r = m.BH * (nproc_r > 0 ? nproc_r : 0) * acc * 4 * mma;
  const double ctas = (double)((N + APPLY_TOK - 1) / APPLY_TOK);
  const double with_state = causal ? (ctas - (double)L / APPLY_TOK > 0 ? ctas - (double)L / APPLY_TOK : 0) : ctas;
  // k-steps of one 128-token tile against the state: 65 slabs x 4 (the round-1 kernel, dq, dk); the symmetric
  // contractions (o, dv) of the generated-operand kernel skip the blocks below the diagonal: 160 + 4
  const bool gen = apply_variant() >= 0;
  const double ks_full = p == 2 ? 65 * 4 : 4, ks_sym = p == 2 ? (gen ? 164 : 260) : 4;
  const double apply_full = m.BH * with_state * (APPLY_TOK / 128) * ks_full * mma;
  const double apply_sym = m.BH * with_state * (APPLY_TOK / 128) * ks_sym * mma;
  // intra-segment tiles (causal only): 128x128x64 GEMMs on the lower-triangular tile pairs of each
  // segment; forward 2 per pair (scores, P.V), dq 3, dkdv 4
  double intra_pair = 0;
  if (causal) {
    const double t = (double)((N + 127) / 128), ts = (double)L / 128;
    const double full_segs = (double)(N / L), rem = t - full_segs * ts;
    intra_pair = m.BH * (full_segs * ts * (ts + 1) / 2 + rem * (rem + 1) / 2) * 2.0 * 128 * 128 * 64;
  }
  *fwd = build_f + apply_sym + 2 * intra_pair;
  *bwd = build_f + build_r + apply_sym + 2 * apply_full + 7 * intra_pair;
  *rebuild = build_f;   // skipped by backward when the forward states were kept
}

}  // namespace fm
