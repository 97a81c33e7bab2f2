// fm_p1scan.cu -- first-order (p = 1) causal FastMax as ONE fused scan per (b,h) and tensor.
//
// For p = 1 the moment state of a (b,h) is a single 64 x 65 matrix, small enough to stay in tensor memory for
// the whole sequence, and the work per token is so small (58 kflop against 1.5 KB of input/output, SURVEY.md
// 7.3-5) that the right bound is HBM bandwidth.  The chunked path of fm_moments.cu / fm_intra.cu spends most of
// its p = 1 time in 512-token quadratic tiles (17 % of the HBM roof); here one CTA walks a (b,h) sequence in
// 128-token tiles and does everything a tile needs in place -- no boundary-state records, no fp32 partials, no
// second pass:
//
//     acc   = A . State_bf16                       inter-tile part (state of all earlier / later tiles)
//     S     = A . B1^T  (two 64-column halves)     tile-local scores -> polynomial + causal mask on CUDA cores
//     acc  += P . [B2 | extra]                     tile-local part (P back in tensor memory as bf16, TS MMA)
//     State += [X | 1]^T . [W | extra]             fp32 accumulator in tensor memory; row 64 = column sums,
//                                                  column 64 = the `extra` column (ones or delta)
//     out   = epilogue(acc, fp32 vectors of the state) ; State -> bf16 copy in shared memory for the next tile
//
// Four uses of the one kernel (reference: compute_F_masked / compute_g_masked, fastmax.py:219-322, and
// gradient_o_{q,k,v}_masked, fastmax.py:433-691, for p = 1; u = do/g, dlt = (do.o)/g from the prep kernel):
//
//   mode  scan     A    B1   B2   state (X, W | extra)   poly on S (rows = A tokens, cols = B tokens)   out
//   FWD   forward  q    k    v    (k, v | 1)             1 + a s        , col <= row                     o = num/den, g = den
//   DQ    forward  u    v    k    (k, v | 1)             s - dlt_row    , col <= row                     dq = a (acc - dlt_row n1)
//   DV    reverse  k    q    u    (q, u | dlt)           1 + a s        , col >= row                     dv = acc + R0
//   DK    reverse  v    u    q    (q, u | dlt)           s - dlt_col    , col >= row                     dk = a (acc - Rdlt)
//
// with n1 = sum k (column 64 of the forward state), R0 = sum u (row 64 of the reverse state), Rdlt = sum dlt q
// (its column 64).  Two CTAs per SM (256 tensor-memory columns and ~100 KB of shared memory each), so a batch of
// 256 (b,h) sequences is resident at once and the per-tile latency chain (~4 MMA round trips) hides behind the
// other CTA and the HBM stream.  bf16, D = 64, causal, p = 1.
#include "fm_common.cuh"
#include "fm_sm100.cuh"
#include "fm_moments.cuh"

#ifndef FM_P1_PF
#define FM_P1_PF 1      // L2 prefetch distance in tiles (0: 2.34, 1: 2.17, 2: 2.20, 3: 2.32, 5: 2.52 ms per step at the headline shape)
#endif
#ifdef FM_TRACE
// developer event trace of CTA 0 (thread 0): clock64 at the steps of every tile; [mode][event][tile]
static __device__ long long fm_trace_p1[4][14][160];
extern "C" int fastmax_debug_trace_fetch_p1(long long* out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, fm_trace_p1, sizeof(fm_trace_p1)) == cudaSuccess ? 0 : -4;
}
#define FM_TRP(ev) do { if (blockIdx.x == 0 && threadIdx.x == 0 && it < 160) fm_trace_p1[MODE][ev][it] = clock64(); } while (0)
#define FM_TRX(ev) do { if (blockIdx.x == 0 && it < 160) fm_trace_p1[MODE][ev][it] = clock64(); } while (0)
#else
#define FM_TRP(ev) ((void)0)
#define FM_TRX(ev) ((void)0)
#endif

namespace fm {
namespace p1 {
using namespace sm100;
typedef __nv_bfloat16 bf16;

enum Mode { FWD = 0, DQ = 1, DV = 2, DK = 3 };

// shared memory (offsets from the 1024-aligned base)
constexpr int OFF_A = 0;                 // resident tile A   [128 tokens][64] bf16, 128B-swizzled rows
constexpr int OFF_B1 = 16384;
constexpr int OFF_B2 = 32768;
// extra panel [128 tokens][64]: element [tok][0] = 1, element [tok][1] = dlt_tok (reverse modes), rest 0.  As the
// second M panel of the update's A operand it adds row 64 (column sums) to the state, as the second N panel of its
// B operand column 64 (sum of X: n1) and column 65 (sum of dlt X: Rdlt); FWD's P.[V | 1] reads it as column 64.
constexpr int OFF_ONES = 49152;
constexpr int OFF_STG = 65536;           // bf16 output tile, 128B-swizzled rows (bulk tensor store source)
constexpr int OFF_SB = 81920;            // state as bf16: panel 0 = [64 y][64 a], panel 1 = [64 y][64] (column 64 in element 0)
constexpr int OFF_MISC = 98304;          // barriers, vectors
constexpr int SMEM = OFF_MISC + 2048 + 1024 /*align slack*/;
constexpr int THREADS = 192;             // 4 compute warps + TMA producer warp + auxiliary MMA issuer warp
// tensor memory (256 columns per CTA)
// scores: half 0 in columns [0, 64), its bf16 P in [0, 32); half 1 in [32, 96) -- written while the P of half 0 is
// still being consumed, which only touches [0, 32) -- and its P in [64, 96), so that columns [0, 64) are free for the
// NEXT tile's first score group as soon as the second polynomial has read its scores
constexpr int C_S = 0;
constexpr int C_ACC = 96;                // 128 x 80 accumulator (FWD uses column 64 for the normaliser)
constexpr int C_ST = 176;                // 128 x 80 state accumulator (rows 0..63: S1 | extra column; row 64: column sums)

struct Params {
  int N, Npad, ntile;
  // a sequence may be cut into `nchunk` ranges of `tiles_per_chunk` tiles, one CTA each (few (b,h), long sequences);
  // a chunk then starts from `init`, the state of everything before (reverse modes: after) it
  int nchunk, tiles_per_chunk;
  const float* init;         // (BH, nchunk, 64*66 + 80) fp32 in the layout of fm_p1full.cu's partial states, or null
  float a;
  const float* dlt;          // (BH, Npad) fp32, zero padded (DQ, DV, DK)
  float* g;                  // (BH, N) (FWD)
  const bf16 *pA, *pB1, *pB2;   // the three input tensors (BH, N, 64), for the L2 prefetch of the tile after next
  // host-built descriptors (fixed shared-memory layout)
  unsigned long long dA[4];           // A tile, K-major, k-step over the 64 features
  unsigned long long dB1[2][4];       // B1 rows [64h, 64h+64), K-major
  unsigned long long dB2[2][4];       // [B2 | extra panel] rows [64h + 16kk, +16), MN-major (N up to 80)
  unsigned long long dSmn[4];         // state bf16 read MN-major (contract y; N up to 80)
  unsigned long long dSk[4];          // state bf16 read K-major (contract a; N = 64)
  unsigned long long dX[8];           // [X | ones panel] MN-major A operand of the state update, k-step = 16 tokens
  unsigned long long dW[8];           // [W | extra panel] MN-major B operand of the state update
};

__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2)
scan_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB1,
            const __grid_constant__ CUtensorMap tmB2, const __grid_constant__ CUtensorMap tmOut,
            const __grid_constant__ Params P) {
  constexpr bool REV = MODE == DV || MODE == DK;
  constexpr bool SCORE_POLY = MODE == FWD || MODE == DV;       // P = 1 + a s   (else E = s - dlt)
  constexpr int NACC = MODE == FWD ? 80 : 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_MISC);
  uint64_t* full_s = bars;           // A and B1 of a tile landed (what the score MMAs need)
  uint64_t* full_v = bars + 7;       // B2 landed
  // MMA groups complete: [0] scores of half 0 (issued at the end of the PREVIOUS tile, so they run under its
  // epilogue), [1] scores of half 1, [2] P.B2' of half 0 + state update, [3] P.B2' of half 1 (and the inter-tile
  // MMA).  One barrier per event, each completing once per tile and waited on by every compute warp, so that no
  // waiter ever skips a phase (a skipped phase is indistinguishable from the one before it).
  uint64_t* mma = bars + 1;          // [0..3] as above, [4] the inter-tile MMA (the bf16 state may then be rewritten)
  uint64_t* pv0_done = bars + 11;    // P.B2' of half 0 complete (its P, columns [0, 32), may be overwritten)
  uint64_t* p1_done = bars + 12;     // every compute thread has read the scores of half 1 (columns [32, 96))
  uint64_t* b1_free = bars + 10;     // B1 may be overwritten (scores and state update complete)
  uint64_t* a_free = bars + 6;       // A may be overwritten (both score groups and the inter-tile MMA complete)
  uint64_t* b2_free = bars + 8;      // B2 may be overwritten (all MMAs of the tile complete)
  uint64_t* tile_go = bars + 9;      // compute warps have begun tile `it`: the previous state copy has been taken and
                                     // the dlt panel of this tile is written (what the auxiliary issuer waits for)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);
  // fp32 vectors of the state, two sets (tile parity): the epilogue of tile `it` reads set it & 1 (the state BEFORE the
  // tile) while the refresh in the middle of the tile writes set (it + 1) & 1
  float* vrow_base = reinterpret_cast<float*>(smem + OFF_MISC + 128);   // [2][80] row 64 of the state: column sums (S0 | count, R0)
  float* vcol_base = vrow_base + 160;                                   // [2][64] column 64 / 65 of the state: n1 / Rdlt
  float* dcol = vcol_base + 128;                                        // [128] dlt of the tile's tokens (DK: per column)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.x / P.nchunk, chunk = blockIdx.x % P.nchunk;
  const int t_lo = chunk * P.tiles_per_chunk, t_hi = min(P.ntile, t_lo + P.tiles_per_chunk);   // this CTA's tiles
  const int ntile = t_hi - t_lo;

  if (warp == 4) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  if (tid == 0) {
    mbar_init(full_s, 1), mbar_init(full_v, 1), mbar_init(b1_free, 1), mbar_init(b2_free, 1), mbar_init(a_free, 1);
    mbar_init(tile_go, 1), mbar_init(pv0_done, 1), mbar_init(p1_done, 1);
    for (int i = 0; i < 5; ++i) mbar_init(&mma[i], 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmA), tma_prefetch_desc(&tmB1), tma_prefetch_desc(&tmB2), tma_prefetch_desc(&tmOut);
  }
  // constant panels and the (empty) initial state
  for (int i = tid; i < (OFF_MISC - OFF_ONES) / 16; i += THREADS)
    reinterpret_cast<uint4*>(smem + OFF_ONES)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < 160 + 128 + 128; i += THREADS) vrow_base[i] = 0.f;
  __syncthreads();
  if (tid < 128)      // element [tok][0] of a 128B-swizzled row sits in 16-byte chunk (0 ^ (tok & 7))
    *reinterpret_cast<uint16_t*>(smem + OFF_ONES + tid * 128 + ((tid & 7) << 4)) = 0x3F80;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (smem_u32(smem) != FM_SMEM_BASE) {
    if (tid == 0) printf("fastmax: unexpected dynamic smem base %u (p1 scan)\n", smem_u32(smem));
    __trap();
  }

  if (warp == 4) {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0) {
      const size_t row0 = (size_t)bh * P.N;
      for (int it = 0; it < ntile; ++it) {
        const int t = REV ? t_hi - 1 - it : t_lo + it;
        const int tok = t * 128;
        if (it > 0) mbar_wait(a_free, (it - 1) & 1, 200 + it);      // each operand is refilled as soon as it is dead
        mbar_arrive_expect_tx(full_s, 2 * 16384);
        for (int h = 0; h < 2; ++h) tma_load_3d(smem + OFF_A + h * 8192, &tmA, full_s, 0, tok + 64 * h, bh);
        if (it > 0) mbar_wait(b1_free, (it - 1) & 1, 100 + it);
        for (int h = 0; h < 2; ++h) tma_load_3d(smem + OFF_B1 + h * 8192, &tmB1, full_s, 0, tok + 64 * h, bh);
        if (it > 0) mbar_wait(b2_free, (it - 1) & 1, 300 + it);
        mbar_arrive_expect_tx(full_v, 16384);
        for (int h = 0; h < 2; ++h) tma_load_3d(smem + OFF_B2 + h * 8192, &tmB2, full_v, 0, tok + 64 * h, bh);
        // pull the tile after next towards L2 while this one is being worked on
        const int it2 = it + FM_P1_PF;
        if (FM_P1_PF > 0 && it2 < ntile) {
          const int t2 = REV ? t_hi - 1 - it2 : t_lo + it2;
          const size_t off = (row0 + (size_t)t2 * 128) * 64;
          const uint32_t nb = (uint32_t)min(128, P.N - t2 * 128) * 128u;       // whole 128-byte rows
          prefetch_l2(P.pA + off, nb), prefetch_l2(P.pB1 + off, nb), prefetch_l2(P.pB2 + off, nb);
        }
      }
    }
  } else if (warp == 5) {
    // ---------------------------------------------------------------- auxiliary MMA issuer
    // Issues what no compute warp has to wait for right away: the state update of a tile (needs only the tile's
    // operands) and the first score group of the NEXT tile (needs the score columns back, i.e. all MMAs of this
    // tile complete).  While a thread issues, it blocks whenever the pipe is busy -- with the other CTA of the SM
    // as well -- and on warp 0 that time sat on every tile's critical path.
    constexpr uint32_t ID_S = make_idesc_bf16(128, 64, false, false);
    constexpr uint32_t ID_UPD = make_idesc_bf16(128, 80, true, true);
    for (int it = 0; it < ntile; ++it) {
      {
        // scores of half 0 of tile it, into columns [0, 64): for it > 0 as soon as the previous tile's second
        // polynomial has read its scores ([32, 96)) and its first P ([0, 32)) has been consumed
        if (lane == 0) {
          if (it > 0) {
            mbar_wait(p1_done, (it - 1) & 1, 400 + it);
            FM_TRX(8);
            mbar_wait(pv0_done, (it - 1) & 1, 450 + it);
            FM_TRX(9);
          }
          mbar_wait(full_s, it & 1, 500 + it);
          FM_TRX(10);
        }
        __syncwarp();
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + C_S, P.dA[kk], P.dB1[0][kk], ID_S, kk > 0);
          umma_commit(&mma[0]);
          FM_TRX(11);
        }
        __syncwarp();
        // state update of tile it
        if (lane == 0) {
          mbar_wait(tile_go, it & 1, 600 + it);
          mbar_wait(full_v, it & 1, 700 + it);
          FM_TRX(12);
        }
        __syncwarp();
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_ss(tmem + C_ST, P.dX[kk], P.dW[kk], ID_UPD, 1);
          umma_commit(&mma[2]);
          FM_TRX(13);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------------------------------------------------------- compute warps (thread = tile row)
    const int q = warp, r = 32 * q + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
    constexpr uint32_t ID_S = make_idesc_bf16(128, 64, false, false);            // scores: A K-major, B K-major
    constexpr uint32_t ID_INTER = (MODE == FWD || MODE == DV) ? make_idesc_bf16(128, NACC, false, true)
                                                              : make_idesc_bf16(128, 64, false, false);
    constexpr uint32_t ID_PV = make_idesc_bf16(128, NACC, false, true);
    const float a = P.a;
    // the state accumulator: zero, or the state of everything before (after) this chunk -- then also its bf16 copy
    // and fp32 vectors, which the first tile reads
    {
      const float* src = P.init ? P.init + ((size_t)bh * P.nchunk + chunk) * (64 * 66 + 80) : nullptr;
      const float sc = (MODE == FWD || MODE == DV) ? a : 1.f;
#pragma unroll 1
      for (int c0 = 0; c0 < 80; c0 += 16) {
        uint32_t z[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) {
          float x = 0.f;
          if (src) {
            if (r < 64) x = c0 + c < 66 ? src[r * 66 + c0 + c] : 0.f;
            else if (r == 64) x = src[64 * 66 + c0 + c];
          }
          z[c] = __float_as_uint(x);
        }
        tmem_st16(lane_addr + C_ST + c0, z);
        if (src && r < 64 && c0 < 64) {      // bf16 copy, rows y = 0..63: 16-byte chunks c0/8 and c0/8 + 1 of the swizzled row
          const uint32_t dst = smem_u32(smem) + OFF_SB + r * 128;
#pragma unroll
          for (int j = 0; j < 2; ++j)
            sts128(dst + ((uint32_t)((c0 / 8 + j) ^ (r & 7)) << 4),
                   pack_bf16x2(sc * __uint_as_float(z[8 * j]), sc * __uint_as_float(z[8 * j + 1])),
                   pack_bf16x2(sc * __uint_as_float(z[8 * j + 2]), sc * __uint_as_float(z[8 * j + 3])),
                   pack_bf16x2(sc * __uint_as_float(z[8 * j + 4]), sc * __uint_as_float(z[8 * j + 5])),
                   pack_bf16x2(sc * __uint_as_float(z[8 * j + 6]), sc * __uint_as_float(z[8 * j + 7])));
        }
        if (src && r < 64 && c0 == 64) {
          vcol_base[r] = __uint_as_float(z[REV ? 1 : 0]);
          if (MODE == FWD)
            *reinterpret_cast<uint16_t*>(smem + OFF_SB + 8192 + r * 128 + ((r & 7) << 4)) =
                __bfloat16_as_ushort(__float2bfloat16_rn(sc * __uint_as_float(z[0])));
        }
        if (src && r == 64) {
#pragma unroll
          for (int c = 0; c < 16; ++c) vrow_base[c0 + c] = __uint_as_float(z[c]);
        }
      }
      tmem_wait_st();
      fence_proxy_async();
      tc_fence_before();
      named_barrier_sync(1, 128);
      tc_fence_after();
    }
    for (int it = 0; it < ntile; ++it) {
      auto wait_mma = [&](int ev) {
        mbar_wait_warp(&mma[ev], it & 1);
        tc_fence_after();
      };
      const int t = REV ? t_hi - 1 - it : t_lo + it;
      const int i_tok = t * 128 + r;                         // this thread's token (row of A)
      const bool valid = i_tok < P.N;
      float dlt_row = 0.f;
      if (MODE != FWD) dlt_row = valid ? P.dlt[(size_t)bh * P.Npad + i_tok] : 0.f;
      FM_TRP(0);
      if (MODE == DV || MODE == DK) {
        // dlt column of the state update's W' (bf16, element 1 of the extra panel's row) and, for DK, the
        // per-column dlt of the polynomial
        *reinterpret_cast<uint16_t*>(smem + OFF_ONES + r * 128 + ((r & 7) << 4) + 2) =
            __bfloat16_as_ushort(__float2bfloat16_rn(dlt_row));
        dcol[r] = dlt_row;
        fence_proxy_async();
        named_barrier_sync(1, 128);
      }
      if (tid == 0) mbar_arrive(tile_go);
      // ---- 1. inter-tile part (only the epilogue needs it: it runs under the polynomials)
      if (q == 0) {
        mbar_wait_warp(full_s, it & 1);                      // A of this tile has landed
        tc_fence_after();
        if (elect_one()) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_ss(tmem + C_ACC, P.dA[kk], (MODE == FWD || MODE == DV) ? P.dSmn[kk] : P.dSk[kk], ID_INTER, kk > 0);
          umma_commit(&mma[4]);
        }
        __syncwarp();
      }
      FM_TRP(1);
      float* vrow = vrow_base + 80 * (it & 1);               // state vectors before this tile ...
      float* vcol = vcol_base + 64 * (it & 1);
      float* vrow_n = vrow_base + 80 * ((it + 1) & 1);       // ... and after it
      float* vcol_n = vcol_base + 64 * ((it + 1) & 1);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        wait_mma(h);
        if (h == 0) FM_TRP(2); else FM_TRP(4);
        // ---- 2. polynomial + tile-local causal mask: rows = A tokens, columns = B tokens 64h .. 64h+63
        const uint32_t sbase = lane_addr + C_S + 32 * h;
        uint32_t pk[32];
        {
          uint32_t sv[64];
          tmem_ld32(sbase, sv);
          tmem_ld32(sbase + 32, sv + 32);
          tmem_wait_ld();
#pragma unroll
          for (int c = 0; c < 64; c += 2) {
            const int col = 64 * h + c;                      // token index of the column inside the tile
            float v0 = __uint_as_float(sv[c]), v1 = __uint_as_float(sv[c + 1]);
            if (SCORE_POLY) {
              v0 = fmaf(a, v0, 1.f), v1 = fmaf(a, v1, 1.f);
            } else if (MODE == DQ) {
              v0 -= dlt_row, v1 -= dlt_row;
            } else {
              v0 -= dcol[col], v1 -= dcol[col + 1];
            }
            const bool k0 = REV ? col >= r : col <= r, k1 = REV ? col + 1 >= r : col + 1 <= r;
            pk[c >> 1] = pack_bf16x2(k0 ? v0 : 0.f, k1 ? v1 : 0.f);
          }
        }
        const uint32_t pcol = h == 0 ? 0 : 64;               // P of half 0 in [0, 32), of half 1 in [64, 96)
        tmem_st32(lane_addr + C_S + pcol, pk);
        tmem_wait_st();
        tc_fence_before();
        named_barrier_sync(1, 128);
        if (h == 1 && tid == 0) mbar_arrive(p1_done);
        if (h == 0) FM_TRP(3); else FM_TRP(5);
        // ---- 3. with the P of half 0: first the scores of half 1 (into columns [32, 96), clear of the P being
        //         consumed), then acc += P . [B2 | extra]
        if (q == 0) {
          if (h == 0) mbar_wait_warp(full_v, it & 1);        // B2 of this tile
          tc_fence_after();
          if (elect_one()) {
            if (h == 0) {
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) umma_ss(tmem + C_S + 32, P.dA[kk], P.dB1[1][kk], ID_S, kk > 0);
              umma_commit(&mma[1]);
            }
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) umma_ts(tmem + C_ACC, tmem + C_S + 64 * h + 8 * kk, P.dB2[h][kk], ID_PV, 1);
            if (h == 0) umma_commit(pv0_done);
            else umma_commit(&mma[3]);
          }
          __syncwarp();
        }
        if (h == 0) {
          // ---- 4. while the scores of half 1 are on their way: the state update of this tile (issued by the auxiliary
          //         warp at the start of the tile) is complete, so refresh the bf16 copy and the fp32 vectors of the
          //         state for the NEXT tile -- once the inter-tile MMA of this tile has read the old copy
          wait_mma(4);
          wait_mma(2);
          if (tid == 0) bulk_wait_group_read0();             // (the previous tile's output has left the staging buffer)
          if (q < 2) {            // rows y = 0..63
            uint32_t v[32], pk2[32];
            const float sc = (MODE == FWD || MODE == DV) ? a : 1.f;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              tmem_ld32(lane_addr + C_ST + 32 * hh, v);
              tmem_wait_ld();
#pragma unroll
              for (int c = 0; c < 32; c += 2)
                pk2[16 * hh + (c >> 1)] = pack_bf16x2(sc * __uint_as_float(v[c]), sc * __uint_as_float(v[c + 1]));
            }
            uint32_t w[8];
            tmem_ld8(lane_addr + C_ST + 64, w);
            tmem_wait_ld();
            const uint32_t dst = smem_u32(smem) + OFF_SB + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              sts128(dst + ((uint32_t)(j ^ (r & 7)) << 4), pk2[4 * j], pk2[4 * j + 1], pk2[4 * j + 2], pk2[4 * j + 3]);
            vcol_n[r] = __uint_as_float(w[REV ? 1 : 0]);       // n1 (column 64) or Rdlt (column 65)
            if (MODE == FWD)       // column 64 of the bf16 state (a n1): element 0 of panel 1's row y
              *reinterpret_cast<uint16_t*>(smem + OFF_SB + 8192 + r * 128 + ((r & 7) << 4)) =
                  __bfloat16_as_ushort(__float2bfloat16_rn(sc * __uint_as_float(w[0])));
          } else if (q == 2) {    // row 64 (lane 0 of this warp): column sums
#pragma unroll 1
            for (int c0 = 0; c0 < 80; c0 += 16) {
              uint32_t v[16];
              tmem_ld16(lane_addr + C_ST + c0, v);         // (warp-collective load; only lane 0's row is kept)
              tmem_wait_ld();
              if (lane == 0) {
#pragma unroll
                for (int c = 0; c < 16; ++c) vrow_n[c0 + c] = __uint_as_float(v[c]);
              }
            }
          }
          fence_proxy_async();                               // the bf16 state is read by the next tile's MMAs
          tc_fence_before();
        }
        if (h == 1 && tid == 0) {
          mbar_arrive(a_free);                               // wait_mma(1) above: inter and both score groups complete
          mbar_arrive(b1_free);                              // ... and the state update: B1 may be refilled
        }
      }
      // ---- 5. all MMAs of the tile complete (the auxiliary warp now issues the next tile's first scores): epilogue
      wait_mma(3);
      FM_TRP(6);
      if (tid == 0) mbar_arrive(b2_free);
      {
        uint32_t pk[32];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t v[32];
          tmem_ld32(lane_addr + C_ACC + 32 * hh, v);
          if (MODE == FWD && hh == 0) {
            uint32_t w[8];
            tmem_ld8(lane_addr + C_ACC + 64, w);
            tmem_wait_ld();
            const float den = __uint_as_float(w[0]) + vrow[64];  // tile-local + a q.n1 + count of earlier tokens
            if (valid) P.g[(size_t)bh * P.N + i_tok] = den;
            dlt_row = 1.f / den;                             // (reused as the row scale)
          } else {
            tmem_wait_ld();
          }
#pragma unroll
          for (int c = 0; c < 32; c += 2) {
            float x0 = __uint_as_float(v[c]), x1 = __uint_as_float(v[c + 1]);
            const int cc = 32 * hh + c;
            if (MODE == FWD) x0 = (x0 + vrow[cc]) * dlt_row, x1 = (x1 + vrow[cc + 1]) * dlt_row;
            else if (MODE == DQ) x0 = a * (x0 - dlt_row * vcol[cc]), x1 = a * (x1 - dlt_row * vcol[cc + 1]);
            else if (MODE == DV) x0 += vrow[cc], x1 += vrow[cc + 1];
            else x0 = a * (x0 - vcol[cc]), x1 = a * (x1 - vcol[cc + 1]);
            pk[16 * hh + (c >> 1)] = pack_bf16x2(x0, x1);
          }
        }
        // (rows straight from registers -- 16 bytes per lane at a 128-byte stride -- took ~2000 cycles per tile; the
        //  previous tile's store left the staging buffer before the barrier after the second polynomial)
        const uint32_t dst = smem_u32(smem) + OFF_STG + r * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          sts128(dst + ((uint32_t)(j ^ (r & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      }
      fence_proxy_async();
      tc_fence_before();
      named_barrier_sync(1, 128);                            // staging and state copy complete and visible; accumulators read
      tc_fence_after();
      if (tid == 0) {
        tma_store_3d(&tmOut, smem + OFF_STG, 0, t * 128, bh);                // rows past N are clipped by the tensor map
        if (t * 128 + 64 < P.N) tma_store_3d(&tmOut, smem + OFF_STG + 8192, 0, t * 128 + 64, bh);
        bulk_commit_group();
      }
      FM_TRP(7);
    }
    if (tid == 0) bulk_wait_group_read0();                   // shared memory must outlive the last store's reads
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

}  // namespace p1

using namespace p1;

template <int MODE>
static int launch_scan(const Problem& pb, int Npad, const void* A, const void* B1, const void* B2, void* out,
                       const float* dlt, float* g, const char* name, const float* init = nullptr, int nchunk = 1,
                       int tiles_per_chunk = 0) {
  const int BH = pb.B * pb.H;
  CUtensorMap tA, tB1, tB2, tO;
  int rc;
  const void* ptr[4] = {A, B1, B2, out};
  CUtensorMap* tm[4] = {&tA, &tB1, &tB2, &tO};
  for (int i = 0; i < 4; ++i)
    if ((rc = make_tmap_bf16_3d(tm[i], ptr[i], 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 64))) return rc;
  Params P{};
  P.N = pb.N, P.Npad = Npad, P.ntile = (pb.N + 127) / 128, P.a = pb.inv_c, P.dlt = dlt, P.g = g;
  P.pA = (const bf16*)A, P.pB1 = (const bf16*)B1, P.pB2 = (const bf16*)B2;
  P.init = init, P.nchunk = nchunk, P.tiles_per_chunk = nchunk > 1 ? tiles_per_chunk : P.ntile;
  const uint32_t base = sm100::FM_SMEM_BASE;
  // operand roles of the state update: X carries the first-order index (k or q), W the value-like tensor (v or u)
  const uint32_t offX = (MODE == FWD || MODE == DV) ? OFF_B1 : OFF_B2;
  const uint32_t offW = (MODE == FWD || MODE == DV) ? OFF_B2 : OFF_B1;
  const uint32_t offE = OFF_ONES;                                             // extra columns of W'
  for (int kk = 0; kk < 4; ++kk) {
    P.dA[kk] = sm100::desc_kmajor_sw128_const(base + OFF_A) + kk * 2;
    for (int h = 0; h < 2; ++h) {
      P.dB1[h][kk] = sm100::desc_kmajor_sw128_const(base + OFF_B1 + h * 8192) + kk * 2;
      // [B2 | extra]: 64-wide N panels OFF_B2 and (FWD only: N = 80) the ones panel; rows 64h + 16kk
      P.dB2[h][kk] = sm100::desc_mnmajor_sw128_const(base + OFF_B2 + h * 8192, OFF_ONES - OFF_B2) + kk * 128;
    }
    P.dSmn[kk] = sm100::desc_mnmajor_sw128_const(base + OFF_SB, 8192) + kk * 128;
    P.dSk[kk] = sm100::desc_kmajor_sw128_const(base + OFF_SB) + kk * 2;
  }
  for (int kk = 0; kk < 8; ++kk) {
    P.dX[kk] = sm100::desc_mnmajor_sw128_const(base + offX, OFF_ONES - offX) + kk * 128;
    P.dW[kk] = sm100::desc_mnmajor_sw128_const(base + offW, offE - offW) + kk * 128;
  }
  auto kern = scan_kernel<MODE>;
  FM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
  FM_LAUNCH(name, pb.stream, kern<<<BH * nchunk, THREADS, SMEM, pb.stream>>>(tA, tB1, tB2, tO, P));
  return FASTMAX_OK;
}

bool p1scan_usable(const Problem& pb) {
  // (the sequence-parallel entries, which hand states between ranks, stay on the chunked path)
  return pb.p == 1 && pb.causal && pb.D == 64 && pb.dtype == FASTMAX_DTYPE_BF16 && pb.N >= 256 &&
         !pb.sp_init_fwd && !pb.sp_init_rev;
}

// One CTA per (b,h) walks the whole sequence when there are enough (b,h) to fill the GPU (two CTAs per SM).  With
// fewer, each sequence is cut into chunks of at least 8 tiles: a streaming reduce pass (fm_p1full.cu) gives every
// chunk's own state, a prefix kernel the state before (after) it, and the chunks scan in parallel from there.
static int scan_chunks(const Problem& pb, int* tiles_per_chunk) {
  const int BH = pb.B * pb.H, ntile = (pb.N + 127) / 128;
  int nc = BH >= 148 ? 1 : std::min(std::min(296 / BH, p1full_max_split()), ntile / 8);
  nc = std::max(nc, 1);
  const int tpc = (ntile + nc - 1) / nc;
  *tiles_per_chunk = tpc;
  return (ntile + tpc - 1) / tpc;
}

size_t p1scan_region_bytes(const Problem& pb) {
  int tpc;
  const int nc = scan_chunks(pb, &tpc);
  return nc > 1 ? p1full_prefix_bytes(pb.B, pb.H, nc) : 0;
}

// region: scratch of p1scan_region_bytes (chunked sequences only)
int p1scan_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g, void* region) {
  int tpc, rc;
  const int nc = scan_chunks(pb, &tpc);
  const float* init = nullptr;
  if (nc > 1 && (rc = p1full_prefix_states(pb, 0, k, v, nullptr, nc, tpc, 0, region, &init))) return rc;
  return launch_scan<FWD>(pb, (int)align_up(pb.N, 64), q, k, v, o, nullptr, g, "p1_scan_fwd", init, nc, tpc);
}

// u = do/g (bf16), dlt = (do.o)/g (fp32, (BH, Npad)) from the prep kernel
int p1scan_bwd(const Problem& pb, int Npad, const void* q, const void* k, const void* v, const void* u, const float* dlt,
               void* dq, void* dk, void* dv, void* region_s, void* region_r) {
  int tpc, rc;
  const int nc = scan_chunks(pb, &tpc);
  const float *init_s = nullptr, *init_r = nullptr;
  if (nc > 1) {
    if ((rc = p1full_prefix_states(pb, Npad, k, v, nullptr, nc, tpc, 0, region_s, &init_s))) return rc;
    if ((rc = p1full_prefix_states(pb, Npad, q, u, dlt, nc, tpc, 1, region_r, &init_r))) return rc;
  }
  if ((rc = launch_scan<DQ>(pb, Npad, u, v, k, dq, dlt, nullptr, "p1_scan_dq", init_s, nc, tpc))) return rc;
  if ((rc = launch_scan<DV>(pb, Npad, k, q, u, dv, dlt, nullptr, "p1_scan_dv", init_r, nc, tpc))) return rc;
  return launch_scan<DK>(pb, Npad, v, u, q, dk, dlt, nullptr, "p1_scan_dk", init_r, nc, tpc);
}

}  // namespace fm
