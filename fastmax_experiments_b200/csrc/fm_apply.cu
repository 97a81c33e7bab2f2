// fm_apply.cu -- inter-segment part of the FastMax chunked scan: contraction of a token tile with the
// boundary moment state, with the Khatri-Rao operand GENERATED into tensor memory.
//
//   out[i][n] = out_scale * inv_g_i * ( coef64 * acc[i][n] + vec[n] + partial[i][n] )
//   acc[i][n] = sum_x sum_k (rho X[i][x] * Y[i][k]) * B_x[k][n]   +   sum_k Y[i][k] * B_64[k][n]
//
// B_x = slab x of the state record (fm_moments.cuh), read MN-major (k = slab row y, n = column a: forward
// o and dv) or K-major (k = column a, n = row r: dq and dk).  The products (rho X[x]) * Y[k] are formed on
// the CUDA cores from the thread's own token row, rounded to bf16 and written to tensor memory as the A
// operand of tcgen05.mma (TS form); every slab accumulates into ONE 128x64 fp32 accumulator per token tile,
// so nothing is read back from tensor memory until the tile is finished.  (The round-1 kernel formed
// T_x = Y . B_x for every slab and contracted x on the CUDA cores: 16 KiB of tensor-memory reads per token
// and apply, which -- not the tensor pipe -- bounded it.)  This is the transpose of the build kernel:
//   build: D[(x,y)][a] += A[(x,y)][token] * V[token][a]        apply: D[token][n] += A[token][(x,k)] * S[(x,k)][n]
//
// Symmetric uses (X == Y: forward o on S, dv on R) need only the slab rows y >= 16*floor(x/16): S2[x][y][a]
// is symmetric in (x,y), so the 16x16 blocks below the diagonal are skipped and the blocks above it count
// twice (weight folded into the generated operand): 160 MMA k-steps per tile instead of 256.
//
//   forward o : X = Y = q, state S, MN-major        dq : X = q, Y = do/g, state S, K-major
//   dv        : X = Y = k, state R, MN-major        dk : X = k, Y = v,    state R, K-major
//
// CTA = 2 token tiles of 128 rows, 21 warps:
//   * 16 generator warps = 2 tiles x 2 sets x 4 tensor-memory lane quadrants; thread = token row.  The unit
//     of hand-over is a PAIR of slabs (K = 128, 64 columns of A): the sets take pairs alternately, so one
//     pair costs a warp one buffer wait, four 16-column tcgen05.st, one wait::st and one arrive.  (Measured
//     with one warp per tile quadrant and a hand-over per slab: ~700 cycles per slab in the generator,
//     against 256 cycles of MMA -- the generators, not the buffer depth, bounded the kernel.)
//   * 4 MMA issuer warps = 2 tiles x 2 (even / odd pairs); converged warps, descriptors from the constant
//     bank, tensor-memory addresses as immediates (fm_sm100.cuh).  An issuer's own sequence per pair (two
//     barrier waits, fence, elect, commit) costs ~500 cycles before any MMA -- measured with everything else
//     compiled out: 1.2 ms per launch -- and its MMA issue blocks while the pipe is busy, so with one issuer
//     per tile the pipe idled during every hand-over; two per tile overlap them.  One commit per pair: the
//     `done` barrier of a ring stage releases both the stage and the A buffers of that pair.
//   * 1 TMA producer warp streaming slab pairs through a 6-stage ring (16 KiB per stage); symmetric uses
//     fetch only the rows they need, one box per slab (four tensor maps with 64/48/32/16-row boxes).
//     (Letting the issuer warps feed the ring coupled the tiles' MMA issue to each other's ring waits: 2x slower.)
//   Every barrier is waited on by the same warps in every phase (a waiter that skips a phase is fooled by
//   the parity of the older one): the barriers are indexed by ring stage (pair mod 6), and a stage always
//   belongs to the same generator set and the same issuers.
#include "fm_common.cuh"
#include "fm_sm100.cuh"
#include "fm_moments.cuh"

FM_TIMING_DEFINE(apply)
#ifndef FM_EXP
#define FM_EXP 0      // developer experiments (wrong results): 1 no MMAs, 2 no operand stores, 4 no slab loads, 8 no products
#endif

#ifdef FM_TRACE
// developer event trace of ONE CTA (blockIdx = (1, 0)): clock64 at the hand-over points of every slab pair
static __device__ long long fm_trace_tab[8][40];
extern "C" int fastmax_debug_trace_fetch_apply(long long* out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(out, fm_trace_tab, sizeof(long long) * 8 * 40) == cudaSuccess ? 0 : -4;
}
#define FM_TR(ev, s) do { if (blockIdx.x == 1 && blockIdx.y == 0) fm_trace_tab[ev][s] = clock64(); } while (0)
#else
#define FM_TR(ev, s) ((void)0)
#endif

namespace fm {
namespace mom {
using namespace sm100;
typedef __nv_bfloat16 bf16;

namespace gen {
constexpr int T = 2;                        // token tiles per CTA
constexpr int NBUF = 3;                     // A buffers per tile (one slab pair each: 64 columns)
constexpr int NSTAGE = 6;                   // ring stages (one slab pair each: 16 KiB)
constexpr int NPAIR = 33;                   // 32 pairs of second-order slabs + slab 64 alone
constexpr int ACOL = 64 * T;                // first A-buffer column
constexpr int GEN_WARPS = 16;
constexpr int THREADS = 21 * 32;
constexpr int RING_BYTES = NSTAGE * 16384;
constexpr int STG_BYTES = T * 16384;        // bf16 output staging (TMA store source)
constexpr int BAR_OFF = RING_BYTES + STG_BYTES;
constexpr int SMEM = BAR_OFF + 1024 + 1024 /*align slack*/;
// barrier indices
constexpr int B_FULL = 0, B_DONE = NSTAGE, B_AREADY = 2 * NSTAGE /* [T][NSTAGE] */, B_ACC = B_AREADY + T * NSTAGE,
              NBARS = B_ACC + T;
static_assert(NBARS * 8 <= 512, "barrier block");
static_assert(64 * T + 64 * T * NBUF <= 512, "tensor memory");
}  // namespace gen

struct GenParams {
  int N, L, nseg, causal;
  int state_reverse;        // 1: records are exclusive suffixes (the last segment has none)
  int edge_state;           // sequence parallel: first (forward states) / last (reverse) local segment has a state
  int vec_sel;              // -1 none, 0: vec[0:64), 1: vec[64:128)
  float rho;                // scale of X inside the generated operand (= coef of the 2nd-order term / coef64)
  float coef64;             // scale of the accumulator (= coef of the first-order term)
  float out_scale;
  const bf16* X;            // (BH, N, 64) per-token scalars
  const bf16* Y;            // (BH, N, 64) per-token vectors
  const float* partial;     // blocked fp32 intra-segment partial (fm_moments.cuh), nullable
  const float* gden;        // (BH, N) fp32: rows are divided by it (forward), nullable
  const uint8_t* state;
  unsigned long long dB[gen::NSTAGE][2][4];   // B-operand descriptors [ring stage][slab of the pair][k-step]
};

template <int NK>
__device__ __forceinline__ void st_a(uint32_t taddr, const uint32_t* a) {
  if constexpr (NK == 4) tmem_st32(taddr, a);
  else if constexpr (NK == 3) tmem_st16(taddr, a), tmem_st8(taddr + 16, a + 16);
  else if constexpr (NK == 2) tmem_st16(taddr, a);
  else tmem_st8(taddr, a);
}

// k-steps [J0, J0 + NJ) of one slab of the generated operand: A[row][8j+w] = xs(j) * Y[16(GX+j) + 2w, +1]
template <bool SYM, int GX, int J0, int NJ>
__device__ __forceinline__ void gen_store(const uint32_t (&Yp)[32], float xf, uint32_t taddr) {
  const uint32_t xs1 = pack_bf16x2(xf, xf);
  const uint32_t xs2 = SYM ? pack_bf16x2(xf + xf, xf + xf) : xs1;   // blocks above the diagonal count twice
  uint32_t a[8 * NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int w = 0; w < 8; ++w)
      a[8 * j + w] = (FM_EXP & 8) ? Yp[8 * (GX + J0 + j) + w] : mul_bf16x2(J0 + j == 0 ? xs1 : xs2, Yp[8 * (GX + J0 + j) + w]);
  if (FM_EXP & 2) return;
  if constexpr (NJ == 2) tmem_st16(taddr + 8 * J0, a);
  else tmem_st8(taddr + 8 * J0, a);
}

template <bool SYM, int GX>
__device__ __forceinline__ void gen_slab(const uint32_t (&Yp)[32], float xf, uint32_t taddr) {
  constexpr int NK = 4 - GX;
  if constexpr (NK >= 2) gen_store<SYM, GX, 0, 2>(Yp, xf, taddr);
  if constexpr (NK == 4) gen_store<SYM, GX, 2, 2>(Yp, xf, taddr);
  if constexpr (NK == 3) gen_store<SYM, GX, 2, 1>(Yp, xf, taddr);
  if constexpr (NK == 1) gen_store<SYM, GX, 0, 1>(Yp, xf, taddr);
}

// the slab pairs [S0, S0 + NS) of this warp set (every other pair): generate, store, hand over
template <bool SYM, int GX>
__device__ __forceinline__ void gen_phase(const uint32_t (&Yp)[32], const uint32_t* xrow, bool valid, float rho,
                                          uint32_t a_tile, uint64_t* bars, int t, int g, int lane) {
  constexpr int S0 = SYM ? 8 * GX : 0, NS = SYM ? 8 : 32;
  uint32_t xw_next = valid ? __ldg(xrow + S0 + g) : 0u;
#pragma unroll 1
  for (int s = S0 + g; s < S0 + NS; s += 2) {
    const uint32_t xw = xw_next;
    xw_next = (valid && s + 2 < S0 + NS) ? __ldg(xrow + s + 2) : 0u;
    const uint32_t abuf = a_tile + 64 * (s % gen::NBUF);
    const bool tr = t == 0 && lane == 0 && (threadIdx.x >> 5 & 3) == 0;
    if (tr) FM_TR(0, s);
    // the buffer was last read by the MMAs of pair s - NBUF, whose completion is the `done` of that pair's stage
    if (s >= gen::NBUF)
      mbar_wait_warp(&bars[gen::B_DONE + (s - gen::NBUF) % gen::NSTAGE], ((s - gen::NBUF) / gen::NSTAGE) & 1);
    if (tr) FM_TR(1, s);
    tc_fence_after();
    gen_slab<SYM, GX>(Yp, __uint_as_float(xw << 16) * rho, abuf);
    gen_slab<SYM, GX>(Yp, __uint_as_float(xw & 0xffff0000u) * rho, abuf + 32);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bars[gen::B_AREADY + t * gen::NSTAGE + s % gen::NSTAGE]);
    if (tr) FM_TR(2, s);
  }
}

// MMA issuer (TT, PAR): the pairs s = PAR (mod 2) of tile TT.  A converged warp; `live` = the tile has rows
// (a tile past the end of the sequence only keeps the ring protocol going).
template <bool SYM, int TT, int PAR>
__device__ __forceinline__ void issue_pairs(const GenParams& P, uint64_t* bars, bool live, int lane) {
  constexpr uint32_t IDESC = make_idesc_bf16(128, 64, false, SYM);
  constexpr uint32_t BAR0 = FM_SMEM_BASE + gen::BAR_OFF;
#pragma unroll 1
  for (int s0 = PAR; s0 < gen::NPAIR; s0 += gen::NSTAGE) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {                     // ring stage u = PAR + 2i (compile time), A buffer u % 3
      const int u = PAR + 2 * i;
      const int buf = u % gen::NBUF;
      const int s = s0 + 2 * i;
      if (s < gen::NPAIR) {
        const int nk = (SYM && s < 32) ? 4 - (s >> 3) : 4;
        const bool last = s == gen::NPAIR - 1;
        if (lane == 0) {
          uint64_t* fb = &bars[gen::B_FULL + u];
          uint64_t* ab = &bars[gen::B_AREADY + TT * gen::NSTAGE + u];
          const uint32_t ph = (s / gen::NSTAGE) & 1;
          if (TT == 0) FM_TR(3, s);
          // both barriers are usually complete: probe them together so the two probe latencies overlap
          const bool ok_f = (FM_EXP & 4) ? true : mbar_try_wait(fb, ph);
          const bool ok_a = live ? mbar_try_wait(ab, ph) : true;
          if (!ok_f) mbar_wait(fb, ph, 2000 + s);
          if (!ok_a) mbar_wait(ab, ph, 3000 + s);
          if (!live) mbar_arrive(&bars[gen::B_DONE + u]);
          if (TT == 0) FM_TR(4, s);
        }
        __syncwarp();
        if (live) {
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              if (e == 1 && last) break;
#pragma unroll
              for (int j = 0; j < 4; ++j)
                if (j < nk && !(FM_EXP & 1))
                  umma_ts(64 * TT, gen::ACOL + 64 * (TT * gen::NBUF + buf) + 32 * e + 8 * j, P.dB[u][e][j], IDESC, 1u);
            }
            umma_commit_addr(BAR0 + 8 * (gen::B_DONE + u));
          }
          __syncwarp();
          if (TT == 0 && lane == 0) FM_TR(5, s);
        }
      }
    }
  }
  // the tile's accumulator is final when BOTH of its issuers' MMAs have completed
  if (live) {
    tc_fence_after();
    if (elect_one()) umma_commit_addr(BAR0 + 8 * (gen::B_ACC + TT));
    __syncwarp();
  }
}

template <bool SYM>
__global__ void __launch_bounds__(gen::THREADS, 1)
apply_gen_kernel(const __grid_constant__ CUtensorMap tmS64, const __grid_constant__ CUtensorMap tmS48,
                 const __grid_constant__ CUtensorMap tmS32, const __grid_constant__ CUtensorMap tmS16,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ GenParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;                                // [NSTAGE][2 slabs][8 KiB]
  uint8_t* stg = smem + gen::RING_BYTES;               // [T][16 KiB] bf16 output tiles (128B-swizzled rows)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + gen::BAR_OFF);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + gen::BAR_OFF + 512);
  float* vec_s = reinterpret_cast<float*>(smem + gen::BAR_OFF + 768);   // [64]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y;
  constexpr int CT = 128 * gen::T;
  const int t0 = blockIdx.x * CT;
  const int seg = t0 / P.L;
  const bool has_state =
      !P.causal || (P.state_reverse ? (seg < P.nseg - 1 || P.edge_state) : (seg > 0 || P.edge_state));
  const int rec_idx = P.causal ? bh * P.nseg + seg : bh;

  if (warp == 16) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  if (tid == 0) {
    for (int i = 0; i < gen::NSTAGE; ++i) mbar_init(&bars[gen::B_FULL + i], 1), mbar_init(&bars[gen::B_DONE + i], gen::T);
    for (int i = 0; i < gen::T * gen::NSTAGE; ++i) mbar_init(&bars[gen::B_AREADY + i], 4);
    for (int i = 0; i < gen::T; ++i) mbar_init(&bars[gen::B_ACC + i], 2);
    fence_barrier_init();
    tma_prefetch_desc(&tmS64);
    tma_prefetch_desc(&tmS16);
    tma_prefetch_desc(&tmOut);
  }
  if (tid < 64) {
    float v = 0.f;
    if (has_state && P.vec_sel >= 0)
      v = reinterpret_cast<const float*>(P.state + (size_t)rec_idx * REC_BYTES + VEC_OFF)[P.vec_sel * 64 + tid];
    vec_s[tid] = v;
  }
  // generator warp = (tile t, set g, lane quadrant q); its threads fetch their token row before the set-up barrier
  const int t = warp >> 3, g = (warp >> 2) & 1, q = warp & 3;
  const int i = t0 + 128 * t + 32 * q + lane;
  const bool is_gen = warp < gen::GEN_WARPS;
  const bool valid = is_gen && i < P.N;
  const size_t grow = (size_t)bh * P.N + (valid ? i : 0);
  uint32_t Yp[32];
  if (is_gen) {
    const uint4* yrow = reinterpret_cast<const uint4*>(P.Y + grow * 64);
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      const uint4 v = (valid && has_state) ? __ldg(yrow + c8) : make_uint4(0u, 0u, 0u, 0u);
      Yp[4 * c8] = v.x, Yp[4 * c8 + 1] = v.y, Yp[4 * c8 + 2] = v.z, Yp[4 * c8 + 3] = v.w;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  assert_bases(tmem, smem);
  // accumulators start at zero (every MMA accumulates: the tile's two issuers run independently); warp (t,g,q)
  // clears columns [32g, 32g+32) of its lane quadrant of tile t
  if (is_gen) {
    uint32_t z[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) z[c] = 0u;
    tmem_st32(tmem + ((uint32_t)(32 * q) << 16) + 64 * t + 32 * g, z);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (is_gen) {
    // ---------------------------------------------------------------- generators + epilogue
    const bool tile_live = t0 + 128 * t < P.N;
    const float inv_g = (valid && P.gden) ? 1.f / __ldg(P.gden + grow) : 1.f;
    const float4* part = nullptr;
    if (P.partial && tile_live)
      part = reinterpret_cast<const float4*>(P.partial + ((size_t)bh * ((P.N + 127) / 128) + (size_t)(t0 / 128 + t)) *
                                                             PART_TILE_FLOATS) + (32 * q + lane);
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * q) << 16);
    if (has_state && tile_live) {
      const uint32_t a_tile = lane_addr + gen::ACOL + 64 * (t * gen::NBUF);
      const uint32_t* xrow = reinterpret_cast<const uint32_t*>(P.X + grow * 64);
      gen_phase<SYM, 0>(Yp, xrow, valid, P.rho, a_tile, bars, t, g, lane);
      if constexpr (SYM) {
        gen_phase<SYM, 1>(Yp, xrow, valid, P.rho, a_tile, bars, t, g, lane);
        gen_phase<SYM, 2>(Yp, xrow, valid, P.rho, a_tile, bars, t, g, lane);
        gen_phase<SYM, 3>(Yp, xrow, valid, P.rho, a_tile, bars, t, g, lane);
      }
      // the intra-segment partial is wanted next: pull this warp's half of the columns into L2
      // (the tile's block is [16 column chunks][128 rows] float4: 2 KiB per chunk, the set takes 8 chunks;
      //  thread r touches line r of its set's 16 KiB)
      if (part) {
        const float* pf = reinterpret_cast<const float*>(part - (32 * q + lane)) + g * 4096 + (32 * q + lane) * 32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
      }
      if (g == 0) {
        // slab 64 alone (pair 32: stage 32 % 6 = 2, buffer 32 % 3 = 2): the plain Y row (first-order term)
        constexpr int s = gen::NPAIR - 1;
        mbar_wait_warp(&bars[gen::B_DONE + (s - gen::NBUF) % gen::NSTAGE], ((s - gen::NBUF) / gen::NSTAGE) & 1);
        tc_fence_after();
        tmem_st32(a_tile + 64 * (s % gen::NBUF), Yp);
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[gen::B_AREADY + t * gen::NSTAGE + s % gen::NSTAGE]);
      }
      mbar_wait_warp(&bars[gen::B_ACC + t], 0);
      tc_fence_after();
    }
    if (tile_live) {
      // epilogue: set g takes columns [32g, 32g + 32) of the warp's 32 rows
      const float scale = P.out_scale * inv_g;
      const float c64 = P.coef64;
      const int rloc = 32 * q + lane;
      const uint32_t dst = smem_u32(stg) + t * 16384 + rloc * 128;
      uint32_t v[32];
      if (has_state) {
        tmem_ld32(lane_addr + 64 * t + 32 * g, v);
        tmem_wait_ld();
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] = 0u;
      }
      uint32_t pk[16];
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 pv = part ? __ldg(part + (8 * g + c4) * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
        const int c = 32 * g + 4 * c4;
        const float r0 = (fmaf(c64, __uint_as_float(v[4 * c4 + 0]), vec_s[c + 0]) + pv.x) * scale;
        const float r1 = (fmaf(c64, __uint_as_float(v[4 * c4 + 1]), vec_s[c + 1]) + pv.y) * scale;
        const float r2 = (fmaf(c64, __uint_as_float(v[4 * c4 + 2]), vec_s[c + 2]) + pv.z) * scale;
        const float r3 = (fmaf(c64, __uint_as_float(v[4 * c4 + 3]), vec_s[c + 3]) + pv.w) * scale;
        pk[2 * c4] = pack_bf16x2(r0, r1), pk[2 * c4 + 1] = pack_bf16x2(r2, r3);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j)
        sts128(dst + ((uint32_t)((4 * g + j) ^ (rloc & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      fence_proxy_async();
      named_barrier_sync(1 + t, 256);
      if (g == 0 && q == 0 && lane == 0) {
        const int ob = t0 + 128 * t;
        tma_store_3d(&tmOut, stg + t * 16384, 0, ob, bh);                    // rows past N are clipped by the tensor map
        if (ob + 64 < P.N) tma_store_3d(&tmOut, stg + t * 16384 + 8192, 0, ob + 64, bh);
        bulk_commit_group();
        bulk_wait_group_read0();
      }
    }
  } else if (warp < 20) {
    // ---------------------------------------------------------------- MMA issuers (tile, even / odd pairs)
    const int tt = (warp - 16) >> 1, par = (warp - 16) & 1;
    if (has_state) {
      const bool live = t0 + 128 * tt < P.N;
      if (tt == 0) {
        if (par == 0) issue_pairs<SYM, 0, 0>(P, bars, live, lane);
        else issue_pairs<SYM, 0, 1>(P, bars, live, lane);
      } else {
        if (par == 0) issue_pairs<SYM, 1, 0>(P, bars, live, lane);
        else issue_pairs<SYM, 1, 1>(P, bars, live, lane);
      }
    }
  } else {
    // ---------------------------------------------------------------- TMA producer
    if (lane == 0 && has_state && !(FM_EXP & 4)) {
      for (int s = 0; s < gen::NPAIR; ++s) {
        const int st = s % gen::NSTAGE;
        FM_TR(6, s);
        if (s >= gen::NSTAGE) mbar_wait(&bars[gen::B_DONE + st], ((s / gen::NSTAGE) - 1) & 1, 1000 + s);
        FM_TR(7, s);
        uint8_t* dst = ring + st * 16384;
        uint64_t* fb = &bars[gen::B_FULL + st];
        if (s == gen::NPAIR - 1) {
          mbar_arrive_expect_tx(fb, 8192);
          tma_load_3d(dst, &tmS64, fb, 0, 64 * 64, rec_idx);
        } else if (SYM) {
          // rows y >= 16 gx of both slabs, one box each
          const int gx = s >> 3;
          const CUtensorMap* tm = gx == 0 ? &tmS64 : (gx == 1 ? &tmS48 : (gx == 2 ? &tmS32 : &tmS16));
          mbar_arrive_expect_tx(fb, 2 * (64 - 16 * gx) * 128);
          tma_load_3d(dst, tm, fb, 0, (2 * s) * 64 + 16 * gx, rec_idx);
          tma_load_3d(dst + 8192, tm, fb, 0, (2 * s + 1) * 64 + 16 * gx, rec_idx);
        } else {
          mbar_arrive_expect_tx(fb, 16384);
          tma_load_3d(dst, &tmS64, fb, 0, (2 * s) * 64, rec_idx);
          tma_load_3d(dst + 8192, &tmS64, fb, 0, (2 * s + 1) * 64, rec_idx);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem, 512);
}

}  // namespace mom

using namespace mom;

template <bool SYM>
static int launch_gen_t(const Problem& pb, int BH, int nrec, const char* name, GenParams P, void* out) {
  CUtensorMap tmS[4], tmOut;
  int rc;
  for (int i = 0; i < 4; ++i)
    if ((rc = make_tmap_bf16_3d(&tmS[i], P.state, 64, NSLAB * 64, (uint64_t)BH * nrec, 128, REC_BYTES, 64, 64 - 16 * i)))
      return rc;
  if ((rc = make_tmap_bf16_3d(&tmOut, out, 64, P.N, BH, 128, (uint64_t)P.N * 128, 64, 64))) return rc;
  for (int st = 0; st < gen::NSTAGE; ++st)
    for (int e = 0; e < 2; ++e)
      for (int j = 0; j < 4; ++j) {
        const uint32_t base = sm100::FM_SMEM_BASE + st * 16384 + e * 8192;
        P.dB[st][e][j] = SYM ? sm100::desc_mnmajor_sw128_const(base, 8192) + j * 128     // k-step = 16 rows = 2048 B
                             : sm100::desc_kmajor_sw128_const(base) + j * 2;             // k-step = 16 columns = 32 B
      }
  auto kern = apply_gen_kernel<SYM>;
  FM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, gen::SMEM));
  dim3 grid((P.N + 128 * gen::T - 1) / (128 * gen::T), BH);
  FM_LAUNCH(name, pb.stream,
            kern<<<grid, gen::THREADS, gen::SMEM, pb.stream>>>(tmS[0], tmS[1], tmS[2], tmS[3], tmOut, P));
  return FASTMAX_OK;
}

int launch_apply_gen(const Problem& pb, int BH, int nrec, const char* name, const ApplyDesc& d, int variant) {
  (void)variant;
  GenParams P{};
  P.N = pb.N, P.L = d.L, P.nseg = d.nseg, P.causal = pb.causal;
  P.state_reverse = d.state_reverse, P.edge_state = d.edge_state, P.vec_sel = d.vec_sel;
  P.rho = d.coef_reg / d.coef64, P.coef64 = d.coef64, P.out_scale = d.out_scale;
  P.X = (const bf16*)d.X, P.Y = (const bf16*)d.Y, P.partial = d.partial, P.gden = d.gden;
  P.state = (const uint8_t*)d.state;
  return d.b_kmajor ? launch_gen_t<false>(pb, BH, nrec, name, P, d.out) : launch_gen_t<true>(pb, BH, nrec, name, P, d.out);
}

}  // namespace fm
