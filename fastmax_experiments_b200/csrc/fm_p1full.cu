// fm_p1full.cu -- first-order (p = 1) NON-causal FastMax as two streaming passes per tensor pair.
//
// Without a mask every row sees the same moment state, the total over the sequence (reference: the `mask is False`
// branch of fastattention_einops.forward, fastmax.py:85-97, and gradient_o_{q,k,v}, fastmax.py:324-431, for p = 1):
//
//     S  = sum_j [k_j | 1]^T [v_j | 1]          (64+1) x (64+1):  S1 | n1 ; S0 | count
//     R  = sum_i [q_i | 1]^T [u_i | 1 | dlt_i]  the same for the backward pass (u = do/g, dlt = (do.o)/g)
//
//     o_i  = (S0 + a q_i.S1) / (N + a q_i.n1)          g_i = N + a q_i.n1
//     dq_i = a (u_i.S1^T - dlt_i n1)
//     dv_j = R0 + a k_j.R1
//     dk_j = a (v_j.R1^T - Rdlt)
//
// so the work is HBM streaming: one REDUCE pass over (k, v) resp. (q, u) that accumulates the state on the tensor
// pipe (the state update of fm_p1scan.cu, nothing else), and one APPLY pass per output that multiplies 128-token
// tiles with the bf16 state and finishes the rows on CUDA cores (fm_p1scan.cu's inter-tile MMA + epilogue).  Both
// are tile-parallel -- a sequence is cut into `nsplit` ranges for the reduce (fp32 partial states, summed by a tiny
// finalize kernel) and into groups of tiles for the apply -- so the grid fills the GPU for any (B*H, N).  The chunked
// path this replaces for p = 1 (transpose + build + the intra kernels' state term) ran at 29-32 % of the HBM roof.
#include "fm_common.cuh"
#include "fm_sm100.cuh"
#include "fm_moments.cuh"

namespace fm {
namespace p1f {
using namespace sm100;
typedef __nv_bfloat16 bf16;

enum Mode { FWD = 0, DQ = 1, DV = 2, DK = 3 };

// fp32 partial state of one (b,h, split): rows y = 0..63 x columns 0..65 (64 = sum x, 65 = sum dlt x), then row 64
// (column sums of w: 64 values; 16 more of the extra panel, unused)
constexpr int PART_FLOATS = 64 * 66 + 80;
constexpr int PART_BYTES = PART_FLOATS * 4;
// final state of one (b,h): 16 KiB bf16 in the shared-memory operand layout (panel 0 = [64 y][64 c] 128B-swizzled rows,
// panel 1 = [64 y][64], column 64 in element 0), then fp32 row 64 [80], column 64 [64], column 65 [64]
constexpr int FINAL_VEC = 16384;
constexpr int FINAL_BYTES = 16384 + (80 + 64 + 64) * 4 + 192;   // 17408, a multiple of 256
constexpr int MAX_SPLIT = 16;

__device__ __forceinline__ void bulk_wait_group_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }

// ============================================================================================ reduce
constexpr int R_STAGES = 2;
constexpr int R_OFF_ONES = R_STAGES * 32768;          // one extra panel per stage: element [tok][0] = 1, [tok][1] = dlt
constexpr int R_OFF_MISC = R_OFF_ONES + R_STAGES * 16384;
constexpr int R_SMEM = R_OFF_MISC + 1024 + 1024;
constexpr int THREADS = 192;                          // 4 compute warps, TMA producer warp, MMA issuer warp

struct ReduceParams {
  int N, Npad, ntile, nsplit, tiles_per_split;
  const float* dlt;           // (BH, Npad) or null
  float* partial;             // (BH, nsplit, PART_FLOATS)
  unsigned long long dX[R_STAGES][8];   // [X | extra panel] MN-major A operand, k-step = 16 tokens
  unsigned long long dW[R_STAGES][8];   // [W | extra panel] MN-major B operand
};

template <bool REV>
__global__ void __launch_bounds__(THREADS, 2)
reduce_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW,
              const __grid_constant__ ReduceParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + R_OFF_MISC);
  uint64_t* full = bars;            // [stage] X and W landed
  uint64_t* freeb = bars + 2;       // [stage] the stage's MMAs are complete
  uint64_t* dready = bars + 4;      // [stage] the dlt column of the stage's extra panel is written (REV)
  uint64_t* done = bars + 6;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, bh = blockIdx.y;
  const int t0 = split * P.tiles_per_split, t1 = min(P.ntile, t0 + P.tiles_per_split);
  const int nt = t1 - t0;                               // >= 1 by construction of the grid

  if (warp == 4) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  if (tid == 0) {
    for (int s = 0; s < R_STAGES; ++s) mbar_init(&full[s], 1), mbar_init(&freeb[s], 1), mbar_init(&dready[s], 1);
    mbar_init(done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmX), tma_prefetch_desc(&tmW);
  }
  for (int i = tid; i < R_STAGES * 16384 / 16; i += THREADS)
    reinterpret_cast<uint4*>(smem + R_OFF_ONES)[i] = make_uint4(0u, 0u, 0u, 0u);
  __syncthreads();
  if (tid < 128)
    for (int s = 0; s < R_STAGES; ++s)    // element [tok][0] of a 128B-swizzled row sits in 16-byte chunk (0 ^ (tok & 7))
      *reinterpret_cast<uint16_t*>(smem + R_OFF_ONES + s * 16384 + tid * 128 + ((tid & 7) << 4)) = 0x3F80;
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (smem_u32(smem) != FM_SMEM_BASE) {
    if (tid == 0) printf("fastmax: unexpected dynamic smem base %u (p1 reduce)\n", smem_u32(smem));
    __trap();
  }

  if (warp == 4) {
    if (lane == 0) {
      for (int i = 0; i < nt; ++i) {
        const int s = i % R_STAGES, tok = (t0 + i) * 128;
        if (i >= R_STAGES) mbar_wait(&freeb[s], ((i / R_STAGES) - 1) & 1, 100 + i);
        mbar_arrive_expect_tx(&full[s], 32768);
        for (int h = 0; h < 2; ++h) {
          tma_load_3d(smem + s * 32768 + h * 8192, &tmX, &full[s], 0, tok + 64 * h, bh);
          tma_load_3d(smem + s * 32768 + 16384 + h * 8192, &tmW, &full[s], 0, tok + 64 * h, bh);
        }
      }
    }
  } else if (warp == 5) {
    constexpr uint32_t ID_UPD = make_idesc_bf16(128, 80, true, true);
    for (int i = 0; i < nt; ++i) {
      const int s = i % R_STAGES;
      const uint32_t ph = (i / R_STAGES) & 1;
      if (lane == 0) {
        mbar_wait(&full[s], ph, 200 + i);
        if (REV) mbar_wait(&dready[s], ph, 300 + i);
      }
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) umma_ss(tmem, P.dX[s][kk], P.dW[s][kk], ID_UPD, (i > 0 || kk > 0) ? 1 : 0);
        umma_commit(&freeb[s]);
        if (i == nt - 1) umma_commit(done);
      }
      __syncwarp();
    }
  } else {
    const int r = tid;
    if (REV) {
      for (int i = 0; i < nt; ++i) {
        const int s = i % R_STAGES;
        const int tok = (t0 + i) * 128 + r;
        const float d = tok < P.N ? P.dlt[(size_t)bh * P.Npad + tok] : 0.f;
        if (i >= R_STAGES) mbar_wait_warp(&freeb[s], ((i / R_STAGES) - 1) & 1);
        *reinterpret_cast<uint16_t*>(smem + R_OFF_ONES + s * 16384 + r * 128 + ((r & 7) << 4) + 2) =
            __bfloat16_as_ushort(__float2bfloat16_rn(d));
        fence_proxy_async();
        named_barrier_sync(1, 128);
        if (tid == 0) mbar_arrive(&dready[s]);
      }
    }
    // the accumulated state -> fp32 partial
    mbar_wait_warp(done, 0);
    tc_fence_after();
    float* out = P.partial + ((size_t)bh * P.nsplit + split) * PART_FLOATS;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
    if (warp < 2) {
#pragma unroll 1
      for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld32(lane_addr + c0, v);
        tmem_wait_ld();
#pragma unroll
        for (int c = 0; c < 32; ++c) out[r * 66 + c0 + c] = __uint_as_float(v[c]);
      }
      uint32_t w[8];
      tmem_ld8(lane_addr + 64, w);
      tmem_wait_ld();
      out[r * 66 + 64] = __uint_as_float(w[0]);
      out[r * 66 + 65] = __uint_as_float(w[1]);
    } else if (warp == 2) {
#pragma unroll 1
      for (int c0 = 0; c0 < 80; c0 += 16) {
        uint32_t v[16];
        tmem_ld16(lane_addr + c0, v);              // (warp-collective load; only lane 0's row -- row 64 -- is kept)
        tmem_wait_ld();
        if (lane == 0) {
#pragma unroll
          for (int c = 0; c < 16; ++c) out[64 * 66 + c0 + c] = __uint_as_float(v[c]);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 128);
}

// sum of the partial states of a (b,h) -> bf16 operand panels + fp32 vectors; FIN_SPLIT blocks per (b,h)
constexpr int FIN_SPLIT = 8;
__global__ void __launch_bounds__(256) finalize_kernel(const float* __restrict__ partial, int nsplit,
                                                       uint8_t* __restrict__ final_state) {
  const int bh = blockIdx.x / FIN_SPLIT, part = blockIdx.x % FIN_SPLIT;
  const float* src = partial + (size_t)bh * nsplit * PART_FLOATS;
  uint8_t* dst = final_state + (size_t)bh * FINAL_BYTES;
  float* vec = reinterpret_cast<float*>(dst + FINAL_VEC);
  constexpr int PER = (PART_FLOATS + FIN_SPLIT - 1) / FIN_SPLIT;
  const int hi = min(PART_FLOATS, (part + 1) * PER);
  for (int idx = part * PER + threadIdx.x; idx < hi; idx += 256) {
    float s = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) s += src[(size_t)sp * PART_FLOATS + idx];
    if (idx < 64 * 66) {
      const int y = idx / 66, c = idx % 66;
      if (c < 64) {
        *reinterpret_cast<bf16*>(dst + y * 128 + (((c >> 3) ^ (y & 7)) << 4) + (c & 7) * 2) = __float2bfloat16_rn(s);
      } else if (c == 64) {
        // panel 1 is read as columns 64..79 of the state: row y = [column 64, 0, 0, ...] (element 0 sits in chunk y & 7)
        uint4* row = reinterpret_cast<uint4*>(dst + 8192 + y * 128);
        for (int j = 0; j < 8; ++j)
          row[j] = make_uint4(j == (y & 7) ? (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(s)) : 0u, 0u, 0u, 0u);
        vec[80 + y] = s;
      } else {
        vec[144 + y] = s;
      }
    } else {
      vec[idx - 64 * 66] = s;
    }
  }
}

// exclusive prefix (dir 0) or suffix (dir 1) over the chunks of a (b,h): what a causal scan over chunk c starts from
__global__ void __launch_bounds__(256) prefix_kernel(const float* __restrict__ partial, int nsplit, int dir,
                                                     float* __restrict__ init) {
  const int bh = blockIdx.x;
  const float* src = partial + (size_t)bh * nsplit * PART_FLOATS;
  float* dst = init + (size_t)bh * nsplit * PART_FLOATS;
  for (int idx = threadIdx.x; idx < PART_FLOATS; idx += 256) {
    float s = 0.f;
    for (int j = 0; j < nsplit; ++j) {
      const int c = dir ? nsplit - 1 - j : j;
      dst[(size_t)c * PART_FLOATS + idx] = s;
      s += src[(size_t)c * PART_FLOATS + idx];
    }
  }
}

// ============================================================================================ apply
constexpr int A_STAGES = 3;
constexpr int A_OFF_SB = 0;
constexpr int A_OFF_A = 16384;
constexpr int A_OFF_STG = A_OFF_A + A_STAGES * 16384;
constexpr int A_OFF_MISC = A_OFF_STG + 2 * 16384;
constexpr int A_SMEM = A_OFF_MISC + 2048 + 1024;

struct ApplyFullParams {
  int N, Npad, ntile, group;
  float a;
  const float* dlt;           // (BH, Npad) (DQ)
  float* g;                   // (BH, N) (FWD)
  const uint8_t* final_state; // (BH, FINAL_BYTES)
  unsigned long long dA[A_STAGES][4];   // A tile, K-major
  unsigned long long dSmn[4];           // state read MN-major (contract y; N up to 80)
  unsigned long long dSk[4];            // state read K-major (contract c; N = 64)
};

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2)
applyfull_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ ApplyFullParams P) {
  constexpr int NACC = MODE == FWD ? 80 : 64;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_OFF_MISC);
  uint64_t* a_full = bars;          // [3]
  uint64_t* a_free = bars + 3;      // [3]
  uint64_t* acc_full = bars + 6;    // [2]
  uint64_t* acc_free = bars + 8;    // [2] (one arrival per compute warp)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* vrow = reinterpret_cast<float*>(smem + A_OFF_MISC + 128);   // [80] row 64 of the state
  float* vc64 = vrow + 80;                                          // [64] column 64
  float* vc65 = vc64 + 64;                                          // [64] column 65
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int bh = blockIdx.y;
  const int t0 = blockIdx.x * P.group, t1 = min(P.ntile, t0 + P.group);
  const int nt = t1 - t0;

  if (warp == 4) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  if (tid == 0) {
    for (int s = 0; s < A_STAGES; ++s) mbar_init(&a_full[s], 1), mbar_init(&a_free[s], 1);
    for (int s = 0; s < 2; ++s) mbar_init(&acc_full[s], 1), mbar_init(&acc_free[s], 4);
    fence_barrier_init();
    tma_prefetch_desc(&tmA), tma_prefetch_desc(&tmOut);
  }
  {
    // the (b,h) state: 16 KiB of operand panels as they are, and the vectors
    const uint8_t* src = P.final_state + (size_t)bh * FINAL_BYTES;
    for (int i = tid; i < 16384 / 16; i += THREADS)
      reinterpret_cast<uint4*>(smem + A_OFF_SB)[i] = __ldg(reinterpret_cast<const uint4*>(src) + i);
    for (int i = tid; i < 208; i += THREADS) vrow[i] = __ldg(reinterpret_cast<const float*>(src + FINAL_VEC) + i);
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  if (smem_u32(smem) != FM_SMEM_BASE) {
    if (tid == 0) printf("fastmax: unexpected dynamic smem base %u (p1 apply)\n", smem_u32(smem));
    __trap();
  }

  if (warp == 4) {
    if (lane == 0) {
      for (int i = 0; i < nt; ++i) {
        const int s = i % A_STAGES, tok = (t0 + i) * 128;
        if (i >= A_STAGES) mbar_wait(&a_free[s], ((i / A_STAGES) - 1) & 1, 100 + i);
        mbar_arrive_expect_tx(&a_full[s], 16384);
        for (int h = 0; h < 2; ++h) tma_load_3d(smem + A_OFF_A + s * 16384 + h * 8192, &tmA, &a_full[s], 0, tok + 64 * h, bh);
      }
    }
  } else if (warp == 5) {
    constexpr uint32_t ID = (MODE == FWD || MODE == DV) ? make_idesc_bf16(128, NACC, false, true)
                                                        : make_idesc_bf16(128, 64, false, false);
    for (int i = 0; i < nt; ++i) {
      const int s = i % A_STAGES, ab = i & 1;
      if (lane == 0) {
        mbar_wait(&a_full[s], (i / A_STAGES) & 1, 200 + i);
        if (i >= 2) mbar_wait(&acc_free[ab], ((i >> 1) - 1) & 1, 300 + i);
      }
      __syncwarp();
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_ss(tmem + 128 * ab, P.dA[s][kk], (MODE == FWD || MODE == DV) ? P.dSmn[kk] : P.dSk[kk], ID, kk > 0);
        umma_commit(&a_free[s]);
        umma_commit(&acc_full[ab]);
      }
      __syncwarp();
    }
  } else {
    const int r = tid;
    const uint32_t lane_addr = tmem + ((uint32_t)(32 * warp) << 16);
    const float a = P.a;
    for (int i = 0; i < nt; ++i) {
      const int ab = i & 1;
      const int tile = t0 + i, i_tok = tile * 128 + r;
      const bool valid = i_tok < P.N;
      float row = 0.f;                                      // DQ: dlt of the row; FWD: 1 / denominator
      if (MODE == DQ) row = valid ? P.dlt[(size_t)bh * P.Npad + i_tok] : 0.f;
      mbar_wait_warp(&acc_full[ab], (i >> 1) & 1);
      tc_fence_after();
      uint32_t pk[32];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        uint32_t v[32];
        tmem_ld32(lane_addr + 128 * ab + 32 * hh, v);
        if (MODE == FWD && hh == 0) {
          uint32_t w[8];
          tmem_ld8(lane_addr + 128 * ab + 64, w);
          tmem_wait_ld();
          const float den = fmaf(a, __uint_as_float(w[0]), (float)P.N);     // N + a q.n1
          if (valid) P.g[(size_t)bh * P.N + i_tok] = den;
          row = 1.f / den;
        } else {
          tmem_wait_ld();
        }
        if (hh == 1) {                                      // the accumulator is in registers: hand it back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_free[ab]);
        }
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float x0 = __uint_as_float(v[c]), x1 = __uint_as_float(v[c + 1]);
          const int cc = 32 * hh + c;
          if (MODE == FWD) x0 = fmaf(a, x0, vrow[cc]) * row, x1 = fmaf(a, x1, vrow[cc + 1]) * row;
          else if (MODE == DQ) x0 = a * (x0 - row * vc64[cc]), x1 = a * (x1 - row * vc64[cc + 1]);
          else if (MODE == DV) x0 = fmaf(a, x0, vrow[cc]), x1 = fmaf(a, x1, vrow[cc + 1]);
          else x0 = a * (x0 - vc65[cc]), x1 = a * (x1 - vc65[cc + 1]);
          pk[16 * hh + (c >> 1)] = pack_bf16x2(x0, x1);
        }
      }
      // the store issued from this staging buffer two tiles ago has read it
      if (tid == 0) bulk_wait_group_read1();
      named_barrier_sync(1, 128);
      const uint32_t dst = smem_u32(smem) + A_OFF_STG + ab * 16384 + r * 128;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        sts128(dst + ((uint32_t)(j ^ (r & 7)) << 4), pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
      fence_proxy_async();
      named_barrier_sync(2, 128);
      if (tid == 0) {
        tma_store_3d(&tmOut, smem + A_OFF_STG + ab * 16384, 0, tile * 128, bh);     // rows past N are clipped by the tensor map
        if (tile * 128 + 64 < P.N) tma_store_3d(&tmOut, smem + A_OFF_STG + ab * 16384 + 8192, 0, tile * 128 + 64, bh);
        bulk_commit_group();
      }
    }
    if (tid == 0) bulk_wait_group_read0();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) tmem_dealloc(tmem, 256);
}

}  // namespace p1f

using namespace p1f;

int p1full_max_split() { return MAX_SPLIT; }
size_t p1full_prefix_bytes(int B, int H, int nsplit) { return (size_t)B * H * nsplit * PART_BYTES * 2; }

static int reduce_geometry(const Problem& pb, int* nsplit, int* tiles_per_split) {
  const int BH = pb.B * pb.H, ntile = (pb.N + 127) / 128;
  int ns = (592 + BH - 1) / BH;                       // two CTAs per SM on 148 SMs, twice over
  ns = std::max(1, std::min(std::min(ns, MAX_SPLIT), ntile));
  const int tps = (ntile + ns - 1) / ns;
  *nsplit = (ntile + tps - 1) / tps;                  // every split has at least one tile
  *tiles_per_split = tps;
  return ntile;
}

// bytes of the region a (b,h)-batched state needs: final states, then the partials
size_t p1full_state_bytes(int B, int H) { return (size_t)B * H * (FINAL_BYTES + (size_t)MAX_SPLIT * PART_BYTES); }

bool p1full_usable(const Problem& pb) {
  return pb.p == 1 && !pb.causal && pb.D == 64 && pb.dtype == FASTMAX_DTYPE_BF16 && pb.N >= 128 &&
         !pb.sp_init_fwd && !pb.sp_init_rev;
}

static int launch_reduce(const Problem& pb, const void* x, const void* w, const float* dlt, ReduceParams& P);

// per-chunk states of (x, w [, dlt]) for a sequence cut into `nsplit` chunks of `tiles_per_split` tiles, and their
// exclusive prefix (dir 0) / suffix (dir 1) -> *init: (BH, nsplit, PART_FLOATS) inside `region` (p1full_prefix_bytes)
int p1full_prefix_states(const Problem& pb, int Npad, const void* x, const void* w, const float* dlt, int nsplit,
                         int tiles_per_split, int dir, void* region, const float** init) {
  const int BH = pb.B * pb.H;
  ReduceParams P{};
  P.N = pb.N, P.Npad = Npad, P.dlt = dlt, P.ntile = (pb.N + 127) / 128;
  P.nsplit = nsplit, P.tiles_per_split = tiles_per_split;
  P.partial = (float*)region;
  float* out = P.partial + (size_t)BH * nsplit * PART_FLOATS;
  int rc;
  if ((rc = launch_reduce(pb, x, w, dlt, P))) return rc;
  FM_LAUNCH("p1_prefix_states", pb.stream, prefix_kernel<<<BH, 256, 0, pb.stream>>>(P.partial, nsplit, dir, out));
  *init = out;
  return FASTMAX_OK;
}

// state of (x, w [, dlt]) -> final states at `region`
int p1full_reduce(const Problem& pb, int Npad, const void* x, const void* w, const float* dlt, void* region) {
  const int BH = pb.B * pb.H;
  ReduceParams P{};
  P.N = pb.N, P.Npad = Npad, P.dlt = dlt;
  P.ntile = reduce_geometry(pb, &P.nsplit, &P.tiles_per_split);
  uint8_t* fin = (uint8_t*)region;
  P.partial = (float*)(fin + (size_t)BH * FINAL_BYTES);
  int rc;
  if ((rc = launch_reduce(pb, x, w, dlt, P))) return rc;
  FM_LAUNCH("p1_full_finalize", pb.stream, finalize_kernel<<<BH * FIN_SPLIT, 256, 0, pb.stream>>>(P.partial, P.nsplit, fin));
  return FASTMAX_OK;
}

static int launch_reduce(const Problem& pb, const void* x, const void* w, const float* dlt, ReduceParams& P) {
  const int BH = pb.B * pb.H;
  CUtensorMap tX, tW;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tX, x, 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&tW, w, 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 64))) return rc;
  const uint32_t base = sm100::FM_SMEM_BASE;
  for (int s = 0; s < R_STAGES; ++s)
    for (int kk = 0; kk < 8; ++kk) {
      const uint32_t xs = s * 32768, ws = s * 32768 + 16384, es = R_OFF_ONES + s * 16384;
      P.dX[s][kk] = sm100::desc_mnmajor_sw128_const(base + xs, es - xs) + kk * 128;
      P.dW[s][kk] = sm100::desc_mnmajor_sw128_const(base + ws, es - ws) + kk * 128;
    }
  dim3 grid(P.nsplit, BH);
  if (dlt) {
    FM_CUDA(cudaFuncSetAttribute(reduce_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM));
    FM_LAUNCH("p1_full_reduce_rev", pb.stream, reduce_kernel<true><<<grid, THREADS, R_SMEM, pb.stream>>>(tX, tW, P));
  } else {
    FM_CUDA(cudaFuncSetAttribute(reduce_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, R_SMEM));
    FM_LAUNCH("p1_full_reduce", pb.stream, reduce_kernel<false><<<grid, THREADS, R_SMEM, pb.stream>>>(tX, tW, P));
  }
  return FASTMAX_OK;
}

template <int MODE>
static int launch_applyfull(const Problem& pb, int Npad, const void* A, void* out, const float* dlt, float* g,
                            const void* region, const char* name) {
  const int BH = pb.B * pb.H;
  CUtensorMap tA, tO;
  int rc;
  if ((rc = make_tmap_bf16_3d(&tA, A, 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 64))) return rc;
  if ((rc = make_tmap_bf16_3d(&tO, out, 64, pb.N, BH, 128, (uint64_t)pb.N * 128, 64, 64))) return rc;
  ApplyFullParams P{};
  P.N = pb.N, P.Npad = Npad, P.ntile = (pb.N + 127) / 128, P.a = pb.inv_c, P.dlt = dlt, P.g = g;
  P.final_state = (const uint8_t*)region;
  // tiles per CTA: enough CTAs for four waves of two per SM, at most 16 tiles so the state read stays a small share
  P.group = (int)std::max<long long>(1, std::min<long long>(16, (long long)P.ntile * BH / 1184));
  const uint32_t base = sm100::FM_SMEM_BASE;
  for (int kk = 0; kk < 4; ++kk) {
    for (int s = 0; s < A_STAGES; ++s) P.dA[s][kk] = sm100::desc_kmajor_sw128_const(base + A_OFF_A + s * 16384) + kk * 2;
    P.dSmn[kk] = sm100::desc_mnmajor_sw128_const(base + A_OFF_SB, 8192) + kk * 128;
    P.dSk[kk] = sm100::desc_kmajor_sw128_const(base + A_OFF_SB) + kk * 2;
  }
  auto kern = applyfull_kernel<MODE>;
  FM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM));
  dim3 grid((P.ntile + P.group - 1) / P.group, BH);
  FM_LAUNCH(name, pb.stream, kern<<<grid, THREADS, A_SMEM, pb.stream>>>(tA, tO, P));
  return FASTMAX_OK;
}

int p1full_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g, void* region) {
  int rc;
  if ((rc = p1full_reduce(pb, 0, k, v, nullptr, region))) return rc;
  return launch_applyfull<FWD>(pb, 0, q, o, nullptr, g, region, "p1_full_apply_fwd");
}

// u = do/g (bf16), dlt = (do.o)/g (fp32, (BH, Npad)) from the prep kernel; region_s: the (k, v) state (rebuilt here
// unless the caller kept it from the forward pass), region_r: scratch for the (q, u) state
int p1full_bwd(const Problem& pb, int Npad, const void* q, const void* k, const void* v, const void* u, const float* dlt,
               void* dq, void* dk, void* dv, void* region_s, bool have_s, void* region_r) {
  int rc;
  if (!have_s && (rc = p1full_reduce(pb, Npad, k, v, nullptr, region_s))) return rc;
  if ((rc = launch_applyfull<DQ>(pb, Npad, u, dq, dlt, nullptr, region_s, "p1_full_apply_dq"))) return rc;
  if ((rc = p1full_reduce(pb, Npad, q, u, dlt, region_r))) return rc;
  if ((rc = launch_applyfull<DV>(pb, Npad, k, dv, nullptr, nullptr, region_r, "p1_full_apply_dv"))) return rc;
  return launch_applyfull<DK>(pb, Npad, v, dk, nullptr, nullptr, region_r, "p1_full_apply_dk");
}

}  // namespace fm
