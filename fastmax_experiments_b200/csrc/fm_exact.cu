// fm_exact.cu -- FASTMAX_ALGO_EXACT: fp32 CUDA-core tiles of the dense FastMax formula.
//
//   o_i = sum_{j (<= i)} f(q_i.k_j / c) v_j / g_i ,   g_i = sum_j f(q_i.k_j / c)
//
// which SURVEY.md section 4 verified to be an exact restatement of the reference's factorised
// forward (fastmax.py:185-322) and backward (fastmax.py:384-691).  All arithmetic is fp32 FFMA
// (inputs may be fp32 or bf16), so this path carries the fp32 parity bar (<= 1e-4) and serves
// every shape the tcgen05 moment-scan path does not cover (any D <= 128, p in {1,2}, causal or
// not, ragged N).  Cost is O(N^2 D): the dispatcher only picks it for small problems or
// shapes outside the fast path.
//
// Tiling: 64 query rows x 64 key rows per step, 256 threads as 16x16; each thread owns a 4x4
// block of the score tile and a 4 x (D/16) block of the output tile.  Bounds: HBM traffic is
// negligible, the kernel is FFMA-issue bound (~1 LDS per 2 FFMA).
#include "fm_common.cuh"

namespace fm {
namespace {

constexpr int BM = 64;       // tile rows (queries)
constexpr int BN = 64;       // tile cols (keys)
constexpr int NT = 256;      // threads
constexpr int PP = BN + 1;   // pitch of the score tiles

template <typename T>
__device__ __forceinline__ void load_tile(float* dst, int pitch, int dmax, const T* src, int row0,
                                          int N, int D, float rowscale_default,
                                          const float* rowscale /* nullable, indexed by global row */) {
  // dst[r][d] = src[(row0+r)*D + d] * scale(row) ; zero outside (row >= N or d >= D)
  for (int idx = threadIdx.x; idx < BM * dmax; idx += NT) {
    int r = idx / dmax, d = idx - r * dmax;
    int row = row0 + r;
    float x = 0.f;
    if (row < N && d < D) {
      x = to_f32<T>(src[(size_t)row * D + d]);
      x *= rowscale ? rowscale[row] : rowscale_default;
    }
    dst[r * pitch + d] = x;
  }
}

// ---------------------------------------------------------------------------------------------
template <int DMAX, typename T>
__global__ void __launch_bounds__(NT) exact_fwd_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                        const T* __restrict__ v, T* __restrict__ o,
                                                        float* __restrict__ g, int N, int D, int p,
                                                        int causal, float inv_c, int seg_len,
                                                        float* __restrict__ o_part) {
  // seg_len > 0 (causal only): keys are restricted to the query's own segment and the
  // unnormalised numerator / partial normaliser go to o_part / g (fp32) for the moment path.
  constexpr int DP = DMAX + 1;
  constexpr int NC = DMAX / 16;
  extern __shared__ float smem[];
  float* Qs = smem;
  float* Ks = Qs + BM * DP;
  float* Vs = Ks + BN * DP;
  float* Ps = Vs + BN * DP;

  const int bh = blockIdx.y;
  const int row0 = blockIdx.x * BM;
  const size_t base = (size_t)bh * N * D;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;

  load_tile<T>(Qs, DP, DMAX, q + base, row0, N, D, 1.f, nullptr);

  float acc[4][NC];
  float gs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[r][c] = 0.f;

  const int jend = causal ? min(row0 + BM, N) : N;
  const int jbeg = seg_len > 0 ? (row0 / seg_len) * seg_len : 0;
  for (int col0 = jbeg; col0 < jend; col0 += BN) {
    __syncthreads();
    load_tile<T>(Ks, DP, DMAX, k + base, col0, N, D, 1.f, nullptr);
    load_tile<T>(Vs, DP, DMAX, v + base, col0, N, D, 1.f, nullptr);
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[r][c] = 0.f;
    for (int d = 0; d < D; ++d) {
      float qv[4], kv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) qv[r] = Qs[(4 * ty + r) * DP + d];
#pragma unroll
      for (int c = 0; c < 4; ++c) kv[c] = Ks[(tx + 16 * c) * DP + d];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) s[r][c] = fmaf(qv[r], kv[c], s[r][c]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int i = row0 + 4 * ty + r, j = col0 + tx + 16 * c;
        bool valid = (j < N) && (!causal || j <= i);
        Ps[(4 * ty + r) * PP + tx + 16 * c] = valid ? poly_f(s[r][c] * inv_c, p) : 0.f;
      }
    __syncthreads();
    for (int jj = 0; jj < BN; ++jj) {
      float pv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        pv[r] = Ps[(4 * ty + r) * PP + jj];
        gs[r] += pv[r];
      }
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float vv = Vs[jj * DP + tx + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][c] = fmaf(pv[r], vv, acc[r][c]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int i = row0 + 4 * ty + r;
    if (i >= N) continue;
    float inv = 1.f / gs[r];
    if (tx == 0) g[(size_t)bh * N + i] = gs[r];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      int d = tx + 16 * c;
      if (d < D) {
        if (o_part) o_part[base + (size_t)i * D + d] = acc[r][c];
        else o[base + (size_t)i * D + d] = from_f32<T>(acc[r][c] * inv);
      }
    }
  }
}

// rows of the tile: inv_g[r] = 1/g_i, delta[r] = (do_i . o_i)/g_i   (256 threads, 4 per row)
template <typename T>
__device__ __forceinline__ void row_stats(float* inv_g, float* delta, const T* d_o, const T* o,
                                          const float* g, int row0, int N, int D) {
  int r = threadIdx.x >> 2, part = threadIdx.x & 3;
  int row = row0 + r;
  float acc = 0.f, ig = 0.f;
  if (row < N) {
    ig = 1.f / g[row];
    for (int d = part; d < D; d += 4)
      acc = fmaf(to_f32<T>(d_o[(size_t)row * D + d]), to_f32<T>(o[(size_t)row * D + d]), acc);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  if (part == 0) {
    inv_g[r] = ig;
    delta[r] = acc * ig;
  }
}

// ---------------------------------------------------------------------------------------------
// dq_i = (1/c) sum_j f'(s_ij) (u_i.v_j - delta_i) k_j ,  u_i = do_i/g_i
template <int DMAX, typename T>
__global__ void __launch_bounds__(NT) exact_bwd_dq_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                           const T* __restrict__ v, const T* __restrict__ o,
                                                           const float* __restrict__ g, const T* __restrict__ d_o,
                                                           T* __restrict__ dq, int N, int D, int p, int causal,
                                                           float inv_c, int seg_len, float* __restrict__ dq_part) {
  constexpr int DP = DMAX + 1;
  constexpr int NC = DMAX / 16;
  extern __shared__ float smem[];
  float* Qs = smem;
  float* Us = Qs + BM * DP;
  float* Ks = Us + BM * DP;
  float* Vs = Ks + BN * DP;
  float* Es = Vs + BN * DP;
  float* inv_g = Es + BM * PP;
  float* delta = inv_g + BM;

  const int bh = blockIdx.y;
  const int row0 = blockIdx.x * BM;
  const size_t base = (size_t)bh * N * D;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;

  row_stats<T>(inv_g, delta, d_o + base, o + base, g + (size_t)bh * N, row0, N, D);
  __syncthreads();
  load_tile<T>(Qs, DP, DMAX, q + base, row0, N, D, 1.f, nullptr);
  // Us[r][d] = do * (1/g)
  for (int idx = threadIdx.x; idx < BM * DMAX; idx += NT) {
    int r = idx / DMAX, d = idx - r * DMAX, row = row0 + r;
    Us[r * DP + d] = (row < N && d < D) ? to_f32<T>(d_o[base + (size_t)row * D + d]) * inv_g[r] : 0.f;
  }

  float acc[4][NC];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) acc[r][c] = 0.f;

  const int jend = causal ? min(row0 + BM, N) : N;
  const int jbeg = seg_len > 0 ? (row0 / seg_len) * seg_len : 0;
  for (int col0 = jbeg; col0 < jend; col0 += BN) {
    __syncthreads();
    load_tile<T>(Ks, DP, DMAX, k + base, col0, N, D, 1.f, nullptr);
    load_tile<T>(Vs, DP, DMAX, v + base, col0, N, D, 1.f, nullptr);
    __syncthreads();
    float s[4][4], dp[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[r][c] = 0.f, dp[r][c] = 0.f;
    for (int d = 0; d < D; ++d) {
      float qv[4], uv[4], kv[4], vv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) qv[r] = Qs[(4 * ty + r) * DP + d], uv[r] = Us[(4 * ty + r) * DP + d];
#pragma unroll
      for (int c = 0; c < 4; ++c) kv[c] = Ks[(tx + 16 * c) * DP + d], vv[c] = Vs[(tx + 16 * c) * DP + d];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          s[r][c] = fmaf(qv[r], kv[c], s[r][c]);
          dp[r][c] = fmaf(uv[r], vv[c], dp[r][c]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int i = row0 + 4 * ty + r, j = col0 + tx + 16 * c;
        bool valid = (j < N) && (!causal || j <= i);
        Es[(4 * ty + r) * PP + tx + 16 * c] =
            valid ? poly_df(s[r][c] * inv_c, p) * (dp[r][c] - delta[4 * ty + r]) : 0.f;
      }
    __syncthreads();
    for (int jj = 0; jj < BN; ++jj) {
      float ev[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) ev[r] = Es[(4 * ty + r) * PP + jj];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float kv = Ks[jj * DP + tx + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r) acc[r][c] = fmaf(ev[r], kv, acc[r][c]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int i = row0 + 4 * ty + r;
    if (i >= N) continue;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      int d = tx + 16 * c;
      if (d < D) {
        if (dq_part) dq_part[base + (size_t)i * D + d] = acc[r][c];   // unscaled; the apply kernel folds 1/c
        else dq[base + (size_t)i * D + d] = from_f32<T>(acc[r][c] * inv_c);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// dv_j = sum_{i>=j} f(s_ij) u_i ;  dk_j = (1/c) sum_{i>=j} f'(s_ij)(u_i.v_j - delta_i) q_i
template <int DMAX, typename T>
__global__ void __launch_bounds__(NT) exact_bwd_dkdv_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                             const T* __restrict__ v, const T* __restrict__ o,
                                                             const float* __restrict__ g, const T* __restrict__ d_o,
                                                             T* __restrict__ dk, T* __restrict__ dv, int N, int D,
                                                             int p, int causal, float inv_c, int seg_len,
                                                             float* __restrict__ dk_part, float* __restrict__ dv_part) {
  constexpr int DP = DMAX + 1;
  constexpr int NC = DMAX / 16;
  extern __shared__ float smem[];
  float* Qs = smem;
  float* Us = Qs + BM * DP;
  float* Ks = Us + BM * DP;
  float* Vs = Ks + BN * DP;
  float* Ps = Vs + BN * DP;
  float* Es = Ps + BM * PP;
  float* inv_g = Es + BM * PP;
  float* delta = inv_g + BM;

  const int bh = blockIdx.y;
  const int col0 = blockIdx.x * BN;  // key tile owned by this CTA
  const size_t base = (size_t)bh * N * D;
  const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;

  load_tile<T>(Ks, DP, DMAX, k + base, col0, N, D, 1.f, nullptr);
  load_tile<T>(Vs, DP, DMAX, v + base, col0, N, D, 1.f, nullptr);

  float ak[4][NC], av[4][NC];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < NC; ++c) ak[r][c] = 0.f, av[r][c] = 0.f;

  const int iend = seg_len > 0 ? min(N, (col0 / seg_len + 1) * seg_len) : N;
  for (int row0 = causal ? col0 : 0; row0 < iend; row0 += BM) {
    __syncthreads();
    row_stats<T>(inv_g, delta, d_o + base, o + base, g + (size_t)bh * N, row0, N, D);
    load_tile<T>(Qs, DP, DMAX, q + base, row0, N, D, 1.f, nullptr);
    __syncthreads();
    for (int idx = threadIdx.x; idx < BM * DMAX; idx += NT) {
      int r = idx / DMAX, d = idx - r * DMAX, row = row0 + r;
      Us[r * DP + d] = (row < N && d < D) ? to_f32<T>(d_o[base + (size_t)row * D + d]) * inv_g[r] : 0.f;
    }
    __syncthreads();
    float s[4][4], dp[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[r][c] = 0.f, dp[r][c] = 0.f;
    for (int d = 0; d < D; ++d) {
      float qv[4], uv[4], kv[4], vv[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) qv[r] = Qs[(4 * ty + r) * DP + d], uv[r] = Us[(4 * ty + r) * DP + d];
#pragma unroll
      for (int c = 0; c < 4; ++c) kv[c] = Ks[(tx + 16 * c) * DP + d], vv[c] = Vs[(tx + 16 * c) * DP + d];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          s[r][c] = fmaf(qv[r], kv[c], s[r][c]);
          dp[r][c] = fmaf(uv[r], vv[c], dp[r][c]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        int i = row0 + 4 * ty + r, j = col0 + tx + 16 * c;
        bool valid = (i < N) && (j < N) && (!causal || j <= i);
        float x = s[r][c] * inv_c;
        Ps[(4 * ty + r) * PP + tx + 16 * c] = valid ? poly_f(x, p) : 0.f;
        Es[(4 * ty + r) * PP + tx + 16 * c] = valid ? poly_df(x, p) * (dp[r][c] - delta[4 * ty + r]) : 0.f;
      }
    __syncthreads();
    // output rows are key rows j = 4*ty + r ; contraction over the 64 query rows of the tile
    for (int ii = 0; ii < BM; ++ii) {
      float pv[4], ev[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) pv[r] = Ps[ii * PP + 4 * ty + r], ev[r] = Es[ii * PP + 4 * ty + r];
#pragma unroll
      for (int c = 0; c < NC; ++c) {
        float uv = Us[ii * DP + tx + 16 * c];
        float qv = Qs[ii * DP + tx + 16 * c];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          av[r][c] = fmaf(pv[r], uv, av[r][c]);
          ak[r][c] = fmaf(ev[r], qv, ak[r][c]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    int j = col0 + 4 * ty + r;
    if (j >= N) continue;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      int d = tx + 16 * c;
      if (d < D) {
        if (dk_part) {
          dv_part[base + (size_t)j * D + d] = av[r][c];
          dk_part[base + (size_t)j * D + d] = ak[r][c];               // unscaled
        } else {
          dv[base + (size_t)j * D + d] = from_f32<T>(av[r][c]);
          dk[base + (size_t)j * D + d] = from_f32<T>(ak[r][c] * inv_c);
        }
      }
    }
  }
}

template <int DMAX> constexpr size_t fwd_smem() { return sizeof(float) * (3 * 64 * (DMAX + 1) + 64 * PP); }
template <int DMAX> constexpr size_t dq_smem() { return sizeof(float) * (4 * 64 * (DMAX + 1) + 64 * PP + 128); }
template <int DMAX> constexpr size_t dkdv_smem() { return sizeof(float) * (4 * 64 * (DMAX + 1) + 2 * 64 * PP + 128); }

template <int DMAX, typename T>
int launch_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g, int seg_len,
               float* o_part) {
  auto kern = exact_fwd_kernel<DMAX, T>;
  FM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<DMAX>()));
  dim3 grid((pb.N + BM - 1) / BM, pb.B * pb.H);
  FM_LAUNCH("exact_fwd", pb.stream,
            kern<<<grid, NT, fwd_smem<DMAX>(), pb.stream>>>((const T*)q, (const T*)k, (const T*)v, (T*)o, g, pb.N,
                                                             pb.D, pb.p, pb.causal, pb.inv_c, seg_len, o_part));
  return FASTMAX_OK;
}

template <int DMAX, typename T>
int launch_bwd(const Problem& pb, const void* q, const void* k, const void* v, const void* o, const float* g,
               const void* d_o, void* dq, void* dk, void* dv, int seg_len, float* dq_part, float* dk_part,
               float* dv_part) {
  auto k1 = exact_bwd_dq_kernel<DMAX, T>;
  auto k2 = exact_bwd_dkdv_kernel<DMAX, T>;
  FM_CUDA(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dq_smem<DMAX>()));
  FM_CUDA(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dkdv_smem<DMAX>()));
  dim3 grid((pb.N + BM - 1) / BM, pb.B * pb.H);
  FM_LAUNCH("exact_bwd_dq", pb.stream,
            k1<<<grid, NT, dq_smem<DMAX>(), pb.stream>>>((const T*)q, (const T*)k, (const T*)v, (const T*)o, g,
                                                          (const T*)d_o, (T*)dq, pb.N, pb.D, pb.p, pb.causal,
                                                          pb.inv_c, seg_len, dq_part));
  FM_LAUNCH("exact_bwd_dkdv", pb.stream,
            k2<<<grid, NT, dkdv_smem<DMAX>(), pb.stream>>>((const T*)q, (const T*)k, (const T*)v, (const T*)o, g,
                                                            (const T*)d_o, (T*)dk, (T*)dv, pb.N, pb.D, pb.p,
                                                            pb.causal, pb.inv_c, seg_len, dk_part, dv_part));
  return FASTMAX_OK;
}

}  // namespace

#define FM_DISPATCH_EXACT(FN, ...)                                                        \
  do {                                                                                    \
    const bool bf = pb.dtype == FASTMAX_DTYPE_BF16;                                       \
    if (pb.D <= 32) return bf ? FN<32, __nv_bfloat16>(__VA_ARGS__) : FN<32, float>(__VA_ARGS__);   \
    if (pb.D <= 64) return bf ? FN<64, __nv_bfloat16>(__VA_ARGS__) : FN<64, float>(__VA_ARGS__);   \
    if (pb.D <= 128) return bf ? FN<128, __nv_bfloat16>(__VA_ARGS__) : FN<128, float>(__VA_ARGS__); \
    return FASTMAX_ERR_UNSUPPORTED;                                                       \
  } while (0)

int exact_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g) {
  FM_DISPATCH_EXACT(launch_fwd, pb, q, k, v, o, g, 0, nullptr);
}

int exact_bwd(const Problem& pb, const void* q, const void* k, const void* v, const void* o, const float* g,
              const void* d_o, void* dq, void* dk, void* dv) {
  FM_DISPATCH_EXACT(launch_bwd, pb, q, k, v, o, g, d_o, dq, dk, dv, 0, nullptr, nullptr, nullptr);
}

// Segment-restricted fp32 partials (bring-up / fallback intra-segment path of the moment scan).
int exact_fwd_partial(const Problem& pb, const void* q, const void* k, const void* v, int seg_len, float* o_part,
                      float* g_part) {
  FM_DISPATCH_EXACT(launch_fwd, pb, q, k, v, nullptr, g_part, seg_len, o_part);
}

int exact_bwd_partial(const Problem& pb, const void* q, const void* k, const void* v, const void* o,
                      const float* g, const void* d_o, int seg_len, float* dq_part, float* dk_part,
                      float* dv_part) {
  FM_DISPATCH_EXACT(launch_bwd, pb, q, k, v, o, g, d_o, nullptr, nullptr, nullptr, seg_len, dq_part, dk_part,
                    dv_part);
}

}  // namespace fm
