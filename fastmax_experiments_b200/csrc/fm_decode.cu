// fm_decode.cu -- constant-size recurrent state for inference.
//
// The reference's generation path (generate/base.py:83-88 -> lit_gpt/model.py:427-430, 478-487) keeps a KV
// cache of max_seq_length rows and, for every new token, evaluates the NON-causal FastMax of the one query
// over the whole zero-padded cache: O(max_seq_length * D) memory and O(max_seq_length * D^2) work per token
// and layer.  The factorisation makes the cache unnecessary: everything a query needs from the keys seen so
// far is the moment state
//     S0 = sum_j v'_j ,  S1 = sum_j k_j (x) v'_j ,  S2 = sum_j k_j (x) k_j (x) v'_j ,     v' = [v, 1]
// (fp32, (1 + D + D^2)(D + 1) numbers per (b,h): 1.1 MB at D = 64), updated with the new keys and applied to
// the new queries:
//     [F_i, g_i] = S0 + q_i.S1 / c + (q_i (x) q_i) : S2 / (2 c^2) ,   o_i = F_i / g_i .
// Zero-padded cache rows contribute nothing to S1, S2 and the value part of S0; the only place they show up in
// the reference is the count term of the normaliser, which compute_g_unmasked takes from the number of QUERIES
// (fastmax.py:269-271, quirk Q2) -- so the count term is an argument here (`count_term`), not state.
//
// CUDA-core kernels, fp32 accumulation, fp32 or bf16 q/k/v/o, any D <= 128, p in {1, 2}.  Decode is bound by
// reading the state once per token ((b,h) x 1.1 MB); prefill reuses each state row for 8 queries.
#include "fm_common.cuh"

namespace fm {
namespace dec {

__host__ __device__ inline size_t state_floats(int D, int p) { return (size_t)(1 + D + (p == 2 ? D * D : 0)) * (D + 1); }

// state += sum_t [1, k_t, k_t (x) k_t] (x) [v_t, 1].  grid (1 + D, BH): block x == 0 owns S0 and S1, block
// 1 + x owns the D x (D+1) slice S2[x][.][.].  Threads stride over the (y, a) plane; tokens are staged in
// shared memory 32 at a time.
template <typename T>
__global__ void __launch_bounds__(256) update_kernel(float* __restrict__ state, const T* __restrict__ k,
                                                      const T* __restrict__ v, int Tn, int D, int p) {
  extern __shared__ float sm[];          // [32][D] k  |  [32][D+1] v'
  const int A = D + 1, bh = blockIdx.y, bx = blockIdx.x;
  float* ks = sm;
  float* vs = sm + 32 * D;
  float* st = state + (size_t)bh * state_floats(D, p);
  const T* kb = k + (size_t)bh * Tn * D;
  const T* vb = v + (size_t)bh * Tn * D;
  // this thread's (y, a) elements: at most ceil(D*A / 256) <= 65 for D = 128 -- kept in registers in chunks
  const int plane = D * A;
  const int limit = bx == 0 ? plane + A : plane;
  for (int base = 0; base < limit; base += 256 * 8) {         // (uniform trip count: the loop holds barriers)
    const int e0 = base + threadIdx.x;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int t0 = 0; t0 < Tn; t0 += 32) {
      const int nt = min(32, Tn - t0);
      __syncthreads();
      for (int i = threadIdx.x; i < nt * D; i += 256) {
        const int tt = i / D, d = i % D;
        ks[tt * D + d] = to_f32<T>(kb[(size_t)(t0 + tt) * D + d]);
        vs[tt * A + d] = to_f32<T>(vb[(size_t)(t0 + tt) * D + d]);
      }
      for (int i = threadIdx.x; i < nt; i += 256) vs[i * A + D] = 1.f;
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int e = e0 + 256 * i;
        if (bx == 0) {
          if (e < A) {                                   // S0[a]
            for (int tt = 0; tt < nt; ++tt) acc[i] += vs[tt * A + e];
          } else if (e < plane + A) {                    // S1[y][a]
            const int y = (e - A) / A, a = (e - A) % A;
            for (int tt = 0; tt < nt; ++tt) acc[i] = fmaf(ks[tt * D + y], vs[tt * A + a], acc[i]);
          }
        } else if (e < plane) {                          // S2[x][y][a]
          const int x = bx - 1, y = e / A, a = e % A;
          for (int tt = 0; tt < nt; ++tt) acc[i] = fmaf(ks[tt * D + x] * ks[tt * D + y], vs[tt * A + a], acc[i]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int e = e0 + 256 * i;
      if (bx == 0) {
        if (e < plane + A) st[e] += acc[i];
      } else if (e < plane) {
        st[(size_t)A + plane + (size_t)(bx - 1) * plane + e] += acc[i];
      }
    }
  }
}

// o_i = F_i / g_i for TQ = 8 queries per block.  Thread = (y group of 4, column a): every S2 element is read
// once per block and used for the 8 queries; partial sums over the y groups are combined in shared memory.
constexpr int TQ = 8;
template <typename T>
__global__ void __launch_bounds__(256) attend_kernel(const float* __restrict__ state, const T* __restrict__ q,
                                                      T* __restrict__ o, int Tn, int D, int p, float a1, float count_term) {
  extern __shared__ float sm[];          // [TQ][D] q  |  [4][TQ][A] partial sums
  const int A = D + 1, bh = blockIdx.y, i0 = blockIdx.x * TQ;
  const int nq = min(TQ, Tn - i0);
  float* qs = sm;
  float* red = sm + TQ * D;
  const float* st = state + (size_t)bh * state_floats(D, p);
  const float* S1 = st + A;
  const float* S2 = st + A + (size_t)D * A;
  for (int i = threadIdx.x; i < TQ * D; i += 256) {
    const int tq = i / D, d = i % D;
    qs[i] = tq < nq ? to_f32<T>(q[((size_t)bh * Tn + i0 + tq) * D + d]) * a1 : 0.f;    // q / c
  }
  __syncthreads();
  const int yg = threadIdx.x >> 6, al = threadIdx.x & 63;       // 4 y groups x 64 columns per pass
  for (int a0 = 0; a0 < A; a0 += 64) {
    const int a = a0 + al;
    float acc[TQ];
#pragma unroll
    for (int t = 0; t < TQ; ++t) acc[t] = 0.f;
    if (a < A) {
      for (int y = yg; y < D; y += 4) {
        // first order: q[y] S1[y][a];  second order: q[y]/2 * sum_x q[x] S2[x][y][a]  (S2 symmetric in x, y)
        const float s1 = S1[(size_t)y * A + a];
        float inner[TQ];
#pragma unroll
        for (int t = 0; t < TQ; ++t) inner[t] = s1;
        if (p == 2) {
          for (int x = 0; x < D; ++x) {
            const float s2 = 0.5f * S2[((size_t)x * D + y) * A + a];
#pragma unroll
            for (int t = 0; t < TQ; ++t) inner[t] = fmaf(qs[t * D + x], s2, inner[t]);
          }
        }
#pragma unroll
        for (int t = 0; t < TQ; ++t) acc[t] = fmaf(qs[t * D + y], inner[t], acc[t]);
      }
    }
#pragma unroll
    for (int t = 0; t < TQ; ++t)
      if (a < A) red[(yg * TQ + t) * A + a] = acc[t];
  }
  __syncthreads();
  // combine: F[a] = S0[a] + sum_yg ; g = count_term + sum_yg [column D]
  for (int i = threadIdx.x; i < nq * D; i += 256) {
    const int t = i / D, a = i % D;
    float f = st[a], g = count_term;
#pragma unroll
    for (int gq = 0; gq < 4; ++gq) f += red[(gq * TQ + t) * A + a], g += red[(gq * TQ + t) * A + D];
    o[((size_t)bh * Tn + i0 + t) * D + a] = from_f32<T>(f / g);
  }
}

}  // namespace dec

size_t decode_state_bytes(int B, int H, int D, int p) { return (size_t)B * H * dec::state_floats(D, p) * sizeof(float); }

int decode_update(void* state, const void* k, const void* v, int B, int H, int T, int D, int p, int dtype, cudaStream_t stream) {
  dim3 grid(p == 2 ? 1 + D : 1, B * H);
  const size_t smem = (size_t)(32 * D + 32 * (D + 1)) * sizeof(float);
  if (dtype == FASTMAX_DTYPE_BF16)
    FM_LAUNCH("decode_update", stream, (dec::update_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(
                  (float*)state, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, T, D, p)));
  else
    FM_LAUNCH("decode_update", stream, (dec::update_kernel<float><<<grid, 256, smem, stream>>>(
                  (float*)state, (const float*)k, (const float*)v, T, D, p)));
  return FASTMAX_OK;
}

int decode_attend(const void* state, const void* q, void* o, int B, int H, int T, int D, int p, float inv_c,
                  float count_term, int dtype, cudaStream_t stream) {
  dim3 grid((T + dec::TQ - 1) / dec::TQ, B * H);
  const size_t smem = (size_t)(dec::TQ * D + 4 * dec::TQ * (D + 1)) * sizeof(float);
  if (dtype == FASTMAX_DTYPE_BF16)
    FM_LAUNCH("decode_attend", stream, (dec::attend_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(
                  (const float*)state, (const __nv_bfloat16*)q, (__nv_bfloat16*)o, T, D, p, inv_c, count_term)));
  else
    FM_LAUNCH("decode_attend", stream, (dec::attend_kernel<float><<<grid, 256, smem, stream>>>(
                  (const float*)state, (const float*)q, (float*)o, T, D, p, inv_c, count_term)));
  return FASTMAX_OK;
}

}  // namespace fm
