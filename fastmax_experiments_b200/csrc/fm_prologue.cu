// fm_prologue.cu -- fused q/k/v assembly in front of the FastMax op (HBM-bound, one pass).
//
// Replaces lit_gpt/model.py:392-425 between the qkv projection and the attention call:
//   qkv.view(B, T, G, q_per_kv + 2, hs).permute(0, 2, 3, 1, 4)          :398-401
//   q, k, v = qkv.split((q_per_kv, 1, 1), dim=2)                          :404
//   [k, v expand over the group: NOT materialised here -- fastmax_*_gqa reads the (b, kv head) rows]   :408-416
//   reshape to (B, heads, T, hs)                                          :418-420
//   apply_rope on the first rope_n features of q and k, concat the rest  :422-425, :702-708
// The reference does this with four full-size copies (permute+reshape of q, k, v, the RoPE products and the
// concatenations); here every 16-byte chunk of the projection output is read once and written once, already in the
// (B, heads, T, hs) layout the kernels consume.  The backward kernel is the transpose (un-rotate, gather into
// dqkv), so autograd through the prologue is two launches.
//
// Arithmetic follows torch's eager rounding so the result is bit-exact against the reference on the same inputs:
//   same dtype for x and cos/sin (bf16 model under bf16-true, fp32 model): every elementwise op rounds to that
//   dtype -- roped = rn(rn(x * cos) + rn(rot * sin)), no fused multiply-add;
//   bf16 x with fp32 cos/sin: type promotion, the three ops run in fp32 and the result is rounded once.
#include "fm_common.cuh"

namespace fm {
namespace prologue {
typedef __nv_bfloat16 bf16;

struct Params {
  int B, T, G, qpk, hs, rope_n;
  const void* qkv;      // forward: input (B, T, G, qpk + 2, hs); backward: output dqkv
  const void* cos;      // (T, rope_n)
  const void* sin;
  void* q;              // forward: output (B, G * qpk, T, hs); backward: input dq
  void* k;              // (B, G, T, hs)
  void* v;
};

template <typename T> struct Chunk;                       // one 16-byte chunk = EL elements
template <> struct Chunk<bf16> { static constexpr int EL = 8; };
template <> struct Chunk<float> { static constexpr int EL = 4; };

template <typename T> __device__ __forceinline__ void load_chunk(const T* p, float* x) {
  const uint4 w = *reinterpret_cast<const uint4*>(p);
  if constexpr (sizeof(T) == 2) {
    const uint32_t u[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) x[2 * j] = __uint_as_float(u[j] << 16), x[2 * j + 1] = __uint_as_float(u[j] & 0xffff0000u);
  } else {
    x[0] = __uint_as_float(w.x), x[1] = __uint_as_float(w.y), x[2] = __uint_as_float(w.z), x[3] = __uint_as_float(w.w);
  }
}
template <typename T> __device__ __forceinline__ void store_chunk(T* p, const float* x) {
  uint4 w;
  if constexpr (sizeof(T) == 2) {
    uint32_t u[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      u[j] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x[2 * j])) |
             ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x[2 * j + 1])) << 16);
    w = make_uint4(u[0], u[1], u[2], u[3]);
  } else {
    w = make_uint4(__float_as_uint(x[0]), __float_as_uint(x[1]), __float_as_uint(x[2]), __float_as_uint(x[3]));
  }
  *reinterpret_cast<uint4*>(p) = w;
}
// rounding of an intermediate: to the tensors' dtype when cos/sin share it (torch eager), none under promotion
// (autograd casts each branch's gradient back to the bf16 input's dtype, so backward always rounds per term)
template <typename T, typename CS, bool BWD> __device__ __forceinline__ float rnd(float x) {
  if constexpr (sizeof(T) == 2 && (sizeof(CS) == 2 || BWD)) return __bfloat162float(__float2bfloat16_rn(x));
  else return x;
}

// BWD = false:  y_i = x_i c_i - x_{i+h} s_i (i < h),  y_i = x_i c_i + x_{i-h} s_i (h <= i < 2h)      apply_rope :702-708
// BWD = true :  dx_i = dy_i c_i + dy_{i+h} s_{i+h} (i < h),  dx_i = dy_i c_i - dy_{i-h} s_{i-h}        (its autograd)
//
// One thread moves TWO 16-byte chunks of a head row: inside the rotary part a chunk and its partner h elements
// further (each is the other's "rotated" operand, so nothing is read twice), outside it two neighbouring pass-through
// chunks.  32 bytes in flight per thread and 2048 threads per SM cover the HBM latency-bandwidth product.
template <typename T, typename CS, bool BWD>
__global__ void __launch_bounds__(256) qkv_kernel(const Params P) {
  constexpr int EL = Chunk<T>::EL;
  const int spr = P.hs / EL / 2;                              // two-chunk slots per head row
  const int rope_slots = P.rope_n / EL / 2;                   // ... of which rotary
  const int rows_per_tok = P.G * (P.qpk + 2);
  const size_t total = (size_t)P.B * P.T * rows_per_tok * spr;
  const size_t idx = (size_t)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int s = (int)(idx % spr);
  size_t row = idx / spr;                                     // (b, t, g, j) in the projection's order
  const size_t lin_row = row;
  const int j = (int)(row % (P.qpk + 2));
  row /= (P.qpk + 2);
  const int g = (int)(row % P.G);
  row /= P.G;
  const int t = (int)(row % P.T);
  const int b = (int)(row / P.T);
  T* lin = (T*)P.qkv + lin_row * P.hs;                        // this row inside qkv / dqkv
  T* head;                                                    // ... and inside q / k / v (dq / dk / dv)
  if (j < P.qpk) head = (T*)P.q + (((size_t)b * P.G * P.qpk + (size_t)g * P.qpk + j) * P.T + t) * P.hs;
  else head = (T*)(j == P.qpk ? P.k : P.v) + (((size_t)b * P.G + g) * P.T + t) * P.hs;
  const T* src = BWD ? head : lin;
  T* dst = BWD ? lin : head;
  const int h = P.rope_n / 2;
  const bool rotary = s < rope_slots;
  // element offsets of the two chunks
  const int e0 = rotary ? s * EL : P.rope_n + (s - rope_slots) * 2 * EL;
  const int e1 = rotary ? e0 + h : e0 + EL;
  float x0[EL], x1[EL];
  load_chunk<T>(src + e0, x0);
  load_chunk<T>(src + e1, x1);
  if (rotary && j <= P.qpk) {                                 // q and k rows only
    float c0[EL], c1[EL], s0[EL], s1[EL];
    const CS* ct = (const CS*)P.cos + (size_t)t * P.rope_n;
    const CS* st = (const CS*)P.sin + (size_t)t * P.rope_n;
    // the caller's tables are read as given (build_rope_cache happens to repeat the angles over the two halves, :697)
    if constexpr (sizeof(CS) == sizeof(T)) {
      load_chunk<CS>(ct + e0, c0), load_chunk<CS>(ct + e1, c1);
      load_chunk<CS>(st + e0, s0), load_chunk<CS>(st + e1, s1);
    } else {                                                  // fp32 tables with bf16 tensors: two chunks of four
      load_chunk<CS>(ct + e0, c0), load_chunk<CS>(ct + e0 + 4, c0 + 4), load_chunk<CS>(ct + e1, c1), load_chunk<CS>(ct + e1 + 4, c1 + 4);
      load_chunk<CS>(st + e0, s0), load_chunk<CS>(st + e0 + 4, s0 + 4), load_chunk<CS>(st + e1, s1), load_chunk<CS>(st + e1 + 4, s1 + 4);
    }
#pragma unroll
    for (int i = 0; i < EL; ++i) {
      const float lo = x0[i], hi = x1[i];
      if constexpr (!BWD) {
        // rotated = (-x2, x1):  y_lo = lo c_lo - hi s_lo ,  y_hi = hi c_hi + lo s_hi
        x0[i] = __fadd_rn(rnd<T, CS, BWD>(__fmul_rn(lo, c0[i])), -rnd<T, CS, BWD>(__fmul_rn(hi, s0[i])));
        x1[i] = __fadd_rn(rnd<T, CS, BWD>(__fmul_rn(hi, c1[i])), rnd<T, CS, BWD>(__fmul_rn(lo, s1[i])));
      } else {
        // dx_lo = dy_lo c_lo + dy_hi s_hi ,  dx_hi = dy_hi c_hi - dy_lo s_lo
        x0[i] = __fadd_rn(rnd<T, CS, BWD>(__fmul_rn(lo, c0[i])), rnd<T, CS, BWD>(__fmul_rn(hi, s1[i])));
        x1[i] = __fadd_rn(rnd<T, CS, BWD>(__fmul_rn(hi, c1[i])), -rnd<T, CS, BWD>(__fmul_rn(lo, s0[i])));
      }
    }
  }
  store_chunk<T>(dst + e0, x0);
  store_chunk<T>(dst + e1, x1);
}

}  // namespace prologue

// dtype_cs: dtype of the cos / sin tables (FASTMAX_DTYPE_*); backward = 1 runs the transpose (q, k, v hold the
// incoming gradients, qkv receives dqkv)
int qkv_prologue(const void* qkv, const void* cos, const void* sin, void* q, void* k, void* v, int B, int T, int G,
                 int qpk, int hs, int rope_n, int dtype, int dtype_cs, int backward, cudaStream_t stream) {
  using namespace prologue;
  const int el = dtype == FASTMAX_DTYPE_BF16 ? 8 : 4;
  if (B <= 0 || T <= 0 || G <= 0 || qpk <= 0 || hs <= 0 || rope_n < 0 || rope_n > hs) return FASTMAX_ERR_BAD_ARG;
  // whole 16-byte chunks, moved in pairs: the rotary half and the pass-through remainder are whole chunk pairs
  if (hs % (2 * el) || (rope_n / 2) % el || rope_n % 2 || (hs - rope_n) % (2 * el)) return FASTMAX_ERR_UNSUPPORTED;
  if (dtype == FASTMAX_DTYPE_F32 && dtype_cs != FASTMAX_DTYPE_F32) return FASTMAX_ERR_UNSUPPORTED;
  Params P{B, T, G, qpk, hs, rope_n, qkv, cos, sin, q, k, v};
  const size_t total = (size_t)B * T * G * (qpk + 2) * (hs / el / 2);
  const unsigned grid = (unsigned)((total + 255) / 256);
#define FM_PRO(TT, CC)                                                                                         \
  do {                                                                                                         \
    if (backward) FM_LAUNCH("qkv_prologue_bwd", stream, (qkv_kernel<TT, CC, true><<<grid, 256, 0, stream>>>(P))); \
    else FM_LAUNCH("qkv_prologue", stream, (qkv_kernel<TT, CC, false><<<grid, 256, 0, stream>>>(P)));          \
  } while (0)
  if (dtype == FASTMAX_DTYPE_BF16 && dtype_cs == FASTMAX_DTYPE_BF16) FM_PRO(bf16, bf16);
  else if (dtype == FASTMAX_DTYPE_BF16) FM_PRO(bf16, float);
  else FM_PRO(float, float);
#undef FM_PRO
  return FASTMAX_OK;
}

}  // namespace fm
