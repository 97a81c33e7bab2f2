// fm_common.cuh -- shared host/device helpers for libfastmax_b200.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/fastmax_b200.h"

namespace fm {

// ---- per-thread status -------------------------------------------------------------------
struct ThreadStatus {
  char cuda_err[256];
  int launches;
};
ThreadStatus& status();

inline int cuda_fail(cudaError_t e, const char* what) {
  ThreadStatus& s = status();
  snprintf(s.cuda_err, sizeof(s.cuda_err), "%s: %s", what, cudaGetErrorString(e));
  return FASTMAX_ERR_CUDA;
}

#define FM_CUDA(call)                                            \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return fm::cuda_fail(e__, #call);    \
  } while (0)

// Launch wrapper: counts the launch, optionally brackets it with CUDA events on the launching
// stream (fastmax_profile_enable), and turns launch errors into FASTMAX_ERR_CUDA.
void prof_begin(const char* name, cudaStream_t stream);
void prof_end(cudaStream_t stream);

#define FM_LAUNCH(name, stream, ...)                             \
  do {                                                           \
    fm::prof_begin(name, stream);                                \
    __VA_ARGS__;                                                 \
    fm::prof_end(stream);                                        \
    fm::status().launches++;                                     \
    cudaError_t e__ = cudaGetLastError();                        \
    if (e__ != cudaSuccess) return fm::cuda_fail(e__, name);     \
  } while (0)

// ---- problem description passed between the API layer and the kernels ---------------------
struct Problem {
  int B, H, N, D, p, causal, dtype;
  float inv_c;
  cudaStream_t stream;
  // sequence parallelism (fastmax_*_sp): this rank holds tokens [sp_offset, sp_offset + N) of a longer
  // sequence; sp_init_fwd = moment state of all earlier tokens, sp_init_rev = reverse state of all later
  // tokens (one record per (b,h); null = none)
  const void* sp_init_fwd = nullptr;
  const void* sp_init_rev = nullptr;
  long long sp_offset = 0;
  // forward states kept for backward (fastmax_fwd_ex / fastmax_bwd_ex): forward writes its boundary
  // states here instead of into the workspace; backward reads them instead of rebuilding them
  void* saved_states = nullptr;
  // grouped-query attention (lit_gpt/model.py:408-416): k, v are (B, Hkv, N, D) and every group of H / Hkv query
  // heads shares one (b, kv head).  0 = no grouping (Hkv == H).  Moment path only (fastmax_*_gqa).
  int Hkv = 0;
  int kv_heads() const { return Hkv > 0 ? Hkv : H; }
};

// index of the (b, kv head) that query head bh = b * H + h reads: head h belongs to group h / (H / Hkv), the
// order in which model.py:418-420 flattens (B, n_query_groups, q_per_kv, T, hs)
__host__ __device__ __forceinline__ int kv_index(int bh, int H, int Hkv) {
  return Hkv == H ? bh : (bh / H) * Hkv + (bh % H) / (H / Hkv);
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- device helpers -------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T x);
template <> __device__ __forceinline__ float to_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 x) { return __bfloat162float(x); }

template <typename T> __device__ __forceinline__ T from_f32(float x);
template <> __device__ __forceinline__ float from_f32<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float x) { return __float2bfloat16_rn(x); }

// f(s) and f'(s) of the reference (fastmax.py:357-360): p=1: 1+s ; p=2: 1+s+s^2/2
__device__ __forceinline__ float poly_f(float s, int p) { return p == 2 ? fmaf(s, fmaf(0.5f, s, 1.f), 1.f) : 1.f + s; }
__device__ __forceinline__ float poly_df(float s, int p) { return p == 2 ? 1.f + s : 1.f; }

// ---- kernel families (defined in their own .cu files) ---------------------------------------
int exact_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g);
int exact_bwd(const Problem& pb, const void* q, const void* k, const void* v, const void* o,
              const float* g, const void* d_o, void* dq, void* dk, void* dv);

int exact_fwd_partial(const Problem& pb, const void* q, const void* k, const void* v, int seg_len, float* o_part,
                      float* g_part);
int exact_bwd_partial(const Problem& pb, const void* q, const void* k, const void* v, const void* o,
                      const float* g, const void* d_o, int seg_len, float* dq_part, float* dk_part,
                      float* dv_part);

// tcgen05 moment-scan path (fm_moments.cu): bf16, D = 64
int intra_fwd(const Problem& pb, int L, int Npad, const void* q, const void* k, const void* v, float* o_part,
              float* g, const uint8_t* state, void* o_final = nullptr);
int intra_dq(const Problem& pb, int L, int Npad, const void* q, const void* u, const void* k, const void* v,
             const float* dlt, float* dq_part, const uint8_t* state, void* dq_final = nullptr);
int intra_dkdv(const Problem& pb, int L, int Npad, const void* k, const void* v, const void* q, const void* u,
               const float* dlt, float* dk_part, float* dv_part, const uint8_t* rstate, void* dk_final = nullptr,
               void* dv_final = nullptr);
// (p = 1 with *_final given: the kernels also apply the first-order state term and write the final bf16
//  outputs; the *_part arguments are then unused)
// second-generation intra-segment kernels for p = 2 causal (fm_seg.cu): same contract as intra_fwd
bool seg_fwd_usable(const Problem& pb, int L);
int seg_fwd(const Problem& pb, int L, const void* q, const void* k, const void* v, float* o_part, float* g,
            const uint8_t* state);
// fused q/k/v assembly in front of the op (fm_prologue.cu; model.py:392-425) and its transpose
int qkv_prologue(const void* qkv, const void* cos, const void* sin, void* q, void* k, void* v, int B, int T, int G,
                 int qpk, int hs, int rope_n, int dtype, int dtype_cs, int backward, cudaStream_t stream);
// ... and of the dk / dv intra-segment kernel (fm_seg_bwd.cu): same contract as intra_dkdv
bool seg_dkdv_usable(const Problem& pb, int L);
int seg_dkdv(const Problem& pb, int L, int Npad, const void* k, const void* v, const void* q, const void* u,
             const float* dlt, float* dk_part, float* dv_part, const uint8_t* rstate);
bool seg_dq_usable(const Problem& pb, int L);
int seg_dq(const Problem& pb, int L, int Npad, const void* q, const void* u, const void* k, const void* v,
           const float* dlt, float* dq_part, const uint8_t* state);
namespace mom { struct ApplyDesc; }
int launch_apply_gen(const Problem& pb, int BH, int nrec, const char* name, const mom::ApplyDesc& d, int variant);
// first-order causal FastMax as one fused scan per (b,h) (fm_p1scan.cu)
bool p1scan_usable(const Problem& pb);
size_t p1scan_region_bytes(const Problem& pb);      // scratch per state region (0 unless sequences are cut into chunks)
int p1scan_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g, void* region);
int p1scan_bwd(const Problem& pb, int Npad, const void* q, const void* k, const void* v, const void* u, const float* dlt,
               void* dq, void* dk, void* dv, void* region_s, void* region_r);
int p1full_max_split();
size_t p1full_prefix_bytes(int B, int H, int nsplit);
int p1full_prefix_states(const Problem& pb, int Npad, const void* x, const void* w, const float* dlt, int nsplit,
                         int tiles_per_split, int dir, void* region, const float** init);
// first-order non-causal FastMax as streaming reduce + apply passes (fm_p1full.cu); `region*`: p1full_state_bytes each
bool p1full_usable(const Problem& pb);
size_t p1full_state_bytes(int B, int H);
int p1full_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g, void* region);
int p1full_bwd(const Problem& pb, int Npad, const void* q, const void* k, const void* v, const void* u, const float* dlt,
               void* dq, void* dk, void* dv, void* region_s, bool have_s, void* region_r);
// head size 128 (bf16) as dense tcgen05 tiles (fm_dense128.cu)
bool dense128_usable(const Problem& pb);
int dense128_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g);
size_t dense128_bwd_workspace_bytes(int B, int H, int N);     // u (bf16) and dlt (fp32)
// fp32 inputs of head size <= 64 through the same kernels as [hi | lo] bf16 pairs (three-term products)
bool dense_split_usable(const Problem& pb);
size_t dense_split_fwd_workspace_bytes(int B, int H, int N);
size_t dense_split_bwd_workspace_bytes(int B, int H, int N);
int dense_split_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g, void* ws);
int dense_split_bwd(const Problem& pb, const void* q, const void* k, const void* v, const void* o, const float* g,
                    const void* d_o, void* dq, void* dk, void* dv, void* ws);
int dense128_bwd(const Problem& pb, const void* q, const void* k, const void* v, const void* o, const float* g,
                 const void* d_o, void* dq, void* dk, void* dv, void* ws);
unsigned sm100_smem_base();     // FM_SMEM_BASE of fm_sm100.cuh (defined in fm_tmap.cu)
// recurrent inference state (fm_decode.cu)
size_t decode_state_bytes(int B, int H, int D, int p);
int decode_update(void* state, const void* k, const void* v, int B, int H, int T, int D, int p, int dtype, cudaStream_t stream);
int decode_attend(const void* state, const void* q, void* o, int B, int H, int T, int D, int p, float inv_c,
                  float count_term, int dtype, cudaStream_t stream);
int moments_seg_len(int B, int H, int N, int p);
size_t moments_state_bytes(int B, int H);
size_t moments_saved_state_bytes(int B, int H, int N, int p, int causal, int Hkv = 0);
size_t moments_local_state_workspace_bytes(int B, int H, int N);
int moments_local_state(const Problem& pb, const void* x, const void* w, const void* w2, void* state_out, void* ws);
int moments_local_state_rev(const Problem& pb, const void* q, const void* o, const float* g, const void* d_o,
                            void* state_out, void* ws);
size_t moments_local_state_rev_workspace_bytes(int B, int H, int N);
int moments_reduce_states(const void* gathered, int G, int rank, int dir, void* out, int B, int H, cudaStream_t stream);
size_t moments_workspace_bytes(int B, int H, int N, int p, int causal, bool bwd, int Hkv = 0);
void moments_flops(int B, int H, int N, int p, int causal, double* fwd, double* bwd, double* rebuild);
int moments_fwd(const Problem& pb, const void* q, const void* k, const void* v, void* o, float* g, void* ws);
int moments_bwd(const Problem& pb, const void* q, const void* k, const void* v, const void* o, const float* g,
                const void* d_o, void* dq, void* dk, void* dv, void* ws);

}  // namespace fm
