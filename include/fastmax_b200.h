/*
 * fastmax_b200.h -- C ABI of the B200-native FastMax attention library
 * (libfastmax_b200.so, built from fastmax_experiments_b200/csrc/).
 *
 * This is the drop-in boundary of the hot path.  What each entry point replaces
 * in the reference (ryansynk/fastmax-experiments):
 *
 *   fastmax_fwd   <- fastattention_einops.forward   attention_mechanisms/fastmax.py:42-111
 *                    (compute_F_{un,}masked :185-250, compute_g_{un,}masked :253-322, o=F/g :97)
 *                    and the never-vendored  fastmax_cuda.forwardpass   lit_gpt/model.py:82-84
 *   fastmax_bwd   <- fastattention_einops.backward  attention_mechanisms/fastmax.py:114-182
 *                    (gradient_o_{q,k,v}_{un,}masked :384-691)
 *                    and                     fastmax_cuda.backwardpass  lit_gpt/model.py:116
 *   fastmax_*_workspace_bytes / fastmax_plan_* -- no counterpart (the reference
 *                    allocates its (B,H,N,D,D,D) intermediates implicitly through torch).
 *   fastmax_sp_*  <- no counterpart: sequence-parallel state exchange (BASELINE.json
 *                    north_star).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types; every function returns
 *     FASTMAX_OK (0) or a negative FASTMAX_ERR_* code and never throws.
 *   - all tensor pointers are DEVICE pointers, layout (B, H, N, D) contiguous
 *     ("b h n d" of fastmax.py:61-63), element type given by `dtype`;
 *     g is float32 (B, H, N) regardless of dtype.
 *   - stream-ordered on `stream` (a cudaStream_t passed as void*), re-entrant, no
 *     global mutable state besides lazily created per-device kernel attributes.
 *   - caller owns every buffer including the workspace (size from *_workspace_bytes,
 *     256-byte aligned).  Workspace contents are scratch; nothing persists.
 *   - score divisor: inv_c = 1/c with c = normalize_term*sqrt(D), or 1 if
 *     tensors_normalized (fastmax.py:78-82).  p in {1,2} (fastmax.py:68-69).
 *     causal != 0 is the reference's `mask is not False`.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns FASTMAX_ERR_CUDA.
 */
#ifndef FASTMAX_B200_H_
#define FASTMAX_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FASTMAX_B200_ABI_VERSION 4

/* error codes */
#define FASTMAX_OK              0
#define FASTMAX_ERR_BAD_ARG    -1   /* null pointer, non-positive size, p not in {1,2} ... */
#define FASTMAX_ERR_UNSUPPORTED -2  /* shape/dtype outside what the kernels cover (D > 128, ...) */
#define FASTMAX_ERR_WORKSPACE  -3   /* workspace too small / misaligned */
#define FASTMAX_ERR_CUDA       -4   /* a CUDA runtime/driver call failed (see fastmax_last_cuda_error) */

/* element types of q,k,v,o,do,dq,dk,dv */
#define FASTMAX_DTYPE_F32   0
#define FASTMAX_DTYPE_BF16  1

/* algorithm selection (fastmax_plan.algo / `algo` arguments) */
#define FASTMAX_ALGO_AUTO     0  /* cost-model choice                                          */
#define FASTMAX_ALGO_EXACT    1  /* fp32 CUDA-core tiles of the dense formula (any D<=128)      */
#define FASTMAX_ALGO_MOMENTS  2  /* tcgen05 chunked scan: intra-segment scores + moment states  */

typedef struct fastmax_plan {
  int32_t algo;           /* resolved FASTMAX_ALGO_*                                   */
  int32_t seg_len;        /* tokens per scan segment (MOMENTS), 0 for EXACT            */
  int32_t n_seg;          /* segments per (b,h) sequence                               */
  int32_t reserved;
  uint64_t fwd_workspace_bytes;
  uint64_t bwd_workspace_bytes;
  double fwd_tensor_flops; /* MMA flops the plan executes (not the algorithmic count)  */
  double bwd_tensor_flops;
  double rebuild_tensor_flops; /* part of bwd_tensor_flops that fastmax_bwd_ex skips when the
                                  forward states were kept (fastmax_fwd_ex saved_states)       */
} fastmax_plan;

int         fastmax_abi_version(void);
const char* fastmax_error_string(int code);
/* text of the last CUDA error seen by the calling thread ("" if none) */
const char* fastmax_last_cuda_error(void);

/* Resolve the algorithm / segmentation for a problem and report workspace needs. */
int fastmax_plan_make(fastmax_plan* plan, int B, int H, int N, int D, int p, int causal,
                      int dtype, int algo);

size_t fastmax_fwd_workspace_bytes(int B, int H, int N, int D, int p, int causal, int dtype, int algo);
size_t fastmax_bwd_workspace_bytes(int B, int H, int N, int D, int p, int causal, int dtype, int algo);

/* o (B,H,N,D) dtype, g (B,H,N) float32.  g is the normaliser the reference saves for
 * backward (fastmax.py:106). */
int fastmax_fwd(const void* q, const void* k, const void* v, void* o, float* g,
                int B, int H, int N, int D, int p, int causal, float inv_c, int dtype, int algo,
                void* workspace, size_t workspace_bytes, void* stream);

/* dq,dk,dv (B,H,N,D) dtype from saved q,k,v,o,g and the upstream gradient d_o. */
int fastmax_bwd(const void* q, const void* k, const void* v, const void* o, const float* g,
                const void* d_o, void* dq, void* dk, void* dv,
                int B, int H, int N, int D, int p, int causal, float inv_c, int dtype, int algo,
                void* workspace, size_t workspace_bytes, void* stream);

/* number of kernel launches the last fastmax_fwd / fastmax_bwd call of this thread issued
 * (bench.py reports it as gpu_launches) */
int fastmax_last_launch_count(void);

/* ---- keeping the forward moment states for backward (speed / memory knob, no reference counterpart) ----
 * By default fastmax_bwd rebuilds the forward-direction boundary states from k, v (nothing but
 * q,k,v,o,g is saved, like the reference: fastmax.py:106).  A caller that can spare
 * fastmax_saved_state_bytes() of memory per call (545 KiB per (b,h) per segment) passes the same buffer
 * to fastmax_fwd_ex and fastmax_bwd_ex and saves one build pass (~14 % of a fwd+bwd step at the
 * headline shape).  saved_states == NULL makes the _ex entry points identical to fastmax_fwd / _bwd.
 * fastmax_saved_state_bytes returns 0 when the plan does not use moment states. */
size_t fastmax_saved_state_bytes(int B, int H, int N, int D, int p, int causal, int dtype, int algo);
int fastmax_fwd_ex(const void* q, const void* k, const void* v, void* o, float* g, void* saved_states,
                   int B, int H, int N, int D, int p, int causal, float inv_c, int dtype, int algo,
                   void* workspace, size_t workspace_bytes, void* stream);
int fastmax_bwd_ex(const void* q, const void* k, const void* v, const void* o, const float* g,
                   const void* d_o, void* dq, void* dk, void* dv, const void* saved_states,
                   int B, int H, int N, int D, int p, int causal, float inv_c, int dtype, int algo,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- grouped-query attention and the fused q/k/v prologue --------------------------------
 * lit_gpt/model.py:408-416 copies k, v over the q_per_kv query heads of a group (expand + reshape) before the
 * attention op sees them.  fastmax_*_gqa take k, v as (B, Hkv, N, D) instead: the forward moment states -- and the
 * memory that keeping them costs -- exist once per (b, kv head) (H / Hkv times less state-update work), every
 * kernel reads the shared rows in place, and dk, dv come back summed over the group, (B, Hkv, N, D): what autograd
 * does with the expand(), without the expanded copies.  Query head h reads kv head h / (H / Hkv), the order of
 * model.py:418-420.  Supported: bf16, D = 64, p = 2 (FASTMAX_ALGO_MOMENTS), causal or not; otherwise
 * FASTMAX_ERR_UNSUPPORTED (the Python op then expands k, v and calls fastmax_fwd / fastmax_bwd).
 * saved_states may be NULL (backward then rebuilds the forward states), else fastmax_gqa_saved_state_bytes bytes. */
size_t fastmax_gqa_workspace_bytes(int B, int H, int Hkv, int N, int D, int p, int causal, int dtype, int backward);
size_t fastmax_gqa_saved_state_bytes(int B, int H, int Hkv, int N, int D, int p, int causal, int dtype);
int fastmax_fwd_gqa(const void* q, const void* k, const void* v, void* o, float* g, void* saved_states,
                    int B, int H, int Hkv, int N, int D, int p, int causal, float inv_c, int dtype,
                    void* workspace, size_t workspace_bytes, void* stream);
int fastmax_bwd_gqa(const void* q, const void* k, const void* v, const void* o, const float* g, const void* d_o,
                    void* dq, void* dk, void* dv, const void* saved_states,
                    int B, int H, int Hkv, int N, int D, int p, int causal, float inv_c, int dtype,
                    void* workspace, size_t workspace_bytes, void* stream);

/* q/k/v assembly between the qkv projection and the attention call, lit_gpt/model.py:398-425 in one pass:
 * qkv (B, T, n_query_groups, q_per_kv + 2, head_size) -> q (B, n_query_groups * q_per_kv, T, head_size),
 * k, v (B, n_query_groups, T, head_size) (not expanded: see fastmax_*_gqa), rotary embedding (apply_rope,
 * model.py:702-708) on the first rope_n_elem features of q and k with the caller's cos / sin tables
 * (T, rope_n_elem).  Bit-exact against the reference's eager arithmetic (per-op rounding when the tables share the
 * tensors' dtype, one rounding under fp32 promotion).  _bwd is its transpose: dq, dk, dv -> dqkv.
 * Whole 16-byte chunks moved in pairs: rope_n_elem / 2 a multiple of 8 (bf16) / 4 (fp32) elements, head_size and
 * head_size - rope_n_elem multiples of 16 / 8, else UNSUPPORTED. */
int fastmax_qkv_prologue(const void* qkv, const void* cos, const void* sin, void* q, void* k, void* v,
                         int B, int T, int n_query_groups, int q_per_kv, int head_size, int rope_n_elem,
                         int dtype, int dtype_cos_sin, void* stream);
int fastmax_qkv_prologue_bwd(const void* dq, const void* dk, const void* dv, const void* cos, const void* sin,
                             void* dqkv, int B, int T, int n_query_groups, int q_per_kv, int head_size,
                             int rope_n_elem, int dtype, int dtype_cos_sin, void* stream);

/* ---- sequence parallelism (no reference counterpart; BASELINE.json north_star) ----------
 * A long sequence is split by contiguous token ranges over the GPUs of one box.  Outputs are linear
 * in the moment state of the preceding tokens, so one exchange per scan direction suffices:
 *   1. fastmax_sp_local_state{,_rev}: each rank reduces the moment state of its own tokens
 *      (fastmax_sp_state_bytes bytes, one 545 KiB record per (b,h));
 *   2. the caller all-gathers the records (NCCL) into [G][B*H] order;
 *   3. fastmax_sp_reduce: exclusive prefix (dir 0) or exclusive suffix (dir 1) over the ranks;
 *   4. fastmax_fwd_sp / fastmax_bwd_sp: the local chunked scan started from that state.
 * Supported: bf16, D = 64, causal (FASTMAX_ALGO_MOMENTS).  N is the LOCAL token count,
 * token_offset the number of tokens held by earlier ranks.  init_* may be NULL (rank 0 / last rank). */
size_t fastmax_sp_state_bytes(int B, int H);
size_t fastmax_sp_local_state_workspace_bytes(int B, int H, int N, int reverse);
int fastmax_sp_local_state(const void* k, const void* v, void* state_out, int B, int H, int N, int D, int p,
                           void* workspace, size_t workspace_bytes, void* stream);
int fastmax_sp_local_state_rev(const void* q, const void* o, const float* g, const void* d_o, void* state_out,
                               int B, int H, int N, int D, int p, void* workspace, size_t workspace_bytes,
                               void* stream);
int fastmax_sp_reduce(const void* gathered, int n_ranks, int rank, int dir, void* state_out, int B, int H,
                      void* stream);
int fastmax_fwd_sp(const void* q, const void* k, const void* v, void* o, float* g, const void* init_state,
                   long long token_offset, int B, int H, int N, int D, int p, float inv_c, int dtype,
                   void* workspace, size_t workspace_bytes, void* stream);
int fastmax_bwd_sp(const void* q, const void* k, const void* v, const void* o, const float* g, const void* d_o,
                   void* dq, void* dk, void* dv, const void* init_state_fwd, const void* init_state_rev,
                   long long token_offset, int B, int H, int N, int D, int p, float inv_c, int dtype,
                   void* workspace, size_t workspace_bytes, void* stream);

/* ---- recurrent inference state (replaces the KV cache of generate/base.py:83-88, lit_gpt/model.py:427-430) ----
 * The reference's generation path evaluates, for every new token, the NON-causal FastMax of the new queries
 * over the whole zero-padded KV cache (model.py:478-487 with input_pos given).  The factorisation needs only
 * the moment state of the keys seen so far: (1 + D + D^2)(D + 1) floats per (b,h) (D^2 term for p = 2 only),
 * fastmax_decode_state_bytes() in total.  fastmax_decode_update adds T new key/value rows (B,H,T,D) to it;
 * fastmax_decode_attend evaluates T query rows against it.  `count_term` is the count term of the normaliser:
 * the reference uses the number of QUERY rows of the call (fastmax.py:269-271, quirk Q2), fastmax_hack.py:21
 * (linearmax) the number of cache rows; pass whichever the caller reproduces.  fp32 state and accumulation,
 * q/k/v/o in `dtype`, D <= 128. */
size_t fastmax_decode_state_bytes(int B, int H, int D, int p);
int fastmax_decode_init(void* state, int B, int H, int D, int p, void* stream);
int fastmax_decode_update(void* state, const void* k, const void* v, int B, int H, int T, int D, int p, int dtype,
                          void* stream);
int fastmax_decode_attend(const void* state, const void* q, void* o, int B, int H, int T, int D, int p, float inv_c,
                          float count_term, int dtype, void* stream);

/* Optional per-launch timing for measurement harnesses (no reference counterpart).
 * While enabled, every kernel launch of the calling thread is bracketed by CUDA events on the
 * launching stream.  fastmax_profile_fetch waits for them and writes up to `cap` durations (ms)
 * plus the '\n'-separated kernel names; returns the number of records or a negative error.
 * Enabling (or disabling) clears earlier records. */
void fastmax_profile_enable(int on);
int  fastmax_profile_fetch(char* names, size_t names_cap, float* ms, int cap);

#ifdef __cplusplus
}
#endif
#endif /* FASTMAX_B200_H_ */
