// fm_api.cu -- the C ABI of include/fastmax_b200.h: argument checks, planning, dispatch.
#include "fm_common.cuh"

#include <atomic>
#include <mutex>
#include <vector>
#include <string>

namespace fm {
ThreadStatus& status() {
  static thread_local ThreadStatus s = {{0}, 0};
  return s;
}

// ---- optional per-launch CUDA-event profiling (bench.py uses it for the roofline object) -----
struct ProfRecord {
  const char* name;
  cudaEvent_t a, b;
};
// process-wide (autograd runs backward on its own thread); a thread remembers the record it opened
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;
static std::vector<ProfRecord> g_prof;
static thread_local cudaEvent_t g_prof_open = nullptr;

void prof_begin(const char* name, cudaStream_t stream) {
  if (!g_prof_on.load(std::memory_order_relaxed)) return;
  ProfRecord r{name, nullptr, nullptr};
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  cudaEventRecord(r.a, stream);
  g_prof_open = r.b;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof.push_back(r);
}
void prof_end(cudaStream_t stream) {
  if (!g_prof_open) return;
  cudaEventRecord(g_prof_open, stream);
  g_prof_open = nullptr;
}
}  // namespace fm

using fm::Problem;

// The tcgen05 kernels build their shared-memory descriptors on the host for a fixed dynamic-shared-memory
// base (fm_sm100.cuh: FM_SMEM_BASE) and trap if the device disagrees -- which would poison the CUDA context.
// Probe the base once per process with a one-warp kernel instead: if this driver / tool (compute-sanitizer, a
// debugger) reserves a different amount, the plan falls back to the CUDA-core kernels (AUTO) or reports
// FASTMAX_ERR_UNSUPPORTED (explicit MOMENTS).  Without a device (planning on a build host) the answer is "yes".
__global__ void probe_smem_base_kernel(unsigned* out) {
  extern __shared__ uint8_t probe_smem[];
  if (threadIdx.x == 0)
    *out = static_cast<unsigned>(__cvta_generic_to_shared(
        reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(probe_smem) + 1023) & ~uintptr_t(1023))));
}

static bool tensor_path_usable() {
  static int cached = -1;
  if (cached >= 0) return cached != 0;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return true;                               // no device: planning only (not cached: a device may appear later)
  }
  unsigned* d = nullptr;
  unsigned h = 0;
  cudaStream_t st = nullptr;
  bool ok = cudaMalloc(&d, sizeof(unsigned)) == cudaSuccess &&
            cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
  if (ok) {
    probe_smem_base_kernel<<<1, 32, 4096, st>>>(d);
    ok = cudaMemcpyAsync(&h, d, sizeof(unsigned), cudaMemcpyDeviceToHost, st) == cudaSuccess &&
         cudaStreamSynchronize(st) == cudaSuccess;
  }
  if (st) cudaStreamDestroy(st);
  if (d) cudaFree(d);
  if (!ok) {
    cudaGetLastError();
    return true;                               // could not probe (e.g. stream capture in progress): keep the default
  }
  cached = h == fm::sm100_smem_base() ? 1 : 0;
  if (!cached)
    fprintf(stderr, "fastmax: dynamic shared memory starts at %u, not %u: tcgen05 path disabled, using the CUDA-core kernels\n",
            h, fm::sm100_smem_base());
  return cached != 0;
}

static int check_common(int B, int H, int N, int D, int p, int dtype) {
  if (B <= 0 || H <= 0 || N <= 0 || D <= 0) return FASTMAX_ERR_BAD_ARG;
  if (p != 1 && p != 2) return FASTMAX_ERR_BAD_ARG;  // reference raises ValueError (fastmax.py:428)
  if (dtype != FASTMAX_DTYPE_F32 && dtype != FASTMAX_DTYPE_BF16) return FASTMAX_ERR_BAD_ARG;
  if (D > 128) return FASTMAX_ERR_UNSUPPORTED;
  return FASTMAX_OK;
}

extern "C" {

int fastmax_abi_version(void) { return FASTMAX_B200_ABI_VERSION; }

const char* fastmax_error_string(int code) {
  switch (code) {
    case FASTMAX_OK: return "ok";
    case FASTMAX_ERR_BAD_ARG: return "bad argument";
    case FASTMAX_ERR_UNSUPPORTED: return "unsupported shape or dtype";
    case FASTMAX_ERR_WORKSPACE: return "workspace too small or misaligned";
    case FASTMAX_ERR_CUDA: return "CUDA error";
    default: return "unknown error";
  }
}

const char* fastmax_last_cuda_error(void) { return fm::status().cuda_err; }
int fastmax_last_launch_count(void) { return fm::status().launches; }

void fastmax_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(fm::g_prof_mu);
  for (auto& r : fm::g_prof) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  fm::g_prof.clear();
  fm::g_prof_on.store(on != 0);
}

int fastmax_profile_fetch(char* names, size_t names_cap, float* ms, int cap) {
  int n = 0;
  size_t used = 0;
  if (names && names_cap) names[0] = 0;
  std::lock_guard<std::mutex> lk(fm::g_prof_mu);
  for (auto& r : fm::g_prof) {
    if (n >= cap) break;
    if (cudaEventSynchronize(r.b) != cudaSuccess) return FASTMAX_ERR_CUDA;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) return FASTMAX_ERR_CUDA;
    ms[n] = t;
    size_t len = strlen(r.name);
    if (names && used + len + 2 <= names_cap) {
      memcpy(names + used, r.name, len);
      names[used + len] = '\n';
      names[used + len + 1] = 0;
      used += len + 1;
    }
    ++n;
  }
  return n;
}

int fastmax_plan_make(fastmax_plan* plan, int B, int H, int N, int D, int p, int causal, int dtype, int algo) {
  if (!plan) return FASTMAX_ERR_BAD_ARG;
  int rc = check_common(B, H, N, D, p, dtype);
  if (rc) return rc;
  memset(plan, 0, sizeof(*plan));
  (void)causal;
  if (algo != FASTMAX_ALGO_AUTO && algo != FASTMAX_ALGO_EXACT && algo != FASTMAX_ALGO_MOMENTS)
    return FASTMAX_ERR_BAD_ARG;
  const bool mom_ok = dtype == FASTMAX_DTYPE_BF16 && D == 64 && tensor_path_usable();
  if (algo == FASTMAX_ALGO_MOMENTS && !mom_ok) return FASTMAX_ERR_UNSUPPORTED;
  // the tcgen05 path also covers short sequences (one segment = dense tensor-core tiles, no moment
  // states); below a couple of tiles the fp32 CUDA-core kernels are as fast and exact
  const bool use_mom = algo == FASTMAX_ALGO_MOMENTS || (algo == FASTMAX_ALGO_AUTO && mom_ok && N >= 256);
  if (use_mom) {
    plan->algo = FASTMAX_ALGO_MOMENTS;
    plan->seg_len = fm::moments_seg_len(B, H, N, p);
    plan->n_seg = (N + plan->seg_len - 1) / plan->seg_len;
    plan->fwd_workspace_bytes = fm::moments_workspace_bytes(B, H, N, p, causal, false);
    plan->bwd_workspace_bytes = fm::moments_workspace_bytes(B, H, N, p, causal, true);
    fm::moments_flops(B, H, N, p, causal, &plan->fwd_tensor_flops, &plan->bwd_tensor_flops,
                      &plan->rebuild_tensor_flops);
    return FASTMAX_OK;
  }
  plan->algo = FASTMAX_ALGO_EXACT;
  plan->seg_len = 0;
  plan->n_seg = 1;
  plan->fwd_workspace_bytes = 0;
  plan->bwd_workspace_bytes = 0;
  {
    // head size 128 in bf16: dense tcgen05 tiles (fm_dense128.cu); its backward needs u and dlt
    Problem pb{B, H, N, D, p, causal ? 1 : 0, dtype, 1.f, nullptr};
    if (fm::dense128_usable(pb) && tensor_path_usable()) plan->bwd_workspace_bytes = fm::dense128_bwd_workspace_bytes(B, H, N);
    // fp32, head size <= 64: the same tiles on [hi | lo] bf16 pairs (AUTO only: EXACT keeps the fp32 CUDA-core kernels)
    if (algo == FASTMAX_ALGO_AUTO && fm::dense_split_usable(pb) && tensor_path_usable()) {
      plan->fwd_workspace_bytes = fm::dense_split_fwd_workspace_bytes(B, H, N);
      plan->bwd_workspace_bytes = fm::dense_split_bwd_workspace_bytes(B, H, N);
    }
  }
  return FASTMAX_OK;
}

size_t fastmax_fwd_workspace_bytes(int B, int H, int N, int D, int p, int causal, int dtype, int algo) {
  fastmax_plan pl;
  if (fastmax_plan_make(&pl, B, H, N, D, p, causal, dtype, algo)) return 0;
  return (size_t)pl.fwd_workspace_bytes;
}

size_t fastmax_bwd_workspace_bytes(int B, int H, int N, int D, int p, int causal, int dtype, int algo) {
  fastmax_plan pl;
  if (fastmax_plan_make(&pl, B, H, N, D, p, causal, dtype, algo)) return 0;
  return (size_t)pl.bwd_workspace_bytes;
}

int fastmax_fwd(const void* q, const void* k, const void* v, void* o, float* g, int B, int H, int N, int D,
                int p, int causal, float inv_c, int dtype, int algo, void* workspace, size_t workspace_bytes,
                void* stream) {
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!q || !k || !v || !o || !g) return FASTMAX_ERR_BAD_ARG;
  fastmax_plan pl;
  int rc = fastmax_plan_make(&pl, B, H, N, D, p, causal, dtype, algo);
  if (rc) return rc;
  if (pl.fwd_workspace_bytes > workspace_bytes || (pl.fwd_workspace_bytes && !workspace))
    return FASTMAX_ERR_WORKSPACE;
  Problem pb{B, H, N, D, p, causal ? 1 : 0, dtype, inv_c, (cudaStream_t)stream};
  if (pl.algo == FASTMAX_ALGO_MOMENTS) {
    if (((uintptr_t)workspace & 255) != 0) return FASTMAX_ERR_WORKSPACE;
    return fm::moments_fwd(pb, q, k, v, o, g, workspace);
  }
  if (fm::dense128_usable(pb) && tensor_path_usable()) return fm::dense128_fwd(pb, q, k, v, o, g);
  if (algo == FASTMAX_ALGO_AUTO && fm::dense_split_usable(pb) && tensor_path_usable()) {
    if (((uintptr_t)workspace & 255) != 0) return FASTMAX_ERR_WORKSPACE;
    return fm::dense_split_fwd(pb, q, k, v, o, g, workspace);
  }
  return fm::exact_fwd(pb, q, k, v, o, g);
}

int fastmax_bwd(const void* q, const void* k, const void* v, const void* o, const float* g, const void* d_o,
                void* dq, void* dk, void* dv, int B, int H, int N, int D, int p, int causal, float inv_c,
                int dtype, int algo, void* workspace, size_t workspace_bytes, void* stream) {
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!q || !k || !v || !o || !g || !d_o || !dq || !dk || !dv) return FASTMAX_ERR_BAD_ARG;
  fastmax_plan pl;
  int rc = fastmax_plan_make(&pl, B, H, N, D, p, causal, dtype, algo);
  if (rc) return rc;
  if (pl.bwd_workspace_bytes > workspace_bytes || (pl.bwd_workspace_bytes && !workspace))
    return FASTMAX_ERR_WORKSPACE;
  Problem pb{B, H, N, D, p, causal ? 1 : 0, dtype, inv_c, (cudaStream_t)stream};
  if (pl.algo == FASTMAX_ALGO_MOMENTS) {
    if (((uintptr_t)workspace & 255) != 0) return FASTMAX_ERR_WORKSPACE;
    return fm::moments_bwd(pb, q, k, v, o, g, d_o, dq, dk, dv, workspace);
  }
  if (fm::dense128_usable(pb) && tensor_path_usable()) {
    if (((uintptr_t)workspace & 255) != 0) return FASTMAX_ERR_WORKSPACE;
    return fm::dense128_bwd(pb, q, k, v, o, g, d_o, dq, dk, dv, workspace);
  }
  if (algo == FASTMAX_ALGO_AUTO && fm::dense_split_usable(pb) && tensor_path_usable()) {
    if (((uintptr_t)workspace & 255) != 0) return FASTMAX_ERR_WORKSPACE;
    return fm::dense_split_bwd(pb, q, k, v, o, g, d_o, dq, dk, dv, workspace);
  }
  return fm::exact_bwd(pb, q, k, v, o, g, d_o, dq, dk, dv);
}

// ---- forward states kept for backward ---------------------------------------------------------
size_t fastmax_saved_state_bytes(int B, int H, int N, int D, int p, int causal, int dtype, int algo) {
  fastmax_plan pl;
  if (fastmax_plan_make(&pl, B, H, N, D, p, causal, dtype, algo)) return 0;
  if (pl.algo != FASTMAX_ALGO_MOMENTS) return 0;
  return fm::moments_saved_state_bytes(B, H, N, p, causal);
}

int fastmax_fwd_ex(const void* q, const void* k, const void* v, void* o, float* g, void* saved_states, int B, int H,
                   int N, int D, int p, int causal, float inv_c, int dtype, int algo, void* workspace,
                   size_t workspace_bytes, void* stream) {
  if (!saved_states)
    return fastmax_fwd(q, k, v, o, g, B, H, N, D, p, causal, inv_c, dtype, algo, workspace, workspace_bytes, stream);
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!q || !k || !v || !o || !g) return FASTMAX_ERR_BAD_ARG;
  fastmax_plan pl;
  int rc = fastmax_plan_make(&pl, B, H, N, D, p, causal, dtype, algo);
  if (rc) return rc;
  if (pl.algo != FASTMAX_ALGO_MOMENTS) return FASTMAX_ERR_UNSUPPORTED;
  if (pl.fwd_workspace_bytes > workspace_bytes || !workspace || ((uintptr_t)workspace & 255)) return FASTMAX_ERR_WORKSPACE;
  Problem pb{B, H, N, D, p, causal ? 1 : 0, dtype, inv_c, (cudaStream_t)stream};
  pb.saved_states = saved_states;
  return fm::moments_fwd(pb, q, k, v, o, g, workspace);
}

int fastmax_bwd_ex(const void* q, const void* k, const void* v, const void* o, const float* g, const void* d_o,
                   void* dq, void* dk, void* dv, const void* saved_states, int B, int H, int N, int D, int p,
                   int causal, float inv_c, int dtype, int algo, void* workspace, size_t workspace_bytes,
                   void* stream) {
  if (!saved_states)
    return fastmax_bwd(q, k, v, o, g, d_o, dq, dk, dv, B, H, N, D, p, causal, inv_c, dtype, algo, workspace,
                       workspace_bytes, stream);
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!q || !k || !v || !o || !g || !d_o || !dq || !dk || !dv) return FASTMAX_ERR_BAD_ARG;
  fastmax_plan pl;
  int rc = fastmax_plan_make(&pl, B, H, N, D, p, causal, dtype, algo);
  if (rc) return rc;
  if (pl.algo != FASTMAX_ALGO_MOMENTS) return FASTMAX_ERR_UNSUPPORTED;
  if (pl.bwd_workspace_bytes > workspace_bytes || !workspace || ((uintptr_t)workspace & 255)) return FASTMAX_ERR_WORKSPACE;
  Problem pb{B, H, N, D, p, causal ? 1 : 0, dtype, inv_c, (cudaStream_t)stream};
  pb.saved_states = const_cast<void*>(saved_states);
  return fm::moments_bwd(pb, q, k, v, o, g, d_o, dq, dk, dv, workspace);
}

// ---- grouped-query attention -------------------------------------------------------------------
static int gqa_check(int B, int H, int Hkv, int N, int D, int p, int causal, int dtype, fastmax_plan* pl) {
  if (Hkv <= 0 || H % Hkv != 0) return FASTMAX_ERR_BAD_ARG;
  int rc = fastmax_plan_make(pl, B, H, N, D, p, causal, dtype, FASTMAX_ALGO_MOMENTS);
  if (rc) return rc;
  if (p != 2) return FASTMAX_ERR_UNSUPPORTED;   // first order runs on the scan kernels, which are not group-aware
  return FASTMAX_OK;
}

size_t fastmax_gqa_workspace_bytes(int B, int H, int Hkv, int N, int D, int p, int causal, int dtype, int backward) {
  fastmax_plan pl;
  if (gqa_check(B, H, Hkv, N, D, p, causal, dtype, &pl)) return 0;
  return fm::moments_workspace_bytes(B, H, N, p, causal, backward != 0, Hkv);
}

size_t fastmax_gqa_saved_state_bytes(int B, int H, int Hkv, int N, int D, int p, int causal, int dtype) {
  fastmax_plan pl;
  if (gqa_check(B, H, Hkv, N, D, p, causal, dtype, &pl)) return 0;
  return fm::moments_saved_state_bytes(B, H, N, p, causal, Hkv);
}

int fastmax_fwd_gqa(const void* q, const void* k, const void* v, void* o, float* g, void* saved_states, int B, int H,
                    int Hkv, int N, int D, int p, int causal, float inv_c, int dtype, void* workspace,
                    size_t workspace_bytes, void* stream) {
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!q || !k || !v || !o || !g) return FASTMAX_ERR_BAD_ARG;
  fastmax_plan pl;
  int rc = gqa_check(B, H, Hkv, N, D, p, causal, dtype, &pl);
  if (rc) return rc;
  if (fm::moments_workspace_bytes(B, H, N, p, causal, false, Hkv) > workspace_bytes || !workspace ||
      ((uintptr_t)workspace & 255))
    return FASTMAX_ERR_WORKSPACE;
  Problem pb{B, H, N, D, p, causal ? 1 : 0, dtype, inv_c, (cudaStream_t)stream};
  pb.saved_states = saved_states;
  pb.Hkv = Hkv;
  return fm::moments_fwd(pb, q, k, v, o, g, workspace);
}

int fastmax_bwd_gqa(const void* q, const void* k, const void* v, const void* o, const float* g, const void* d_o,
                    void* dq, void* dk, void* dv, const void* saved_states, int B, int H, int Hkv, int N, int D, int p,
                    int causal, float inv_c, int dtype, void* workspace, size_t workspace_bytes, void* stream) {
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!q || !k || !v || !o || !g || !d_o || !dq || !dk || !dv) return FASTMAX_ERR_BAD_ARG;
  fastmax_plan pl;
  int rc = gqa_check(B, H, Hkv, N, D, p, causal, dtype, &pl);
  if (rc) return rc;
  if (fm::moments_workspace_bytes(B, H, N, p, causal, true, Hkv) > workspace_bytes || !workspace ||
      ((uintptr_t)workspace & 255))
    return FASTMAX_ERR_WORKSPACE;
  Problem pb{B, H, N, D, p, causal ? 1 : 0, dtype, inv_c, (cudaStream_t)stream};
  pb.saved_states = const_cast<void*>(saved_states);
  pb.Hkv = Hkv;
  return fm::moments_bwd(pb, q, k, v, o, g, d_o, dq, dk, dv, workspace);
}

int fastmax_qkv_prologue(const void* qkv, const void* cos, const void* sin, void* q, void* k, void* v, int B, int T,
                         int n_query_groups, int q_per_kv, int head_size, int rope_n_elem, int dtype, int dtype_cos_sin,
                         void* stream) {
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!qkv || !q || !k || !v || (rope_n_elem > 0 && (!cos || !sin))) return FASTMAX_ERR_BAD_ARG;
  if ((dtype != FASTMAX_DTYPE_BF16 && dtype != FASTMAX_DTYPE_F32) ||
      (dtype_cos_sin != FASTMAX_DTYPE_BF16 && dtype_cos_sin != FASTMAX_DTYPE_F32))
    return FASTMAX_ERR_BAD_ARG;
  return fm::qkv_prologue(qkv, cos, sin, q, k, v, B, T, n_query_groups, q_per_kv, head_size, rope_n_elem, dtype,
                          dtype_cos_sin, 0, (cudaStream_t)stream);
}

int fastmax_qkv_prologue_bwd(const void* dq, const void* dk, const void* dv, const void* cos, const void* sin, void* dqkv,
                             int B, int T, int n_query_groups, int q_per_kv, int head_size, int rope_n_elem, int dtype,
                             int dtype_cos_sin, void* stream) {
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!dqkv || !dq || !dk || !dv || (rope_n_elem > 0 && (!cos || !sin))) return FASTMAX_ERR_BAD_ARG;
  if ((dtype != FASTMAX_DTYPE_BF16 && dtype != FASTMAX_DTYPE_F32) ||
      (dtype_cos_sin != FASTMAX_DTYPE_BF16 && dtype_cos_sin != FASTMAX_DTYPE_F32))
    return FASTMAX_ERR_BAD_ARG;
  return fm::qkv_prologue(dqkv, cos, sin, const_cast<void*>(dq), const_cast<void*>(dk), const_cast<void*>(dv), B, T,
                          n_query_groups, q_per_kv, head_size, rope_n_elem, dtype, dtype_cos_sin, 1, (cudaStream_t)stream);
}

// ---- recurrent inference state ---------------------------------------------------------------
size_t fastmax_decode_state_bytes(int B, int H, int D, int p) {
  if (B <= 0 || H <= 0 || D <= 0 || D > 128 || (p != 1 && p != 2)) return 0;
  return fm::decode_state_bytes(B, H, D, p);
}

int fastmax_decode_init(void* state, int B, int H, int D, int p, void* stream) {
  fm::status().launches = 0;
  const size_t nb = fastmax_decode_state_bytes(B, H, D, p);
  if (!state || !nb) return FASTMAX_ERR_BAD_ARG;
  FM_CUDA(cudaMemsetAsync(state, 0, nb, (cudaStream_t)stream));
  return FASTMAX_OK;
}

int fastmax_decode_update(void* state, const void* k, const void* v, int B, int H, int T, int D, int p, int dtype,
                          void* stream) {
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!state || !k || !v) return FASTMAX_ERR_BAD_ARG;
  int rc = check_common(B, H, T, D, p, dtype);
  if (rc) return rc;
  return fm::decode_update(state, k, v, B, H, T, D, p, dtype, (cudaStream_t)stream);
}

int fastmax_decode_attend(const void* state, const void* q, void* o, int B, int H, int T, int D, int p, float inv_c,
                          float count_term, int dtype, void* stream) {
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!state || !q || !o) return FASTMAX_ERR_BAD_ARG;
  int rc = check_common(B, H, T, D, p, dtype);
  if (rc) return rc;
  return fm::decode_attend(state, q, o, B, H, T, D, p, inv_c, count_term, dtype, (cudaStream_t)stream);
}

// ---- sequence parallelism ------------------------------------------------------------------
static int check_sp(int B, int H, int N, int D, int p, int dtype) {
  int rc = check_common(B, H, N, D, p, dtype);
  if (rc) return rc;
  if (dtype != FASTMAX_DTYPE_BF16 || D != 64 || !tensor_path_usable()) return FASTMAX_ERR_UNSUPPORTED;
  return FASTMAX_OK;
}

size_t fastmax_sp_state_bytes(int B, int H) { return fm::moments_state_bytes(B, H); }

size_t fastmax_sp_local_state_workspace_bytes(int B, int H, int N, int reverse) {
  return reverse ? fm::moments_local_state_rev_workspace_bytes(B, H, N) : fm::moments_local_state_workspace_bytes(B, H, N);
}

int fastmax_sp_local_state(const void* k, const void* v, void* state_out, int B, int H, int N, int D, int p,
                           void* workspace, size_t workspace_bytes, void* stream) {
  fm::status().launches = 0;
  if (!k || !v || !state_out || !workspace) return FASTMAX_ERR_BAD_ARG;
  int rc = check_sp(B, H, N, D, p, FASTMAX_DTYPE_BF16);
  if (rc) return rc;
  if (workspace_bytes < fm::moments_local_state_workspace_bytes(B, H, N)) return FASTMAX_ERR_WORKSPACE;
  Problem pb{B, H, N, D, p, 0, FASTMAX_DTYPE_BF16, 0.f, (cudaStream_t)stream};
  return fm::moments_local_state(pb, k, v, k, state_out, workspace);
}

int fastmax_sp_local_state_rev(const void* q, const void* o, const float* g, const void* d_o, void* state_out,
                               int B, int H, int N, int D, int p, void* workspace, size_t workspace_bytes,
                               void* stream) {
  fm::status().launches = 0;
  if (!q || !o || !g || !d_o || !state_out || !workspace) return FASTMAX_ERR_BAD_ARG;
  int rc = check_sp(B, H, N, D, p, FASTMAX_DTYPE_BF16);
  if (rc) return rc;
  if (workspace_bytes < fm::moments_local_state_rev_workspace_bytes(B, H, N)) return FASTMAX_ERR_WORKSPACE;
  Problem pb{B, H, N, D, p, 0, FASTMAX_DTYPE_BF16, 0.f, (cudaStream_t)stream};
  return fm::moments_local_state_rev(pb, q, o, g, d_o, state_out, workspace);
}

int fastmax_sp_reduce(const void* gathered, int n_ranks, int rank, int dir, void* state_out, int B, int H,
                      void* stream) {
  fm::status().launches = 0;
  if (!gathered || !state_out || n_ranks <= 0 || rank < 0 || rank >= n_ranks || B <= 0 || H <= 0 || (dir != 0 && dir != 1))
    return FASTMAX_ERR_BAD_ARG;
  return fm::moments_reduce_states(gathered, n_ranks, rank, dir, state_out, B, H, (cudaStream_t)stream);
}

int fastmax_fwd_sp(const void* q, const void* k, const void* v, void* o, float* g, const void* init_state,
                   long long token_offset, int B, int H, int N, int D, int p, float inv_c, int dtype,
                   void* workspace, size_t workspace_bytes, void* stream) {
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!q || !k || !v || !o || !g || token_offset < 0) return FASTMAX_ERR_BAD_ARG;
  int rc = check_sp(B, H, N, D, p, dtype);
  if (rc) return rc;
  const size_t need = fm::moments_workspace_bytes(B, H, N, p, 1, false);
  if (need > workspace_bytes || !workspace || ((uintptr_t)workspace & 255)) return FASTMAX_ERR_WORKSPACE;
  Problem pb{B, H, N, D, p, 1, dtype, inv_c, (cudaStream_t)stream};
  pb.sp_init_fwd = init_state, pb.sp_offset = token_offset;
  return fm::moments_fwd(pb, q, k, v, o, g, workspace);
}

int fastmax_bwd_sp(const void* q, const void* k, const void* v, const void* o, const float* g, const void* d_o,
                   void* dq, void* dk, void* dv, const void* init_state_fwd, const void* init_state_rev,
                   long long token_offset, int B, int H, int N, int D, int p, float inv_c, int dtype,
                   void* workspace, size_t workspace_bytes, void* stream) {
  fm::status().launches = 0;
  fm::status().cuda_err[0] = 0;
  if (!q || !k || !v || !o || !g || !d_o || !dq || !dk || !dv || token_offset < 0) return FASTMAX_ERR_BAD_ARG;
  int rc = check_sp(B, H, N, D, p, dtype);
  if (rc) return rc;
  const size_t need = fm::moments_workspace_bytes(B, H, N, p, 1, true);
  if (need > workspace_bytes || !workspace || ((uintptr_t)workspace & 255)) return FASTMAX_ERR_WORKSPACE;
  Problem pb{B, H, N, D, p, 1, dtype, inv_c, (cudaStream_t)stream};
  pb.sp_init_fwd = init_state_fwd, pb.sp_init_rev = init_state_rev, pb.sp_offset = token_offset;
  return fm::moments_bwd(pb, q, k, v, o, g, d_o, dq, dk, dv, workspace);
}

}  // extern "C"
