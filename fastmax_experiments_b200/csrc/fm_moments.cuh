// fm_moments.cuh -- layout of the moment-state records shared by the build / apply / norm kernels.
//
// FastMax p=2 factorises  sum_j f(q.k_j/c) v'_j  (v' = [v, 1]) over the running moments
//   S0 = sum v'_j,  S1 = sum k_j (x) v'_j,  S2 = sum k_j (x) k_j (x) v'_j        
// One *state record* holds these for one (b,h) at one segment boundary (exclusive prefix of the
// tokens before the segment, or exclusive suffix for the reverse scan of the backward pass):
//
//   slab x in [0,64):  bf16 [y][a]  = sum X[x] Y[y] W[a]      (S2; forward: X=Y=k, W=v)
//   slab 64:           bf16 [y][a]  = sum Y[y] W[a]           (S1)
//   slab 65:           bf16 [y][x'] = sum Y[y] W2[x']         (normaliser part of S2; fwd: W2=k)
//   vec  fp32: [0,64)   = sum W[a]     (S0 value part)
//              [64,128) = sum W2[x']   (normaliser part of S1)
//              [128, 128+4096) = fp32 copy of slab 65
// Each slab is a 64x64 bf16 row-major tile (128-byte rows) so one TMA box with 128B swizzle
// brings it to shared memory as a canonical UMMA B operand -- MN-major when the contraction
// runs over y (forward o, dv) and K-major when it runs over a (dq, dk).
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace fm {
namespace mom {

constexpr int D = 64;                       // head size of the tcgen05 path
constexpr int NSLAB = 66;
constexpr size_t SLAB_BYTES = 64 * 64 * 2;  // 8 KiB
constexpr size_t VEC_OFF = NSLAB * SLAB_BYTES;
constexpr int VEC_FLOATS = 128 + 64 * 64;
constexpr size_t REC_BYTES = ((VEC_OFF + VEC_FLOATS * 4 + 1023) / 1024) * 1024;
constexpr int TILE_TOK = 128;               // rows of one MMA tile
constexpr int APPLY_TOK = 512;              // tokens per apply CTA (4 tiles)
// fp32 partials between the intra and the apply kernels: per 128-token tile one 32 KiB block laid out
// [16-byte column chunk c4 = 0..15][row 0..127], so that a warp of row-owning threads reads or writes 512
// contiguous bytes per instruction (a row-major partial means 16 bytes per lane at a 256-byte stride).
constexpr int PART_TILE_FLOATS = 128 * 64;
constexpr int BLK_TOK = 64;                 // tokens per build step

// One inter-segment contraction (fm_apply.cu):
//   out = out_scale * [1/g] * ( coef_reg * sum_x X[x] (Y . B_x) + coef64 * (Y . B_64) + vec + partial )
struct ApplyDesc {
  int L, nseg;
  int state_reverse;        // 1: records are exclusive suffixes (the last segment has none)
  int b_kmajor;             // 1: contraction over the slab's contiguous index (dq, dk); 0: over its rows (o, dv; X == Y)
  int edge_state;           // sequence parallel: first (forward states) / last (reverse) local segment has a state
  int vec_sel;              // -1 none, 0: vec[0:64), 1: vec[64:128)
  float coef_reg, coef64, out_scale;
  const void* X;            // (BH, N, 64) bf16 per-token scalars
  const void* Y;            // (BH, N, 64) bf16 per-token vectors
  const float* partial;     // blocked fp32 intra-segment partial, nullable
  const float* gden;        // (BH, N) fp32 normaliser the rows are divided by (forward), nullable
  void* out;                // (BH, N, 64) bf16
  const void* state;
};

}  // namespace mom
}  // namespace fm
