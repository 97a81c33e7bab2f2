// fm_tmap.cu -- host-side TMA tensor-map construction (driver entry point resolved at run time,
// so the library links against cudart only).
#include "fm_common.cuh"
#include "fm_sm100.cuh"

namespace fm {

unsigned sm100_smem_base() { return sm100::FM_SMEM_BASE; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

static int make_tmap_3d(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t cols, uint64_t rows,
                        uint64_t batch, uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_cols,
                        uint32_t box_rows);

int make_tmap_bf16_3d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                      uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_cols,
                      uint32_t box_rows) {
  return make_tmap_3d(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, cols, rows, batch, row_stride_bytes,
                      batch_stride_bytes, box_cols, box_rows);
}

int make_tmap_f32_3d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint64_t batch,
                     uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_cols,
                     uint32_t box_rows) {
  return make_tmap_3d(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, cols, rows, batch, row_stride_bytes,
                      batch_stride_bytes, box_cols, box_rows);
}

static int make_tmap_3d(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t cols, uint64_t rows,
                        uint64_t batch, uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_cols,
                        uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    snprintf(status().cuda_err, sizeof(status().cuda_err), "cuTensorMapEncodeTiled entry point not found");
    return FASTMAX_ERR_CUDA;
  }
  cuuint64_t dims[3] = {cols, rows, batch};
  cuuint64_t strides[2] = {row_stride_bytes, batch_stride_bytes};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(status().cuda_err, sizeof(status().cuda_err),
             "cuTensorMapEncodeTiled failed: %d (cols %llu rows %llu batch %llu rs %llu bs %llu box %u x %u)", (int)r,
             (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)batch,
             (unsigned long long)row_stride_bytes, (unsigned long long)batch_stride_bytes, box_cols, box_rows);
    return FASTMAX_ERR_CUDA;
  }
  return FASTMAX_OK;
}

// 4-D bf16 map over the state records viewed as (col 64, row 64, slab NSLAB, record): box = (64 cols x box_rows x
// box_slabs x 1), 128B swizzle.  One bulk store then covers `box_rows` rows of `box_slabs` consecutive slabs.
int make_tmap_state_4d(CUtensorMap* map, const void* base, uint64_t nslab, uint64_t nrec, uint64_t rec_bytes,
                       uint32_t box_rows, uint32_t box_slabs) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    snprintf(status().cuda_err, sizeof(status().cuda_err), "cuTensorMapEncodeTiled entry point not found");
    return FASTMAX_ERR_CUDA;
  }
  cuuint64_t dims[4] = {64, 64, nslab, nrec};
  cuuint64_t strides[3] = {128, 64 * 128, rec_bytes};
  cuuint32_t box[4] = {64, box_rows, box_slabs, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    snprintf(status().cuda_err, sizeof(status().cuda_err), "cuTensorMapEncodeTiled (4-D state map) failed: %d", (int)r);
    return FASTMAX_ERR_CUDA;
  }
  return FASTMAX_OK;
}

}  // namespace fm
