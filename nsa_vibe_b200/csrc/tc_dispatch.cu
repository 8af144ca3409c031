// Dispatch into the tcgen05 / TMA kernel family and the host side of TMA (tensor-map encoding).
// Shapes outside the specialisation of these kernels run on the SIMT kernels in generic.cu (same device,
// same semantics) -- never on a CPU path.
#include <cudaTypedefs.h>

#include "launchers.h"
#include "tc_common.cuh"

namespace nsa {

// ---- tensor maps (driver entry point fetched through the runtime: no link-time dependency on libcuda) -------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
  }
  return fn;
}

static CUtensorMapDataType tm_dtype(int dtype) {
  return dtype == NSA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

int make_tmap_rows(CUtensorMap* out, const void* base, int dtype, int D, int rows_present, int row_stride_elems,
                   long long slab_stride_elems, int slabs, int box_rows) {
  auto enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return NSA_ERR_CUDA; }
  NSA_REQUIRE(((uintptr_t)base & 15) == 0, "TMA needs 16-byte aligned tensors");
  NSA_REQUIRE(D * 2 == 128, "TMA tiles here are 128-byte rows (D=64, 2-byte elements), got D=%d", D);
  if (rows_present < 1) rows_present = 1;  // never read: every consumer masks by its own row count
  cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)rows_present, (cuuint64_t)slabs};
  cuuint64_t gstr[2] = {(cuuint64_t)row_stride_elems * 2, (cuuint64_t)slab_stride_elems * 2};
  cuuint32_t box[3] = {(cuuint32_t)D, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, tm_dtype(dtype), 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(rows) failed with CUresult %d", (int)r); return NSA_ERR_CUDA; }
  return NSA_OK;
}

int make_tmap_q(CUtensorMap* out, const void* base, int dtype, int D, int h, int G, long long n_tokens, int box_tokens) {
  auto enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return NSA_ERR_CUDA; }
  NSA_REQUIRE(((uintptr_t)base & 15) == 0, "TMA needs 16-byte aligned tensors");
  NSA_REQUIRE(D * 2 == 128, "TMA tiles here are 128-byte rows (D=64, 2-byte elements), got D=%d", D);
  // Q [tokens][G][h][D]: dims innermost first
  cuuint64_t gdim[4] = {(cuuint64_t)D, (cuuint64_t)h, (cuuint64_t)G, (cuuint64_t)n_tokens};
  cuuint64_t gstr[3] = {(cuuint64_t)D * 2, (cuuint64_t)h * D * 2, (cuuint64_t)G * h * D * 2};
  cuuint32_t box[4] = {(cuuint32_t)D, 1, 1, (cuuint32_t)box_tokens};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, tm_dtype(dtype), 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(Q) failed with CUresult %d", (int)r); return NSA_ERR_CUDA; }
  return NSA_OK;
}

// 4-D map over Q [tokens][G][h][D] whose box holds ALL heads of box_tokens tokens of one group: (D, h, 1, box_tokens)
// -> smem rows ordered (token, head), 128 B each.
int make_tmap_q_heads(CUtensorMap* out, const void* base, int dtype, int D, int h, int G, long long n_tokens, int box_tokens) {
  auto enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return NSA_ERR_CUDA; }
  NSA_REQUIRE(((uintptr_t)base & 15) == 0, "TMA needs 16-byte aligned tensors");
  NSA_REQUIRE(D * 2 == 128, "TMA tiles here are 128-byte rows (D=64, 2-byte elements), got D=%d", D);
  NSA_REQUIRE(h >= 1 && h <= 256 && box_tokens >= 1 && box_tokens <= 256, "Q box h=%d tokens=%d", h, box_tokens);
  cuuint64_t gdim[4] = {(cuuint64_t)D, (cuuint64_t)h, (cuuint64_t)G, (cuuint64_t)n_tokens};
  cuuint64_t gstr[3] = {(cuuint64_t)D * 2, (cuuint64_t)h * D * 2, (cuuint64_t)G * h * D * 2};
  cuuint32_t box[4] = {(cuuint32_t)D, (cuuint32_t)h, 1, (cuuint32_t)box_tokens};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, tm_dtype(dtype), 4, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(Q heads) failed with CUresult %d", (int)r); return NSA_ERR_CUDA; }
  return NSA_OK;
}

// ---- capability checks ---------------------------------------------------------------------------------------
bool tc_sel_supported(const nsa_dims_t& dm);
bool tc_dense_supported(const nsa_dims_t& dm, int branch);
int launch_dense_tc(const nsa_dims_t& dm, int branch, const void* Q, const void* K, const void* V, void* O, float* lse,
                    cudaStream_t stream);
int launch_sel_tc(const nsa_dims_t& dm, const void* Q, const void* K, const void* V, const int32_t* ranges, void* O, float* lse,
                  cudaStream_t stream);

bool tc_branch_supported(const nsa_dims_t& dm, int branch) {
  if (dm.impl == NSA_IMPL_SIMT) return false;
  if (branch == 1) return tc_sel_supported(dm);
  return tc_dense_supported(dm, branch);
}
bool tc_decode_supported(const nsa_dims_t& dm) { (void)dm; return false; }
int64_t tc_decode_workspace(const nsa_dims_t& dm) { (void)dm; return 0; }

int launch_branch_tc(const nsa_dims_t& dm, int branch, const void* Q, const void* K, const void* V, const int32_t* ranges,
                     void* O_b, float* lse_b, cudaStream_t stream) {
  if (branch == 1) return launch_sel_tc(dm, Q, K, V, ranges, O_b, lse_b, stream);
  return launch_dense_tc(dm, branch, Q, K, V, O_b, lse_b, stream);
}
int launch_decode_tc(const nsa_dims_t&, const void*, const void*, const void*, const void*, const void*, const void*,
                     const void*, const nsa_gate_params_t&, void*, int32_t*, void*, cudaStream_t) {
  set_error("tcgen05 decode kernel not built");
  return NSA_ERR_UNSUPPORTED;
}

}  // namespace nsa
