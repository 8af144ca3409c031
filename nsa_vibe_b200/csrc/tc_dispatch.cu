// Dispatch into the tcgen05 / TMA kernel family and the host side of TMA (tensor-map encoding).
// Shapes outside the specialisation of these kernels run on the SIMT kernels in generic.cu (same device,
// same semantics) -- never on a CPU path.
#include <cudaTypedefs.h>
#include <stdlib.h>

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

// 3-D map over a [tiles][rows_per_tile][D] row-major tensor (2-byte elements, 128-byte rows) with 128-B swizzle: box =
// (D, box_rows, 1).  Used for STORES of whole warps' row groups: rows of a box beyond rows_per_tile are clipped by the TMA unit.
int make_tmap_tiles(CUtensorMap* out, const void* base, int dtype, int D, int rows_per_tile, long long n_tiles, int box_rows) {
  auto enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return NSA_ERR_CUDA; }
  NSA_REQUIRE(((uintptr_t)base & 15) == 0, "TMA needs 16-byte aligned tensors");
  NSA_REQUIRE(D * 2 == 128, "TMA tiles here are 128-byte rows (D=64, 2-byte elements), got D=%d", D);
  if (n_tiles < 1) n_tiles = 1;
  cuuint64_t gdim[3] = {(cuuint64_t)D, (cuuint64_t)rows_per_tile, (cuuint64_t)n_tiles};
  cuuint64_t gstr[2] = {(cuuint64_t)D * 2, (cuuint64_t)rows_per_tile * D * 2};
  cuuint32_t box[3] = {(cuuint32_t)D, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, tm_dtype(dtype), 3, const_cast<void*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled(tiles) failed with CUresult %d", (int)r); return NSA_ERR_CUDA; }
  return NSA_OK;
}

// ---- capability checks ---------------------------------------------------------------------------------------
bool tc_dense_supported(const nsa_dims_t& dm, int branch);
bool tc_gather_supported(const nsa_dims_t& dm, int branch_mask);
int launch_gather_tc(const nsa_dims_t& dm, int branch_mask, const void* Q, const void* const* K, const void* const* V,
                     const int32_t* ranges, void* const* O_br, float* const* lse, const float* gates, void* O,
                     const nsa_gate_params_t* gp_fuse, int S_sel, int32_t* ranges_out, cudaStream_t stream,
                     const nsa_decode_state_t* state = nullptr);
bool tc_gather_fuse_supported(const nsa_dims_t& dm, int S_sel);
int launch_dense_tc(const nsa_dims_t& dm, int branch, const void* Q, const void* K, const void* V, void* O, float* lse,
                    cudaStream_t stream);

bool tc_branch_supported(const nsa_dims_t& dm, int branch) {
  if (dm.impl == NSA_IMPL_SIMT) return false;
  if (branch == 1) return tc_gather_supported(dm, 2);
  return tc_dense_supported(dm, branch);
}
// ---- decode step on the tensor-core gather kernel ---------------------------------------------------------------
// workspace layout: [ranges: B*G*n_sel*2 int32 (used when ranges_out is NULL)] [gates: B*G*3 fp32]
bool tc_decode_supported(const nsa_dims_t& dm) { return dm.S == 1 && tc_gather_supported(dm, 7); }
int64_t tc_decode_workspace(const nsa_dims_t& dm) { return (int64_t)dm.B * dm.G * 3 * sizeof(float) + 16; }

int launch_decode_tc(const nsa_dims_t& dm, const void* Q, const void* K_sel, const void* V_sel, const void* K_win,
                     const void* V_win, const void* K_cmp, const void* V_cmp, const nsa_gate_params_t& gp, void* O,
                     int32_t* ranges_out, void* workspace, cudaStream_t stream) {
  NSA_REQUIRE(workspace, "decode_fwd(tc): needs a workspace of nsa_workspace_bytes(NSA_WS_DECODE) bytes");
  const size_t rbytes = ((size_t)dm.B * dm.G * dm.n_sel * 2 * sizeof(int32_t) + 15) & ~(size_t)15;
  int32_t* ranges = ranges_out ? ranges_out : reinterpret_cast<int32_t*>(workspace);
  float* gates = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + rbytes);
  const int t = dm.t0;
  const int cover = t + 1 > dm.l_sel ? t + 1 : dm.l_sel;  // meta covers max(t+1, l_sel) tokens (nsa_attention.py:609-632)
  const int S_sel = ceil_div(cover, dm.l_sel);
  const void* Ks[3] = {K_cmp, K_sel, K_win};
  const void* Vs[3] = {V_cmp, V_sel, V_win};
  static const bool no_fuse = getenv("NSA_B200_DECODE_UNFUSED") != nullptr;  // A/B switch (benchmarks only)
  if (!no_fuse && tc_gather_fuse_supported(dm, S_sel))  // one launch: scoring, selection, gate, three branches, combine
    return launch_gather_tc(dm, 7, Q, Ks, Vs, nullptr, nullptr, nullptr, nullptr, O, &gp, S_sel, ranges_out, stream);
  nsa_dims_t d2 = dm;
  d2.norm_mode = NSA_NORM_FULL_ROW;  // decode sees only emitted blocks: softmax over all of them (:650-651)
  if (int rc = launch_score_generic(d2, Q, K_cmp, S_sel, t + 1, 1, dm.n_sel, nullptr, ranges, stream)) return rc;
  if (int rc = launch_gate_fwd(dm, Q, gp, gates, stream)) return rc;
  return launch_gather_tc(dm, 7, Q, Ks, Vs, ranges, nullptr, nullptr, gates, O, nullptr, 0, nullptr, stream);
}

// Device-stepped decode (CUDA-graph replay): the kernel reads position and row counts from the device record, so everything is
// validated against the slab CAPACITIES here: the fused scorer ranks <= kGMaxSel selection blocks and <= 1024 compressed keys.
bool tc_decode_stepped_supported(const nsa_dims_t& dm) {
  if (dm.S != 1 || dm.cap_sel < 1 || dm.cap_win < 1 || dm.cap_cmp < 1) return false;
  nsa_dims_t d2 = dm;
  d2.S_sel_kv = dm.cap_sel; d2.S_win_kv = dm.cap_win; d2.S_cmp = dm.cap_cmp; d2.t0 = dm.cap_sel - 1;
  const int S_sel_max = ceil_div(dm.cap_sel > dm.l_sel ? dm.cap_sel : dm.l_sel, dm.l_sel);
  return tc_gather_fuse_supported(d2, S_sel_max);
}

int launch_decode_tc_stepped(const nsa_dims_t& dm, const void* Q, const void* K_sel, const void* V_sel, const void* K_win,
                             const void* V_win, const void* K_cmp, const void* V_cmp, const nsa_gate_params_t& gp, void* O,
                             int32_t* ranges_out, const nsa_decode_state_t* state, cudaStream_t stream) {
  const void* Ks[3] = {K_cmp, K_sel, K_win};
  const void* Vs[3] = {V_cmp, V_sel, V_win};
  return launch_gather_tc(dm, 7, Q, Ks, Vs, nullptr, nullptr, nullptr, nullptr, O, &gp, 1, ranges_out, stream, state);
}

int launch_branch_tc(const nsa_dims_t& dm, int branch, const void* Q, const void* K, const void* V, const int32_t* ranges,
                     void* O_b, float* lse_b, cudaStream_t stream) {
  if (branch == 1) {
    const void* Ks[3] = {nullptr, K, nullptr};
    const void* Vs[3] = {nullptr, V, nullptr};
    void* Os[3] = {nullptr, O_b, nullptr};
    float* Ls[3] = {nullptr, lse_b, nullptr};
    return launch_gather_tc(dm, 2, Q, Ks, Vs, ranges, Os, Ls, nullptr, nullptr, nullptr, 0, nullptr, stream);
  }
  return launch_dense_tc(dm, branch, Q, K, V, O_b, lse_b, stream);
}

}  // namespace nsa
