// Standalone selection kernels: p_grp (fp32, HBM) -> ranges (int32).  HBM-bound: 4*S_sel bytes read
// and 8*K bytes written per row; one warp per row, rows strided over a grid sized to the machine.
#include "select.cuh"
#include "launchers.h"

namespace nsa {

constexpr int kSelWarps = 8;

__global__ void __launch_bounds__(kSelWarps * 32, 4)
select_kernel(const float* __restrict__ p_grp, int n_rows, int S_rows, int G, int S_sel, int l_sel, int n_sel,
              int mode, int nf, int K, int t0, int32_t* __restrict__ ranges) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* sc = smem + (size_t)warp * S_sel;
  for (int row = blockIdx.x * kSelWarps + warp; row < n_rows; row += gridDim.x * kSelWarps) {
    // rows are (b, s, g) with t = t0 + s under either rule; a decode step has S_rows = 1, i.e. t = t0
    const int t = t0 + (row / G) % S_rows;
    const float* src = p_grp + (size_t)row * S_sel;
    if (S_sel > 128 && S_sel <= 1024) {  // warp-uniform
      select_row_warp_1024(src, sc, S_sel, l_sel, n_sel, mode, nf, K, t, ranges + (size_t)row * K * 2);
    } else {
      for (int j = lane; j < S_sel; j += 32) sc[j] = __ldg(src + j);
      __syncwarp();
      select_row_warp(sc, S_sel, l_sel, n_sel, mode, nf, K, t, ranges + (size_t)row * K * 2);
    }
    __syncwarp();
  }
}

int launch_select(const float* p_grp, int n_rows, int S_rows, int G, int S_sel, int l_sel, int n_sel, int mode, int nf,
                  int K, int t0, int32_t* ranges, cudaStream_t stream) {
  if (n_rows == 0 || K == 0) return NSA_OK;
  NSA_REQUIRE(S_sel >= 1 && S_sel <= kSelMaxWords * 1024, "select: S_sel=%d outside [1,%d]", S_sel, kSelMaxWords * 1024);
  size_t smem = (size_t)kSelWarps * S_sel * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("select: smem attr: %s", cudaGetErrorString(e)); return NSA_ERR_CUDA; }
  }
  int blocks = ceil_div(n_rows, kSelWarps);
  int max_blocks = 148 * 8;
  if (blocks > max_blocks) blocks = max_blocks;
  select_kernel<<<blocks, kSelWarps * 32, smem, stream>>>(p_grp, n_rows, S_rows, G, S_sel, l_sel, n_sel, mode, nf, K, t0,
                                                         ranges);
  return check_launch("select_kernel");
}

}  // namespace nsa
