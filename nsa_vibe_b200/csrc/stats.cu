// Observability reductions on the device: the gate-health statistics of _compute_gate_stats
// (nsa/core/nsa_attention.py:127-165) and the selection statistics of _update_sel_stats_from_ranges (:455-507) as one pass each over
// the tensors the hot path already produced (gates [rows,3] fp32, ranges [rows,K,2] int32), leaving ONE small record for the host to
// read -- the reference runs ~15 ATen kernels and six .item() synchronisations per call.  HBM bound: 12 / 8K bytes per row.
// Per-row arithmetic is fp32 in the reference's order (entropy = -(g0*log(g0+1e-8) + g1*log(g1+1e-8) + g2*log(g2+1e-8))); sums are
// accumulated in fp64, so their order does not show in the fp32 results.
#include "common.cuh"
#include "launchers.h"

namespace nsa {

namespace {

// total order on floats as signed ints (for atomicMin / atomicMax)
__device__ __forceinline__ int f2ord(float f) {
  const int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void stats_init_kernel(nsa_stats_t* out) {
  if (threadIdx.x == 0) {
    for (int i = 0; i < 6; ++i) out->gate_sum[i] = 0.0;
    out->entropy_min_ord = 0x7fffffff;
    out->max_gate_max_ord = (int)0x80000000;
    out->k_sum = 0;
    out->k_max = 0;
    out->rows_at_max = 0;
  }
}

__global__ void __launch_bounds__(256)
gate_stats_kernel(const float* __restrict__ gates, long long n_rows, nsa_stats_t* out) {
  double s_ent = 0.0, s_max = 0.0, s_col = 0.0, s_g0 = 0.0, s_g1 = 0.0, s_g2 = 0.0;
  float e_min = INFINITY, m_max = -INFINITY;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
    const float g0 = gates[3 * r], g1 = gates[3 * r + 1], g2 = gates[3 * r + 2];
    const float t0 = __fmul_rn(g0, logf(__fadd_rn(g0, 1e-8f))), t1 = __fmul_rn(g1, logf(__fadd_rn(g1, 1e-8f))),
                t2 = __fmul_rn(g2, logf(__fadd_rn(g2, 1e-8f)));
    const float ent = -__fadd_rn(__fadd_rn(t0, t1), t2);
    const float mx = fmaxf(g0, fmaxf(g1, g2));
    s_ent += ent;
    s_max += mx;
    s_col += (ent < 0.1f && mx > 0.95f) ? 1.0 : 0.0;
    s_g0 += g0;
    s_g1 += g1;
    s_g2 += g2;
    e_min = fminf(e_min, ent);
    m_max = fmaxf(m_max, mx);
  }
  s_ent = warp_sum_d(s_ent);
  s_max = warp_sum_d(s_max);
  s_col = warp_sum_d(s_col);
  s_g0 = warp_sum_d(s_g0);
  s_g1 = warp_sum_d(s_g1);
  s_g2 = warp_sum_d(s_g2);
  e_min = -warp_max(-e_min);
  m_max = warp_max(m_max);
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(&out->gate_sum[0], s_ent);
    atomicAdd(&out->gate_sum[1], s_max);
    atomicAdd(&out->gate_sum[2], s_col);
    atomicAdd(&out->gate_sum[3], s_g0);
    atomicAdd(&out->gate_sum[4], s_g1);
    atomicAdd(&out->gate_sum[5], s_g2);
    if (e_min < INFINITY) atomicMin(&out->entropy_min_ord, f2ord(e_min));
    if (m_max > -INFINITY) atomicMax(&out->max_gate_max_ord, f2ord(m_max));
  }
}

// L[row] = sum_k max(end - start, 0); sum and max over rows
__global__ void __launch_bounds__(256)
sel_stats_kernel(const int32_t* __restrict__ ranges, long long n_rows, int K, int32_t* __restrict__ L, nsa_stats_t* out) {
  long long s = 0;
  int mx = 0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
    const int2* p = reinterpret_cast<const int2*>(ranges) + r * K;
    int len = 0;
    for (int k = 0; k < K; ++k) {
      const int2 v = p[k];
      len += v.y > v.x ? v.y - v.x : 0;
    }
    L[r] = len;
    s += len;
    mx = len > mx ? len : mx;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    const int m2 = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = m2 > mx ? m2 : mx;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(&out->k_sum), (unsigned long long)s);
    atomicMax(&out->k_max, mx);
  }
}

__global__ void __launch_bounds__(256)
sel_stats_count_kernel(const int32_t* __restrict__ L, long long n_rows, nsa_stats_t* out) {
  const int kmax = out->k_max;
  long long c = 0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) c += L[r] == kmax;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(reinterpret_cast<unsigned long long*>(&out->rows_at_max), (unsigned long long)c);
}

// blocks[0] = max over rows of the number of 64-key blocks the tcgen05 selected-branch kernels cut the row's ranges into
// (each range clamped to [0, S_kv) and cut separately, as tc_gather.cu / tc_sel2.cu do)
__global__ void __launch_bounds__(256)
ranges_max_blocks_kernel(const int32_t* __restrict__ ranges, long long n_rows, int K, int S_kv, int32_t* __restrict__ out) {
  int mx = 0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < n_rows; r += (long long)gridDim.x * blockDim.x) {
    const int2* rr = reinterpret_cast<const int2*>(ranges + r * K * 2);
    int n = 0;
    for (int k = 0; k < K; ++k) {
      int2 v = rr[k];
      if (v.x < 0) v.x = 0;
      if (v.y > S_kv) v.y = S_kv;
      if (v.y > v.x) n += (v.y - v.x + 63) >> 6;
    }
    mx = n > mx ? n : mx;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int m2 = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = m2 > mx ? m2 : mx;
  }
  if ((threadIdx.x & 31) == 0 && mx > 0) atomicMax(out, mx);
}

int grid_for(long long n) {
  long long g = (n + 255) / 256;
  return (int)(g < 1 ? 1 : (g > 148 * 8 ? 148 * 8 : g));
}

}  // namespace

int launch_stats(const float* gates, long long n_gate_rows, const int32_t* ranges, long long n_range_rows, int K, int32_t* row_len,
                 nsa_stats_t* out, cudaStream_t stream) {
  NSA_REQUIRE(out, "stats: NULL output record");
  NSA_REQUIRE(n_gate_rows >= 0 && n_range_rows >= 0 && K >= 0, "stats: negative size");
  NSA_REQUIRE(!(ranges && n_range_rows > 0) || row_len, "stats: selection statistics need the row_len workspace (n_range_rows int32)");
  stats_init_kernel<<<1, 32, 0, stream>>>(out);
  if (int rc = check_launch("stats_init_kernel")) return rc;
  if (gates && n_gate_rows > 0) {
    gate_stats_kernel<<<grid_for(n_gate_rows), 256, 0, stream>>>(gates, n_gate_rows, out);
    if (int rc = check_launch("gate_stats_kernel")) return rc;
  }
  if (ranges && n_range_rows > 0) {
    sel_stats_kernel<<<grid_for(n_range_rows), 256, 0, stream>>>(ranges, n_range_rows, K, row_len, out);
    if (int rc = check_launch("sel_stats_kernel")) return rc;
    sel_stats_count_kernel<<<grid_for(n_range_rows), 256, 0, stream>>>(row_len, n_range_rows, out);
    if (int rc = check_launch("sel_stats_count_kernel")) return rc;
  }
  return NSA_OK;
}

int launch_ranges_max_blocks(const int32_t* ranges, long long n_rows, int K, int S_kv, int32_t* out, cudaStream_t stream) {
  NSA_REQUIRE(out && (ranges || n_rows == 0) && n_rows >= 0 && K >= 0, "ranges_max_blocks: bad arguments");
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(int32_t), stream);
  if (e != cudaSuccess) { set_error("ranges_max_blocks: cudaMemsetAsync: %s", cudaGetErrorString(e)); return NSA_ERR_CUDA; }
  if (n_rows == 0 || K == 0) return NSA_OK;
  ranges_max_blocks_kernel<<<grid_for(n_rows), 256, 0, stream>>>(ranges, n_rows, K, S_kv, out);
  return check_launch("ranges_max_blocks_kernel");
}

}  // namespace nsa
