// Shared helpers for the NSA sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>
#include <atomic>

#include "../../include/nsa_b200.h"

namespace nsa {

constexpr int kWarp = 32;
constexpr float kLog2e = 1.4426950408889634f;

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define NSA_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      nsa::set_error(__VA_ARGS__);             \
      return NSA_ERR_INVALID_ARG;              \
    }                                          \
  } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// cudaFuncSetAttribute is per device (context): `done` remembers, one bit per device ordinal, where the opt-in has been
// made, so a process that drives several GPUs (or several host threads) gets it on each.  max_carveout also asks for the
// largest shared-memory carve-out, which a kernel needs when TWO of its CTAs are meant to share an SM: the driver otherwise
// sizes the carve-out for one (ncu: launch__shared_mem_config_size 135 KB, launch__occupancy_limit_shared_mem 1).
template <typename Kern>
inline int ensure_smem_attr(Kern kern, int bytes, std::atomic<unsigned long long>& done, const char* who, bool max_carveout = false) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { set_error("%s: cudaGetDevice: %s", who, cudaGetErrorString(e)); return NSA_ERR_CUDA; }
  const unsigned long long bit = dev >= 0 && dev < 64 ? 1ull << dev : 0ull;
  if (bit && (done.load(std::memory_order_acquire) & bit)) return NSA_OK;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && max_carveout)
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) { set_error("%s: shared-memory attribute: %s", who, cudaGetErrorString(e)); return NSA_ERR_CUDA; }
  if (bit) done.fetch_or(bit, std::memory_order_release);
  return NSA_OK;
}

// num_cmp(t): compressed tokens visible at absolute position t (nsa/core/packing.py:15-23)
__host__ __device__ inline int num_cmp_at(int t, int l, int d, int S_cmp) {
  if (t + 1 < l) return 0;
  int n = (t + 1 - l) / d + 1;
  return n < S_cmp ? n : S_cmp;
}

__host__ __device__ inline size_t elt_size(int dtype) { return dtype == NSA_F32 ? 4 : 2; }

// dtype-erased element access (SIMT fallback kernels only; the tcgen05 kernels are typed)
__device__ __forceinline__ float ld_elt(const void* p, size_t i, int dtype) {
  if (dtype == NSA_F32) return reinterpret_cast<const float*>(p)[i];
  if (dtype == NSA_BF16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]);
  return __half2float(reinterpret_cast<const __half*>(p)[i]);
}
__device__ __forceinline__ void st_elt(void* p, size_t i, int dtype, float v) {
  if (dtype == NSA_F32) reinterpret_cast<float*>(p)[i] = v;
  else if (dtype == NSA_BF16) reinterpret_cast<__nv_bfloat16*>(p)[i] = __float2bfloat16(v);
  else reinterpret_cast<__half*>(p)[i] = __float2half(v);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace nsa
