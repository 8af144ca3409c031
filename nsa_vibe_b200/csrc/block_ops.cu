// Caller-side row kernels of the transformer block around the hot path: RMSNorm fused with the residual add that
// feeds it and with the cast its consumers (the projections under autocast) would otherwise do, forward and backward.
// The reference's RMSNorm (nsa/model/llama_block_nsa.py:13-22) is six ATen passes forward (pow, mean, add, rsqrt, two multiplies)
// plus one fp32->bf16 cast per consuming nn.Linear, and about twice that backward; here a row is read once per direction.
//   forward:  s = x (+ r);  rstd = rsqrt(mean(s^2) + eps);  y = (s * rstd) * w            [HBM bound: 4..10 bytes per element]
//   backward: g = dy * w;  xh = s * rstd;  dx = rstd * (g - xh * mean(g * xh)) (+ ds);  dw = sum_rows dy * xh
// All arithmetic in fp32; a row's elements are rounded once on the way out.  dw is reduced in a fixed order (per-warp column
// sums -> per-CTA partials -> one column pass), so the result is bitwise reproducible for a given grid.
#include "common.cuh"
#include "launchers.h"

namespace nsa {

namespace {

__device__ __forceinline__ void ld4(const void* p, size_t i, int dtype, float v[4]) {
  if (dtype == NSA_F32) {
    const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(p) + i);
    if (dtype == NSA_BF16) {
      const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.x));
      const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&t.y));
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else {
      const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&t.x));
      const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&t.y));
      v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
  }
}

__device__ __forceinline__ void st4(void* p, size_t i, int dtype, const float v[4]) {
  if (dtype == NSA_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = make_float4(v[0], v[1], v[2], v[3]);
  } else {
    uint2 t;
    if (dtype == NSA_BF16) {
      *reinterpret_cast<__nv_bfloat162*>(&t.x) = __floats2bfloat162_rn(v[0], v[1]);
      *reinterpret_cast<__nv_bfloat162*>(&t.y) = __floats2bfloat162_rn(v[2], v[3]);
    } else {
      *reinterpret_cast<__half2*>(&t.x) = __floats2half2_rn(v[0], v[1]);
      *reinterpret_cast<__half2*>(&t.y) = __floats2half2_rn(v[2], v[3]);
    }
    *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(p) + i) = t;
  }
}

constexpr int kNormThreads = 256;
constexpr int kNormWarps = kNormThreads / 32;

// One warp per row (grid-stride over rows); lane owns the 4-element vectors lane, lane+32, ... of the row.
__global__ void __launch_bounds__(kNormThreads)
rmsnorm_fwd_kernel(const void* __restrict__ x, const void* __restrict__ r, const void* __restrict__ w, void* __restrict__ s_out,
                   void* __restrict__ y, float* __restrict__ rstd_out, int rows, int dim, float eps, int x_dtype, int r_dtype,
                   int w_dtype, int y_dtype) {
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kNormWarps + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kNormWarps;
  const int nvec = dim >> 2;
  const float inv_dim = 1.0f / (float)dim;
  for (int row = warp; row < rows; row += nwarps) {
    const size_t base = (size_t)row * dim;
    float ss = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      float a[4];
      ld4(x, base + 4 * v, x_dtype, a);
      if (r) {
        float b[4];
        ld4(r, base + 4 * v, r_dtype, b);
#pragma unroll
        for (int k = 0; k < 4; ++k) a[k] += b[k];
        if (x_dtype != NSA_F32) {  // the sum is what the caller keeps (s_out): the norm sees the rounded value, as in torch
          st4(s_out, base + 4 * v, x_dtype, a);
          ld4(s_out, base + 4 * v, x_dtype, a);
        } else {
          st4(s_out, base + 4 * v, x_dtype, a);
        }
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) ss = fmaf(a[k], a[k], ss);
    }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss * inv_dim + eps);
    if (lane == 0 && rstd_out) rstd_out[row] = rstd;
    const void* src = r ? s_out : x;  // second pass hits L1/L2: the row was just touched by this warp
    for (int v = lane; v < nvec; v += 32) {
      float a[4], g[4];
      ld4(src, base + 4 * v, x_dtype, a);
      ld4(w, 4 * v, w_dtype, g);
#pragma unroll
      for (int k = 0; k < 4; ++k) a[k] = (a[k] * rstd) * g[k];
      st4(y, base + 4 * v, y_dtype, a);
    }
  }
}

// dynamic shared memory: kNormWarps x dim fp32 column sums (dw of the rows each warp handled)
__global__ void __launch_bounds__(kNormThreads)
rmsnorm_bwd_kernel(const void* __restrict__ dy, const void* __restrict__ s, const void* __restrict__ w,
                   const float* __restrict__ rstd_in, const void* __restrict__ ds, void* __restrict__ dx,
                   float* __restrict__ dw_partial, int rows, int dim, int x_dtype, int w_dtype, int y_dtype) {
  extern __shared__ float acc[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * kNormWarps + wib;
  const int nwarps = gridDim.x * kNormWarps;
  const int nvec = dim >> 2;
  const float inv_dim = 1.0f / (float)dim;
  float* mine = acc + (size_t)wib * dim;
  for (int v = lane; v < nvec; v += 32) *reinterpret_cast<float4*>(mine + 4 * v) = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = warp; row < rows; row += nwarps) {
    const size_t base = (size_t)row * dim;
    const float rstd = rstd_in[row];
    float c = 0.f;
    for (int v = lane; v < nvec; v += 32) {
      float a[4], g[4], d[4];
      ld4(s, base + 4 * v, x_dtype, a);
      ld4(dy, base + 4 * v, y_dtype, d);
      ld4(w, 4 * v, w_dtype, g);
      float4 m = *reinterpret_cast<float4*>(mine + 4 * v);
      float* mp = reinterpret_cast<float*>(&m);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float xh = a[k] * rstd;
        c = fmaf(d[k] * g[k], xh, c);
        if (dw_partial) mp[k] = fmaf(d[k], xh, mp[k]);
      }
      *reinterpret_cast<float4*>(mine + 4 * v) = m;
    }
    c = warp_sum(c) * inv_dim;
    for (int v = lane; v < nvec; v += 32) {
      float a[4], g[4], d[4], o[4];
      ld4(s, base + 4 * v, x_dtype, a);
      ld4(dy, base + 4 * v, y_dtype, d);
      ld4(w, 4 * v, w_dtype, g);
#pragma unroll
      for (int k = 0; k < 4; ++k) o[k] = rstd * (d[k] * g[k] - (a[k] * rstd) * c);
      if (ds) {
        float e[4];
        ld4(ds, base + 4 * v, x_dtype, e);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k] += e[k];
      }
      st4(dx, base + 4 * v, x_dtype, o);
    }
  }
  if (!dw_partial) return;
  __syncthreads();
  for (int cidx = threadIdx.x; cidx < dim; cidx += kNormThreads) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kNormWarps; ++k) t += acc[(size_t)k * dim + cidx];
    dw_partial[(size_t)blockIdx.x * dim + cidx] = t;
  }
}

// Register-resident variants for dim <= 32*4*NV (NV <= 8, i.e. dim <= 1024), typed so that the load phase is straight-line code:
// a lane issues its NV 16-byte (8-byte for 16-bit types) loads of every operand of the row back to back, the row is read from
// HBM exactly once, and nothing waits on a second pass.  (With run-time dtype switches inside the unrolled loops ptxas kept one
// branch per load and the loads of a row went out one after the other: 65 us instead of 30 us for the backward at 16384 x 768.)
// The weight row is converted to fp32 in shared memory once per CTA.
template <typename T> struct Elt4;
template <> struct Elt4<float> {
  static __device__ __forceinline__ void ld(const void* p, size_t i, float v[4]) { ld4(p, i, NSA_F32, v); }
  static __device__ __forceinline__ void st(void* p, size_t i, const float v[4]) { st4(p, i, NSA_F32, v); }
  static __device__ __forceinline__ float rnd(float x) { return x; }
};
template <> struct Elt4<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const void* p, size_t i, float v[4]) { ld4(p, i, NSA_BF16, v); }
  static __device__ __forceinline__ void st(void* p, size_t i, const float v[4]) { st4(p, i, NSA_BF16, v); }
  static __device__ __forceinline__ float rnd(float x) { return __bfloat162float(__float2bfloat16(x)); }
};
template <> struct Elt4<__half> {
  static __device__ __forceinline__ void ld(const void* p, size_t i, float v[4]) { ld4(p, i, NSA_F16, v); }
  static __device__ __forceinline__ void st(void* p, size_t i, const float v[4]) { st4(p, i, NSA_F16, v); }
  static __device__ __forceinline__ float rnd(float x) { return __half2float(__float2half(x)); }
};

__device__ __forceinline__ void stage_weight(float* wsm, const void* w, int dim, int w_dtype) {
  for (int c = threadIdx.x; c < dim; c += blockDim.x) wsm[c] = ld_elt(w, c, w_dtype);
  __syncthreads();
}

// RT = dtype of the residual (XT or YT)
template <int NV, typename XT, typename YT, typename RT>
__global__ void __launch_bounds__(kNormThreads, 3)
rmsnorm_fwd_reg_kernel(const void* __restrict__ x, const void* __restrict__ r, const void* __restrict__ w, void* __restrict__ s_out,
                       void* __restrict__ y, float* __restrict__ rstd_out, int rows, int dim, float eps, int w_dtype) {
  __shared__ __align__(16) float wsm[128 * NV];
  stage_weight(wsm, w, dim, w_dtype);
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kNormWarps + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kNormWarps;
  const int nvec = dim >> 2;
  const float inv_dim = 1.0f / (float)dim;
  for (int row = warp; row < rows; row += nwarps) {
    const size_t base = (size_t)row * dim;
    float a[NV][4];
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (lane + 32 * j < nvec) Elt4<XT>::ld(x, base + 4 * (lane + 32 * j), a[j]);
    if (r) {
      float b[NV][4];
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (lane + 32 * j < nvec) Elt4<RT>::ld(r, base + 4 * (lane + 32 * j), b[j]);
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (lane + 32 * j < nvec) {
#pragma unroll
          for (int k = 0; k < 4; ++k) a[j][k] += b[j][k];
          Elt4<XT>::st(s_out, base + 4 * (lane + 32 * j), a[j]);
#pragma unroll
          for (int k = 0; k < 4; ++k) a[j][k] = Elt4<XT>::rnd(a[j][k]);  // the norm sees the stored running sum, as in torch
        }
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (lane + 32 * j < nvec) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ss = fmaf(a[j][k], a[j][k], ss);
      }
    ss = warp_sum(ss);
    const float rstd = rsqrtf(ss * inv_dim + eps);
    if (lane == 0 && rstd_out) rstd_out[row] = rstd;
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (lane + 32 * j < nvec) {
        const float4 g = *reinterpret_cast<const float4*>(wsm + 4 * (lane + 32 * j));
        float o[4] = {(a[j][0] * rstd) * g.x, (a[j][1] * rstd) * g.y, (a[j][2] * rstd) * g.z, (a[j][3] * rstd) * g.w};
        Elt4<YT>::st(y, base + 4 * (lane + 32 * j), o);
      }
  }
}

template <int NV, typename XT, typename YT>
__global__ void __launch_bounds__(kNormThreads, 2)
rmsnorm_bwd_reg_kernel(const void* __restrict__ dy, const void* __restrict__ s, const void* __restrict__ w,
                       const float* __restrict__ rstd_in, const void* __restrict__ ds, void* __restrict__ dx,
                       float* __restrict__ dw_partial, int rows, int dim, int w_dtype) {
  extern __shared__ float acc[];  // kNormWarps x dim column sums, then dim fp32 weights
  float* wsm = acc + (size_t)kNormWarps * dim;
  stage_weight(wsm, w, dim, w_dtype);
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * kNormWarps + wib;
  const int nwarps = gridDim.x * kNormWarps;
  const int nvec = dim >> 2;
  const float inv_dim = 1.0f / (float)dim;
  float dwr[NV][4];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
#pragma unroll
    for (int k = 0; k < 4; ++k) dwr[j][k] = 0.f;
  }
  for (int row = warp; row < rows; row += nwarps) {
    const size_t base = (size_t)row * dim;
    float a[NV][4], d[NV][4];  // after the first loop: a = xh, d = dy * w
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (lane + 32 * j < nvec) Elt4<XT>::ld(s, base + 4 * (lane + 32 * j), a[j]);
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (lane + 32 * j < nvec) Elt4<YT>::ld(dy, base + 4 * (lane + 32 * j), d[j]);
    const float rstd = rstd_in[row];
    float c = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (lane + 32 * j < nvec) {
        const float4 g4 = *reinterpret_cast<const float4*>(wsm + 4 * (lane + 32 * j));
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          a[j][k] *= rstd;  // xh
          dwr[j][k] = fmaf(d[j][k], a[j][k], dwr[j][k]);
          d[j][k] *= g[k];
          c = fmaf(d[j][k], a[j][k], c);
        }
      }
    c = warp_sum(c) * inv_dim;
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (lane + 32 * j < nvec) {
#pragma unroll
        for (int k = 0; k < 4; ++k) a[j][k] = rstd * (d[j][k] - a[j][k] * c);
      }
    if (ds) {
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (lane + 32 * j < nvec) Elt4<XT>::ld(ds, base + 4 * (lane + 32 * j), d[j]);
#pragma unroll
      for (int j = 0; j < NV; ++j)
        if (lane + 32 * j < nvec) {
#pragma unroll
          for (int k = 0; k < 4; ++k) a[j][k] += d[j][k];
        }
    }
#pragma unroll
    for (int j = 0; j < NV; ++j)
      if (lane + 32 * j < nvec) Elt4<XT>::st(dx, base + 4 * (lane + 32 * j), a[j]);
  }
  if (!dw_partial) return;
#pragma unroll
  for (int j = 0; j < NV; ++j)
    if (lane + 32 * j < nvec)
      *reinterpret_cast<float4*>(acc + (size_t)wib * dim + 4 * (lane + 32 * j)) = make_float4(dwr[j][0], dwr[j][1], dwr[j][2], dwr[j][3]);
  __syncthreads();
  for (int cidx = threadIdx.x; cidx < dim; cidx += kNormThreads) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kNormWarps; ++k) t += acc[(size_t)k * dim + cidx];
    dw_partial[(size_t)blockIdx.x * dim + cidx] = t;
  }
}

// typed dispatch of the register-resident kernels; false = no specialisation for this (dim, dtypes): the generic kernels run
template <typename XT, typename YT, typename RT>
bool fwd_reg_nv(int nv, int grid, cudaStream_t st, const void* x, const void* r, const void* w, void* s_out, void* y, float* rstd,
                int rows, int dim, float eps, int w_dtype) {
  if (nv <= 2) rmsnorm_fwd_reg_kernel<2, XT, YT, RT><<<grid, kNormThreads, 0, st>>>(x, r, w, s_out, y, rstd, rows, dim, eps, w_dtype);
  else if (nv <= 4) rmsnorm_fwd_reg_kernel<4, XT, YT, RT><<<grid, kNormThreads, 0, st>>>(x, r, w, s_out, y, rstd, rows, dim, eps, w_dtype);
  else if (nv <= 6) rmsnorm_fwd_reg_kernel<6, XT, YT, RT><<<grid, kNormThreads, 0, st>>>(x, r, w, s_out, y, rstd, rows, dim, eps, w_dtype);
  else if (nv <= 8) rmsnorm_fwd_reg_kernel<8, XT, YT, RT><<<grid, kNormThreads, 0, st>>>(x, r, w, s_out, y, rstd, rows, dim, eps, w_dtype);
  else return false;
  return true;
}

template <typename XT, typename YT>
bool fwd_reg_r(int r_dtype, int x_dtype, int nv, int grid, cudaStream_t st, const void* x, const void* r, const void* w, void* s_out,
               void* y, float* rstd, int rows, int dim, float eps, int w_dtype) {
  if (!r || r_dtype == x_dtype) return fwd_reg_nv<XT, YT, XT>(nv, grid, st, x, r, w, s_out, y, rstd, rows, dim, eps, w_dtype);
  return fwd_reg_nv<XT, YT, YT>(nv, grid, st, x, r, w, s_out, y, rstd, rows, dim, eps, w_dtype);  // caller checked r_dtype == y_dtype
}

template <typename XT, typename YT>
bool bwd_reg_nv(int nv, int grid, size_t smem, cudaStream_t st, const void* dy, const void* s, const void* w, const float* rstd,
                const void* ds, void* dx, float* part, int rows, int dim, int w_dtype) {
  if (nv <= 2) rmsnorm_bwd_reg_kernel<2, XT, YT><<<grid, kNormThreads, smem, st>>>(dy, s, w, rstd, ds, dx, part, rows, dim, w_dtype);
  else if (nv <= 4) rmsnorm_bwd_reg_kernel<4, XT, YT><<<grid, kNormThreads, smem, st>>>(dy, s, w, rstd, ds, dx, part, rows, dim, w_dtype);
  else if (nv <= 6) rmsnorm_bwd_reg_kernel<6, XT, YT><<<grid, kNormThreads, smem, st>>>(dy, s, w, rstd, ds, dx, part, rows, dim, w_dtype);
  else if (nv <= 8) rmsnorm_bwd_reg_kernel<8, XT, YT><<<grid, kNormThreads, smem, st>>>(dy, s, w, rstd, ds, dx, part, rows, dim, w_dtype);
  else return false;
  return true;
}

// dw[c] = sum_p partial[p, c]: block = 32 columns x 8 slices of the partials, fixed summation order
__global__ void __launch_bounds__(256)
rmsnorm_dw_kernel(const float* __restrict__ partial, void* __restrict__ dw, int nparts, int dim, int w_dtype) {
  __shared__ float red[8][33];
  const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + cx;
  float t = 0.f;
  if (c < dim)
    for (int p = py; p < nparts; p += 8) t += partial[(size_t)p * dim + c];
  red[py][cx] = t;
  __syncthreads();
  if (py == 0 && c < dim) {
    float u = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) u += red[k][cx];
    st_elt(dw, c, w_dtype, u);
  }
}

int norm_grid(int rows, int sms) {
  const int want = ceil_div(rows, kNormWarps);
  const int cap = 4 * sms;
  return want < cap ? (want > 0 ? want : 1) : cap;
}

bool ok_dtype(int d) { return d == NSA_F32 || d == NSA_BF16 || d == NSA_F16; }

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace

int rmsnorm_partials(int rows) { return norm_grid(rows, sm_count()); }

int launch_rmsnorm_fwd(const void* x, const void* r, const void* w, void* s_out, void* y, float* rstd, int rows, int dim, float eps,
                       int x_dtype, int r_dtype, int w_dtype, int y_dtype, cudaStream_t stream) {
  NSA_REQUIRE(x && w && y, "rmsnorm_fwd: NULL pointer");
  NSA_REQUIRE(!r || s_out, "rmsnorm_fwd: a residual needs s_out");
  NSA_REQUIRE(rows >= 0 && dim >= 4 && dim % 4 == 0, "rmsnorm_fwd: rows=%d dim=%d (dim must be a multiple of 4)", rows, dim);
  NSA_REQUIRE(ok_dtype(x_dtype) && ok_dtype(w_dtype) && ok_dtype(y_dtype) && (!r || ok_dtype(r_dtype)), "rmsnorm_fwd: dtype");
  if (rows == 0) return NSA_OK;
  const int grid = norm_grid(rows, sm_count());
  const int nv = ceil_div(dim / 4, 32);
  bool done = false;
  if (!r || r_dtype == x_dtype || r_dtype == y_dtype) {
#define NSA_NORM_FWD(XD, YD, XT, YT)               \
  if (!done && x_dtype == XD && y_dtype == YD)     \
    done = fwd_reg_r<XT, YT>(r_dtype, x_dtype, nv, grid, stream, x, r, w, s_out, y, rstd, rows, dim, eps, w_dtype);
    NSA_NORM_FWD(NSA_F32, NSA_BF16, float, __nv_bfloat16)
    NSA_NORM_FWD(NSA_F32, NSA_F32, float, float)
    NSA_NORM_FWD(NSA_BF16, NSA_BF16, __nv_bfloat16, __nv_bfloat16)
    NSA_NORM_FWD(NSA_F16, NSA_F16, __half, __half)
    NSA_NORM_FWD(NSA_F32, NSA_F16, float, __half)
#undef NSA_NORM_FWD
  }
  if (!done)
    rmsnorm_fwd_kernel<<<grid, kNormThreads, 0, stream>>>(x, r, w, s_out, y, rstd, rows, dim, eps, x_dtype, r_dtype, w_dtype, y_dtype);
  return check_launch("rmsnorm_fwd_kernel");
}

int launch_rmsnorm_bwd(const void* dy, const void* s, const void* w, const float* rstd, const void* ds, void* dx, void* dw,
                       float* dw_partial, int rows, int dim, int x_dtype, int w_dtype, int y_dtype, cudaStream_t stream) {
  NSA_REQUIRE(dy && s && w && rstd && dx, "rmsnorm_bwd: NULL pointer");
  NSA_REQUIRE(!dw || dw_partial, "rmsnorm_bwd: dw needs the dw_partial workspace ([nsa_rmsnorm_partials(rows), dim] fp32)");
  NSA_REQUIRE(rows >= 0 && dim >= 4 && dim % 4 == 0, "rmsnorm_bwd: rows=%d dim=%d (dim must be a multiple of 4)", rows, dim);
  NSA_REQUIRE(ok_dtype(x_dtype) && ok_dtype(w_dtype) && ok_dtype(y_dtype), "rmsnorm_bwd: dtype");
  const size_t smem = (size_t)kNormWarps * dim * sizeof(float);
  NSA_REQUIRE(smem <= 200 * 1024, "rmsnorm_bwd: dim=%d too wide for the column-sum staging", dim);
  const int grid = norm_grid(rows, sm_count());
  if (rows > 0) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(rmsnorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) { set_error("rmsnorm_bwd: smem attr: %s", cudaGetErrorString(e)); return NSA_ERR_CUDA; }
    }
    float* part = dw ? dw_partial : nullptr;
    const int nv = ceil_div(dim / 4, 32);
    const size_t smem_reg = smem + (size_t)dim * sizeof(float);
    bool done = false;
#define NSA_NORM_BWD(XD, YD, XT, YT)           \
  if (!done && x_dtype == XD && y_dtype == YD) \
    done = bwd_reg_nv<XT, YT>(nv, grid, smem_reg, stream, dy, s, w, rstd, ds, dx, part, rows, dim, w_dtype);
    NSA_NORM_BWD(NSA_F32, NSA_BF16, float, __nv_bfloat16)
    NSA_NORM_BWD(NSA_F32, NSA_F32, float, float)
    NSA_NORM_BWD(NSA_BF16, NSA_BF16, __nv_bfloat16, __nv_bfloat16)
    NSA_NORM_BWD(NSA_F16, NSA_F16, __half, __half)
    NSA_NORM_BWD(NSA_F32, NSA_F16, float, __half)
#undef NSA_NORM_BWD
    if (!done) rmsnorm_bwd_kernel<<<grid, kNormThreads, smem, stream>>>(dy, s, w, rstd, ds, dx, part, rows, dim, x_dtype, w_dtype, y_dtype);
    const int rc = check_launch("rmsnorm_bwd_kernel");
    if (rc != NSA_OK) return rc;
  }
  if (dw) {
    rmsnorm_dw_kernel<<<ceil_div(dim, 32), 256, 0, stream>>>(dw_partial, dw, rows > 0 ? grid : 0, dim, w_dtype);
    return check_launch("rmsnorm_dw_kernel");
  }
  return NSA_OK;
}

}  // namespace nsa
