// Producers of the hot path's inputs: RoPE fused with the [B,S,V*D] <-> [B,V,S,D] re-layout the caches use, and
// the phi average pool over RoPE(K_raw) / V_raw that emits the compressed tokens.  The reference does these with a dozen ATen
// kernels per tensor (rope.py:16-51: pow, sin, cos, two casts, reshape, four multiplies, two add/sub, stack; then
// permute().contiguous(), nsa_attention.py:403-405; avg_pool1d between two transposes, compress_pool.py:9-38); here each is one
// HBM-bound pass.  Numerics follow the reference's order: angle = (pos / scale) * base^(-2i/dim) in fp32, sin/cos rounded to
// the tensor's dtype, every product and sum rounded to that dtype; the pool accumulates in fp32 and divides by l.
#include "common.cuh"
#include "launchers.h"

namespace nsa {

template <typename T> struct PrT;
template <> struct PrT<float> {
  static __device__ __forceinline__ void ld2(const float* p, float& a, float& b) {
    const float2 t = *reinterpret_cast<const float2*>(p);
    a = t.x;
    b = t.y;
  }
  static __device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }
  static __device__ __forceinline__ float ld(const float* p) { return *p; }
  static __device__ __forceinline__ float rnd(float v) { return v; }
  static __device__ __forceinline__ void st(float* p, float v) { *p = v; }
};
template <> struct PrT<__nv_bfloat16> {
  static __device__ __forceinline__ void ld2(const __nv_bfloat16* p, float& a, float& b) {
    const float2 t = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(p));
    a = t.x;
    b = t.y;
  }
  static __device__ __forceinline__ void st2(__nv_bfloat16* p, float a, float b) {
    *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b);
  }
  static __device__ __forceinline__ float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
  static __device__ __forceinline__ float rnd(float v) { return __bfloat162float(__float2bfloat16(v)); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16(v); }
};
template <> struct PrT<__half> {
  static __device__ __forceinline__ void ld2(const __half* p, float& a, float& b) {
    const float2 t = __half22float2(*reinterpret_cast<const __half2*>(p));
    a = t.x;
    b = t.y;
  }
  static __device__ __forceinline__ void st2(__half* p, float a, float b) { *reinterpret_cast<__half2*>(p) = __floats2half2_rn(a, b); }
  static __device__ __forceinline__ float ld(const __half* p) { return __half2float(*p); }
  static __device__ __forceinline__ float rnd(float v) { return __half2float(__float2half(v)); }
  static __device__ __forceinline__ void st(__half* p, float v) { *p = __float2half(v); }
};

// sin / cos of the rotation of pair `pair` (of a vector of width dim) at position pos, rounded to T like sin.to(x.dtype)
template <typename T>
__device__ __forceinline__ void rope_sincos(int pos, int pair, int dim, float base, float scale, float& sn, float& cs) {
  // ATen divides a tensor by a scalar as a multiplication by its fp32 reciprocal (BinaryDivTrueKernel.cu): mirrored here so
  // that fp32 results agree with the reference's `base ** (-2 * idx / dim)` and `pos / scale` to the last bit
  const float inv_freq = powf(base, __fmul_rn(__fmul_rn(-2.0f, (float)pair), __fdiv_rn(1.0f, (float)dim)));
  const float ang = __fmul_rn(__fmul_rn((float)pos, __fdiv_rn(1.0f, scale)), inv_freq);
  sn = PrT<T>::rnd(sinf(ang));
  cs = PrT<T>::rnd(cosf(ang));
}

// (x0, x1) -> rotated pair; inverse applies the transposed rotation (the backward of the forward one)
template <typename T>
__device__ __forceinline__ void rope_rotate(float x0, float x1, float sn, float cs, bool inverse, float& y0, float& y1) {
  if (inverse) sn = -sn;
  y0 = PrT<T>::rnd(__fsub_rn(PrT<T>::rnd(__fmul_rn(x0, cs)), PrT<T>::rnd(__fmul_rn(x1, sn))));
  y1 = PrT<T>::rnd(__fadd_rn(PrT<T>::rnd(__fmul_rn(x0, sn)), PrT<T>::rnd(__fmul_rn(x1, cs))));
}

// Element (b, s, v, e): layout 0 = [B,S,V,D], layout 1 = [B,V,S,D].  A thread owns one rotation pair of the token (column
// c = v*D/2 + p) and walks kRopeRows rows, so the pair's frequency (a powf) is computed once per thread and a row costs one
// sincosf; consecutive threads touch consecutive pairs, i.e. contiguous 128-byte segments in either layout.
constexpr int kRopeRows = 16;

template <typename T>
__global__ void __launch_bounds__(256)
rope_shape_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int S, int V, int D, int src_layout, int dst_layout,
                  int rot_dim, int t0, float base, float scale, int inverse) {
  const int C = V * (D / 2);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const int v = c / (D / 2), p = c - v * (D / 2);
  const int b = blockIdx.z;
  const int s0 = blockIdx.y * kRopeRows;
  const int s1 = s0 + kRopeRows < S ? s0 + kRopeRows : S;
  // ATen divides a tensor by a scalar as a multiplication by its fp32 reciprocal (BinaryDivTrueKernel.cu): mirrored here so
  // that fp32 results agree with the reference's `base ** (-2 * idx / dim)` and `pos / scale` to the last bit
  const int pair = rot_dim == D ? p : c;  // Q is rotated as ONE vector of width V*D (nsa_attention.py:1002-1009)
  const float inv_freq = rot_dim > 0 ? powf(base, __fmul_rn(__fmul_rn(-2.0f, (float)pair), __fdiv_rn(1.0f, (float)rot_dim))) : 0.f;
  const float inv_scale = __fdiv_rn(1.0f, scale);
  for (int s = s0; s < s1; ++s) {
    const size_t so = (src_layout == 0 ? (((size_t)b * S + s) * V + v) : (((size_t)b * V + v) * S + s)) * D + 2 * p;
    const size_t dof = (dst_layout == 0 ? (((size_t)b * S + s) * V + v) : (((size_t)b * V + v) * S + s)) * D + 2 * p;
    const float x0 = PrT<T>::ld(x + so), x1 = PrT<T>::ld(x + so + 1);
    float y0 = x0, y1 = x1;
    if (rot_dim > 0) {
      const float ang = __fmul_rn(__fmul_rn((float)(t0 + s), inv_scale), inv_freq);
      float sn, cs;
      sincosf(ang, &sn, &cs);
      rope_rotate<T>(x0, x1, PrT<T>::rnd(sn), PrT<T>::rnd(cs), inverse != 0, y0, y1);
    }
    PrT<T>::st(y + dof, y0);
    PrT<T>::st(y + dof + 1, y1);
  }
}

// y[bg, c, :] = (1/l) sum_{r<l} rot(x[bg, c*d + r, :])   (rope = 0: plain average).  One thread per output pair.
template <typename T>
__global__ void __launch_bounds__(256)
phi_avgpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int BG, int S, int S_cmp, int D, int l, int d, int rope, int t0,
                       float base, float scale, const float* __restrict__ w) {  // w [D][l]: learnable phi (depthwise conv), else mean
  const long long n = (long long)BG * S_cmp * (D / 2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % (D / 2));
    const long long r = i / (D / 2);
    const int c = (int)(r % S_cmp), bg = (int)(r / S_cmp);
    float a0 = 0.f, a1 = 0.f;
    // the pair's frequency once per output (the expression of rope_sincos), eight rows' loads in flight at a time; the rows are
    // folded in ascending order as before, so the result is unchanged.  (One load, one powf and the sin / cos per row in a serial
    // loop made this kernel latency-bound: 31 us for 17 MB at 64k.)
    const float inv_freq = rope ? powf(base, __fmul_rn(__fmul_rn(-2.0f, (float)p), __fdiv_rn(1.0f, (float)D))) : 0.f;
    const float inv_scale = __fdiv_rn(1.0f, scale);
    const T* row0 = x + ((size_t)bg * S + (size_t)c * d) * D + 2 * p;
    for (int k0 = 0; k0 < l; k0 += 8) {
      float v0[8], v1[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (k0 + u < l) PrT<T>::ld2(row0 + (size_t)(k0 + u) * D, v0[u], v1[u]);
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (k0 + u < l) {
          const int k = k0 + u;
          float x0 = v0[u], x1 = v1[u];
          if (rope) {
            const float ang = __fmul_rn(__fmul_rn((float)(t0 + c * d + k), inv_scale), inv_freq);
            float w0, w1;
            rope_rotate<T>(x0, x1, PrT<T>::rnd(sinf(ang)), PrT<T>::rnd(cosf(ang)), false, w0, w1);
            x0 = w0;
            x1 = w1;
          }
          if (w) {
            a0 = fmaf(w[(size_t)(2 * p) * l + k], x0, a0);
            a1 = fmaf(w[(size_t)(2 * p + 1) * l + k], x1, a1);
          } else {
            a0 = __fadd_rn(a0, x0);
            a1 = __fadd_rn(a1, x1);
          }
        }
    }
    T* out = y + ((size_t)bg * S_cmp + c) * D + 2 * p;
    PrT<T>::st(out, w ? a0 : __fdiv_rn(a0, (float)l));
    PrT<T>::st(out + 1, w ? a1 : __fdiv_rn(a1, (float)l));
  }
}

// dx[bg, s, :] = rot(s)^T ( (1/l) sum_{c : c*d <= s < c*d + l} dy[bg, c, :] ).  One thread per input pair.
template <typename T>
__global__ void __launch_bounds__(256)
phi_avgpool_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int BG, int S, int S_cmp, int D, int l, int d, int rope,
                       int t0, float base, float scale, const float* __restrict__ w) {
  const long long n = (long long)BG * S * (D / 2);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int p = (int)(i % (D / 2));
    const long long r = i / (D / 2);
    const int s = (int)(r % S), bg = (int)(r / S);
    int c_lo = s - l + 1 > 0 ? (s - l + 1 + d - 1) / d : 0;
    int c_hi = s / d;
    if (c_hi > S_cmp - 1) c_hi = S_cmp - 1;
    float a0 = 0.f, a1 = 0.f;
    for (int c = c_lo; c <= c_hi; ++c) {
      const T* row = dy + ((size_t)bg * S_cmp + c) * D + 2 * p;
      if (w) {  // tap r = s - c*d of the window that starts at c*d
        a0 = fmaf(w[(size_t)(2 * p) * l + (s - c * d)], PrT<T>::ld(row), a0);
        a1 = fmaf(w[(size_t)(2 * p + 1) * l + (s - c * d)], PrT<T>::ld(row + 1), a1);
      } else {
        a0 += PrT<T>::ld(row);
        a1 += PrT<T>::ld(row + 1);
      }
    }
    a0 = PrT<T>::rnd(w ? a0 : a0 / (float)l);
    a1 = PrT<T>::rnd(w ? a1 : a1 / (float)l);
    float g0 = a0, g1 = a1;
    if (rope) {
      float sn, cs;
      rope_sincos<T>(t0 + s, p, D, base, scale, sn, cs);
      rope_rotate<T>(a0, a1, sn, cs, true, g0, g1);
    }
    T* out = dx + ((size_t)bg * S + s) * D + 2 * p;
    PrT<T>::st(out, g0);
    PrT<T>::st(out + 1, g1);
  }
}

// dw[e][r] = sum over (bg, c) of dy[bg, c, e] * rot(x[bg, c*d + r, :])[e]: one CTA per (rotation pair, tap), fixed-order
// reduction (per-thread partial sums over a strided walk, then a shared-memory tree), so the result is reproducible.
template <typename T>
__global__ void __launch_bounds__(256)
phi_conv_dw_kernel(const T* __restrict__ x, const T* __restrict__ dy, float* __restrict__ dw, int BG, int S, int S_cmp, int D, int l,
                   int d, int rope, int t0, float base, float scale) {
  __shared__ float red0[256], red1[256];
  const int p = blockIdx.x, r = blockIdx.y;
  float a0 = 0.f, a1 = 0.f;
  const long long n = (long long)BG * S_cmp;
  for (long long i = threadIdx.x; i < n; i += blockDim.x) {
    const int c = (int)(i % S_cmp), bg = (int)(i / S_cmp);
    const int s = c * d + r;
    const T* row = x + ((size_t)bg * S + s) * D + 2 * p;
    float v0 = PrT<T>::ld(row), v1 = PrT<T>::ld(row + 1);
    if (rope) {
      float sn, cs, w0, w1;
      rope_sincos<T>(t0 + s, p, D, base, scale, sn, cs);
      rope_rotate<T>(v0, v1, sn, cs, false, w0, w1);
      v0 = w0;
      v1 = w1;
    }
    const T* g = dy + ((size_t)bg * S_cmp + c) * D + 2 * p;
    a0 = fmaf(PrT<T>::ld(g), v0, a0);
    a1 = fmaf(PrT<T>::ld(g + 1), v1, a1);
  }
  red0[threadIdx.x] = a0;
  red1[threadIdx.x] = a1;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) { red0[threadIdx.x] += red0[threadIdx.x + o]; red1[threadIdx.x] += red1[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    dw[(size_t)(2 * p) * l + r] = red0[0];
    dw[(size_t)(2 * p + 1) * l + r] = red1[0];
  }
}

static int pr_blocks(long long n) {
  long long b = (n + 255) / 256;
  if (b > 148LL * 16) b = 148LL * 16;
  return b < 1 ? 1 : (int)b;
}

int launch_rope_shape(const void* x, void* y, int B, int S, int V, int D, int src_layout, int dst_layout, int rot_dim, int t0,
                      float base, float scale, int inverse, int dtype, cudaStream_t stream) {
  NSA_REQUIRE(x && y, "rope_shape: NULL pointer");
  NSA_REQUIRE(B >= 0 && S >= 0 && V >= 1 && D >= 2 && D % 2 == 0, "rope_shape: B=%d S=%d V=%d D=%d", B, S, V, D);
  NSA_REQUIRE(rot_dim == 0 || rot_dim == D || rot_dim == V * D, "rope_shape: rot_dim=%d is neither 0, D nor V*D", rot_dim);
  NSA_REQUIRE((src_layout | dst_layout | 1) == 1, "rope_shape: layouts are 0 ([B,S,V,D]) or 1 ([B,V,S,D])");
  const long long n = (long long)B * S * V * (D / 2);
  if (n == 0) return NSA_OK;
  if (!(scale > 0.f)) scale = 1.0f;
  const int C = V * (D / 2);
  const int bdx = C >= 256 ? 256 : ((C + 31) / 32) * 32;
  const dim3 grid((C + bdx - 1) / bdx, (S + kRopeRows - 1) / kRopeRows, B);
  NSA_REQUIRE(grid.y <= 65535 && B <= 65535, "rope_shape: S=%d B=%d too large for one launch", S, B);
  if (dtype == NSA_F32)
    rope_shape_kernel<float><<<grid, bdx, 0, stream>>>((const float*)x, (float*)y, B, S, V, D, src_layout, dst_layout, rot_dim, t0, base, scale, inverse);
  else if (dtype == NSA_BF16)
    rope_shape_kernel<__nv_bfloat16><<<grid, bdx, 0, stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, B, S, V, D, src_layout, dst_layout, rot_dim, t0, base, scale, inverse);
  else
    rope_shape_kernel<__half><<<grid, bdx, 0, stream>>>((const __half*)x, (__half*)y, B, S, V, D, src_layout, dst_layout, rot_dim, t0, base, scale, inverse);
  return check_launch("rope_shape_kernel");
}

int launch_phi_avgpool(const void* x, void* y, int BG, int S, int D, int l, int d, int rope, int t0, float base, float scale,
                       int backward, int dtype, cudaStream_t stream, const float* w, const void* dy_for_dw) {
  NSA_REQUIRE(x && y, "phi_avgpool: NULL pointer");
  if (backward == 2) {  // dw of the learnable phi: x = the raw stream, dy_for_dw = dy, y = dw [D][l] fp32
    NSA_REQUIRE(dy_for_dw && BG >= 0 && S >= l && l >= 1 && d >= 1 && D >= 2 && D % 2 == 0, "phi_conv dw: bad arguments");
    const int S_c = (S - l) / d + 1;
    if (!(scale > 0.f)) scale = 1.0f;
    const dim3 grid(D / 2, l);
    if (dtype == NSA_F32) phi_conv_dw_kernel<float><<<grid, 256, 0, stream>>>((const float*)x, (const float*)dy_for_dw, (float*)y, BG, S, S_c, D, l, d, rope, t0, base, scale);
    else if (dtype == NSA_BF16) phi_conv_dw_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)dy_for_dw, (float*)y, BG, S, S_c, D, l, d, rope, t0, base, scale);
    else phi_conv_dw_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)x, (const __half*)dy_for_dw, (float*)y, BG, S, S_c, D, l, d, rope, t0, base, scale);
    return check_launch("phi_conv_dw_kernel");
  }
  NSA_REQUIRE(BG >= 0 && S >= l && l >= 1 && d >= 1 && D >= 2 && D % 2 == 0, "phi_avgpool: BG=%d S=%d D=%d l=%d d=%d", BG, S, D, l, d);
  const int S_cmp = (S - l) / d + 1;
  const long long n = (long long)BG * (backward ? S : S_cmp) * (D / 2);
  if (n == 0) return NSA_OK;
  if (!(scale > 0.f)) scale = 1.0f;
  const int blocks = pr_blocks(n);
#define NSA_PHI(T)                                                                                                              \
  do {                                                                                                                          \
    if (backward) phi_avgpool_bwd_kernel<T><<<blocks, 256, 0, stream>>>((const T*)x, (T*)y, BG, S, S_cmp, D, l, d, rope, t0, base, scale, w); \
    else phi_avgpool_fwd_kernel<T><<<blocks, 256, 0, stream>>>((const T*)x, (T*)y, BG, S, S_cmp, D, l, d, rope, t0, base, scale, w); \
  } while (0)
  if (dtype == NSA_F32) NSA_PHI(float);
  else if (dtype == NSA_BF16) NSA_PHI(__nv_bfloat16);
  else NSA_PHI(__half);
#undef NSA_PHI
  return check_launch("phi_avgpool_kernel");
}


// ---------------------------------------------------------------------------------------------------------------------
// Projection-split producer: one kernel takes the fused projection output of S tokens per sequence
//   y [B, S, H*Dk + G*(3*Dk + 3*Dv)] = (Q | K_sel | V_sel | K_win | V_win | K_raw | V_raw)
// rotates Q (as one H*Dk-wide vector) and the two K streams (per Dk-vector; row s sits at position t + s), and writes Q into
// q_out [B,S,H*Dk] and the six streams straight into rows row[i] + s of their [B,G,cap,D] tensors -- the reference's seven
// rope / view / permute().contiguous() / torch.cat chains (nsa_attention.py:545-586 decode, :998-1016 prefill, kv_cache.py:28-49).
// S = 1 with the cache slabs as destinations is a decode step (it also records the step's read counters, kv_cache.py:51-65);
// inverse = 1 is the backward pass: it reads the seven gradients from the same places and writes dy (transposed rotation).
// A thread owns one rotation pair of the fused row (its frequency, a powf, is computed once) and walks `rows` consecutive tokens.
// ---------------------------------------------------------------------------------------------------------------------
using DecodeProduceArgs = nsa_decode_produce_t;  // include/nsa_b200.h
constexpr int kProduceBatch = 8;

template <typename T>
__global__ void __launch_bounds__(256)
decode_produce_kernel(DecodeProduceArgs a, int rows) {
  const int QW = a.H * a.Dk;
  const int N = QW + a.G * (3 * a.Dk + 3 * a.Dv);
  const int pairs = N / 2;
  // device-stepped decode (CUDA-graph replay): position, rows and counters come from the device record
  int t_pos = a.t, ctr_idx = a.counters_idx, r_sel = 0, r_win = 0, r_raw = 0;
  const bool stepped = a.state != nullptr;
  if (stepped) {
    t_pos = a.state->t; r_sel = t_pos; r_win = a.state->row_win; r_raw = a.state->row_raw; ctr_idx = a.state->ctr_idx;
  }
  if (a.counters && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x < 5 && ctr_idx < a.counters_cap) {
    long long v = a.counter_val[threadIdx.x];
    if (stepped) {  // nsa_attention.py:634-638 / kv_cache.py:51-65: (pred, total, sel, cmp, win)
      const int S_raw = r_raw + 1;
      const long long nc = S_raw < a.l ? 0 : (S_raw - a.l) / a.d + 1;
      const long long nw = S_raw < a.w ? S_raw : a.w;
      const long long ns = (long long)a.n_sel * a.l_sel;
      v = threadIdx.x <= 1 ? nc + ns + nw : (threadIdx.x == 2 ? ns : (threadIdx.x == 3 ? nc : nw));
    }
    a.counters[(size_t)threadIdx.x * a.counters_cap + ctr_idx] = v;
  }
  const int pc = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.z;
  if (pc >= pairs) return;
  const int c = 2 * pc;  // column of y
  // where this pair lives on the split side, and how it is rotated
  T* other;            // element (b, s = 0) of the pair on the split side
  size_t other_pitch;  // elements between consecutive tokens there
  int rot_pair = -1, rot_dim = 0;
  if (c < QW) {
    other = reinterpret_cast<T*>(a.q_out) + (size_t)b * a.S * QW + c;
    other_pitch = QW;
    rot_pair = pc;
    rot_dim = QW;
  } else {
    // segments after Q: (K_sel, V_sel, K_win, V_win, K_raw, V_raw), widths G*Dk / G*Dv alternating
    int off = c - QW, seg = 0;
    for (; seg < 5; ++seg) {
      const int wdt = a.G * ((seg & 1) ? a.Dv : a.Dk);
      if (off < wdt) break;
      off -= wdt;
    }
    const int D = (seg & 1) ? a.Dv : a.Dk;
    const int g = off / D, e = off - g * D;
    if (seg == 0 || seg == 2) {  // RoPE'd keys of the selection and window caches
      rot_pair = e / 2;
      rot_dim = D;
    }
    const int row_seg = stepped ? (seg < 2 ? r_sel : (seg < 4 ? r_win : r_raw)) : a.row[seg];
    if (row_seg + a.S > a.cap[seg]) return;  // a stepped loop that outran its slabs writes nothing (the host re-plans before that)
    other = reinterpret_cast<T*>(a.slab[seg]) + (((size_t)b * a.G + g) * a.cap[seg] + row_seg) * D + e;
    other_pitch = D;
  }
  // ATen divides a tensor by a scalar as a multiplication by its fp32 reciprocal: mirrored (see rope_shape_kernel)
  const float inv_freq = rot_dim > 0 ? powf(a.base, __fmul_rn(__fmul_rn(-2.0f, (float)rot_pair), __fdiv_rn(1.0f, (float)rot_dim))) : 0.f;
  const float inv_scale = __fdiv_rn(1.0f, a.scale);
  T* yp = reinterpret_cast<T*>(const_cast<void*>(a.y)) + (size_t)b * a.S * N + c;
  const int s0 = blockIdx.y * rows;
  const int s1 = s0 + rows < a.S ? s0 + rows : a.S;
  // eight tokens per trip: their loads are issued before any of the (long) sincosf chains, so a thread keeps eight requests in flight
  for (int sb = s0; sb < s1; sb += kProduceBatch) {
    float x0[kProduceBatch], x1[kProduceBatch];
#pragma unroll
    for (int u = 0; u < kProduceBatch; ++u)
      if (sb + u < s1) {
        const T* src = a.inverse ? other + (size_t)(sb + u) * other_pitch : yp + (size_t)(sb + u) * N;
        PrT<T>::ld2(src, x0[u], x1[u]);
      }
#pragma unroll
    for (int u = 0; u < kProduceBatch; ++u)
      if (sb + u < s1) {
        const int s = sb + u;
        float y0 = x0[u], y1 = x1[u];
        if (rot_dim > 0) {
          float sn, cs;
          sincosf(__fmul_rn(__fmul_rn((float)(t_pos + s), inv_scale), inv_freq), &sn, &cs);
          rope_rotate<T>(x0[u], x1[u], PrT<T>::rnd(sn), PrT<T>::rnd(cs), a.inverse != 0, y0, y1);
        }
        T* dst = a.inverse ? yp + (size_t)s * N : other + (size_t)s * other_pitch;
        PrT<T>::st2(dst, y0, y1);
      }
  }
}

// (sin, cos) tables for the producers: out[row][pair] = (rnd(sin a), rnd(cos a)), a = (t0 + row) / scale * base^(-2 pair / rot_dim),
// the same expression and roundings as rope_sincos / decode_produce_kernel.
template <typename T>
__global__ void __launch_bounds__(256)
rope_table_kernel(T* __restrict__ out, int rows, int pairs, int rot_dim, int t0, float base, float scale) {
  const long long n = (long long)rows * pairs;
  const float inv_scale = __fdiv_rn(1.0f, scale);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int pair = (int)(i % pairs);
    const int row = (int)(i / pairs);
    const float inv_freq = powf(base, __fmul_rn(__fmul_rn(-2.0f, (float)pair), __fdiv_rn(1.0f, (float)rot_dim)));
    float sn, cs;
    sincosf(__fmul_rn(__fmul_rn((float)(t0 + row), inv_scale), inv_freq), &sn, &cs);
    PrT<T>::st2(out + 2 * i, PrT<T>::rnd(sn), PrT<T>::rnd(cs));
  }
}

int launch_rope_table(int rows, int pairs, int rot_dim, int t0, float base, float scale, int dtype, void* out, cudaStream_t stream) {
  NSA_REQUIRE(out && rows >= 0 && pairs >= 1 && rot_dim >= 2, "rope_table: rows=%d pairs=%d rot_dim=%d", rows, pairs, rot_dim);
  const long long n = (long long)rows * pairs;
  if (n == 0) return NSA_OK;
  if (!(scale > 0.f)) scale = 1.0f;
  const int blocks = pr_blocks(n);
  if (dtype == NSA_F32) rope_table_kernel<float><<<blocks, 256, 0, stream>>>((float*)out, rows, pairs, rot_dim, t0, base, scale);
  else if (dtype == NSA_BF16) rope_table_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>((__nv_bfloat16*)out, rows, pairs, rot_dim, t0, base, scale);
  else rope_table_kernel<__half><<<blocks, 256, 0, stream>>>((__half*)out, rows, pairs, rot_dim, t0, base, scale);
  return check_launch("rope_table_kernel");
}

// The same producer with 16-byte accesses: a thread owns VE = 16 / sizeof(T) consecutive columns of the fused row (VE / 2 rotation
// pairs, or a plain copy for the V and raw streams), so the address arithmetic, the load and the store of a token are shared by
// 8 elements instead of 2.  ncu on the pair-per-thread kernel at 64k: 129 M warp instructions, 82 per pair and token, issue-active
// 72 %, 167 us for 360 MB (profiles/r1_produce_ncu_raw.csv): instruction-bound at a third of the HBM rate.  Same arithmetic per
// element, so the outputs are bit-identical (tests/test_producers_gpu.py).  Needs Dk, Dv multiples of VE and 16-byte aligned bases.
constexpr int kProduceVecBatch = 4;

template <typename T>
__global__ void __launch_bounds__(512)
decode_produce_vec_kernel(DecodeProduceArgs a, int rows) {
  constexpr int VE = 16 / (int)sizeof(T), NP = VE / 2;
  const int QW = a.H * a.Dk;
  const int N = QW + a.G * (3 * a.Dk + 3 * a.Dv);
  const int vecs = N / VE;
  int t_pos = a.t, ctr_idx = a.counters_idx, r_sel = 0, r_win = 0, r_raw = 0;
  const bool stepped = a.state != nullptr;
  if (stepped) {
    t_pos = a.state->t; r_sel = t_pos; r_win = a.state->row_win; r_raw = a.state->row_raw; ctr_idx = a.state->ctr_idx;
  }
  if (a.counters && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x < 5 && ctr_idx < a.counters_cap) {
    long long v = a.counter_val[threadIdx.x];
    if (stepped) {  // nsa_attention.py:634-638 / kv_cache.py:51-65: (pred, total, sel, cmp, win)
      const int S_raw = r_raw + 1;
      const long long nc = S_raw < a.l ? 0 : (S_raw - a.l) / a.d + 1;
      const long long nw = S_raw < a.w ? S_raw : a.w;
      const long long ns = (long long)a.n_sel * a.l_sel;
      v = threadIdx.x <= 1 ? nc + ns + nw : (threadIdx.x == 2 ? ns : (threadIdx.x == 3 ? nc : nw));
    }
    a.counters[(size_t)threadIdx.x * a.counters_cap + ctr_idx] = v;
  }
  const int vc = blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.z;
  if (vc >= vecs) return;
  const int c = VE * vc;  // first column of y
  T* other;
  size_t other_pitch;
  int rot_pair = -1, rot_dim = 0;
  if (c < QW) {
    other = reinterpret_cast<T*>(a.q_out) + (size_t)b * a.S * QW + c;
    other_pitch = QW;
    rot_pair = c / 2;
    rot_dim = QW;
  } else {
    int off = c - QW, seg = 0;
    for (; seg < 5; ++seg) {
      const int wdt = a.G * ((seg & 1) ? a.Dv : a.Dk);
      if (off < wdt) break;
      off -= wdt;
    }
    const int D = (seg & 1) ? a.Dv : a.Dk;
    const int g = off / D, e = off - g * D;
    if (seg == 0 || seg == 2) {
      rot_pair = e / 2;
      rot_dim = D;
    }
    const int row_seg = stepped ? (seg < 2 ? r_sel : (seg < 4 ? r_win : r_raw)) : a.row[seg];
    if (row_seg + a.S > a.cap[seg]) return;  // a stepped loop that outran its slabs writes nothing (the host re-plans before that)
    other = reinterpret_cast<T*>(a.slab[seg]) + (((size_t)b * a.G + g) * a.cap[seg] + row_seg) * D + e;
    other_pitch = D;
  }
  // rotation tables (nsa_rope_table), validated on the host: row (t_pos + s - rope_t0), this thread's NP pairs = 16 bytes
  const T* tab = nullptr;
  size_t tab_pitch = 0;
  if (rot_dim > 0 && a.rope_q && a.rope_k && !stepped) {
    tab_pitch = (size_t)rot_dim;  // rot_dim / 2 pairs x (sin, cos)
    tab = reinterpret_cast<const T*>(c < QW ? a.rope_q : a.rope_k) + (size_t)(t_pos - a.rope_t0) * tab_pitch + 2 * rot_pair;
  }
  float inv_freq[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j)  // ATen's reciprocal-multiply form, see rope_shape_kernel
    inv_freq[j] = (rot_dim > 0 && !tab) ? powf(a.base, __fmul_rn(__fmul_rn(-2.0f, (float)(rot_pair + j)), __fdiv_rn(1.0f, (float)rot_dim))) : 0.f;
  const float inv_scale = __fdiv_rn(1.0f, a.scale);
  T* yp = reinterpret_cast<T*>(const_cast<void*>(a.y)) + (size_t)b * a.S * N + c;
  const int s0 = blockIdx.y * rows;
  const int s1 = s0 + rows < a.S ? s0 + rows : a.S;
  for (int sb = s0; sb < s1; sb += kProduceVecBatch) {
    uint4 v[kProduceVecBatch], tv[kProduceVecBatch];
#pragma unroll
    for (int u = 0; u < kProduceVecBatch; ++u)
      if (sb + u < s1) {
        v[u] = *reinterpret_cast<const uint4*>(a.inverse ? other + (size_t)(sb + u) * other_pitch : yp + (size_t)(sb + u) * N);
        if (tab) tv[u] = *reinterpret_cast<const uint4*>(tab + (size_t)(sb + u) * tab_pitch);
      }
#pragma unroll
    for (int u = 0; u < kProduceVecBatch; ++u)
      if (sb + u < s1) {
        const int s = sb + u;
        if (rot_dim > 0) {
          T* el = reinterpret_cast<T*>(&v[u]);
          if (tab) {  // (sin, cos) of this thread's NP pairs at this position: one 16-byte load
            const T* sc = reinterpret_cast<const T*>(&tv[u]);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              float y0, y1;
              rope_rotate<T>(PrT<T>::ld(el + 2 * j), PrT<T>::ld(el + 2 * j + 1), PrT<T>::ld(sc + 2 * j), PrT<T>::ld(sc + 2 * j + 1),
                             a.inverse != 0, y0, y1);
              PrT<T>::st(el + 2 * j, y0);
              PrT<T>::st(el + 2 * j + 1, y1);
            }
          } else {
            const float pos = __fmul_rn((float)(t_pos + s), inv_scale);
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              float sn, cs, y0, y1;
              sincosf(__fmul_rn(pos, inv_freq[j]), &sn, &cs);
              rope_rotate<T>(PrT<T>::ld(el + 2 * j), PrT<T>::ld(el + 2 * j + 1), PrT<T>::rnd(sn), PrT<T>::rnd(cs), a.inverse != 0, y0, y1);
              PrT<T>::st(el + 2 * j, y0);
              PrT<T>::st(el + 2 * j + 1, y1);
            }
          }
        }
        *reinterpret_cast<uint4*>(a.inverse ? yp + (size_t)s * N : other + (size_t)s * other_pitch) = v[u];
      }
  }
}

int launch_decode_produce(const nsa_decode_produce_t& a, cudaStream_t stream) {
  const int dtype = a.dtype;
  NSA_REQUIRE(a.y && a.q_out, "produce: NULL pointer");
  NSA_REQUIRE(a.S >= 1, "produce: S=%d", a.S);
  NSA_REQUIRE(!a.state || (a.S == 1 && !a.inverse && a.l >= 1 && a.d >= 1), "produce: a device step record needs S = 1, forward, l, d");
  for (int i = 0; i < 6; ++i)
    NSA_REQUIRE(a.slab[i] && (a.state || (a.row[i] >= 0 && a.row[i] + a.S <= a.cap[i])), "produce: slab %d rows [%d, %d) cap %d", i,
                a.row[i], a.row[i] + a.S, a.cap[i]);
  NSA_REQUIRE(a.B >= 0 && a.H >= 1 && a.G >= 1 && a.Dk >= 2 && a.Dv >= 2 && a.Dk % 2 == 0 && a.Dv % 2 == 0, "produce: bad geometry");
  NSA_REQUIRE(!a.counters || a.state || (a.counters_idx >= 0 && a.counters_idx < a.counters_cap), "produce: counter index");
  if (a.B == 0) return NSA_OK;
  DecodeProduceArgs b = a;
  if (!(b.scale > 0.f)) b.scale = 1.0f;
  // rotation tables only when both are there, aligned and cover [t, t + S); else the kernels evaluate sincosf
  if (!(b.rope_q && b.rope_k && !b.state && b.t >= b.rope_t0 && (long long)b.t + b.S <= (long long)b.rope_t0 + b.rope_rows &&
        ((uintptr_t)b.rope_q & 15) == 0 && ((uintptr_t)b.rope_k & 15) == 0)) {
    b.rope_q = nullptr;
    b.rope_k = nullptr;
  }
  const int pairs = (a.H * a.Dk + a.G * (3 * a.Dk + 3 * a.Dv)) / 2;
  // rows per thread: amortise the powf over up to 16 tokens, but keep at least ~4 CTAs per SM in flight
  int rows = 16;
  while (rows > 1 && (long long)((pairs + 255) / 256) * ((a.S + rows - 1) / rows) * a.B < 4 * 148) rows /= 2;
  NSA_REQUIRE(a.B <= 65535 && (a.S + rows - 1) / rows <= 65535, "produce: B=%d S=%d", a.B, a.S);
  {  // 16-byte form when the geometry allows it
    const int ve = dtype == NSA_F32 ? 4 : 8;
    bool vec_ok = a.Dk % ve == 0 && a.Dv % ve == 0 && ((uintptr_t)a.y & 15) == 0 && ((uintptr_t)a.q_out & 15) == 0;
    for (int i = 0; i < 6; ++i) vec_ok = vec_ok && ((uintptr_t)a.slab[i] & 15) == 0;
    static const bool vec_env = !(getenv("NSA_B200_PRODUCE_VEC") && atoi(getenv("NSA_B200_PRODUCE_VEC")) == 0);
    if (vec_ok && vec_env) {
      const int vecs = 2 * pairs / ve;
      const int bdx = vecs <= 512 ? ((vecs + 31) / 32) * 32 : 256;  // one CTA row per token row when the fused row fits (288 vectors at m7c)
      int vrows = 16;
      while (vrows > 1 && (long long)((vecs + bdx - 1) / bdx) * ((a.S + vrows - 1) / vrows) * a.B < 4 * 148) vrows /= 2;
      const dim3 vgrid((vecs + bdx - 1) / bdx, (a.S + vrows - 1) / vrows, a.B);
      NSA_REQUIRE(vgrid.y <= 65535, "produce: S=%d too long for one launch", a.S);
      if (dtype == NSA_F32) decode_produce_vec_kernel<float><<<vgrid, bdx, 0, stream>>>(b, vrows);
      else if (dtype == NSA_BF16) decode_produce_vec_kernel<__nv_bfloat16><<<vgrid, bdx, 0, stream>>>(b, vrows);
      else decode_produce_vec_kernel<__half><<<vgrid, bdx, 0, stream>>>(b, vrows);
      return check_launch("decode_produce_vec_kernel");
    }
  }
  const dim3 grid((pairs + 255) / 256, (a.S + rows - 1) / rows, a.B);
  if (dtype == NSA_F32) decode_produce_kernel<float><<<grid, 256, 0, stream>>>(b, rows);
  else if (dtype == NSA_BF16) decode_produce_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(b, rows);
  else decode_produce_kernel<__half><<<grid, 256, 0, stream>>>(b, rows);
  return check_launch("decode_produce_kernel");
}

// ---- device-stepped decode: emission and state advance ---------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
decode_emit_kernel(nsa_decode_emit_t a) {
  const int S_raw = a.state->row_raw + 1;
  if (S_raw < a.l || (S_raw - a.l) % a.d != 0) return;  // not an emission step (nsa_attention.py:587-588)
  const int c_row = a.state->S_cmp;
  if (c_row >= a.cap_cmp || S_raw > a.cap_raw) return;
  const int s0 = S_raw - a.l;
  const int hk = a.Dk / 2, hv = a.Dv / 2;
  const long long n = (long long)a.BG * (hk + hv);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int bg = (int)(i / (hk + hv));
    int p = (int)(i - (long long)bg * (hk + hv));
    const bool is_k = p < hk;
    if (!is_k) p -= hk;
    const int D = is_k ? a.Dk : a.Dv;
    const T* x = reinterpret_cast<const T*>(is_k ? a.K_raw : a.V_raw) + ((size_t)bg * a.cap_raw + s0) * D + 2 * p;
    const float* wt = is_k ? a.w_k : a.w_v;  // learnable phi (depthwise conv taps [D][l]) or NULL = mean
    float a0 = 0.f, a1 = 0.f;
    for (int k = 0; k < a.l; ++k) {  // same arithmetic as phi_avgpool_fwd_kernel: the window at absolute positions s0 .. s0 + l - 1
      float v0 = PrT<T>::ld(x + (size_t)k * D), v1 = PrT<T>::ld(x + (size_t)k * D + 1);
      if (is_k) {
        float sn, cs, w0, w1;
        rope_sincos<T>(s0 + k, p, D, a.base, a.scale, sn, cs);
        rope_rotate<T>(v0, v1, sn, cs, false, w0, w1);
        v0 = w0;
        v1 = w1;
      }
      if (wt) {
        a0 = fmaf(wt[(size_t)(2 * p) * a.l + k], v0, a0);
        a1 = fmaf(wt[(size_t)(2 * p + 1) * a.l + k], v1, a1);
      } else {
        a0 = __fadd_rn(a0, v0);
        a1 = __fadd_rn(a1, v1);
      }
    }
    T* out = reinterpret_cast<T*>(is_k ? a.K_cmp : a.V_cmp) + ((size_t)bg * a.cap_cmp + c_row) * D + 2 * p;
    PrT<T>::st(out, wt ? a0 : __fdiv_rn(a0, (float)a.l));
    PrT<T>::st(out + 1, wt ? a1 : __fdiv_rn(a1, (float)a.l));
  }
}

int launch_decode_emit(const nsa_decode_emit_t& a0, cudaStream_t stream) {
  nsa_decode_emit_t a = a0;
  NSA_REQUIRE(a.state && a.K_raw && a.V_raw && a.K_cmp && a.V_cmp, "decode_emit: NULL pointer");
  NSA_REQUIRE(a.BG >= 0 && a.l >= 1 && a.d >= 1 && a.Dk >= 2 && a.Dv >= 2 && a.Dk % 2 == 0 && a.Dv % 2 == 0, "decode_emit: bad geometry");
  NSA_REQUIRE(a.dtype == NSA_F32 || a.dtype == NSA_BF16 || a.dtype == NSA_F16, "decode_emit: dtype %d", a.dtype);
  if (a.BG == 0) return NSA_OK;
  if (!(a.scale > 0.f)) a.scale = 1.0f;
  const int blocks = pr_blocks((long long)a.BG * (a.Dk / 2 + a.Dv / 2));
  if (a.dtype == NSA_F32) decode_emit_kernel<float><<<blocks, 256, 0, stream>>>(a);
  else if (a.dtype == NSA_BF16) decode_emit_kernel<__nv_bfloat16><<<blocks, 256, 0, stream>>>(a);
  else decode_emit_kernel<__half><<<blocks, 256, 0, stream>>>(a);
  return check_launch("decode_emit_kernel");
}

__global__ void decode_advance_kernel(nsa_decode_state_t* st, int l, int d) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int S_raw = st->row_raw + 1;
  if (S_raw >= l && (S_raw - l) % d == 0) st->S_cmp += 1;
  st->t += 1;
  st->row_win += 1;
  st->row_raw += 1;
  st->ctr_idx += 1;
}

int launch_decode_advance(nsa_decode_state_t* state, int l, int d, cudaStream_t stream) {
  NSA_REQUIRE(state && l >= 1 && d >= 1, "decode_advance: bad arguments");
  decode_advance_kernel<<<1, 32, 0, stream>>>(state, l, d);
  return check_launch("decode_advance_kernel");
}

}  // namespace nsa
