// Shape- and dtype-generic SIMT kernels for the NSA hot path (fp32 / bf16 / fp16, any Dk,Dv <= 128,
// h <= 16).  They are the parity workhorse (fp32 against the oracle at 5e-5) and serve every shape the
// tcgen05 kernels do not specialise.  Mapping: one warp per (b, s, g) query-row group; lanes split the head
// dimension, so every K/V row is one coalesced warp load shared by all h heads of the group (GQA reuse).
#include "common.cuh"
#include "select.cuh"
#include "gate.cuh"
#include <string.h>
#include "launchers.h"

namespace nsa {

// ------------------------------------------------------------------------------------------------
// Scoring: p_cmp softmax -> Eq.9 -> Eq.10, optional fused selection.
// ------------------------------------------------------------------------------------------------
constexpr int kScoreWarps = 4;

__global__ void __launch_bounds__(kScoreWarps * 32)
score_generic_kernel(nsa_dims_t dm, const void* __restrict__ Q, const void* __restrict__ Kc, int S_sel, int S_total,
                     int sel_mode, int nf, int Kr, float* __restrict__ p_grp, int32_t* __restrict__ ranges) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per_warp = dm.S_cmp + S_sel + dm.h * dm.Dk;
  float* lg = smem + (size_t)warp * per_warp;  // [S_cmp] logits -> probabilities
  float* pg = lg + dm.S_cmp;                   // [S_sel]
  float* qs = pg + S_sel;                      // [h][Dk]
  const int n_rows = dm.B * dm.S * dm.G;
  const int dt = dm.dtype;
  for (int row = blockIdx.x * kScoreWarps + warp; row < n_rows; row += gridDim.x * kScoreWarps) {
    const int g = row % dm.G, s = (row / dm.G) % dm.S, b = row / (dm.G * dm.S);
    const int t = dm.t0 + s;
    const int nkeys = dm.norm_mode == NSA_NORM_CAUSAL ? num_cmp_at(t, dm.l, dm.d, dm.S_cmp) : dm.S_cmp;
    for (int j = lane; j < S_sel; j += 32) pg[j] = 0.f;
    const size_t qbase = (size_t)row * dm.h * dm.Dk;
    for (int i = lane; i < dm.h * dm.Dk; i += 32) qs[i] = ld_elt(Q, qbase + i, dt);
    __syncwarp();
    const size_t kbase = (size_t)(b * dm.G + g) * dm.cap_cmp * dm.Dk;
    for (int hh = 0; hh < dm.h; ++hh) {
      const float* qh = qs + hh * dm.Dk;
      float m = -INFINITY;
      for (int i = lane; i < nkeys; i += 32) {
        float a = 0.f;
        const size_t kr = kbase + (size_t)i * dm.Dk;
        for (int k = 0; k < dm.Dk; ++k) a = fmaf(qh[k], ld_elt(Kc, kr + k, dt), a);
        a *= dm.scale;
        lg[i] = a;
        m = fmaxf(m, a);
      }
      m = warp_max(m);
      float sum = 0.f;
      for (int i = lane; i < nkeys; i += 32) {
        float e = expf(lg[i] - m);
        lg[i] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = nkeys > 0 ? 1.0f / sum : 0.f;
      __syncwarp();
      // Eq.9 (block_index.py:43-71): compressed block i = [i*d, i*d+l) gives overlap/l to selection block j;
      // accumulated in ascending i like the reference's CPU scatter_add (selection_scorer.py:111-115).
      for (int j = lane; j < S_sel; j += 32) {
        const int b0 = j * dm.l_sel, b1 = b0 + dm.l_sel;
        int i_lo = b0 - dm.l + 1 <= 0 ? 0 : (b0 - dm.l + 1 + dm.d - 1) / dm.d;
        int i_hi = (b1 - 1) / dm.d;
        if (i_hi > nkeys - 1) i_hi = nkeys - 1;
        float acc = 0.f;
        for (int i = i_lo; i <= i_hi; ++i) {
          const int a0 = i * dm.d, a1 = a0 + dm.l;
          const int ov = min(a1, b1) - max(a0, b0);
          if (ov > 0) acc += (lg[i] * inv) * ((float)ov / (float)dm.l);
        }
        pg[j] += acc;  // Eq.10: sum over the heads of the group (nsa_attention.py:1091)
      }
      __syncwarp();
    }
    if (p_grp) {
      float* dst = p_grp + (size_t)row * S_sel;
      for (int j = lane; j < S_sel; j += 32) dst[j] = pg[j];
    }
    if (ranges) {
      __syncwarp();
      select_row_warp(pg, S_sel, dm.l_sel, dm.n_sel, sel_mode, nf, Kr, t, ranges + (size_t)row * Kr * 2);
    }
    __syncwarp();
  }
}

int launch_score_generic(const nsa_dims_t& dm, const void* Q, const void* Kc, int S_sel, int S_total, int sel_mode, int Kr,
                         float* p_grp, int32_t* ranges, cudaStream_t stream) {
  const int n_rows = dm.B * dm.S * dm.G;
  if (n_rows == 0) return NSA_OK;
  NSA_REQUIRE(S_sel >= 1 && S_sel <= kSelMaxWords * 1024, "score: S_sel=%d unsupported", S_sel);
  size_t smem = (size_t)kScoreWarps * (dm.S_cmp + S_sel + dm.h * dm.Dk) * sizeof(float);
  NSA_REQUIRE(smem <= 200 * 1024, "score(simt): S_cmp=%d needs %zu B of shared memory", dm.S_cmp, smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(score_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("score: smem attr: %s", cudaGetErrorString(e)); return NSA_ERR_CUDA; }
  }
  int blocks = ceil_div(n_rows, kScoreWarps);
  if (blocks > 148 * 16) blocks = 148 * 16;
  const int nf = forced_code_default(sel_mode, S_total, dm.l_sel);
  score_generic_kernel<<<blocks, kScoreWarps * 32, smem, stream>>>(dm, Q, Kc, S_sel, S_total, sel_mode, nf, Kr, p_grp,
                                                                  ranges);
  return check_launch("score_generic_kernel");
}

// ------------------------------------------------------------------------------------------------
// The scorer's stages as stand-alone functions (the reference's free functions compute_pcmp_all, map_pcmp_to_pslc(_batched),
// convert_indices_to_ranges_batched(_v2): selection_scorer.py:42-61, :64-116, :380-605).  The hot path never materialises
// these tensors (p_cmp alone is 12.9 GB per 64k sequence); callers and tests that want them get them from here.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kScoreWarps * 32)
pcmp_all_kernel(nsa_dims_t dm, const void* __restrict__ Q, const void* __restrict__ Kc, float* __restrict__ p) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* lg = smem + (size_t)warp * (dm.S_cmp + dm.Dk);
  float* qh = lg + dm.S_cmp;
  const long long n = (long long)dm.B * dm.S * dm.G * dm.h;
  const int dt = dm.dtype;
  for (long long rh = (long long)blockIdx.x * kScoreWarps + warp; rh < n; rh += (long long)gridDim.x * kScoreWarps) {
    const long long row = rh / dm.h;
    const int g = (int)(row % dm.G), b = (int)(row / ((long long)dm.G * dm.S));
    for (int i = lane; i < dm.Dk; i += 32) qh[i] = ld_elt(Q, (size_t)rh * dm.Dk + i, dt);
    __syncwarp();
    const size_t kbase = (size_t)(b * dm.G + g) * dm.cap_cmp * dm.Dk;
    float m = -INFINITY;
    for (int i = lane; i < dm.S_cmp; i += 32) {
      float a = 0.f;
      for (int k = 0; k < dm.Dk; ++k) a = fmaf(qh[k], ld_elt(Kc, kbase + (size_t)i * dm.Dk + k, dt), a);
      a *= dm.scale;
      lg[i] = a;
      m = fmaxf(m, a);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int i = lane; i < dm.S_cmp; i += 32) {
      const float e = expf(lg[i] - m);
      lg[i] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
    for (int i = lane; i < dm.S_cmp; i += 32) p[(size_t)rh * dm.S_cmp + i] = lg[i] * inv;
    __syncwarp();
  }
}

int launch_pcmp_all(const nsa_dims_t& dm, const void* Q, const void* Kc, float* p_cmp, cudaStream_t stream) {
  const long long n = (long long)dm.B * dm.S * dm.G * dm.h;
  if (n == 0 || dm.S_cmp == 0) return NSA_OK;
  const size_t smem = (size_t)kScoreWarps * (dm.S_cmp + dm.Dk) * sizeof(float);
  NSA_REQUIRE(smem <= 200 * 1024, "pcmp_all: S_cmp=%d needs %zu B of shared memory (this stand-alone stage serves S_cmp <= ~12k; the "
              "hot path scores any length without materialising p_cmp)", dm.S_cmp, smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(pcmp_all_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("pcmp_all: smem attr: %s", cudaGetErrorString(e)); return NSA_ERR_CUDA; }
  }
  long long blocks = (n + kScoreWarps - 1) / kScoreWarps;
  if (blocks > 148 * 16) blocks = 148 * 16;
  pcmp_all_kernel<<<(int)blocks, kScoreWarps * 32, smem, stream>>>(dm, Q, Kc, p_cmp);
  return check_launch("pcmp_all_kernel");
}

// Eq.9 for any d | l, d | l_sel: p_slc[row][j] = sum_i p_cmp[row][i] * overlap(i, j) / l, ascending i (the COO order of the
// reference's CPU scatter_add, block_index.py:81-85, selection_scorer.py:102-115).  One thread per (row, j).
__global__ void __launch_bounds__(256)
map_pslc_kernel(const float* __restrict__ p, long long n_rows, int S_cmp, int S_sel, int l, int d, int l_sel, float* __restrict__ out) {
  const long long total = n_rows * S_sel;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const long long row = idx / S_sel;
    const int j = (int)(idx - row * S_sel);
    const int b0 = j * l_sel, b1 = b0 + l_sel;
    int i_lo = b0 - l + 1 <= 0 ? 0 : (b0 - l + 1 + d - 1) / d;
    int i_hi = (b1 - 1) / d;
    if (i_hi > S_cmp - 1) i_hi = S_cmp - 1;
    float acc = 0.f;
    for (int i = i_lo; i <= i_hi; ++i) {
      const int a0 = i * d, a1 = a0 + l;
      const int ov = min(a1, b1) - max(a0, b0);
      if (ov > 0) acc += p[row * S_cmp + i] * ((float)ov / (float)l);
    }
    out[idx] = acc;
  }
}

int launch_map_pslc(const float* p_cmp, long long n_rows, int S_cmp, int S_sel, int l, int d, int l_sel, float* p_slc, cudaStream_t stream) {
  const long long total = n_rows * S_sel;
  if (total == 0) return NSA_OK;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  map_pslc_kernel<<<(int)blocks, 256, 0, stream>>>(p_cmp, n_rows, S_cmp, S_sel, l, d, l_sel, p_slc);
  return check_launch("map_pslc_kernel");
}

// Block ids (ascending per row, negative = padding) -> merged [start, end) token ranges, clamped to t + 1, [0,0] padded: the
// loop of convert_indices_to_ranges_batched (selection_scorer.py:380-431), which ranges v2 (:434-605) reproduces.  One thread
// per (b, t, g) row; the hot path never calls this (its selection kernel emits ranges straight from the pick bitmap).
__global__ void __launch_bounds__(128)
indices_to_ranges_kernel(const int32_t* __restrict__ idx, long long n_rows, int S, int G, int K, int S_sel, int l_sel, int t0,
                         int32_t* __restrict__ out) {
  for (long long row = blockIdx.x * (long long)blockDim.x + threadIdx.x; row < n_rows; row += (long long)gridDim.x * blockDim.x) {
    const int t = t0 + (int)((row / G) % S);
    const int32_t* r = idx + row * K;
    int32_t* o = out + row * K * 2;
    int n = 0, prev = -1, last_s = -1, last_e = -1;
    for (int k = 0; k < K; ++k) {
      const int bid = r[k];
      if (bid < 0 || bid >= S_sel || bid == prev) continue;
      prev = bid;
      const int s0 = bid * l_sel;
      int e0 = s0 + l_sel;
      if (e0 > t + 1) e0 = t + 1;
      if (e0 <= s0) continue;
      if (last_s < 0) { last_s = s0; last_e = e0; }
      else if (s0 == last_e) last_e = e0;
      else { o[2 * n] = last_s; o[2 * n + 1] = last_e; ++n; last_s = s0; last_e = e0; }
    }
    if (last_s >= 0) { o[2 * n] = last_s; o[2 * n + 1] = last_e; ++n; }
    for (; n < K; ++n) { o[2 * n] = 0; o[2 * n + 1] = 0; }
  }
}

int launch_indices_to_ranges(const int32_t* indices, int B, int S, int G, int K, int S_sel, int l_sel, int t0, int32_t* ranges,
                             cudaStream_t stream) {
  const long long n_rows = (long long)B * S * G;
  if (n_rows == 0 || K == 0) return NSA_OK;
  long long blocks = (n_rows + 127) / 128;
  if (blocks > 148 * 16) blocks = 148 * 16;
  indices_to_ranges_kernel<<<(int)blocks, 128, 0, stream>>>(indices, n_rows, S, G, K, S_sel, l_sel, t0, ranges);
  return check_launch("indices_to_ranges_kernel");
}

// ------------------------------------------------------------------------------------------------
// Attention over key ranges, forward.  branch_mask bit b enables branch b (0 cmp, 1 sel, 2 win).
// ------------------------------------------------------------------------------------------------

constexpr int kAttnWarps = 4;

// number of [start,end) pieces and the piece itself for one branch of one row (absolute key rows in the cache)
__device__ __forceinline__ int branch_pieces(const nsa_dims_t& dm, int branch) { return branch == 1 ? dm.n_ranges : 1; }
__device__ __forceinline__ void branch_piece(const nsa_dims_t& dm, int branch, int t, const int32_t* rrow, int i, int& a0,
                                             int& a1) {
  if (branch == 0) {  // compressed tokens [0, num_cmp(t))            (attention_kernels.py:117-126)
    a0 = 0;
    a1 = num_cmp_at(t, dm.l, dm.d, dm.S_cmp);
  } else if (branch == 1) {  // selected ranges, clamped to the cache  (attention_kernels.py:724-725)
    a0 = rrow[2 * i];
    a1 = rrow[2 * i + 1];
    if (a0 < 0) a0 = 0;
    if (a1 > dm.S_sel_kv) a1 = dm.S_sel_kv;
  } else {  // sliding window [t-w+1, t] in cache rows                 (attention_kernels.py:159-161)
    int lo = t - dm.w + 1;
    if (lo < dm.win_off) lo = dm.win_off;
    if (lo < 0) lo = 0;
    a0 = lo - dm.win_off;
    a1 = t + 1 - dm.win_off;
    if (a1 > dm.S_win_kv) a1 = dm.S_win_kv;
    if (dm.w <= 0) a1 = a0;
  }
}
__device__ __forceinline__ int branch_cap(const nsa_dims_t& dm, int branch) {
  return branch == 0 ? dm.cap_cmp : (branch == 1 ? dm.cap_sel : dm.cap_win);
}

template <int HMAX, int EPT>
__global__ void __launch_bounds__(kAttnWarps * 32)
fwd_generic_kernel(nsa_dims_t dm, FwdArgs a) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qgp = smem + (size_t)warp * (dm.Dk + 2 * dm.gate_hidden);
  float* xs = qgp + dm.Dk;
  const int n_rows = dm.B * dm.S * dm.G;
  const int dt = dm.dtype, h = dm.h, Dk = dm.Dk, Dv = dm.Dv;
  const size_t rows_h = (size_t)n_rows * h;
  for (int row = blockIdx.x * kAttnWarps + warp; row < n_rows; row += gridDim.x * kAttnWarps) {
    const int g = row % dm.G, s = (row / dm.G) % dm.S, b = row / (dm.G * dm.S);
    const int t = dm.t0 + s;
    float q[HMAX][EPT];
    const size_t qbase = (size_t)row * h * Dk;
#pragma unroll
    for (int hh = 0; hh < HMAX; ++hh)
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int k = lane + 32 * e;
        q[hh][e] = (hh < h && k < Dk) ? ld_elt(a.Q, qbase + (size_t)hh * Dk + k, dt) : 0.f;
      }
    // ---- gates ---------------------------------------------------------------------------------
    Gate3 gt;
    if (a.gates_in) {
      gt = {a.gates_in[(size_t)row * 3], a.gates_in[(size_t)row * 3 + 1], a.gates_in[(size_t)row * 3 + 2]};
    } else {
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int k = lane + 32 * e;
        float m = 0.f;
#pragma unroll
        for (int hh = 0; hh < HMAX; ++hh) m += q[hh][e];
        if (k < Dk) qgp[k] = m / (float)h;  // q_gp = mean over heads (nsa_attention.py:1357)
      }
      __syncwarp();
      gt = gate_forward_warp(qgp, xs, nullptr, a.gp, Dk, dm.gate_hidden, dm.gate_tau, dm.gate_mode, nullptr);
      __syncwarp();
    }
    if (a.gates_out && lane == 0) {
      a.gates_out[(size_t)row * 3] = gt.c;
      a.gates_out[(size_t)row * 3 + 1] = gt.s;
      a.gates_out[(size_t)row * 3 + 2] = gt.w;
    }
    float out[HMAX][EPT];
#pragma unroll
    for (int hh = 0; hh < HMAX; ++hh)
#pragma unroll
      for (int e = 0; e < EPT; ++e) out[hh][e] = 0.f;

    const int32_t* rrow = a.ranges ? a.ranges + (size_t)row * dm.n_ranges * 2 : nullptr;
#pragma unroll 1
    for (int br = 0; br < 3; ++br) {
      if (!(a.branch_mask & (1 << br))) continue;
      const float gb = br == 0 ? gt.c : (br == 1 ? gt.s : gt.w);
      float m[HMAX], ls[HMAX], acc[HMAX][EPT];
#pragma unroll
      for (int hh = 0; hh < HMAX; ++hh) {
        m[hh] = -INFINITY;
        ls[hh] = 0.f;
#pragma unroll
        for (int e = 0; e < EPT; ++e) acc[hh][e] = 0.f;
      }
      const size_t slab = (size_t)(b * dm.G + g) * branch_cap(dm, br);
      const int np = (br == 1 && !rrow) ? 0 : branch_pieces(dm, br);
      for (int pi = 0; pi < np; ++pi) {
        int a0, a1;
        branch_piece(dm, br, t, rrow, pi, a0, a1);
        for (int key = a0; key < a1; ++key) {
          float kk[EPT], vv[EPT];
          const size_t kr = (slab + key) * Dk, vr = (slab + key) * Dv;
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            const int k = lane + 32 * e;
            kk[e] = k < Dk ? ld_elt(a.K[br], kr + k, dt) : 0.f;
            vv[e] = k < Dv ? ld_elt(a.V[br], vr + k, dt) : 0.f;
          }
#pragma unroll
          for (int hh = 0; hh < HMAX; ++hh) {
            if (hh < h) {
              float part = 0.f;
#pragma unroll
              for (int e = 0; e < EPT; ++e) part = fmaf(q[hh][e], kk[e], part);
              const float sc = warp_sum(part) * dm.scale;
              const float mn = fmaxf(m[hh], sc);
              const float al = expf(m[hh] - mn);  // exp(-inf) = 0 on the first key
              const float p = expf(sc - mn);
              ls[hh] = ls[hh] * al + p;
#pragma unroll
              for (int e = 0; e < EPT; ++e) acc[hh][e] = fmaf(p, vv[e], acc[hh][e] * al);
              m[hh] = mn;
            }
          }
        }
      }
#pragma unroll
      for (int hh = 0; hh < HMAX; ++hh) {
        if (hh < h) {
          const float inv = ls[hh] > 0.f ? 1.0f / ls[hh] : 0.f;  // empty row -> zeros (attention_kernels.py:769-771)
          if (a.lse && lane == 0)
            a.lse[(size_t)br * rows_h + (size_t)row * h + hh] = ls[hh] > 0.f ? m[hh] + logf(ls[hh]) : -INFINITY;
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            const float o = acc[hh][e] * inv;
            const int k = lane + 32 * e;
            if (a.O_br && k < Dv) st_elt(a.O_br, ((size_t)br * rows_h + (size_t)row * h + hh) * Dv + k, dt, o);
            out[hh][e] = fmaf(gb, o, out[hh][e]);
          }
        }
      }
    }
    if (a.O) {
#pragma unroll
      for (int hh = 0; hh < HMAX; ++hh)
#pragma unroll
        for (int e = 0; e < EPT; ++e) {
          const int k = lane + 32 * e;
          if (hh < h && k < Dv) st_elt(a.O, ((size_t)row * h + hh) * Dv + k, dt, out[hh][e]);
        }
    }
    __syncwarp();
  }
}

template <int HMAX, int EPT>
static int launch_fwd_t(const nsa_dims_t& dm, const FwdArgs& a, cudaStream_t stream) {
  const int n_rows = dm.B * dm.S * dm.G;
  int blocks = ceil_div(n_rows, kAttnWarps);
  if (blocks > 148 * 32) blocks = 148 * 32;
  size_t smem = (size_t)kAttnWarps * (dm.Dk + 2 * dm.gate_hidden) * sizeof(float);
  fwd_generic_kernel<HMAX, EPT><<<blocks, kAttnWarps * 32, smem, stream>>>(dm, a);
  return check_launch("fwd_generic_kernel");
}

#define NSA_DISPATCH_HE(FN, ...)                                                      \
  do {                                                                                \
    const int dmax = dm.Dk > dm.Dv ? dm.Dk : dm.Dv;                                   \
    const int ept = dmax <= 32 ? 1 : (dmax <= 64 ? 2 : 4);                            \
    const int hm = dm.h <= 4 ? 4 : (dm.h <= 8 ? 8 : 16);                              \
    if (hm == 4 && ept == 1) return FN<4, 1>(__VA_ARGS__);                            \
    if (hm == 4 && ept == 2) return FN<4, 2>(__VA_ARGS__);                            \
    if (hm == 4 && ept == 4) return FN<4, 4>(__VA_ARGS__);                            \
    if (hm == 8 && ept == 1) return FN<8, 1>(__VA_ARGS__);                            \
    if (hm == 8 && ept == 2) return FN<8, 2>(__VA_ARGS__);                            \
    if (hm == 8 && ept == 4) return FN<8, 4>(__VA_ARGS__);                            \
    if (hm == 16 && ept == 1) return FN<16, 1>(__VA_ARGS__);                          \
    if (hm == 16 && ept == 2) return FN<16, 2>(__VA_ARGS__);                          \
    return FN<16, 4>(__VA_ARGS__);                                                    \
  } while (0)

int launch_fwd_generic(const nsa_dims_t& dm, const FwdArgs& a, cudaStream_t stream) {
  if (dm.B * dm.S * dm.G == 0) return NSA_OK;
  NSA_REQUIRE(dm.h >= 1 && dm.h <= 16, "attention(simt): h=%d outside [1,16]", dm.h);
  NSA_REQUIRE(dm.Dk >= 1 && dm.Dk <= 128 && dm.Dv >= 1 && dm.Dv <= 128, "attention(simt): Dk=%d Dv=%d outside [1,128]",
              dm.Dk, dm.Dv);
  NSA_REQUIRE(dm.gate_hidden >= 0 && dm.gate_hidden <= 1024, "gate_hidden=%d", dm.gate_hidden);
  NSA_DISPATCH_HE(launch_fwd_t, dm, a, stream);
}

// ------------------------------------------------------------------------------------------------
// Attention over key ranges, analytical backward (flash-style recompute from the saved LSE).
//   dV = P^T dO, dP = dO V^T, dS = P o (dP - rowsum(dO o O)), dQ = dS K scale, dK = dS^T Q scale
// (the reference's _selection_attention_backward, kernels/triton_sel_kernel/__init__.py:163-231, without
// its first-key-only line).  dK/dV are scatter-added with fp32 atomics because many rows share a key.
// ------------------------------------------------------------------------------------------------

template <int HMAX, int EPT>
__global__ void __launch_bounds__(kAttnWarps * 32)
bwd_generic_kernel(nsa_dims_t dm, BwdArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rows = dm.B * dm.S * dm.G;
  const int dt = dm.dtype, h = dm.h, Dk = dm.Dk, Dv = dm.Dv;
  const size_t rows_h = (size_t)n_rows * h;
  for (int row = blockIdx.x * kAttnWarps + warp; row < n_rows; row += gridDim.x * kAttnWarps) {
    const int g = row % dm.G, s = (row / dm.G) % dm.S, b = row / (dm.G * dm.S);
    const int t = dm.t0 + s;
    float q[HMAX][EPT], dO[HMAX][EPT], dq[HMAX][EPT];
#pragma unroll
    for (int hh = 0; hh < HMAX; ++hh)
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int k = lane + 32 * e;
        q[hh][e] = (hh < h && k < Dk) ? ld_elt(a.Q, ((size_t)row * h + hh) * Dk + k, dt) : 0.f;
        dO[hh][e] = (hh < h && k < Dv) ? ld_elt(a.dO, ((size_t)row * h + hh) * Dv + k, dt) : 0.f;
        dq[hh][e] = 0.f;
      }
    const int32_t* rrow = a.ranges ? a.ranges + (size_t)row * dm.n_ranges * 2 : nullptr;
#pragma unroll 1
    for (int br = 0; br < 3; ++br) {
      if (!(a.branch_mask & (1 << br))) continue;
      const float gb = a.gates ? a.gates[(size_t)row * 3 + br] : 1.0f;
      float delta[HMAX], lse[HMAX];
      float dg = 0.f;
#pragma unroll
      for (int hh = 0; hh < HMAX; ++hh) {
        float part = 0.f;
        if (hh < h) {
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            const int k = lane + 32 * e;
            if (k < Dv) part = fmaf(dO[hh][e], ld_elt(a.O_br, ((size_t)br * rows_h + (size_t)row * h + hh) * Dv + k, dt), part);
          }
        }
        part = warp_sum(part);
        dg += part;              // d gate_b = sum_{h,d} dO * O_b
        delta[hh] = gb * part;   // rowsum(dO_b o O_b) with dO_b = g_b dO
        lse[hh] = hh < h ? a.lse[(size_t)br * rows_h + (size_t)row * h + hh] : -INFINITY;
      }
      if (a.dgates && lane == 0) a.dgates[(size_t)row * 3 + br] = dg;
      if (gb == 0.f) continue;  // forced-branch gates: nothing flows into this branch
      const size_t slab = (size_t)(b * dm.G + g) * branch_cap(dm, br);
      const int np = (br == 1 && !rrow) ? 0 : branch_pieces(dm, br);
      for (int pi = 0; pi < np; ++pi) {
        int a0, a1;
        branch_piece(dm, br, t, rrow, pi, a0, a1);
        for (int key = a0; key < a1; ++key) {
          float kk[EPT], vv[EPT], dk[EPT], dv[EPT];
          const size_t kr = (slab + key) * Dk, vr = (slab + key) * Dv;
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            const int k = lane + 32 * e;
            kk[e] = k < Dk ? ld_elt(a.K[br], kr + k, dt) : 0.f;
            vv[e] = k < Dv ? ld_elt(a.V[br], vr + k, dt) : 0.f;
            dk[e] = 0.f;
            dv[e] = 0.f;
          }
#pragma unroll
          for (int hh = 0; hh < HMAX; ++hh) {
            if (hh < h) {
              float ps = 0.f, pd = 0.f;
#pragma unroll
              for (int e = 0; e < EPT; ++e) {
                ps = fmaf(q[hh][e], kk[e], ps);
                pd = fmaf(dO[hh][e], vv[e], pd);
              }
              const float sc = warp_sum(ps) * dm.scale;
              const float dp = gb * warp_sum(pd);
              const float p = isinf(lse[hh]) ? 0.f : expf(sc - lse[hh]);
              const float ds = p * (dp - delta[hh]) * dm.scale;
              const float pg = p * gb;
#pragma unroll
              for (int e = 0; e < EPT; ++e) {
                dq[hh][e] = fmaf(ds, kk[e], dq[hh][e]);
                dk[e] = fmaf(ds, q[hh][e], dk[e]);
                dv[e] = fmaf(pg, dO[hh][e], dv[e]);
              }
            }
          }
#pragma unroll
          for (int e = 0; e < EPT; ++e) {
            const int k = lane + 32 * e;
            if (k < Dk) atomicAdd(a.dK[br] + kr + k, dk[e]);
            if (k < Dv) atomicAdd(a.dV[br] + vr + k, dv[e]);
          }
        }
      }
    }
#pragma unroll
    for (int hh = 0; hh < HMAX; ++hh)
#pragma unroll
      for (int e = 0; e < EPT; ++e) {
        const int k = lane + 32 * e;
        if (hh < h && k < Dk) a.dQ[((size_t)row * h + hh) * Dk + k] += dq[hh][e];  // row owned by this warp
      }
  }
}

template <int HMAX, int EPT>
static int launch_bwd_t(const nsa_dims_t& dm, const BwdArgs& a, cudaStream_t stream) {
  const int n_rows = dm.B * dm.S * dm.G;
  int blocks = ceil_div(n_rows, kAttnWarps);
  if (blocks > 148 * 32) blocks = 148 * 32;
  bwd_generic_kernel<HMAX, EPT><<<blocks, kAttnWarps * 32, 0, stream>>>(dm, a);
  return check_launch("bwd_generic_kernel");
}

int launch_bwd_generic(const nsa_dims_t& dm, const BwdArgs& a, cudaStream_t stream) {
  if (dm.B * dm.S * dm.G == 0) return NSA_OK;
  NSA_REQUIRE(dm.h >= 1 && dm.h <= 16, "attention bwd(simt): h=%d outside [1,16]", dm.h);
  NSA_REQUIRE(dm.Dk <= 128 && dm.Dv <= 128, "attention bwd(simt): Dk=%d Dv=%d", dm.Dk, dm.Dv);
  NSA_DISPATCH_HE(launch_bwd_t, dm, a, stream);
}

// ------------------------------------------------------------------------------------------------
// Gate forward (standalone) and backward.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kAttnWarps * 32)
gate_fwd_kernel(nsa_dims_t dm, const void* __restrict__ Q, nsa_gate_params_t gp, float* __restrict__ gates) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qgp = smem + (size_t)warp * (dm.Dk + 2 * dm.gate_hidden);
  float* xs = qgp + dm.Dk;
  const int n_rows = dm.B * dm.S * dm.G;
  for (int row = blockIdx.x * kAttnWarps + warp; row < n_rows; row += gridDim.x * kAttnWarps) {
    for (int k = lane; k < dm.Dk; k += 32) {
      float m = 0.f;
      for (int hh = 0; hh < dm.h; ++hh) m += ld_elt(Q, ((size_t)row * dm.h + hh) * dm.Dk + k, dm.dtype);
      qgp[k] = m / (float)dm.h;
    }
    __syncwarp();
    Gate3 gt = gate_forward_warp(qgp, xs, nullptr, gp, dm.Dk, dm.gate_hidden, dm.gate_tau, dm.gate_mode, nullptr);
    if (lane == 0) {
      gates[(size_t)row * 3] = gt.c;
      gates[(size_t)row * 3 + 1] = gt.s;
      gates[(size_t)row * 3 + 2] = gt.w;
    }
    __syncwarp();
  }
}

int launch_gate_fwd(const nsa_dims_t& dm, const void* Q, const nsa_gate_params_t& gp, float* gates, cudaStream_t stream) {
  const int n_rows = dm.B * dm.S * dm.G;
  if (n_rows == 0) return NSA_OK;
  int blocks = ceil_div(n_rows, kAttnWarps);
  if (blocks > 148 * 16) blocks = 148 * 16;
  size_t smem = (size_t)kAttnWarps * (dm.Dk + 2 * dm.gate_hidden) * sizeof(float);
  gate_fwd_kernel<<<blocks, kAttnWarps * 32, smem, stream>>>(dm, Q, gp, gates);
  return check_launch("gate_fwd_kernel");
}

// Gate + three-branch combine over branch outputs already in HBM (used when the branches ran as separate kernels).
__global__ void __launch_bounds__(kAttnWarps * 32)
combine_kernel(nsa_dims_t dm, const void* __restrict__ Q, nsa_gate_params_t gp, const void* __restrict__ O_br,
               void* __restrict__ O, float* __restrict__ gates) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qgp = smem + (size_t)warp * (dm.Dk + 2 * dm.gate_hidden);
  float* xs = qgp + dm.Dk;
  const int n_rows = dm.B * dm.S * dm.G;
  const size_t per_branch = (size_t)n_rows * dm.h * dm.Dv;
  for (int row = blockIdx.x * kAttnWarps + warp; row < n_rows; row += gridDim.x * kAttnWarps) {
    if (dm.gate_mode == NSA_GATE_MLP) {
      for (int k = lane; k < dm.Dk; k += 32) {
        float m = 0.f;
        for (int hh = 0; hh < dm.h; ++hh) m += ld_elt(Q, ((size_t)row * dm.h + hh) * dm.Dk + k, dm.dtype);
        qgp[k] = m / (float)dm.h;
      }
      __syncwarp();
    }
    Gate3 gt = gate_forward_warp(qgp, xs, nullptr, gp, dm.Dk, dm.gate_hidden, dm.gate_tau, dm.gate_mode, nullptr);
    if (gates && lane == 0) {
      gates[(size_t)row * 3] = gt.c;
      gates[(size_t)row * 3 + 1] = gt.s;
      gates[(size_t)row * 3 + 2] = gt.w;
    }
    const size_t base = (size_t)row * dm.h * dm.Dv;
    for (int i = lane; i < dm.h * dm.Dv; i += 32) {
      const float v = gt.c * ld_elt(O_br, base + i, dm.dtype) + gt.s * ld_elt(O_br, per_branch + base + i, dm.dtype) +
                      gt.w * ld_elt(O_br, 2 * per_branch + base + i, dm.dtype);
      st_elt(O, base + i, dm.dtype, v);
    }
    __syncwarp();
  }
}

// Fast gate + combine for 16-bit tensors (the prefill path after the tensor-core branch kernels): HBM-bound.
// Gate weights live in shared memory (fc1 transposed so lane u reads column u without bank conflicts); one warp per row:
// q_gp by 16-byte loads, lane u = hidden unit u of fc1, three warp sums for fc2, then O = sum_b g_b O_b by 16-byte chunks.
// Algorithmic bytes per row: (1 + 3 + 1) * h * D * 2.
constexpr int kCfWarps = 8;

template <typename T>
__global__ void __launch_bounds__(kCfWarps * 32, 4)
combine_fast_kernel(nsa_dims_t dm, const T* __restrict__ Q, nsa_gate_params_t gp, const T* __restrict__ O_br, T* __restrict__ O,
                    float* __restrict__ gates) {
  extern __shared__ float smem[];
  const int Dk = dm.Dk, H = dm.gate_hidden, h = dm.h;
  float* w1t = smem;                 // [Dk][H]
  float* b1 = w1t + Dk * H;          // [H]
  float* w2 = b1 + H;                // [3][H]
  float* b2 = w2 + 3 * H;            // [4]
  float* qg = b2 + 4;                // [warps][Dk]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool mlp = dm.gate_mode == NSA_GATE_MLP;
  if (mlp) {
    for (int i = threadIdx.x; i < Dk * H; i += blockDim.x) w1t[(i % Dk) * H + i / Dk] = gp.fc1_w[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) b1[i] = gp.fc1_b ? gp.fc1_b[i] : 0.f;
    for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) w2[i] = gp.fc2_w[i];
    if (threadIdx.x < 3) b2[threadIdx.x] = gp.fc2_b ? gp.fc2_b[threadIdx.x] : 0.f;
  }
  __syncthreads();
  float* qgp = qg + warp * Dk;
  const int n_rows = dm.B * dm.S * dm.G;
  const int row_elems = h * dm.Dv;           // multiple of 8 (checked on the host)
  const int chunks = row_elems / 8;          // 16-byte chunks per row
  const size_t per_branch = (size_t)n_rows * row_elems;
  const float inv_tau = 1.0f / fmaxf(dm.gate_tau, 1e-6f);
  for (int row = blockIdx.x * kCfWarps + warp; row < n_rows; row += gridDim.x * kCfWarps) {
    // the branch outputs do not depend on the gate: their loads (up to two 16-byte chunks per lane and branch) are issued before the
    // MLP, whose dependent FMA chain then hides their latency
    const size_t base = (size_t)row * row_elems;
    constexpr int kPre = 2;
    uint4 pa[kPre], pb[kPre], pc[kPre];
#pragma unroll
    for (int j = 0; j < kPre; ++j) {
      const int cidx = lane + 32 * j;
      if (cidx < chunks) {
        pa[j] = *reinterpret_cast<const uint4*>(O_br + base + cidx * 8);
        pb[j] = *reinterpret_cast<const uint4*>(O_br + per_branch + base + cidx * 8);
        pc[j] = *reinterpret_cast<const uint4*>(O_br + 2 * per_branch + base + cidx * 8);
      }
    }
    float g0 = 1.0f / 3.0f, g1 = 1.0f / 3.0f, g2 = 1.0f / 3.0f;
    if (dm.gate_mode == NSA_GATE_CMP) { g0 = 1.f; g1 = 0.f; g2 = 0.f; }
    else if (dm.gate_mode == NSA_GATE_SEL) { g0 = 0.f; g1 = 1.f; g2 = 0.f; }
    else if (dm.gate_mode == NSA_GATE_WIN) { g0 = 0.f; g1 = 0.f; g2 = 1.f; }
    else if (mlp) {
      // q_gp = mean over heads (nsa_attention.py:1357): lane owns k = 2*lane, 2*lane+1 (+64, ...)
      const T* qrow = Q + (size_t)row * h * Dk;
      for (int k = 2 * lane; k < Dk; k += 64) {
        float m0 = 0.f, m1 = 0.f;
        for (int h0 = 0; h0 < h; h0 += 8) {  // eight heads' loads in flight at a time
          uint32_t v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = h0 + u < h ? *reinterpret_cast<const uint32_t*>(qrow + (h0 + u) * Dk + k) : 0u;
#pragma unroll
          for (int u = 0; u < 8; ++u) {  // zero bits are +0.0 in both 16-bit formats
            m0 += (float)reinterpret_cast<const T*>(&v[u])[0];
            m1 += (float)reinterpret_cast<const T*>(&v[u])[1];
          }
        }
        qgp[k] = m0 / (float)h;
        qgp[k + 1] = m1 / (float)h;
      }
      __syncwarp();
      float l0 = 0.f, l1 = 0.f, l2 = 0.f;
      for (int u = lane; u < H; u += 32) {
        float a = b1[u];
        int k = 0;
        for (; k + 4 <= Dk; k += 4) {  // one 16-byte broadcast read serves four products (same summation order as the scalar loop)
          const float4 qv = *reinterpret_cast<const float4*>(qgp + k);
          a = fmaf(w1t[k * H + u], qv.x, a);
          a = fmaf(w1t[(k + 1) * H + u], qv.y, a);
          a = fmaf(w1t[(k + 2) * H + u], qv.z, a);
          a = fmaf(w1t[(k + 3) * H + u], qv.w, a);
        }
        for (; k < Dk; ++k) a = fmaf(w1t[k * H + u], qgp[k], a);
        const float x = a / (1.0f + expf(-a));  // silu
        l0 = fmaf(w2[u], x, l0);
        l1 = fmaf(w2[H + u], x, l1);
        l2 = fmaf(w2[2 * H + u], x, l2);
      }
      l0 = (warp_sum(l0) + b2[0]) * inv_tau;
      l1 = (warp_sum(l1) + b2[1]) * inv_tau;
      l2 = (warp_sum(l2) + b2[2]) * inv_tau;
      const float mx = fmaxf(l0, fmaxf(l1, l2));
      const int am = l0 >= l1 ? (l0 >= l2 ? 0 : 2) : (l1 >= l2 ? 1 : 2);  // first maximum
      const float second = am == 0 ? fmaxf(l1, l2) : (am == 1 ? fmaxf(l0, l2) : fmaxf(l0, l1));
      if (mx - second > 50.0f) {  // hard one-hot (nsa_attention.py:74-81)
        g0 = am == 0 ? 1.f : 0.f; g1 = am == 1 ? 1.f : 0.f; g2 = am == 2 ? 1.f : 0.f;
      } else {
        const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), e2 = expf(l2 - mx);
        const float inv = 1.0f / (e0 + e1 + e2);
        g0 = e0 * inv; g1 = e1 * inv; g2 = e2 * inv;
      }
      __syncwarp();
    }
    if (gates && lane == 0) {
      gates[(size_t)row * 3] = g0;
      gates[(size_t)row * 3 + 1] = g1;
      gates[(size_t)row * 3 + 2] = g2;
    }
    auto blend = [&](const uint4& a, const uint4& b, const uint4& c, int cidx) {
      const T* ea = reinterpret_cast<const T*>(&a);
      const T* eb = reinterpret_cast<const T*>(&b);
      const T* ec = reinterpret_cast<const T*>(&c);
      uint4 o;
      T* po = reinterpret_cast<T*>(&o);
#pragma unroll
      for (int e = 0; e < 8; ++e) po[e] = T(g0 * (float)ea[e] + g1 * (float)eb[e] + g2 * (float)ec[e]);
      *reinterpret_cast<uint4*>(O + base + cidx * 8) = o;
    };
#pragma unroll
    for (int j = 0; j < kPre; ++j)
      if (lane + 32 * j < chunks) blend(pa[j], pb[j], pc[j], lane + 32 * j);
    for (int cidx = lane + 32 * kPre; cidx < chunks; cidx += 32) {  // rows wider than 2 x 32 chunks (h * Dv > 512)
      const uint4 a = *reinterpret_cast<const uint4*>(O_br + base + cidx * 8);
      const uint4 b = *reinterpret_cast<const uint4*>(O_br + per_branch + base + cidx * 8);
      const uint4 c = *reinterpret_cast<const uint4*>(O_br + 2 * per_branch + base + cidx * 8);
      blend(a, b, c, cidx);
    }
  }
}

// The gate half of combine_fast_kernel on its own (same statements, so the gates are bit-identical): q_gp by 16-byte-friendly
// loads, weights in shared memory, one warp per row.  Used by the long no-grad prefill, where it runs on a side stream next to the
// scorer and the merge of the selected branch's partials then blends the three branches itself (sel2_merge_blend_kernel).
template <typename T>
__global__ void __launch_bounds__(kCfWarps * 32, 4)
gate_fast_kernel(nsa_dims_t dm, const T* __restrict__ Q, nsa_gate_params_t gp, float* __restrict__ gates) {
  extern __shared__ float smem[];
  const int Dk = dm.Dk, H = dm.gate_hidden, h = dm.h;
  float* w1t = smem;                 // [Dk][H]
  float* b1 = w1t + Dk * H;          // [H]
  float* w2 = b1 + H;                // [3][H]
  float* b2 = w2 + 3 * H;            // [4]
  float* qg = b2 + 4;                // [warps][Dk]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < Dk * H; i += blockDim.x) w1t[(i % Dk) * H + i / Dk] = gp.fc1_w[i];
  for (int i = threadIdx.x; i < H; i += blockDim.x) b1[i] = gp.fc1_b ? gp.fc1_b[i] : 0.f;
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) w2[i] = gp.fc2_w[i];
  if (threadIdx.x < 3) b2[threadIdx.x] = gp.fc2_b ? gp.fc2_b[threadIdx.x] : 0.f;
  __syncthreads();
  float* qgp = qg + warp * Dk;
  const int n_rows = dm.B * dm.S * dm.G;
  const float inv_tau = 1.0f / fmaxf(dm.gate_tau, 1e-6f);
  for (int row = blockIdx.x * kCfWarps + warp; row < n_rows; row += gridDim.x * kCfWarps) {
    float g0, g1, g2;
    const T* qrow = Q + (size_t)row * h * Dk;
    for (int k = 2 * lane; k < Dk; k += 64) {
      float m0 = 0.f, m1 = 0.f;
      for (int h0 = 0; h0 < h; h0 += 8) {  // eight heads' loads in flight at a time
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = h0 + u < h ? *reinterpret_cast<const uint32_t*>(qrow + (h0 + u) * Dk + k) : 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {  // zero bits are +0.0 in both 16-bit formats
          m0 += (float)reinterpret_cast<const T*>(&v[u])[0];
          m1 += (float)reinterpret_cast<const T*>(&v[u])[1];
        }
      }
      qgp[k] = m0 / (float)h;
      qgp[k + 1] = m1 / (float)h;
    }
    __syncwarp();
    float l0 = 0.f, l1 = 0.f, l2 = 0.f;
    for (int u = lane; u < H; u += 32) {
      float a = b1[u];
      int k = 0;
      for (; k + 4 <= Dk; k += 4) {
        const float4 qv = *reinterpret_cast<const float4*>(qgp + k);
        a = fmaf(w1t[k * H + u], qv.x, a);
        a = fmaf(w1t[(k + 1) * H + u], qv.y, a);
        a = fmaf(w1t[(k + 2) * H + u], qv.z, a);
        a = fmaf(w1t[(k + 3) * H + u], qv.w, a);
      }
      for (; k < Dk; ++k) a = fmaf(w1t[k * H + u], qgp[k], a);
      const float x = a / (1.0f + expf(-a));  // silu
      l0 = fmaf(w2[u], x, l0);
      l1 = fmaf(w2[H + u], x, l1);
      l2 = fmaf(w2[2 * H + u], x, l2);
    }
    l0 = (warp_sum(l0) + b2[0]) * inv_tau;
    l1 = (warp_sum(l1) + b2[1]) * inv_tau;
    l2 = (warp_sum(l2) + b2[2]) * inv_tau;
    const float mx = fmaxf(l0, fmaxf(l1, l2));
    const int am = l0 >= l1 ? (l0 >= l2 ? 0 : 2) : (l1 >= l2 ? 1 : 2);  // first maximum
    const float second = am == 0 ? fmaxf(l1, l2) : (am == 1 ? fmaxf(l0, l2) : fmaxf(l0, l1));
    if (mx - second > 50.0f) {  // hard one-hot (nsa_attention.py:74-81)
      g0 = am == 0 ? 1.f : 0.f; g1 = am == 1 ? 1.f : 0.f; g2 = am == 2 ? 1.f : 0.f;
    } else {
      const float e0 = expf(l0 - mx), e1 = expf(l1 - mx), e2 = expf(l2 - mx);
      const float inv = 1.0f / (e0 + e1 + e2);
      g0 = e0 * inv; g1 = e1 * inv; g2 = e2 * inv;
    }
    __syncwarp();
    if (lane == 0) {
      gates[(size_t)row * 3] = g0;
      gates[(size_t)row * 3 + 1] = g1;
      gates[(size_t)row * 3 + 2] = g2;
    }
  }
}

bool gate_fast_supported(const nsa_dims_t& dm, const nsa_gate_params_t* gp, const void* Q) {
  return gp && dm.gate_mode == NSA_GATE_MLP && gp->fc1_w && gp->fc2_w && dm.dtype != NSA_F32 && dm.Dk % 4 == 0 &&
         ((size_t)dm.Dk * dm.gate_hidden + 4 * dm.gate_hidden + 4 + (size_t)kCfWarps * dm.Dk) * sizeof(float) <= 48 * 1024 &&
         ((uintptr_t)Q & 3) == 0;
}

int launch_gate_fast(const nsa_dims_t& dm, const void* Q, const nsa_gate_params_t& gp, float* gates, cudaStream_t stream) {
  const int n_rows = dm.B * dm.S * dm.G;
  if (n_rows == 0) return NSA_OK;
  NSA_REQUIRE(gate_fast_supported(dm, &gp, Q) && gates, "gate_fast: 16-bit GateMLP only");
  const int H = dm.gate_hidden;
  const size_t smem = ((size_t)dm.Dk * H + H + 3 * H + 4 + (size_t)kCfWarps * dm.Dk) * sizeof(float);
  int blocks = ceil_div(n_rows, kCfWarps);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (dm.dtype == NSA_BF16) gate_fast_kernel<__nv_bfloat16><<<blocks, kCfWarps * 32, smem, stream>>>(dm, (const __nv_bfloat16*)Q, gp, gates);
  else gate_fast_kernel<__half><<<blocks, kCfWarps * 32, smem, stream>>>(dm, (const __half*)Q, gp, gates);
  return check_launch("gate_fast_kernel");
}

template <typename T>
static int launch_combine_fast(const nsa_dims_t& dm, const void* Q, const nsa_gate_params_t& gp, const void* O_br, void* O,
                               float* gates, cudaStream_t stream) {
  const int n_rows = dm.B * dm.S * dm.G;
  const int H = dm.gate_mode == NSA_GATE_MLP ? dm.gate_hidden : 0;
  size_t smem = ((size_t)dm.Dk * H + H + 3 * H + 4 + (size_t)kCfWarps * dm.Dk) * sizeof(float);
  int blocks = ceil_div(n_rows, kCfWarps);
  if (blocks > 148 * 8) blocks = 148 * 8;
  combine_fast_kernel<T><<<blocks, kCfWarps * 32, smem, stream>>>(dm, (const T*)Q, gp, (const T*)O_br, (T*)O, gates);
  return check_launch("combine_fast_kernel");
}

int launch_combine(const nsa_dims_t& dm, const void* Q, const nsa_gate_params_t& gp, const void* O_br, void* O, float* gates,
                   cudaStream_t stream) {
  const int n_rows = dm.B * dm.S * dm.G;
  if (n_rows == 0) return NSA_OK;
  const int H = dm.gate_mode == NSA_GATE_MLP ? dm.gate_hidden : 0;
  if (dm.dtype != NSA_F32 && (dm.h * dm.Dv) % 8 == 0 && dm.Dk % 4 == 0 &&
      ((size_t)dm.Dk * H + 4 * H + 4 + (size_t)kCfWarps * dm.Dk) * sizeof(float) <= 48 * 1024 &&
      ((uintptr_t)O_br & 15) == 0 && ((uintptr_t)O & 15) == 0 && ((uintptr_t)Q & 3) == 0) {
    if (dm.dtype == NSA_BF16) return launch_combine_fast<__nv_bfloat16>(dm, Q, gp, O_br, O, gates, stream);
    return launch_combine_fast<__half>(dm, Q, gp, O_br, O, gates, stream);
  }
  int blocks = ceil_div(n_rows, kAttnWarps);
  if (blocks > 148 * 16) blocks = 148 * 16;
  size_t smem = (size_t)kAttnWarps * (dm.Dk + 2 * dm.gate_hidden) * sizeof(float);
  combine_kernel<<<blocks, kAttnWarps * 32, smem, stream>>>(dm, Q, gp, O_br, O, gates);
  return check_launch("combine_kernel");
}

// Backward: recompute the MLP per row, accumulate parameter gradients in shared memory per CTA, flush with
// one atomicAdd per parameter per CTA.
__global__ void __launch_bounds__(kAttnWarps * 32)
gate_bwd_kernel(nsa_dims_t dm, const void* __restrict__ Q, nsa_gate_params_t gp, const float* __restrict__ dgates,
                float* __restrict__ dQ, float* d_fc1_w, float* d_fc1_b, float* d_fc2_w, float* d_fc2_b) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Dk = dm.Dk, H = dm.gate_hidden;
  float* a_w1 = smem;             // [H*Dk]
  float* a_b1 = a_w1 + H * Dk;    // [H]
  float* a_w2 = a_b1 + H;         // [3*H]
  float* a_b2 = a_w2 + 3 * H;     // [3] (+1 pad)
  float* wbase = a_b2 + 4 + (size_t)warp * (Dk + 3 * H);
  float* qgp = wbase;             // [Dk]
  float* xs = qgp + Dk;           // [H]
  float* pre = xs + H;            // [H]
  float* dpre = pre + H;          // [H]
  const int nacc = H * Dk + H + 3 * H + 4;
  for (int i = threadIdx.x; i < nacc; i += blockDim.x) smem[i] = 0.f;
  __syncthreads();
  const int n_rows = dm.B * dm.S * dm.G;
  const float inv_tau = 1.0f / fmaxf(dm.gate_tau, 1e-6f);
  for (int row = blockIdx.x * kAttnWarps + warp; row < n_rows; row += gridDim.x * kAttnWarps) {
    for (int k = lane; k < Dk; k += 32) {
      float m = 0.f;
      for (int hh = 0; hh < dm.h; ++hh) m += ld_elt(Q, ((size_t)row * dm.h + hh) * Dk + k, dm.dtype);
      qgp[k] = m / (float)dm.h;
    }
    __syncwarp();
    bool peaked = false;
    Gate3 gt = gate_forward_warp(qgp, xs, pre, gp, Dk, H, dm.gate_tau, NSA_GATE_MLP, &peaked);
    __syncwarp();
    if (!peaked) {  // the hard one-hot is a constant: no gradient (torch.where(peaked, one_hot, p))
      const float p[3] = {gt.c, gt.s, gt.w};
      const float dp[3] = {dgates[(size_t)row * 3], dgates[(size_t)row * 3 + 1], dgates[(size_t)row * 3 + 2]};
      const float dot = p[0] * dp[0] + p[1] * dp[1] + p[2] * dp[2];
      float dg[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) dg[c] = p[c] * (dp[c] - dot) * inv_tau;
      if (lane < 3) atomicAdd(a_b2 + lane, dg[lane]);
      for (int u = lane; u < H; u += 32) {
        float dx = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          atomicAdd(a_w2 + c * H + u, dg[c] * xs[u]);
          dx = fmaf(gp.fc2_w[c * H + u], dg[c], dx);
        }
        const float z = pre[u];
        const float sg = 1.0f / (1.0f + expf(-z));
        const float d = dx * sg * (1.0f + z * (1.0f - sg));  // d silu
        dpre[u] = d;
        atomicAdd(a_b1 + u, d);
      }
      __syncwarp();
      for (int u = 0; u < H; ++u) {
        const float d = dpre[u];
        for (int k = lane; k < Dk; k += 32) atomicAdd(a_w1 + u * Dk + k, d * qgp[k]);
      }
      if (dQ) {
        for (int k = lane; k < Dk; k += 32) {
          float dqg = 0.f;
          for (int u = 0; u < H; ++u) dqg = fmaf(gp.fc1_w[(size_t)u * Dk + k], dpre[u], dqg);
          dqg /= (float)dm.h;
          for (int hh = 0; hh < dm.h; ++hh) dQ[((size_t)row * dm.h + hh) * Dk + k] += dqg;  // row owned by this warp
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < H * Dk; i += blockDim.x) atomicAdd(d_fc1_w + i, a_w1[i]);
  for (int i = threadIdx.x; i < H; i += blockDim.x) atomicAdd(d_fc1_b + i, a_b1[i]);
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) atomicAdd(d_fc2_w + i, a_w2[i]);
  for (int i = threadIdx.x; i < 3; i += blockDim.x) atomicAdd(d_fc2_b + i, a_b2[i]);
}

// Fast gate backward for the training shapes (16-bit Q, Dk = 64, hidden <= 32): lane u owns hidden unit u and keeps its row
// of d_fc1_w (64 values) and its d_fc1_b / d_fc2_w entries in registers over all rows of the warp, so a row costs ~200
// instructions per lane instead of 2048 shared-memory atomics; weights sit in shared memory in both orientations.  Partial
// sums are folded per CTA in shared memory and flushed with one atomicAdd per parameter per CTA.
constexpr int kGbWarps = 8;

template <typename T>
__global__ void __launch_bounds__(kGbWarps * 32, 2)
gate_bwd_fast_kernel(nsa_dims_t dm, const T* __restrict__ Q, nsa_gate_params_t gp, const float* __restrict__ dgates,
                     float* __restrict__ dQ, float* d_fc1_w, float* d_fc1_b, float* d_fc2_w, float* d_fc2_b) {
  constexpr int DK = 64;
  extern __shared__ float smem[];
  const int H = dm.gate_hidden, h = dm.h;
  float* w1 = smem;                 // [H][DK]
  float* w1t = w1 + H * DK;         // [DK][H]
  float* w2 = w1t + DK * H;         // [3][H]
  float* acc_s = w2 + 3 * H;        // [H*DK + H + 3*H + 4] CTA accumulators
  float* per_warp = acc_s + H * DK + 4 * H + 4;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qgp = per_warp + warp * (DK + 32);  // [DK]
  float* dpre_s = qgp + DK;                  // [32]
  for (int i = threadIdx.x; i < H * DK; i += blockDim.x) {
    const float v = gp.fc1_w[i];
    w1[i] = v;
    w1t[(i % DK) * H + i / DK] = v;
  }
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) w2[i] = gp.fc2_w[i];
  for (int i = threadIdx.x; i < H * DK + 4 * H + 4; i += blockDim.x) acc_s[i] = 0.f;
  __syncthreads();
  const bool unit = lane < H;
  const float b1 = unit && gp.fc1_b ? gp.fc1_b[lane] : 0.f;
  const float b2[3] = {gp.fc2_b ? gp.fc2_b[0] : 0.f, gp.fc2_b ? gp.fc2_b[1] : 0.f, gp.fc2_b ? gp.fc2_b[2] : 0.f};
  const float w2u[3] = {unit ? w2[lane] : 0.f, unit ? w2[H + lane] : 0.f, unit ? w2[2 * H + lane] : 0.f};
  float a_w1[DK], a_b1 = 0.f, a_w2[3] = {0.f, 0.f, 0.f}, a_b2 = 0.f;
#pragma unroll
  for (int k = 0; k < DK; ++k) a_w1[k] = 0.f;
  const int n_rows = dm.B * dm.S * dm.G;
  const float inv_tau = 1.0f / fmaxf(dm.gate_tau, 1e-6f), inv_h = 1.0f / (float)h;
  // the next row's Q heads and dgates are requested before the current row is processed (a row is a chain of dependent
  // steps; without the prefetch every row also pays a global-memory round trip)
  constexpr int HM = 8;
  uint32_t qn[HM];
  float dn0 = 0.f, dn1 = 0.f, dn2 = 0.f;
  auto fetch = [&](int row) {
    if (row < n_rows) {
#pragma unroll
      for (int hh = 0; hh < HM; ++hh)
        qn[hh] = hh < h ? reinterpret_cast<const uint32_t*>(Q + ((size_t)row * h + hh) * DK)[lane] : 0u;
      dn0 = dgates[(size_t)row * 3];
      dn1 = dgates[(size_t)row * 3 + 1];
      dn2 = dgates[(size_t)row * 3 + 2];
    }
  };
  fetch(blockIdx.x * kGbWarps + warp);
  for (int row = blockIdx.x * kGbWarps + warp; row < n_rows; row += gridDim.x * kGbWarps) {
    // q_gp = mean over heads; lane owns k = 2*lane, 2*lane+1 (one 4-byte load per head)
    float q0 = 0.f, q1 = 0.f;
#pragma unroll
    for (int hh = 0; hh < HM; ++hh) {
      const uint32_t u = qn[hh];  // zero bits for hh >= h: +0.0 in both 16-bit formats
      T lo, hi;
      memcpy(&lo, &u, 2);
      memcpy(&hi, reinterpret_cast<const char*>(&u) + 2, 2);
      q0 += (float)lo;
      q1 += (float)hi;
    }
    const float dp0 = dn0, dp1 = dn1, dp2 = dn2;
    fetch(row + gridDim.x * kGbWarps);
    __syncwarp();
    reinterpret_cast<float2*>(qgp)[lane] = make_float2(q0 * inv_h, q1 * inv_h);
    __syncwarp();
    float pre = b1;
    if (unit) {
#pragma unroll
      for (int k = 0; k < DK; k += 4) {
        const float4 qv = *reinterpret_cast<const float4*>(qgp + k);
        pre = fmaf(w1t[k * H + lane], qv.x, pre);
        pre = fmaf(w1t[(k + 1) * H + lane], qv.y, pre);
        pre = fmaf(w1t[(k + 2) * H + lane], qv.z, pre);
        pre = fmaf(w1t[(k + 3) * H + lane], qv.w, pre);
      }
    }
    const float sg = 1.0f / (1.0f + expf(-pre));
    const float xs = unit ? pre * sg : 0.f;
    float g[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) g[c] = (warp_sum(w2u[c] * xs) + b2[c]) * inv_tau;
    const float mx = fmaxf(g[0], fmaxf(g[1], g[2]));
    const int am = g[0] >= g[1] ? (g[0] >= g[2] ? 0 : 2) : (g[1] >= g[2] ? 1 : 2);
    const float second = am == 0 ? fmaxf(g[1], g[2]) : (am == 1 ? fmaxf(g[0], g[2]) : fmaxf(g[0], g[1]));
    if (mx - second > 50.0f) continue;  // hard one-hot: constant, no gradient (nsa_attention.py:74-81); warp-uniform
    const float e0 = expf(g[0] - mx), e1 = expf(g[1] - mx), e2 = expf(g[2] - mx);
    const float inv = 1.0f / (e0 + e1 + e2);
    const float p[3] = {e0 * inv, e1 * inv, e2 * inv};
    const float dot = p[0] * dp0 + p[1] * dp1 + p[2] * dp2;
    const float dg[3] = {p[0] * (dp0 - dot) * inv_tau, p[1] * (dp1 - dot) * inv_tau, p[2] * (dp2 - dot) * inv_tau};
    if (lane < 3) a_b2 += lane == 0 ? dg[0] : (lane == 1 ? dg[1] : dg[2]);
    float dx = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      a_w2[c] = fmaf(dg[c], xs, a_w2[c]);
      dx = fmaf(w2u[c], dg[c], dx);
    }
    const float d = unit ? dx * sg * (1.0f + pre * (1.0f - sg)) : 0.f;  // d silu
    a_b1 += d;
#pragma unroll
    for (int k = 0; k < DK; k += 4) {
      const float4 qv = *reinterpret_cast<const float4*>(qgp + k);
      a_w1[k] = fmaf(d, qv.x, a_w1[k]);
      a_w1[k + 1] = fmaf(d, qv.y, a_w1[k + 1]);
      a_w1[k + 2] = fmaf(d, qv.z, a_w1[k + 2]);
      a_w1[k + 3] = fmaf(d, qv.w, a_w1[k + 3]);
    }
    if (dQ) {
      dpre_s[lane] = d;
      __syncwarp();
      float dq0 = 0.f, dq1 = 0.f;
      for (int u = 0; u < H; ++u) {
        const float2 wv = reinterpret_cast<const float2*>(w1 + u * DK)[lane];
        const float du = dpre_s[u];
        dq0 = fmaf(wv.x, du, dq0);
        dq1 = fmaf(wv.y, du, dq1);
      }
      dq0 *= inv_h;
      dq1 *= inv_h;
      // row owned by this warp; all heads are read before any is written (one global round trip, not h dependent ones)
      float2 cur[HM];
#pragma unroll
      for (int hh = 0; hh < HM; ++hh)
        if (hh < h) cur[hh] = *(reinterpret_cast<const float2*>(dQ + ((size_t)row * h + hh) * DK) + lane);
#pragma unroll
      for (int hh = 0; hh < HM; ++hh)
        if (hh < h) *(reinterpret_cast<float2*>(dQ + ((size_t)row * h + hh) * DK) + lane) = make_float2(cur[hh].x + dq0, cur[hh].y + dq1);
    }
  }
  if (unit) {
#pragma unroll
    for (int k = 0; k < DK; ++k) atomicAdd(acc_s + k * H + lane, a_w1[k]);  // [DK][H]: lanes on distinct banks
    atomicAdd(acc_s + H * DK + lane, a_b1);
#pragma unroll
    for (int c = 0; c < 3; ++c) atomicAdd(acc_s + H * DK + H + c * H + lane, a_w2[c]);
  }
  if (lane < 3) atomicAdd(acc_s + H * DK + 4 * H + lane, a_b2);
  __syncthreads();
  for (int i = threadIdx.x; i < H * DK; i += blockDim.x) atomicAdd(d_fc1_w + i, acc_s[(i % DK) * H + i / DK]);
  for (int i = threadIdx.x; i < H; i += blockDim.x) atomicAdd(d_fc1_b + i, acc_s[H * DK + i]);
  for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) atomicAdd(d_fc2_w + i, acc_s[H * DK + H + i]);
  for (int i = threadIdx.x; i < 3; i += blockDim.x) atomicAdd(d_fc2_b + i, acc_s[H * DK + 4 * H + i]);
}

template <typename T>
static int launch_gate_bwd_fast(const nsa_dims_t& dm, const void* Q, const nsa_gate_params_t& gp, const float* dgates, float* dQ,
                                float* d_fc1_w, float* d_fc1_b, float* d_fc2_w, float* d_fc2_b, cudaStream_t stream) {
  const int H = dm.gate_hidden, n_rows = dm.B * dm.S * dm.G;
  const size_t smem = ((size_t)3 * H * 64 + 3 * H + 4 * H + 4 + (size_t)kGbWarps * (64 + 32)) * sizeof(float);
  int blocks = ceil_div(n_rows, kGbWarps * 4);
  if (blocks > 148 * 2) blocks = 148 * 2;
  gate_bwd_fast_kernel<T><<<blocks, kGbWarps * 32, smem, stream>>>(dm, (const T*)Q, gp, dgates, dQ, d_fc1_w, d_fc1_b, d_fc2_w, d_fc2_b);
  return check_launch("gate_bwd_fast_kernel");
}

int launch_gate_bwd(const nsa_dims_t& dm, const void* Q, const nsa_gate_params_t& gp, const float* dgates, float* dQ,
                    float* d_fc1_w, float* d_fc1_b, float* d_fc2_w, float* d_fc2_b, cudaStream_t stream) {
  const int n_rows = dm.B * dm.S * dm.G;
  if (n_rows == 0 || dm.gate_mode != NSA_GATE_MLP) return NSA_OK;
  const int H = dm.gate_hidden, Dk = dm.Dk;
  if (Dk == 64 && H >= 1 && H <= 32 && dm.h <= 8 && dm.dtype != NSA_F32) {
    if (dm.dtype == NSA_BF16) return launch_gate_bwd_fast<__nv_bfloat16>(dm, Q, gp, dgates, dQ, d_fc1_w, d_fc1_b, d_fc2_w, d_fc2_b, stream);
    return launch_gate_bwd_fast<__half>(dm, Q, gp, dgates, dQ, d_fc1_w, d_fc1_b, d_fc2_w, d_fc2_b, stream);
  }
  size_t smem = ((size_t)H * Dk + H + 3 * H + 4 + (size_t)kAttnWarps * (Dk + 3 * H)) * sizeof(float);
  NSA_REQUIRE(smem <= 200 * 1024, "gate bwd: hidden=%d Dk=%d needs %zu B of shared memory", H, Dk, smem);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(gate_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("gate bwd: smem attr: %s", cudaGetErrorString(e)); return NSA_ERR_CUDA; }
  }
  int blocks = ceil_div(n_rows, kAttnWarps * 8);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  gate_bwd_kernel<<<blocks, kAttnWarps * 32, smem, stream>>>(dm, Q, gp, dgates, dQ, d_fc1_w, d_fc1_b, d_fc2_w, d_fc2_b);
  return check_launch("gate_bwd_kernel");
}

}  // namespace nsa
