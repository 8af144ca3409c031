// Selected-branch attention on tcgen05 tensor cores: a block-sparse gather kernel.
//
// One CTA walks a tile of consecutive query tokens of one (b, g).  For each token the selected ranges are cut into
// 64-key blocks; TMA (SWIZZLE_128B tiles addressed through the ranges) streams the K blocks, then the V blocks,
// through a 4 x 16 KB shared-memory ring.  All h <= 8 heads of the KV group share every block (GQA).
//
//   S^T[128 keys x 16 slots] = K_pair[128 x 64] . Q^T[64 x 16]     (M=128, N=16 (M=128 needs N%16==0), K=16 x4; keys on TMEM lanes)
//   softmax over <= 1024 keys: exact two-phase (all S^T tiles stay in TMEM, 8 columns per 128 keys)
//   O^T[64 dv x 8 heads]    += V_pair^T[64 x 128] . P^T[128 x 8]    (M=64, N=8, K=16 x8; V is the MN-major A operand)
//
// Swapping the operands (keys as M, heads as N) is what makes M = h = 6 rows of work fill a tensor-core tile.
// Replaces grouped_selection_attention* / Triton sel_fwd / sel_cuda.cpp for bf16/fp16, Dk = Dv = 64.
#include "tc_common.cuh"
#include "launchers.h"

namespace nsa {
using namespace tc;

constexpr int kSelN = 8;           // head slots used (N of the P.V MMA)
constexpr int kSelNQ = 16;         // N of the Q.K^T MMA: tcgen05 kind::f16 with M=128 needs N % 16 == 0; slots 8..15 are zero
constexpr int kSelPairStages = 4;  // ring depth in 128-key pair stages
constexpr int kSelMaxBlk = 16;     // <= 1024 selected keys per row
constexpr int kSelMaxPairs = kSelMaxBlk / 2;
constexpr int kPairBytes = 128 * 128;  // 128 rows x 64 x 2 B
constexpr int kSelTmemCols = 256;      // 8 pairs x 16 columns of S^T + 8 columns of O^T -> next power of two

struct SelSmem {
  // offsets into dynamic shared memory (base 1024-aligned)
  static constexpr int ring = 0;
  static constexpr int P = ring + kSelPairStages * kPairBytes;       // [pairs][128 keys][8 heads] bf16
  static constexpr int Q = P + kSelMaxPairs * 128 * kSelN * 2;       // [8 k-chunks][8 heads][8] bf16 (no swizzle)
  static constexpr int misc = Q + 8 * kSelNQ * 8 * 2;            // two groups of 8 head rows
  static constexpr int total = misc + 1024;
};

struct SelMisc {
  uint64_t full[kSelPairStages], empty[kSelPairStages], s_done, o_done;
  uint32_t tmem_base;
  int nblk;
  int blk_row[kSelMaxBlk], blk_valid[kSelMaxBlk];
  float red_max[4][kSelN], red_sum[4][kSelN];
};

template <typename T>
__global__ void __launch_bounds__(128)
sel_attn_tc_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, nsa_dims_t dm,
                   const T* __restrict__ Q, const int32_t* __restrict__ ranges, T* __restrict__ O, float* __restrict__ lse,
                   int tokens_per_cta) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  SelMisc* ms = reinterpret_cast<SelMisc*>(smem + SelSmem::misc);
  uint8_t* ring = smem + SelSmem::ring;
  uint8_t* Pbuf = smem + SelSmem::P;
  uint8_t* Qbuf = smem + SelSmem::Q;

  // ---- one-time setup ---------------------------------------------------------------------------------
  {  // zero the ring so a never-loaded half of a pair stage holds finite data (it is masked, but 0 * NaN = NaN)
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < kSelPairStages * kPairBytes / 16; i += 128) reinterpret_cast<uint4*>(ring)[i] = z;
  }
  if (tid == 0) {
    for (int i = 0; i < kSelPairStages; ++i) { mbar_init(&ms->full[i], 1); mbar_init(&ms->empty[i], 1); }
    mbar_init(&ms->s_done, 1);
    mbar_init(&ms->o_done, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 0) tmem_alloc(&ms->tmem_base, kSelTmemCols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ms->tmem_base;
  const uint32_t tmem_S = tmem;                                  // pair j -> columns [j*16, j*16+16), first 8 used
  const uint32_t tmem_O = tmem + kSelMaxPairs * kSelNQ;          // 8 columns, M=64 layout
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  const int tiles_per_seq = ceil_div(dm.S, tokens_per_cta);
  const int tile = blockIdx.x % tiles_per_seq;
  const int bg = blockIdx.x / tiles_per_seq;
  const int g = bg % dm.G, b = bg / dm.G;
  const int h = dm.h;
  constexpr uint32_t idesc_qk = make_idesc_f16(128, kSelNQ, TcType<T>::fmt, 0, 0);
  constexpr uint32_t idesc_pv = make_idesc_f16(64, kSelN, TcType<T>::fmt, 1, 1);
  const float sl2 = dm.scale * kLog2e;

  uint32_t n_loads = 0;   // pair loads issued so far by this CTA (ring position)   [thread 0]
  uint32_t n_cons = 0;    // pair loads consumed so far                              [thread 0]
  uint32_t tok_par = 0;   // parity of s_done / o_done

  for (int ti = 0; ti < tokens_per_cta; ++ti) {
    const int s = tile * tokens_per_cta + ti;
    if (s >= dm.S) break;
    const size_t row = ((size_t)b * dm.S + s) * dm.G + g;

    // ---- block list + Q^T ------------------------------------------------------------------------------
    if (tid == 0) {
      const int32_t* rr = ranges + row * dm.n_ranges * 2;
      int n = 0;
      for (int i = 0; i < dm.n_ranges; ++i) {
        int a0 = rr[2 * i], a1 = rr[2 * i + 1];
        if (a0 < 0) a0 = 0;
        if (a1 > dm.S_sel_kv) a1 = dm.S_sel_kv;
        for (int p = a0; p < a1 && n < kSelMaxBlk; p += 64) {
          ms->blk_row[n] = p;
          ms->blk_valid[n] = a1 - p < 64 ? a1 - p : 64;
          ++n;
        }
      }
      ms->nblk = n;
    }
    {  // Q^T into the no-swizzle K-major core-matrix layout: (head, k) -> (k/8)*128 + head*16 + (k%8)*2 bytes
      const T* qrow = Q + row * h * 64;
      for (int i = tid; i < kSelNQ * 8; i += 128) {  // one 16-byte chunk per (head slot, k-chunk)
        const int head = i >> 3, kc = i & 7;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (head < h) v = *reinterpret_cast<const uint4*>(qrow + head * 64 + kc * 8);
        *reinterpret_cast<uint4*>(Qbuf + (head >> 3) * 1024 + kc * 128 + (head & 7) * 16) = v;
      }
    }
    fence_proxy_async();
    __syncthreads();
    const int nblk = ms->nblk;
    const int np = (nblk + 1) >> 1;
    if (nblk == 0) {  // empty row -> zeros, lse = -inf (attention_kernels.py:769-771)
      for (int i = tid; i < h * 64; i += 128) O[row * h * 64 + i] = T(0.f);
      if (lse && tid < h) lse[row * h + tid] = -INFINITY;
      __syncthreads();
      continue;
    }

    // ---- producer / MMA issue (thread 0) -----------------------------------------------------------------
    const uint32_t load_base = n_loads;
    auto issue_load = [&](int idx) {  // idx in [0, 2*np): K pairs then V pairs
      const uint32_t st = n_loads % kSelPairStages;
      mbar_wait(&ms->empty[st], ((n_loads / kSelPairStages) & 1) ^ 1);
      const int pj = idx < np ? idx : idx - np;
      const CUtensorMap* tm = idx < np ? &tmK : &tmV;
      const int nb = (2 * pj + 1 < nblk) ? 2 : 1;
      mbar_expect_tx(&ms->full[st], nb * (kPairBytes / 2));
      tma_load_3d(ring + st * kPairBytes, tm, &ms->full[st], 0, ms->blk_row[2 * pj], bg);
      if (nb == 2) tma_load_3d(ring + st * kPairBytes + kPairBytes / 2, tm, &ms->full[st], 0, ms->blk_row[2 * pj + 1], bg);
      ++n_loads;
    };
    auto pump = [&](int consumed) {
      while ((int)(n_loads - load_base) < 2 * np && (int)(n_loads - load_base) < consumed + kSelPairStages)
        issue_load((int)(n_loads - load_base));
    };
    if (tid == 0) {
      pump(0);
      for (int j = 0; j < np; ++j) {
        const uint32_t st = n_cons % kSelPairStages;
        mbar_wait(&ms->full[st], (n_cons / kSelPairStages) & 1);
        tc_fence_after();
        const uint32_t a_base = smem_u32(ring + st * kPairBytes);
        const uint32_t q_base = smem_u32(Qbuf);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t ad = make_smem_desc(a_base + k * 32, 16, 1024, kSwizzle128B);
          const uint64_t bd = make_smem_desc(q_base + k * 256, 128, 1024, kSwizzleNone);
          umma_f16(tmem_S + j * kSelNQ, ad, bd, idesc_qk, k > 0);
        }
        umma_commit(&ms->empty[st]);
        ++n_cons;
        pump(j + 1);
      }
      umma_commit(&ms->s_done);
    }

    // ---- softmax: thread = key lane of every pair ---------------------------------------------------------
    mbar_wait(&ms->s_done, tok_par);
    tc_fence_after();
    float sc[kSelMaxPairs][kSelN];
    float mx[kSelN];
#pragma unroll
    for (int e = 0; e < kSelN; ++e) mx[e] = -INFINITY;
#pragma unroll
    for (int j = 0; j < kSelMaxPairs; ++j) {
      if (j < np) {
        uint32_t r[8];
        tmem_ld8(tmem_S + lane_base + j * kSelNQ, r);
        tmem_ld_wait();
        const int blk = 2 * j + (tid >> 6);
        const bool ok = blk < nblk && (tid & 63) < ms->blk_valid[blk];
#pragma unroll
        for (int e = 0; e < kSelN; ++e) {
          sc[j][e] = ok ? __uint_as_float(r[e]) : -INFINITY;
          mx[e] = fmaxf(mx[e], sc[j][e]);
        }
      }
    }
#pragma unroll
    for (int e = 0; e < kSelN; ++e) mx[e] = warp_max(mx[e]);
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < kSelN; ++e) ms->red_max[warp][e] = mx[e];
    }
    tc_fence_before();
    __syncthreads();
    float sum[kSelN];
#pragma unroll
    for (int e = 0; e < kSelN; ++e) {
      mx[e] = fmaxf(fmaxf(ms->red_max[0][e], ms->red_max[1][e]), fmaxf(ms->red_max[2][e], ms->red_max[3][e]));
      sum[e] = 0.f;
    }
#pragma unroll
    for (int j = 0; j < kSelMaxPairs; ++j) {
      if (j < np) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < kSelN; e += 2) {
          const float p0 = exp2f((sc[j][e] - mx[e]) * sl2);      // masked keys: exp2(-inf) = 0
          const float p1 = exp2f((sc[j][e + 1] - mx[e + 1]) * sl2);
          sum[e] += p0;
          sum[e + 1] += p1;
          pk[e >> 1] = pack2(T(), p0, p1);
        }
        // P^T for the MN-major no-swizzle B operand: [pair][key][8 heads] -> one 16-byte row per key
        *reinterpret_cast<uint4*>(Pbuf + (j * 128 + tid) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
#pragma unroll
    for (int e = 0; e < kSelN; ++e) sum[e] = warp_sum(sum[e]);
    if (lane == 0) {
#pragma unroll
      for (int e = 0; e < kSelN; ++e) ms->red_sum[warp][e] = sum[e];
    }
    fence_proxy_async();
    __syncthreads();

    // ---- P.V (thread 0 issues) -------------------------------------------------------------------------------
    if (tid == 0) {
      tc_fence_after();
      bool first = true;
      for (int j = 0; j < np; ++j) {
        const uint32_t st = n_cons % kSelPairStages;
        mbar_wait(&ms->full[st], (n_cons / kSelPairStages) & 1);
        tc_fence_after();
        const uint32_t v_base = smem_u32(ring + st * kPairBytes);
        const uint32_t p_base = smem_u32(Pbuf + j * 128 * 16);
        int keys = ms->blk_valid[2 * j];
        if (2 * j + 1 < nblk) keys = 64 + ms->blk_valid[2 * j + 1];
        const int ksteps = (keys + 15) >> 4;  // skip k-steps made only of masked keys
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t ad = make_smem_desc(v_base + k * 2048, 8192, 1024, kSwizzle128B);  // V^T: MN-major A
          const uint64_t bd = make_smem_desc(p_base + k * 256, 128, 2048, kSwizzleNone);    // P^T: MN-major B
          umma_f16(tmem_O, ad, bd, idesc_pv, first ? 0u : 1u);
          first = false;
        }
        umma_commit(&ms->empty[st]);
        ++n_cons;
        pump(np + j + 1);
      }
      umma_commit(&ms->o_done);
    }

    // ---- epilogue: O^T (64 dv x 8 heads, M=64 layout: dv row r on lane 32*(r/16) + r%16) -------------------------
    mbar_wait(&ms->o_done, tok_par);
    tc_fence_after();
    {
      uint32_t r[8];
      tmem_ld8(tmem_O + lane_base, r);
      tmem_ld_wait();
      if (lane < 16) {
        const int dv = warp * 16 + lane;
#pragma unroll
        for (int e = 0; e < kSelN; ++e) {
          if (e < h) {
            const float l = ms->red_sum[0][e] + ms->red_sum[1][e] + ms->red_sum[2][e] + ms->red_sum[3][e];
            O[(row * h + e) * 64 + dv] = T(__uint_as_float(r[e]) / l);
          }
        }
      }
      if (lse && tid < h) {
        const float l = ms->red_sum[0][tid] + ms->red_sum[1][tid] + ms->red_sum[2][tid] + ms->red_sum[3][tid];
        const float m = fmaxf(fmaxf(ms->red_max[0][tid], ms->red_max[1][tid]), fmaxf(ms->red_max[2][tid], ms->red_max[3][tid]));
        lse[row * h + tid] = m * dm.scale + logf(l);
      }
    }
    tok_par ^= 1;
    tc_fence_before();
    __syncthreads();  // Q / P / red buffers and TMEM are reused by the next token
  }

  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kSelTmemCols);
}

template <typename T>
static int launch_sel_t(const nsa_dims_t& dm, const void* Q, const void* K, const void* V, const int32_t* ranges, void* O,
                        float* lse, cudaStream_t stream) {
  CUtensorMap tmK, tmV;
  const int slabs = dm.B * dm.G;
  if (int rc = make_tmap_rows(&tmK, K, dm.dtype, 64, dm.S_sel_kv, 64, (long long)dm.cap_sel * 64, slabs, 64)) return rc;
  if (int rc = make_tmap_rows(&tmV, V, dm.dtype, 64, dm.S_sel_kv, 64, (long long)dm.cap_sel * 64, slabs, 64)) return rc;
  const int tpc = 16;
  const int grid = slabs * ceil_div(dm.S, tpc);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sel_attn_tc_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, SelSmem::total);
    if (e != cudaSuccess) { set_error("sel tc: smem attr: %s", cudaGetErrorString(e)); return NSA_ERR_CUDA; }
    attr_set = true;
  }
  sel_attn_tc_kernel<T><<<grid, 128, SelSmem::total, stream>>>(tmK, tmV, dm, (const T*)Q, ranges, (T*)O, lse, tpc);
  return check_launch("sel_attn_tc_kernel");
}

bool tc_sel_supported(const nsa_dims_t& dm) {
  return (dm.dtype == NSA_BF16 || dm.dtype == NSA_F16) && dm.Dk == 64 && dm.Dv == 64 && dm.h <= kSelN && dm.l_sel % 64 == 0 &&
         (long long)dm.n_sel * dm.l_sel <= 64 * kSelMaxBlk && dm.n_ranges * 1 <= 64;
}

int launch_sel_tc(const nsa_dims_t& dm, const void* Q, const void* K, const void* V, const int32_t* ranges, void* O, float* lse,
                  cudaStream_t stream) {
  static_assert(sizeof(SelMisc) <= 1024, "SelMisc must fit its slot");
  if (dm.B * dm.S * dm.G == 0) return NSA_OK;
  if (dm.dtype == NSA_BF16) return launch_sel_t<__nv_bfloat16>(dm, Q, K, V, ranges, O, lse, stream);
  return launch_sel_t<__half>(dm, Q, K, V, ranges, O, lse, stream);
}

}  // namespace nsa
