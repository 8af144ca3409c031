// Dense ranged attention on tcgen05 tensor cores: the compressed and sliding-window branches of NSA prefill.
//   cmp: row t attends compressed tokens [0, num_cmp(t))            (attention_kernels.py:106-143, packing.py:15-23)
//   win: row t attends cache rows of tokens [max(0, t-w+1), t]      (attention_kernels.py:146-178)
// both with softmax over every allowed key.  Flash-attention structure, B200 style:
//   * one CTA = MT M-tiles of 128 rows, a row being one (token, head) of one (b, g): TOK = 128/h tokens per M-tile, so
//     every K/V tile is shared by all h heads of TOK*MT tokens (GQA reuse) and loaded once by TMA;
//   * S = Q.K^T (M=128, N=128, K=64) into TMEM; softmax warps (thread = TMEM lane = row) take the tile max, write
//     P = exp2(s*c - m*c) as bf16 into 128B-swizzled shared memory; O_t = P.V (M=128, N=64, K=128) into TMEM with V as the
//     MN-major B operand; O_t is folded into fp32 register accumulators with the online-softmax rescale;
//   * the two M-tiles ping-pong: while one runs its softmax (MUFU-bound), the tensor core works for the other.
// Warp roles: warps [0, 4*MT) softmax, warp 4*MT TMA producer, warp 4*MT+1 MMA issuer.
#include <stdlib.h>

#include "tc_common.cuh"
#include "launchers.h"

namespace nsa {
using namespace tc;

constexpr int kDnMaxStages = 3;      // K / V ring stages (DnSmem::KS of them in use)
constexpr int kDnTile = 128 * 128;   // bytes: 128 rows x 64 x 2 B
constexpr int kDnMaxMT = 4;

// Two shapes of the same kernel: (MT, NK) = (4, 64) -- 4 M-tiles x 64-key tiles, 16 softmax warps, TMEM exactly full
// (4 x 64 S columns + 4 x 64 O columns) -- is the default; (2, 128) is kept for comparison (NSA_B200_DENSE_MT=2).
template <int MT, int NK>
struct DnSmem {
  static constexpr int KS = (MT == 2 && NK == 64) ? 2 : 3;  // K and V ring stages; the (2, 64) shape leaves room for a second CTA
  static constexpr int tmem_cols = MT * (NK + 64) <= 256 ? 256 : 512;
  static constexpr int ctas_per_sm = (MT == 2 && NK == 64) ? 2 : 1;
  static constexpr int kvtile = NK * 128;                   // bytes of one K or V tile
  static constexpr int ptile = 128 * NK * 2;                // bytes of one P tile: [NK/64 key halves][128 rows][128 B]
  static constexpr int q = 0;                               // MT x 16 KB
  static constexpr int k = q + MT * kDnTile;
  static constexpr int v = k + KS * kvtile;
  static constexpr int p = v + KS * kvtile;
  static constexpr int misc = p + MT * ptile;
  static constexpr int total = misc + 512 + 1024;
};

struct DnMisc {
  uint64_t q_full;
  uint64_t k_full[kDnMaxStages], k_empty[kDnMaxStages], v_full[kDnMaxStages], v_empty[kDnMaxStages];
  uint64_t s_full[kDnMaxMT], s_empty[kDnMaxMT], p_full[kDnMaxMT], p_empty[kDnMaxMT];
  uint32_t tmem_base;
};

__device__ __forceinline__ float dn_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void dn_ld_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

__device__ __forceinline__ void dn_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// key range of one row in cache-row coordinates
__device__ __forceinline__ void dn_row_range(const nsa_dims_t& dm, int branch, int t, int& lo, int& hi) {
  if (branch == 0) {
    lo = 0;
    hi = num_cmp_at(t, dm.l, dm.d, dm.S_cmp);
  } else {
    int a = t - dm.w + 1;
    if (a < dm.win_off) a = dm.win_off;
    if (a < 0) a = 0;
    lo = a - dm.win_off;
    hi = t + 1 - dm.win_off;
    if (hi > dm.S_win_kv) hi = dm.S_win_kv;
    if (dm.w <= 0 || hi < lo) hi = lo;
  }
}

#ifdef NSA_DENSE_DBG  // compile-time: timeline of one CTA (tag, tile, clock) for tools/dbg_dense.py
#define DDBG(tag, it)                                                                 \
  do {                                                                                \
    if (dbg && blockIdx.x == gridDim.x / 2 + 200) {                                   \
      const unsigned long long i_ = atomicAdd((unsigned long long*)dbg, 1ull);        \
      if (i_ < 4000) { dbg[1 + 2 * i_] = ((long long)(tag) << 32) | (unsigned)(it); dbg[2 + 2 * i_] = clock64(); } \
    }                                                                                 \
  } while (0)
#else
#define DDBG(tag, it) do { } while (0)
#endif

template <typename T, int MT, int NK>
__global__ void __launch_bounds__(32 * (4 * MT + 2), DnSmem<MT, NK>::ctas_per_sm)
dense_attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, nsa_dims_t dm, int branch, T* __restrict__ O,
                     float* __restrict__ lse, int TOK, long long* dbg) {
  using SM = DnSmem<MT, NK>;
  constexpr int kDnKS = SM::KS, kDnVS = SM::KS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  DnMisc* ms = reinterpret_cast<DnMisc*>(smem + SM::misc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kSoftWarps = 4 * MT;

  const int tiles_per_seq = ceil_div(dm.S, MT * TOK);
  const int tile = tiles_per_seq - 1 - blockIdx.x % tiles_per_seq;  // longest (latest, causal) query tiles first
  const int bg = blockIdx.x / tiles_per_seq;
  const int g = bg % dm.G, b = bg / dm.G;
  const int s_base = tile * MT * TOK;
  int s_last = s_base + MT * TOK - 1;
  if (s_last > dm.S - 1) s_last = dm.S - 1;
  // union of the rows' key ranges -> key tiles [kt_lo, kt_lo + n)
  int lo_first, hi_first, lo_last, hi_last;
  dn_row_range(dm, branch, dm.t0 + s_base, lo_first, hi_first);
  dn_row_range(dm, branch, dm.t0 + s_last, lo_last, hi_last);
  const int kt_lo = lo_first / NK;
  const int n = hi_last > lo_first ? ceil_div(hi_last, NK) - kt_lo : 0;

  // ---- setup ---------------------------------------------------------------------------------------------
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < MT * kDnTile / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + SM::q)[i] = z;
  }
  if (tid == 0) {
    mbar_init(&ms->q_full, 1);
    for (int i = 0; i < kDnKS; ++i) { mbar_init(&ms->k_full[i], 1); mbar_init(&ms->k_empty[i], 1); }
    for (int i = 0; i < kDnVS; ++i) { mbar_init(&ms->v_full[i], 1); mbar_init(&ms->v_empty[i], 1); }
    for (int m = 0; m < MT; ++m) {
      mbar_init(&ms->s_full[m], 1);
      mbar_init(&ms->s_empty[m], 4);
      mbar_init(&ms->p_full[m], 4);
      mbar_init(&ms->p_empty[m], 1);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 0) tmem_alloc(&ms->tmem_base, SM::tmem_cols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ms->tmem_base;
  // TMEM columns: S[mt] at mt*NK ; O[mt] (accumulated over all key tiles) at MT*NK + mt*64

  if (warp == kSoftWarps) {
    // ===== TMA producer ====================================================================================
    if (lane == 0 && n > 0) {
      mbar_expect_tx(&ms->q_full, MT * TOK * dm.h * 128);
      for (int m = 0; m < MT; ++m)
        tma_load_4d(smem + SM::q + m * kDnTile, &tmQ, &ms->q_full, 0, 0, g, b * dm.S + s_base + m * TOK);
      for (int i = 0; i < n; ++i) {
        const int ks = i % kDnKS, vs = i % kDnVS;
        mbar_wait(&ms->k_empty[ks], ((i / kDnKS) & 1) ^ 1);
        mbar_expect_tx(&ms->k_full[ks], SM::kvtile);
        tma_load_3d(smem + SM::k + ks * SM::kvtile, &tmK, &ms->k_full[ks], 0, (kt_lo + i) * NK, bg);
        mbar_wait(&ms->v_empty[vs], ((i / kDnVS) & 1) ^ 1);
        mbar_expect_tx(&ms->v_full[vs], SM::kvtile);
        tma_load_3d(smem + SM::v + vs * SM::kvtile, &tmV, &ms->v_full[vs], 0, (kt_lo + i) * NK, bg);
      }
    }
  } else if (warp == kSoftWarps + 1) {
    // ===== MMA issuer: the whole warp runs this code with warp-uniform operands (descriptors stay in uniform registers, a
    // k-step is one uniform add) and one elected lane issues; building each descriptor from scratch under `if (lane == 0)`
    // cost ~80 cycles per tcgen05.mma =======================================================================
    if (n > 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(128, NK, TcType<T>::fmt, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, 64, TcType<T>::fmt, 0, 1);
      // descriptor words: lo = addr>>4 | (LBO>>4)<<16, hi = SBO>>4 | version | swizzle
      constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | ((uint32_t)kSwizzle128B << 29);
      constexpr uint32_t kLoK = (16u >> 4) << 16, kLoMN = (8192u >> 4) << 16;
      const uint32_t smem0 = smem_u32(smem) >> 4;
      auto issue_qk = [&](int m, int i) {
        const uint32_t q_lo = (smem0 + ((SM::q + m * kDnTile) >> 4)) | kLoK;
        const uint32_t k_lo = (smem0 + ((SM::k + (i % kDnKS) * SM::kvtile) >> 4)) | kLoK;
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16_elect(tmem + m * NK, q_lo + k * 2, kHi, k_lo + k * 2, kHi, idesc_qk, k > 0);
        umma_commit_elect(&ms->s_full[m]);
      };
      mbar_wait(&ms->q_full, 0);
      mbar_wait(&ms->k_full[0], 0);
      tc_fence_after();
      for (int m = 0; m < MT; ++m) issue_qk(m, 0);
      umma_commit_elect(&ms->k_empty[0]);
      for (int i = 0; i < n; ++i) {
        const int vs = i % kDnVS;
        mbar_wait(&ms->v_full[vs], (i / kDnVS) & 1);
        if (i + 1 < n) mbar_wait(&ms->k_full[(i + 1) % kDnKS], ((i + 1) / kDnKS) & 1);
        const uint32_t v_lo = (smem0 + ((SM::v + vs * SM::kvtile) >> 4)) | kLoMN;
        for (int m = 0; m < MT; ++m) {
          // the softmax warps wait for the NEXT tile's S; nothing waits for this tile's P.V before their next P write: S first
          if (i + 1 < n) {
            mbar_wait(&ms->s_empty[m], i & 1);
            tc_fence_after();
            issue_qk(m, i + 1);
          }
          if (lane == 0) DDBG(10 + m, i);
          mbar_wait(&ms->p_full[m], i & 1);
          if (lane == 0) DDBG(12 + m, i);
          tc_fence_after();
          const uint32_t p_lo = (smem0 + ((SM::p + m * SM::ptile) >> 4)) | kLoK;
          const uint32_t od = tmem + MT * NK + m * 64;
          const uint32_t acc0 = i > 0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < NK / 16; ++k)  // P: 64-key halves of 16 KB, 32 B per k-step; V: 16 rows = 2048 B per k-step
            umma_f16_elect(od, p_lo + (k >> 2) * (kDnTile >> 4) + (k & 3) * 2, kHi, v_lo + k * (2048 >> 4), kHi, idesc_pv,
                           k > 0 ? 1u : acc0);
          umma_commit_elect(&ms->p_empty[m]);
        }
        umma_commit_elect(&ms->v_empty[vs]);
        if (i + 1 < n) umma_commit_elect(&ms->k_empty[(i + 1) % kDnKS]);
      }
    }
  } else {
    // ===== softmax warps ===================================================================================
    const int mt = warp >> 2;
    const int r = tid & 127;
    const int tok_l = r / dm.h, head = r - tok_l * dm.h;
    const int s = s_base + mt * TOK + tok_l;
    const bool row_ok = tok_l < TOK && s < dm.S;
    // rows that are not stored (padding rows of the M-tile, tokens beyond S) behave like rows that see every key, so they
    // never push their warp onto the masked path
    int lo = 0, hi = 0x3fffffff;
    if (row_ok) dn_row_range(dm, branch, dm.t0 + s, lo, hi);
    const float c = dm.scale * kLog2e;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tm_S = tmem + lane_off + mt * NK;
    const uint32_t tm_O = tmem + lane_off + MT * NK + mt * 64;
    uint8_t* prow = smem + SM::p + mt * SM::ptile + r * 128;
    const int sw = r & 7;

    float m_run = -INFINITY, l_run = 0.f;

    // Online softmax with a LAZY reference: m_run is fixed by the first tile that holds a valid key (two passes over that
    // tile: max, then probabilities) and then only moves when a later tile exceeds it by more than kJump (log2 units), in
    // which case that row redoes the tile against the new reference and rescales.  Every other tile is ONE pass over S:
    // p = exp2(s*c - m_run*c) (values above 1 are fine in fp32 / bf16), so S is read from TMEM once and nothing waits
    // for a tile maximum.  Exact: P, l and acc always share one reference per row.
    constexpr float kJump = 24.f;
    for (int i = 0; i < n; ++i) {
      const int col_base = (kt_lo + i) * NK;
      if (tid == 0) DDBG(1, i);
      mbar_wait(&ms->s_full[mt], i & 1);
      if (tid == 0) DDBG(2, i);
      tc_fence_after();
      // one path per warp: the masked path also handles full rows, so a warp takes it as a whole or not at all
      const bool full_tile = __all_sync(0xffffffffu, col_base >= lo && col_base + NK <= hi);
      // tcgen05.ld is warp-collective: every decision that guards one is made warp-uniform with a ballot
      if (__ballot_sync(0xffffffffu, !(m_run > -INFINITY)) != 0u) {  // a row without a reference: take this tile's maximum
        float cm = -INFINITY;
        uint32_t ua[16];
#pragma unroll 1
        for (int ch = 0; ch < NK / 16; ++ch) {
          tmem_ld16(tm_S + ch * 16, ua);
          dn_ld_wait16(ua);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int col = col_base + ch * 16 + e;
            if (col >= lo && col < hi) cm = fmaxf(cm, __uint_as_float(ua[e]));
          }
        }
        if (!(m_run > -INFINITY)) m_run = cm;
      }
      if (tid == 0) DDBG(3, i);
      mbar_wait(&ms->p_empty[mt], (i & 1) ^ 1);  // P.V of the previous tile has read the P buffer
      if (tid == 0) DDBG(4, i);
      float alpha = 1.f, rowsum = 0.f;
#pragma unroll 1
      for (int rep = 0; rep < 2; ++rep) {
        const bool any = m_run > -INFINITY;
        const float mc = any ? m_run * c : 0.f;
        float cm = -INFINITY;
        rowsum = 0.f;
        // 8 chunks of 16 columns, TMEM loads double-buffered; probabilities overwrite the loaded registers in place
        uint32_t ua[16], ub[16];
        tmem_ld16(tm_S, ua);
#pragma unroll
        for (int ch = 0; ch < NK / 16; ++ch) {
          uint32_t(&cur)[16] = (ch & 1) ? ub : ua;
          uint32_t(&nxt)[16] = (ch & 1) ? ua : ub;
          dn_ld_wait16(cur);
          if (ch < NK / 16 - 1) tmem_ld16(tm_S + (ch + 1) * 16, nxt);
          if (full_tile) {
#pragma unroll
            for (int e = 0; e < 16; e += 2) cm = fmaxf(cm, fmaxf(__uint_as_float(cur[e]), __uint_as_float(cur[e + 1])));
            // issue the 16 exponentials first, consume them afterwards (a MUFU result used by the next instruction
            // stalls the in-order issue for the MUFU latency)
#pragma unroll
            for (int e = 0; e < 16; ++e) cur[e] = __float_as_uint(dn_ex2(fmaf(__uint_as_float(cur[e]), c, -mc)));
          } else {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const int col = col_base + ch * 16 + e;
              const bool ok = any && col >= lo && col < hi;
              if (ok) cm = fmaxf(cm, __uint_as_float(cur[e]));
              cur[e] = ok ? __float_as_uint(dn_ex2(fmaf(__uint_as_float(cur[e]), c, -mc))) : 0u;
            }
          }
          float r0 = 0.f, r1 = 0.f;
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 16; e += 4) {
            r0 += __uint_as_float(cur[e]) + __uint_as_float(cur[e + 1]);
            r1 += __uint_as_float(cur[e + 2]) + __uint_as_float(cur[e + 3]);
          }
          rowsum += r0 + r1;
#pragma unroll
          for (int e = 0; e < 16; e += 2) pk[e >> 1] = pack2(T(), __uint_as_float(cur[e]), __uint_as_float(cur[e + 1]));
          // 16 keys = 2 chunks of 16 B; key chunk index kc = ch*2 + q in [0,16): half = kc >> 3, chunk-in-row = kc & 7
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int kc = ch * 2 + q;
            *reinterpret_cast<uint4*>(prow + (kc >> 3) * kDnTile + (((kc & 7) ^ sw) << 4)) =
                make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
          }
        }
        const bool jump = any && (cm - m_run) * c > kJump;
        if (rep == 1 || __ballot_sync(0xffffffffu, jump) == 0u) break;
        // rare: a row's tile maximum is far above its reference -> new reference for that row; the whole warp redoes the
        // tile (rows that did not move recompute the same values) and the past is rescaled
        if (jump) {
          alpha = dn_ex2((m_run - cm) * c);
          l_run *= alpha;
          m_run = cm;
        }
      }
      if (tid == 0) DDBG(5, i);
      // rare, warp-uniform: a row moved its reference -> scale the O accumulated so far (in TMEM; P.V of tile i-1 is complete,
      // P.V of this tile has not been released yet)
      if (__ballot_sync(0xffffffffu, alpha != 1.f) != 0u && i > 0) {
#pragma unroll 1
        for (int hf = 0; hf < 2; ++hf) {
          uint32_t oa[32];
          tmem_ld32(tm_O + hf * 32, oa);
          dn_ld_wait32(oa);
#pragma unroll
          for (int e = 0; e < 32; ++e) oa[e] = __float_as_uint(__uint_as_float(oa[e]) * alpha);
          tmem_st32(tm_O + hf * 32, oa);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ms->s_empty[mt]);
        mbar_arrive(&ms->p_full[mt]);
      }
      if (tid == 0) DDBG(6, i);
      l_run += rowsum;
    }

    // ---- epilogue: O (TMEM, all key tiles accumulated) / l -> global ------------------------------------------
    if (n > 0) {
      mbar_wait(&ms->p_empty[mt], (n - 1) & 1);  // the last P.V has completed
      tc_fence_after();
    }
    const size_t orow = (((size_t)b * dm.S + s) * dm.G + g) * dm.h + head;
    const float inv = l_run > 0.f ? 1.0f / l_run : 0.f;  // empty row -> zeros
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
      uint32_t oa[32];
      if (n > 0) {  // CTA-uniform
        tmem_ld32(tm_O + hf * 32, oa);
        dn_ld_wait32(oa);
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) oa[e] = 0u;
      }
      if (row_ok) {
        uint4* dst = reinterpret_cast<uint4*>(O + orow * 64 + hf * 32);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 v;
          v.x = pack2(T(), __uint_as_float(oa[q * 8 + 0]) * inv, __uint_as_float(oa[q * 8 + 1]) * inv);
          v.y = pack2(T(), __uint_as_float(oa[q * 8 + 2]) * inv, __uint_as_float(oa[q * 8 + 3]) * inv);
          v.z = pack2(T(), __uint_as_float(oa[q * 8 + 4]) * inv, __uint_as_float(oa[q * 8 + 5]) * inv);
          v.w = pack2(T(), __uint_as_float(oa[q * 8 + 6]) * inv, __uint_as_float(oa[q * 8 + 7]) * inv);
          dst[q] = v;
        }
      }
    }
    if (row_ok && lse) lse[orow] = l_run > 0.f ? m_run * dm.scale + logf(l_run) : -INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, SM::tmem_cols);
}

// ---- host ------------------------------------------------------------------------------------------------------
int make_tmap_q_heads(CUtensorMap* out, const void* base, int dtype, int D, int h, int G, long long n_tokens, int box_tokens);

bool tc_dense_supported(const nsa_dims_t& dm, int branch) {
  if (dm.impl == NSA_IMPL_SIMT) return false;
  if (!((dm.dtype == NSA_BF16 || dm.dtype == NSA_F16) && dm.Dk == 64 && dm.Dv == 64 && dm.h >= 1 && dm.h <= 64 && dm.S >= 1))
    return false;
  if (branch == 0) return dm.S_cmp >= 1;
  if (branch == 2) return dm.S_win_kv >= 1 && dm.w >= 1;
  return false;
}

template <typename T, int MT, int NK>
static int launch_dense_t(const nsa_dims_t& dm, int branch, const void* Q, const void* K, const void* V, void* O, float* lse,
                          cudaStream_t stream) {
  using SM = DnSmem<MT, NK>;
  const int TOK = 128 / dm.h;
  CUtensorMap tmQ, tmK, tmV;
  const int rows = branch == 0 ? dm.S_cmp : dm.S_win_kv;
  const long long cap = branch == 0 ? dm.cap_cmp : dm.cap_win;
  if (int rc = make_tmap_q_heads(&tmQ, Q, dm.dtype, 64, dm.h, dm.G, (long long)dm.B * dm.S, TOK)) return rc;
  if (int rc = make_tmap_rows(&tmK, K, dm.dtype, 64, rows, 64, cap * 64, dm.B * dm.G, NK)) return rc;
  if (int rc = make_tmap_rows(&tmV, V, dm.dtype, 64, rows, 64, cap * 64, dm.B * dm.G, NK)) return rc;
  auto kern = dense_attn_tc_kernel<T, MT, NK>;
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_smem_attr(kern, SM::total, attr_done, "dense tc", SM::ctas_per_sm > 1)) return rc;
  const int grid = dm.B * dm.G * ceil_div(dm.S, MT * TOK);
  static const bool dbg_on = getenv("NSA_B200_DENSE_DBG") != nullptr;
  static long long* dbg_buf = nullptr;
  if (dbg_on) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 8008 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 8008 * sizeof(long long), stream);
  }
  kern<<<grid, 32 * (4 * MT + 2), SM::total, stream>>>(tmQ, tmK, tmV, dm, branch, (T*)O, lse, TOK, dbg_on ? dbg_buf : nullptr);
  if (dbg_on) {  // debug only: timeline of one CTA (tag, tile, clock)
    static int dumps = 0;
    cudaStreamSynchronize(stream);
    if (dumps++ == 2) {
      static long long host[8008];
      cudaMemcpy(host, dbg_buf, sizeof(host), cudaMemcpyDeviceToHost);
      long long n = host[0] < 4000 ? host[0] : 4000;
      for (long long i = 0; i < n; ++i)
        fprintf(stderr, "DDBG %lld %lld %lld\n", host[1 + 2 * i] >> 32, host[1 + 2 * i] & 0xffffffff, host[2 + 2 * i]);
    }
  }
  return check_launch("dense_attn_tc_kernel");
}

template <typename T>
static int launch_dense_mt(const nsa_dims_t& dm, int branch, const void* Q, const void* K, const void* V, void* O, float* lse,
                           cudaStream_t stream) {
  static const int mt_env = getenv("NSA_B200_DENSE_MT") ? atoi(getenv("NSA_B200_DENSE_MT")) : 0;
  // 4 M-tiles per CTA once there are enough rows to fill the machine with 84-token CTAs
  const bool big = mt_env == 4 || (mt_env != 2 && (long long)dm.B * dm.G * dm.S >= 4LL * (128 / dm.h) * 148);
  // the sliding branch walks only ~10 key tiles per CTA: with one CTA per SM its fill and drain (Q load, first K tile, epilogue) are
  // not hidden by anything; the (2, 64) shape fits two CTAs per SM (half the TMEM, two ring stages) that overlap them
  static const int win_env = getenv("NSA_B200_WIN_MT") ? atoi(getenv("NSA_B200_WIN_MT")) : 2;
  if (big && branch == 2 && win_env == 2) return launch_dense_t<T, 2, 64>(dm, branch, Q, K, V, O, lse, stream);
  if (big) return launch_dense_t<T, 4, 64>(dm, branch, Q, K, V, O, lse, stream);
  return launch_dense_t<T, 2, 128>(dm, branch, Q, K, V, O, lse, stream);
}

int launch_dense_tc(const nsa_dims_t& dm, int branch, const void* Q, const void* K, const void* V, void* O, float* lse,
                    cudaStream_t stream) {
  static_assert(sizeof(DnMisc) <= 512, "DnMisc must fit its slot");
  if (dm.B * dm.S * dm.G == 0) return NSA_OK;
  if (dm.dtype == NSA_BF16) return launch_dense_mt<__nv_bfloat16>(dm, branch, Q, K, V, O, lse, stream);
  return launch_dense_mt<__half>(dm, branch, Q, K, V, O, lse, stream);
}

}  // namespace nsa
