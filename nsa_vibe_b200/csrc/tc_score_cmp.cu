// Pass 2 of the compressed-key scorer FUSED with the compressed branch (long prefill).
//
// The scorer's second pass forms p = exp2(s*c - (m*c + log2 l)) for every (row, compressed key) to fold it through Eq.9 / Eq.10
// (selection_scorer.py:42-116, nsa_attention.py:1091); the compressed branch (attention_kernels.py:106-143 with the mask it
// builds, packing.py:15-23) needs softmax over the row's first num_cmp(t) of the SAME logits.  The two differ by a per-row
// constant that cancels in O_cmp = (sum_c p_c V_c) / (sum_c p_c), so one exponential serves both: this kernel is the dense
// flash-attention structure of tc_dense.cu (4 M-tiles x 64-key tiles, S and O in TMEM, P as a 16-bit swizzled tile, O += P.V)
// whose probabilities are ALSO reduced to p_grp.  It removes dense_attn_tc_kernel[cmp] from the 64k step: 1.6 G exponentials and
// 206 GFLOP of recomputed Q.K_cmp^T per sequence.
//
//   pass 1 (score_tc_kernel, stats only): per (token, head) row  offs = m*c + log2 l  (full-row or causal normaliser)
//                                         and  mcc = c * max over the row's causal keys
//   this kernel: p = exp2(s*c - offs) -> Eq.9 stencil -> Eq.10 head sum -> p_grp (columns up to the CTA's selection limit);
//                P = p masked to col < num_cmp(t) -> O_cmp, lse_cmp.  A row whose causal logits all sit more than kRefGap below
//                the full-row reference (future keys dominate it: possible only with the full-row normaliser) takes mcc as the
//                reference of its branch softmax instead -- a second exponential for that warp's tile, never seen on sane data.
//
// Row layout: a row is one (token, head); the h rows of a token sit on consecutive TMEM lanes of ONE warp (TOKW = 32 / h tokens per
// warp, 4 * TOKW per M-tile; lanes >= TOKW * h of a warp are padding: 2 of 32 at h = 6), so the Eq.10 head sum never crosses a
// warp: it goes through a per-warp shared-memory buffer between two __syncwarp()s.  With the tokens packed densely (21 per M-tile
// at h = 6) three tokens straddled warps and every key tile cost two 128-thread named barriers per M-tile: 1.10 ms at 64k.
// Warp roles: warps [0, 16) softmax (thread = TMEM lane = row), warp 16 TMA producer, warp 17 MMA issuer.
#include <stdlib.h>

#include "tc_common.cuh"
#include "launchers.h"

namespace nsa {
using namespace tc;

constexpr int kFcMT = 4;              // M-tiles per CTA
constexpr int kFcNK = 64;             // keys per tile
constexpr int kFcKS = 3, kFcVS = 3;   // K / V ring stages
constexpr int kFcTile = 128 * 128;    // bytes: 128 rows x 64 x 2 B
constexpr int kFcKV = kFcNK * 128;    // bytes of one K or V tile
constexpr int kFcBPT = kFcNK / 4;     // selection blocks completed per key tile (R = l_sel / d = 4)
constexpr int kFcRedLd = kFcBPT + 4;  // row of the head-reduction buffer: 20 floats = 80 B keeps 16-byte accesses aligned and conflict-free
constexpr float kRefGap = 100.f;      // log2 units

struct FcSmem {
  static constexpr int q = 0;
  static constexpr int k = q + kFcMT * kFcTile;
  static constexpr int v = k + kFcKS * kFcKV;
  static constexpr int p = v + kFcVS * kFcKV;
  static constexpr int red = p + kFcMT * kFcTile;                       // [16 warps][32 rows][20] fp32
  static constexpr int misc = red + kFcMT * 128 * kFcRedLd * 4;
  static constexpr int total = misc + 512 + 1024;
};
static_assert(FcSmem::total <= 227 * 1024, "score+cmp kernel: shared memory");

struct FcMisc {
  uint64_t q_full;
  uint64_t k_full[kFcKS], k_empty[kFcKS], v_full[kFcVS], v_empty[kFcVS];
  uint64_t s_full[kFcMT], s_empty[kFcMT], p_full[kFcMT], p_empty[kFcMT];
  uint32_t tmem_base;
};

__device__ __forceinline__ float fc_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void fc_ld_wait16(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}
__device__ __forceinline__ void fc_ld_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// One 16-column chunk of a row: probabilities, Eq.9 block sums (-> rb4[0..3]), the branch's numerators as 16-bit P (-> prow), row sum.
template <typename T, bool PLAIN>
__device__ __forceinline__ void fc_chunk(uint32_t (&cur)[16], int col0, int nk, int hi, bool own_ref, float c, float offs, float ref,
                                         float& carry, float& rowsum, float* rb4, uint8_t* prow, int ch, int sw) {
  float pp[16];  // numerators of the branch's softmax where they differ from the scorer's probabilities (!PLAIN only)
  if (PLAIN) {
#pragma unroll
    for (int e = 0; e < 16; ++e) cur[e] = __float_as_uint(fc_ex2(fmaf(__uint_as_float(cur[e]), c, -offs)));
  } else {
#pragma unroll
    for (int e = 0; e < 16; ++e) {
      const int col = col0 + e;
      const float sv = __uint_as_float(cur[e]);
      const float pv = col < nk ? fc_ex2(fmaf(sv, c, -offs)) : 0.f;
      pp[e] = col < hi ? (own_ref ? fc_ex2(fmaf(sv, c, -ref)) : pv) : 0.f;
      cur[e] = __float_as_uint(pv);
    }
  }
  // Eq.9 (l = 2d, l_sel = 4d): block j = 1/2 p[4j-1] + p[4j] + p[4j+1] + p[4j+2] + 1/2 p[4j+3], ascending compressed index.
  // `carry` holds the UNSCALED p[4j-1]; the halves enter through FMAs: 0.5 * x is exact, so fma(0.5, x, y) rounds like
  // y + 0.5 * x -- the same four roundings per block as the stand-alone scorer.
  float blk[4];
  const float chunk_in = carry;
#pragma unroll
  for (int jj = 0; jj < 4; ++jj) {
    float a = fmaf(0.5f, carry, __uint_as_float(cur[jj * 4]));
    a += __uint_as_float(cur[jj * 4 + 1]);
    a += __uint_as_float(cur[jj * 4 + 2]);
    carry = __uint_as_float(cur[jj * 4 + 3]);
    blk[jj] = fmaf(0.5f, carry, a);
  }
  *reinterpret_cast<float4*>(rb4) = make_float4(blk[0], blk[1], blk[2], blk[3]);
  uint32_t pk[8];
  if (PLAIN) {  // the row sum of the chunk follows from its four block sums
    rowsum += ((blk[0] + blk[1]) + (blk[2] + blk[3])) + 0.5f * (carry - chunk_in);
#pragma unroll
    for (int e = 0; e < 16; e += 2) pk[e >> 1] = pack2(T(), __uint_as_float(cur[e]), __uint_as_float(cur[e + 1]));
  } else {
    float r0 = 0.f, r1 = 0.f;
#pragma unroll
    for (int e = 0; e < 16; e += 4) {
      r0 += pp[e] + pp[e + 1];
      r1 += pp[e + 2] + pp[e + 3];
    }
    rowsum += r0 + r1;
#pragma unroll
    for (int e = 0; e < 16; e += 2) pk[e >> 1] = pack2(T(), pp[e], pp[e + 1]);
  }
#pragma unroll
  for (int q = 0; q < 2; ++q) {  // 16 keys = 2 chunks of 16 B: chunk kc = ch*2 + q of the 128-B row
    const int kc = ch * 2 + q;
    *reinterpret_cast<uint4*>(prow + ((kc ^ sw) << 4)) = make_uint4(pk[q * 4], pk[q * 4 + 1], pk[q * 4 + 2], pk[q * 4 + 3]);
  }
}

template <typename T>
__global__ void __launch_bounds__(32 * (4 * kFcMT + 2), 1)
score_cmp_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                    const __grid_constant__ CUtensorMap tmV, nsa_dims_t dm, int S_sel, const float2* __restrict__ stats,
                    float* __restrict__ p_grp, T* __restrict__ O, float* __restrict__ lse, int TOK) {
  using SM = FcSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  FcMisc* ms = reinterpret_cast<FcMisc*>(smem + SM::misc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int MT = kFcMT, NK = kFcNK, kSoftWarps = 4 * kFcMT;

  const int tiles_per_seq = ceil_div(dm.S, MT * TOK);
  const int tile = tiles_per_seq - 1 - blockIdx.x % tiles_per_seq;  // latest (longest) query tiles first
  const int bg = blockIdx.x / tiles_per_seq;
  const int g = bg % dm.G, b = bg / dm.G;
  const int s_base = tile * MT * TOK;
  int s_last = s_base + MT * TOK - 1;
  if (s_last > dm.S - 1) s_last = dm.S - 1;
  const bool causal_norm = dm.norm_mode == NSA_NORM_CAUSAL;
  const int t_last = dm.t0 + s_last;
  const int hi_last = num_cmp_at(t_last, dm.l, dm.d, dm.S_cmp);           // compressed keys the CTA's last row attends
  const int nk_cta = causal_norm ? hi_last : dm.S_cmp;                    // keys with a non-zero probability
  int need = ((t_last + 1) / dm.l_sel) * 4;                               // keys whose blocks the last row may select (R = 4)
  if (need > nk_cta) need = nk_cta;
  const int n = ceil_div(need > hi_last ? need : hi_last, NK);            // key tiles of this CTA

  // ---- setup ---------------------------------------------------------------------------------------------
  {
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < MT * kFcTile / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + SM::q)[i] = z;
  }
  if (tid == 0) {
    mbar_init(&ms->q_full, 1);
    for (int i = 0; i < kFcKS; ++i) { mbar_init(&ms->k_full[i], 1); mbar_init(&ms->k_empty[i], 1); }
    for (int i = 0; i < kFcVS; ++i) { mbar_init(&ms->v_full[i], 1); mbar_init(&ms->v_empty[i], 1); }
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
  if (warp == 0) tmem_alloc(&ms->tmem_base, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ms->tmem_base;
  // TMEM columns: S[mt] at mt*NK ; O[mt] (accumulated over all key tiles) at MT*NK + mt*64

  if (warp == kSoftWarps) {
    // ===== TMA producer ====================================================================================
    if (lane == 0 && n > 0) {
      const int TOKW = TOK / 4;
      mbar_expect_tx(&ms->q_full, MT * TOK * dm.h * 128);
      for (int m = 0; m < MT; ++m)
        for (int q = 0; q < 4; ++q)  // one box per warp quarter: TOKW tokens x h heads land on that warp's first TOKW * h rows
          tma_load_4d(smem + SM::q + m * kFcTile + q * 32 * 128, &tmQ, &ms->q_full, 0, 0, g, b * dm.S + s_base + m * TOK + q * TOKW);
      for (int i = 0; i < n; ++i) {
        const int ks = i % kFcKS, vs = i % kFcVS;
        mbar_wait(&ms->k_empty[ks], ((i / kFcKS) & 1) ^ 1);
        mbar_expect_tx(&ms->k_full[ks], kFcKV);
        tma_load_3d(smem + SM::k + ks * kFcKV, &tmK, &ms->k_full[ks], 0, i * NK, bg);
        mbar_wait(&ms->v_empty[vs], ((i / kFcVS) & 1) ^ 1);
        mbar_expect_tx(&ms->v_full[vs], kFcKV);
        tma_load_3d(smem + SM::v + vs * kFcKV, &tmV, &ms->v_full[vs], 0, i * NK, bg);
      }
    }
  } else if (warp == kSoftWarps + 1) {
    // ===== MMA issuer (whole warp, warp-uniform operands, one elected lane issues) ==============================
    if (n > 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(128, NK, TcType<T>::fmt, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(128, 64, TcType<T>::fmt, 0, 1);
      constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | ((uint32_t)kSwizzle128B << 29);
      constexpr uint32_t kLoK = (16u >> 4) << 16, kLoMN = (8192u >> 4) << 16;
      const uint32_t smem0 = smem_u32(smem) >> 4;
      auto issue_qk = [&](int m, int i) {
        const uint32_t q_lo = (smem0 + ((SM::q + m * kFcTile) >> 4)) | kLoK;
        const uint32_t k_lo = (smem0 + ((SM::k + (i % kFcKS) * kFcKV) >> 4)) | kLoK;
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
        const int vs = i % kFcVS;
        mbar_wait(&ms->v_full[vs], (i / kFcVS) & 1);
        if (i + 1 < n) mbar_wait(&ms->k_full[(i + 1) % kFcKS], ((i + 1) / kFcKS) & 1);
        const uint32_t v_lo = (smem0 + ((SM::v + vs * kFcKV) >> 4)) | kLoMN;
        for (int m = 0; m < MT; ++m) {
          // the softmax warps wait for the NEXT tile's S; nothing waits for this tile's P.V before their next P write: S first
          if (i + 1 < n) {
            mbar_wait(&ms->s_empty[m], i & 1);
            tc_fence_after();
            issue_qk(m, i + 1);
          }
          mbar_wait(&ms->p_full[m], i & 1);
          tc_fence_after();
          const uint32_t p_lo = (smem0 + ((SM::p + m * kFcTile) >> 4)) | kLoK;
          const uint32_t od = tmem + MT * NK + m * 64;
          const uint32_t acc0 = i > 0 ? 1u : 0u;
#pragma unroll
          for (int k = 0; k < NK / 16; ++k)  // P: 32 B per k-step inside the 128-B rows; V: 16 rows = 2048 B per k-step
            umma_f16_elect(od, p_lo + k * 2, kHi, v_lo + k * (2048 >> 4), kHi, idesc_pv, k > 0 ? 1u : acc0);
          umma_commit_elect(&ms->p_empty[m]);
        }
        umma_commit_elect(&ms->v_empty[vs]);
        if (i + 1 < n) umma_commit_elect(&ms->k_empty[(i + 1) % kFcKS]);
      }
    }
  } else {
    // ===== softmax / Eq.9 / Eq.10 / P warps ================================================================
    const int mt = warp >> 2, w4 = warp & 3;
    const int r = tid & 127;
    const int TOKW = TOK / 4;
    const int tok_w = lane / dm.h, head = lane - tok_w * dm.h;
    const int s = s_base + mt * TOK + w4 * TOKW + tok_w;
    const bool row_ok = tok_w < TOKW && s < dm.S;
    const int t = dm.t0 + s;
    // rows that are not stored (padding rows of the M-tile, tokens beyond S) behave like the CTA's last row, so they never push
    // their warp onto a masked path
    const int hi = row_ok ? num_cmp_at(t, dm.l, dm.d, dm.S_cmp) : hi_last;      // keys of the compressed branch
    const int nk = row_ok ? (causal_norm ? hi : dm.S_cmp) : nk_cta;             // keys with a non-zero probability
    const size_t orow = (((size_t)b * dm.S + s) * dm.G + g) * dm.h + head;
    float offs = INFINITY, ref = INFINITY;
    if (row_ok) {
      const float2 st = stats[orow];
      offs = st.x;
      // branch reference: the scorer's own unless that would underflow every causal probability
      ref = (st.y > -INFINITY && st.x < INFINITY && st.x - st.y > kRefGap) ? st.y : st.x;
    }
    const bool own_ref = ref != offs;
    const float c = dm.scale * kLog2e;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t tm_S = tmem + lane_off + mt * NK;
    const uint32_t tm_O = tmem + lane_off + MT * NK + mt * 64;
    uint8_t* prow = smem + SM::p + mt * kFcTile + r * 128;
    const int sw = r & 7;
    float* red = reinterpret_cast<float*>(smem + SM::red) + (size_t)warp * 32 * kFcRedLd;  // this warp's 32 rows
    float* rb = red + (size_t)lane * kFcRedLd;

    float carry = 0.f;   // half of the last straddling compressed block, owed to the next selection block
    float rowsum = 0.f;  // sum of the branch's (masked) numerators, relative to `ref`
    const int n_out = n + 1;  // one more output tile flushes the carry
    for (int i = 0; i < n_out; ++i) {
      if (i < n) {
        const int col_base = i * NK;
        mbar_wait(&ms->s_full[mt], i & 1);
        tc_fence_after();
        // warp-uniform path choice: every key of the tile is attended by every row of the warp (then it also has a non-zero
        // probability: hi <= nk) and no row needs its own reference
        const bool plain = __all_sync(0xffffffffu, col_base + NK <= hi && !own_ref);
        mbar_wait(&ms->p_empty[mt], (i & 1) ^ 1);  // P.V of the previous tile has read the P buffer
        uint32_t ua[16], ub[16];
        tmem_ld16(tm_S, ua);
#pragma unroll
        for (int ch = 0; ch < NK / 16; ++ch) {
          uint32_t(&cur)[16] = (ch & 1) ? ub : ua;
          uint32_t(&nxt)[16] = (ch & 1) ? ua : ub;
          fc_ld_wait16(cur);
          if (ch < NK / 16 - 1) tmem_ld16(tm_S + (ch + 1) * 16, nxt);
          // two copies of the chunk body: in the plain one the branch's numerators ARE the scorer's probabilities (one register
          // set; a shared body made the compiler copy every exponential into a second set: 56 M moves per 64k sequence)
          if (plain) fc_chunk<T, true>(cur, col_base + ch * 16, nk, hi, own_ref, c, offs, ref, carry, rowsum, rb + ch * 4, prow, ch, sw);
          else fc_chunk<T, false>(cur, col_base + ch * 16, nk, hi, own_ref, c, offs, ref, carry, rowsum, rb + ch * 4, prow, ch, sw);
        }
        tc_fence_before();
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(&ms->s_empty[mt]);
          mbar_arrive(&ms->p_full[mt]);
        }
      } else {
#pragma unroll
        for (int jj = 0; jj < kFcBPT; ++jj) rb[jj] = jj == 0 ? 0.5f * carry : 0.f;
        carry = 0.f;
      }
      __syncwarp();
      // Eq.10: sum the h head rows of each of the warp's tokens, heads in ascending order like the stand-alone scorer; a lane
      // owns four consecutive blocks of one token (one 16-byte load per head, one 16-byte store)
      if (lane < TOKW * (kFcBPT / 4)) {
        const int tk = lane / (kFcBPT / 4), qd = lane % (kFcBPT / 4);
        const int ss = s_base + mt * TOK + w4 * TOKW + tk;
        const int j = i * kFcBPT + qd * 4;
        if (ss < dm.S && j < S_sel) {
          float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
          const float* src = red + (size_t)(tk * dm.h) * kFcRedLd + qd * 4;
          for (int hh = 0; hh < dm.h; ++hh) {
            const float4 v = *reinterpret_cast<const float4*>(src + (size_t)hh * kFcRedLd);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
          }
          float* dst = p_grp + (((size_t)b * dm.S + ss) * dm.G + g) * S_sel + j;
          if (j + 4 <= S_sel && (S_sel & 3) == 0) {
            *reinterpret_cast<float4*>(dst) = a;
          } else {
            dst[0] = a.x;
            if (j + 1 < S_sel) dst[1] = a.y;
            if (j + 2 < S_sel) dst[2] = a.z;
            if (j + 3 < S_sel) dst[3] = a.w;
          }
        }
      }
      __syncwarp();  // single buffer: the next tile's partial sums overwrite it
    }

    // ---- epilogue: O (TMEM, all key tiles accumulated) / rowsum -> global ------------------------------------
    if (n > 0) {
      mbar_wait(&ms->p_empty[mt], (n - 1) & 1);  // the last P.V has completed
      tc_fence_after();
    }
    const float inv = rowsum > 0.f ? 1.0f / rowsum : 0.f;  // no causal key -> zeros
#pragma unroll 1
    for (int hf = 0; hf < 2; ++hf) {
      uint32_t oa[32];
      if (n > 0) {  // CTA-uniform
        tmem_ld32(tm_O + hf * 32, oa);
        fc_ld_wait32(oa);
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
    // natural-log softmax normaliser of the branch: ln sum_c exp(s_c * scale) = (ref + log2 rowsum) * ln 2
    if (row_ok && lse) lse[orow] = rowsum > 0.f ? (ref + log2f(rowsum)) * 0.6931471805599453f : -INFINITY;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---- host ------------------------------------------------------------------------------------------------------
int make_tmap_q_heads(CUtensorMap* out, const void* base, int dtype, int D, int h, int G, long long n_tokens, int box_tokens);

// Served shapes: what the tcgen05 scorer and the dense compressed branch both serve, at sizes where the 4-M-tile kernels are used.
bool tc_score_cmp_supported(const nsa_dims_t& dm) {
  static const bool off = getenv("NSA_B200_FUSE_CMP") && atoi(getenv("NSA_B200_FUSE_CMP")) == 0;  // A/B switch (benchmarks / tests)
  if (off || dm.impl == NSA_IMPL_SIMT) return false;
  if (!((dm.dtype == NSA_BF16 || dm.dtype == NSA_F16) && dm.Dk == 64 && dm.Dv == 64 && dm.l == 2 * dm.d && dm.l_sel == 4 * dm.d)) return false;
  if (dm.h < 1 || dm.h > 32 || dm.S_cmp < 1 || dm.S < 1) return false;  // a token's heads must fit one warp
  return (long long)dm.B * dm.G * dm.S >= 4LL * (128 / dm.h) * 148;
}

int64_t tc_score_cmp_stats_bytes(const nsa_dims_t& dm) {
  return (((int64_t)dm.B * dm.S * dm.G * dm.h * (int64_t)sizeof(float2)) + 255) & ~(int64_t)255;
}

template <typename T>
static int launch_score_cmp_t(const nsa_dims_t& dm, const void* Q, const void* Kc, const void* Vc, int S_sel, const float* stats,
                              float* p_grp, void* O_cmp, float* lse_cmp, cudaStream_t stream) {
  const int TOK = 4 * (32 / dm.h);  // TOKW = 32 / h tokens per warp quarter
  CUtensorMap tmQ, tmK, tmV;
  if (int rc = make_tmap_q_heads(&tmQ, Q, dm.dtype, 64, dm.h, dm.G, (long long)dm.B * dm.S, TOK / 4)) return rc;
  if (int rc = make_tmap_rows(&tmK, Kc, dm.dtype, 64, dm.S_cmp, 64, (long long)dm.cap_cmp * 64, dm.B * dm.G, kFcNK)) return rc;
  if (int rc = make_tmap_rows(&tmV, Vc, dm.dtype, 64, dm.S_cmp, 64, (long long)dm.cap_cmp * 64, dm.B * dm.G, kFcNK)) return rc;
  auto kern = score_cmp_tc_kernel<T>;
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_smem_attr(kern, FcSmem::total, attr_done, "score+cmp tc")) return rc;
  const int grid = dm.B * dm.G * ceil_div(dm.S, kFcMT * TOK);
  kern<<<grid, 32 * (4 * kFcMT + 2), FcSmem::total, stream>>>(tmQ, tmK, tmV, dm, S_sel, reinterpret_cast<const float2*>(stats), p_grp,
                                                            (T*)O_cmp, lse_cmp, TOK);
  return check_launch("score_cmp_tc_kernel");
}

// stats: tc_score_cmp_stats_bytes(dm) bytes written by launch_score_stats_tc; p_grp [B,S,G,S_sel] fp32 staging (columns up to each
// CTA's selection limit are written, as the stand-alone scorer does for nsa_score_select); O_cmp [B,S,G,h,64]; lse_cmp may be NULL.
int launch_score_cmp_tc(const nsa_dims_t& dm, const void* Q, const void* Kc, const void* Vc, int S_sel, const float* stats,
                        float* p_grp, void* O_cmp, float* lse_cmp, cudaStream_t stream) {
  static_assert(sizeof(FcMisc) <= 512, "FcMisc must fit its slot");
  if (dm.B * dm.S * dm.G == 0) return NSA_OK;
  if (dm.dtype == NSA_BF16) return launch_score_cmp_t<__nv_bfloat16>(dm, Q, Kc, Vc, S_sel, stats, p_grp, O_cmp, lse_cmp, stream);
  return launch_score_cmp_t<__half>(dm, Q, Kc, Vc, S_sel, stats, p_grp, O_cmp, lse_cmp, stream);
}

}  // namespace nsa
