// Compressed-key scoring on tcgen05 tensor cores (NSA Eq.8-10):
//   p_cmp = softmax_i(Q . K_cmp^T / sqrt(Dk))   (compute_pcmp_all, nsa/core/selection_scorer.py:42-61)
//   p_slc = Eq.9 overlap stencil                (map_pcmp_to_pslc_batched, selection_scorer.py:89-116)
//   p_grp = sum over the h heads of the group   (nsa_attention.py:1091)
//
// One CTA owns MT M-tiles of 128 rows; a row is one (token, head) of one (b, g), so TOK = 128 / h tokens fill an
// M-tile and all heads of a token sit on adjacent TMEM lanes.  Key tiles of 128 compressed keys stream through a TMA
// ring and are shared by the MT M-tiles.  The softmax is exact two-pass: pass 1 reduces the row max and normaliser
// online, pass 2 recomputes S = Q.K^T (tensor work is cheap here: K = Dk = 64; the kernel is bound by exp throughput),
// forms p = exp2(s*c - (m*c + log2 l)) and folds it through the Eq.9 stencil in ascending compressed index -- the
// order of the reference's CPU scatter_add.  Nothing but p_grp [B,S,G,S_sel] leaves the SM.
//
// Warp roles: warps [0, 4*MT) softmax (thread = TMEM lane = row), warp 4*MT TMA producer, warp 4*MT+1 MMA issuer.
#include <stdlib.h>

#include "tc_common.cuh"
#include "launchers.h"
#include "select.cuh"

namespace nsa {
using namespace tc;

constexpr int kScStages = 4;       // K ring depth (16 KB tiles)
constexpr int kScTile = 128 * 128; // bytes of one 128-row x 64-element tile
constexpr int kScRedLd = 33;       // padded row of the head-reduction buffer
constexpr int kScPolyDefault = 6;  // NSA_B200_POLY_EX2 default: every 6th exponential of pass 1 on the FMA pipe (0 = none; 4, 5, 6, 8 built)

// MT M-tiles per CTA share every K tile.  TMEM (512 columns) holds STG = 4 / MT accumulator stages of 128 columns per
// M-tile: MT = 2 double-buffers S; MT = 4 single-buffers it and relies on 4 softmax warps per scheduler to hide the MMA.
template <int MT>
struct ScSmem {
  static constexpr int STG = MT <= 2 ? 2 : 1;                       // S accumulator stages per M-tile
  static constexpr int RB = MT <= 2 ? 2 : 1;                        // head-reduction buffers per M-tile
  static constexpr int q = 0;
  static constexpr int ring = q + MT * kScTile;
  static constexpr int red = ring + kScStages * kScTile;           // [MT][RB][128][33] fp32
  static constexpr int misc = red + MT * RB * 128 * kScRedLd * 4;
  static constexpr int total = misc + 256 + 1024;                   // + alignment slack
};

template <int MT>
struct ScMisc {
  uint64_t k_full[kScStages], k_empty[kScStages];
  uint64_t s_full[MT][2], s_empty[MT][2];
  uint64_t q_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// exp2 on the FMA pipe (no MUFU): x <= 0 clamped to >= -125, split as n + f with n = rint(x), |f| <= 1/2 (magic-number rounding),
// 2^f by its degree-6 Taylor polynomial in f*ln2 (remainder < 1.3e-7 relative, the accuracy class of ex2.approx), 2^n by an integer
// add into the exponent field.  11 issue slots against 1 (+ 8 cycles of the scheduler's 4-lane MUFU unit): pass 1 of the scorer
// is bound by MUFU throughput with half of its issue slots idle, so every PE-th column goes this way (NSA_B200_POLY_EX2).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.f);
  const float t = x + 12582912.f;            // 1.5 * 2^23: the low mantissa bits of t hold rint(x)
  const float f = x - (t - 12582912.f);
  float p = fmaf(f, 1.5403530e-4f, 1.3333558e-3f);
  p = fmaf(p, f, 9.6181291e-3f);
  p = fmaf(p, f, 5.5504109e-2f);
  p = fmaf(p, f, 2.4022651e-1f);
  p = fmaf(p, f, 6.9314718e-1f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// tcgen05.wait::ld that also names the destination registers, so no use of them can be scheduled above the wait
__device__ __forceinline__ void tmem_ld_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// R = l_sel / d compressed blocks start inside one selection block; l = 2 d, so the last of them straddles into the
// next selection block with weight 1/2 each (block_index.py:43-71).
template <typename T, int MT, int R, int PE>
__global__ void __launch_bounds__(32 * (4 * MT + 2), 1)
score_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK, nsa_dims_t dm, int S_sel,
                float* __restrict__ p_grp, int TOK, int sel_only, float2* __restrict__ stats_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  using SM = ScSmem<MT>;
  ScMisc<MT>* ms = reinterpret_cast<ScMisc<MT>*>(smem + SM::misc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int kSoftWarps = 4 * MT;
  constexpr int BPT = 128 / R;  // selection blocks completed per key tile
  constexpr int STG = SM::STG, RB = SM::RB;

  const int tiles_per_seq = ceil_div(dm.S, MT * TOK);
  const int tile = tiles_per_seq - 1 - blockIdx.x % tiles_per_seq;  // latest (longest, when causal) query tiles first
  const int bg = blockIdx.x / tiles_per_seq;
  const int g = bg % dm.G, b = bg / dm.G;
  const int s_base = tile * MT * TOK;
  const bool causal = dm.norm_mode == NSA_NORM_CAUSAL;
  int s_last = s_base + MT * TOK - 1;
  if (s_last > dm.S - 1) s_last = dm.S - 1;
  const int nk_cta = causal ? num_cmp_at(dm.t0 + s_last, dm.l, dm.d, dm.S_cmp) : dm.S_cmp;
  const int NT = ceil_div(nk_cta, 128);        // key tiles with work
  // Pass 2 (probabilities) covers every key tile when the caller wants p_grp itself.  When p_grp only feeds the selection
  // (sel_only), a row t never looks at blocks j >= (t+1)/l_sel (selection_scorer.py:303-309 masks them), i.e. at compressed
  // keys i >= R*(t+1)/l_sel -- so the CTA stops pass 2 at its last row's limit: half of the pass-2 exponentials on average.
  // Pass 1 keeps every key: the full-row normaliser is the reference's.  Columns past the limit stay unwritten.
  int NT2 = NT;
  if (stats_out) {
    NT2 = 0;  // pass 1 only: the row statistics go to stats_out, pass 2 runs fused with the compressed branch (tc_score_cmp.cu)
  } else if (sel_only) {
    int need = ((dm.t0 + s_last + 1) / dm.l_sel) * R;
    if (need > nk_cta) need = nk_cta;
    NT2 = ceil_div(need, 128);
  }
  int NTO = ceil_div(S_sel, BPT);              // output tiles (tiles >= NT2 only flush the carry / write zeros)
  if (sel_only && NT2 + 1 < NTO) NTO = NT2 + 1;
  if (stats_out) NTO = 0;

  // ---- setup ---------------------------------------------------------------------------------------------
  {  // rows of the Q tiles that TMA does not write (>= TOK*h) must hold finite data
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < MT * kScTile / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + SM::q)[i] = z;
  }
  if (tid == 0) {
    for (int i = 0; i < kScStages; ++i) { mbar_init(&ms->k_full[i], 1); mbar_init(&ms->k_empty[i], 1); }
    for (int m = 0; m < MT; ++m)
      for (int i = 0; i < 2; ++i) { mbar_init(&ms->s_full[m][i], 1); mbar_init(&ms->s_empty[m][i], 4); }
    mbar_init(&ms->q_full, 1);
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
  }
  if (warp == 0) tmem_alloc(&ms->tmem_base, MT * STG * 128);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ms->tmem_base;

  if (warp == kSoftWarps) {
    // ===== TMA producer ====================================================================================
    if (lane == 0) {
      mbar_expect_tx(&ms->q_full, MT * TOK * dm.h * 128);
      for (int m = 0; m < MT; ++m)
        tma_load_4d(smem + SM::q + m * kScTile, &tmQ, &ms->q_full, 0, 0, g, b * dm.S + s_base + m * TOK);
      for (int it = 0; it < NT + NT2; ++it) {
        const int ks = it % kScStages;
        mbar_wait(&ms->k_empty[ks], ((it / kScStages) & 1) ^ 1);
        mbar_expect_tx(&ms->k_full[ks], kScTile);
        tma_load_3d(smem + SM::ring + ks * kScTile, &tmK, &ms->k_full[ks], 0, (it % NT) * 128, bg);
      }
    }
  } else if (warp == kSoftWarps + 1) {
    // ===== MMA issuer: whole warp, warp-uniform operands (descriptors in uniform registers), one elected lane issues ======
    {
      constexpr uint32_t idesc = make_idesc_f16(128, 128, TcType<T>::fmt, 0, 0);
      constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | ((uint32_t)kSwizzle128B << 29);
      constexpr uint32_t kLoK = (16u >> 4) << 16;
      const uint32_t smem0 = smem_u32(smem) >> 4;
      mbar_wait(&ms->q_full, 0);
      for (int it = 0; it < NT + NT2; ++it) {
        const int ks = it % kScStages, st = it % STG;
        mbar_wait(&ms->k_full[ks], (it / kScStages) & 1);
        const uint32_t k_lo = (smem0 + ((SM::ring + ks * kScTile) >> 4)) | kLoK;
        for (int m = 0; m < MT; ++m) {
          mbar_wait(&ms->s_empty[m][st], ((it / STG) & 1) ^ 1);
          tc_fence_after();
          const uint32_t q_lo = (smem0 + ((SM::q + m * kScTile) >> 4)) | kLoK;
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16_elect(tmem + (m * STG + st) * 128, q_lo + k * 2, kHi, k_lo + k * 2, kHi, idesc, k > 0);
          umma_commit_elect(&ms->s_full[m][st]);
        }
        umma_commit_elect(&ms->k_empty[ks]);
      }
    }
  } else {
    // ===== softmax / Eq.9 / Eq.10 warps ====================================================================
    const int mt = warp >> 2;
    const int r = tid & 127;                              // row of the M-tile == TMEM lane
    const int tok_l = r / dm.h;
    const int s = s_base + mt * TOK + tok_l;
    const bool row_ok = tok_l < TOK && s < dm.S;
    const int t = dm.t0 + s;
    // rows that are not stored (padding rows of the M-tile, tokens beyond S) take the key count of the CTA's last row, so
    // they never push their warp onto the masked path
    const int nk = !row_ok ? nk_cta : (causal ? num_cmp_at(t, dm.l, dm.d, dm.S_cmp) : dm.S_cmp);
    const float c = dm.scale * kLog2e;
    const uint32_t tm_row = tmem + ((uint32_t)((warp & 3) * 32) << 16) + mt * STG * 128;
    float* red = reinterpret_cast<float*>(smem + SM::red) + (size_t)mt * RB * 128 * kScRedLd;

    // ---- pass 1: row max and normaliser --------------------------------------------------------------------
    float m_run = -INFINITY, l_run = 0.f;
    // with stats_out also the max over the row's CAUSAL keys (col < num_cmp(t)): the fused pass 2 + compressed branch falls back
    // to it as the reference of the branch's softmax when the full-row one would underflow every causal probability
    float mc_run = -INFINITY;
    const int nkc = !row_ok ? 0 : num_cmp_at(t, dm.l, dm.d, dm.S_cmp);
    for (int kt = 0; kt < NT; ++kt) {
      const int it = kt, st = it % STG;
      mbar_wait(&ms->s_full[mt][st], (it / STG) & 1);
      tc_fence_after();
      uint32_t va[32], vb[32];
      tmem_ld32(tm_row + st * 128, va);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t(&cur)[32] = (ch & 1) ? vb : va;
        uint32_t(&nxt)[32] = (ch & 1) ? va : vb;
        tmem_ld_wait32(cur);
        if (ch < 3) tmem_ld32(tm_row + st * 128 + (ch + 1) * 32, nxt);
        const int col0 = kt * 128 + ch * 32;
        const bool wfull = __all_sync(0xffffffffu, col0 + 32 <= nk);  // one path per warp (no divergent double execution)
        // causal max (stats_out only), votes taken before the per-lane branch below: every causal key of the warp's rows inside
        // this chunk -> the chunk maximum serves; none -> skip; otherwise a masked scan
        const bool c_all = stats_out != nullptr && __all_sync(0xffffffffu, col0 + 32 <= nkc || !row_ok);
        const bool c_any = stats_out != nullptr && __any_sync(0xffffffffu, col0 < nkc);
        if (col0 < nk) {
          float cm = -INFINITY;
          if (wfull) {
#pragma unroll
            for (int i = 0; i < 32; ++i) cm = fmaxf(cm, __uint_as_float(cur[i]));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              if (col0 + i >= nk) cur[i] = 0xff800000u;  // -inf
              cm = fmaxf(cm, __uint_as_float(cur[i]));
            }
          }
          if (c_all) {
            if (row_ok) mc_run = fmaxf(mc_run, cm);
          } else if (c_any) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i < nkc) mc_run = fmaxf(mc_run, __uint_as_float(cur[i]));
          }
          const float m_new = fmaxf(m_run, cm);
          const float mc = m_new * c;
          // all 32 exponentials are issued before the first is consumed: with 2 warps per scheduler a MUFU result used by
          // the next instruction stalls the in-order issue for the MUFU latency
          float e[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float xx = fmaf(__uint_as_float(cur[i]), c, -mc);
            e[i] = (PE > 0 && (i % (PE > 0 ? PE : 1)) == (PE > 0 ? PE : 1) - 1) ? ex2_poly(xx) : ex2f(xx);
          }
          const float corr = ex2f((m_run - m_new) * c);
          float s0 = e[0], s1 = e[1], s2 = e[2], s3 = e[3];
#pragma unroll
          for (int i = 4; i < 32; i += 4) { s0 += e[i]; s1 += e[i + 1]; s2 += e[i + 2]; s3 += e[i + 3]; }
          const float sum = (s0 + s1) + (s2 + s3);
          l_run = fmaf(l_run, corr, sum);
          m_run = m_new;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->s_empty[mt][st]);
    }
    // p = exp2(s*c - offs); rows without keys produce zeros
    const bool has = l_run > 0.f;
    const float offs = has ? fmaf(m_run, c, log2f(l_run)) : 0.f;
    if (stats_out) {  // (offs, causal max in the log2 domain); rows without keys: +inf makes every probability exp2(-inf) = 0
      if (row_ok) {
        const int head = r - tok_l * dm.h;
        stats_out[(((size_t)b * dm.S + s) * dm.G + g) * dm.h + head] = make_float2(has ? offs : INFINITY, mc_run > -INFINITY ? mc_run * c : -INFINITY);
      }
    }

    // ---- pass 2: probabilities -> Eq.9 -> Eq.10 ------------------------------------------------------------
    float carry = 0.f;  // half of the last straddling compressed block, owed to the next selection block
    for (int kt = 0; kt < NTO; ++kt) {
      float* rb = red + (size_t)(kt % RB) * 128 * kScRedLd + (size_t)r * kScRedLd;
      if (kt < NT2) {
        const int it = NT + kt, st = it % STG;
        mbar_wait(&ms->s_full[mt][st], (it / STG) & 1);
        tc_fence_after();
        uint32_t va[32], vb[32];
        tmem_ld32(tm_row + st * 128, va);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t(&cur)[32] = (ch & 1) ? vb : va;
          uint32_t(&nxt)[32] = (ch & 1) ? va : vb;
          tmem_ld_wait32(cur);
          if (ch < 3) tmem_ld32(tm_row + st * 128 + (ch + 1) * 32, nxt);
          const int col0 = kt * 128 + ch * 32;
          float p[32];
          if (__all_sync(0xffffffffu, has && col0 + 32 <= nk)) {
#pragma unroll
            for (int i = 0; i < 32; ++i) p[i] = ex2f(fmaf(__uint_as_float(cur[i]), c, -offs));
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              p[i] = (has && col0 + i < nk) ? ex2f(fmaf(__uint_as_float(cur[i]), c, -offs)) : 0.f;
          }
#pragma unroll
          for (int jj = 0; jj < 32 / R; ++jj) {
            float a = carry;
#pragma unroll
            for (int q = 0; q < R - 1; ++q) a += p[jj * R + q];
            const float half = 0.5f * p[jj * R + R - 1];
            a += half;
            carry = half;
            rb[ch * (32 / R) + jj] = a;
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ms->s_empty[mt][st]);
      } else {
#pragma unroll 4
        for (int jj = 0; jj < BPT; ++jj) rb[jj] = jj == 0 ? carry : 0.f;
        carry = 0.f;
      }
      named_bar_sync(1 + mt, 128);
      // Eq.10: sum the h head rows of each token; 32 lanes write 32 consecutive blocks of one token
      const float* rd = red + (size_t)(kt % RB) * 128 * kScRedLd;
      for (int idx = r; idx < TOK * BPT; idx += 128) {
        const int tk = idx / BPT, cc = idx % BPT;
        const int ss = s_base + mt * TOK + tk;
        const int j = kt * BPT + cc;
        if (ss < dm.S && j < S_sel) {
          float a = 0.f;
          for (int hh = 0; hh < dm.h; ++hh) a += rd[(size_t)(tk * dm.h + hh) * kScRedLd + cc];
          p_grp[(((size_t)b * dm.S + ss) * dm.G + g) * S_sel + j] = a;
        }
      }
      if (RB == 1) named_bar_sync(1 + mt, 128);  // single buffer: the next tile's partial sums overwrite it
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, MT * STG * 128);
}

// ---- host ------------------------------------------------------------------------------------------------------
int make_tmap_q_heads(CUtensorMap* out, const void* base, int dtype, int D, int h, int G, long long n_tokens, int box_tokens);

bool tc_score_supported(const nsa_dims_t& dm) {
  if (dm.impl == NSA_IMPL_SIMT) return false;
  return (dm.dtype == NSA_BF16 || dm.dtype == NSA_F16) && dm.Dk == 64 && dm.l == 2 * dm.d && dm.l_sel == 4 * dm.d &&
         dm.h >= 1 && dm.h <= 64 && dm.S_cmp >= 1 && dm.S >= 1;
}

int64_t tc_score_workspace(const nsa_dims_t& dm) {
  // p_grp staging for the fused score+select entry point: [B,S,G,S_sel] fp32 with S_sel <= ceil((t0+S)/l_sel) + 1
  const int64_t S_sel = ceil_div(dm.t0 + dm.S > dm.l_sel ? dm.t0 + dm.S : dm.l_sel, dm.l_sel) + 1;
  return (int64_t)dm.B * dm.S * dm.G * S_sel * 4;
}

template <typename T, int MT>
static int launch_score_t(const nsa_dims_t& dm, const void* Q, const void* Kc, int S_sel, float* p_grp, bool sel_only,
                          cudaStream_t stream, float2* stats_out = nullptr) {
  const int TOK = 128 / dm.h;
  CUtensorMap tmQ, tmK;
  if (int rc = make_tmap_q_heads(&tmQ, Q, dm.dtype, 64, dm.h, dm.G, (long long)dm.B * dm.S, TOK)) return rc;
  if (int rc = make_tmap_rows(&tmK, Kc, dm.dtype, 64, dm.S_cmp, 64, (long long)dm.cap_cmp * 64, dm.B * dm.G, 128)) return rc;
  // every PE-th exponential of pass 1 on the FMA pipe (4-M-tile kernel only; 0 = all on MUFU).  One process-wide setting, so the
  // stand-alone scorer and the split scorer compute identical row statistics.
  static const int pe_env = getenv("NSA_B200_POLY_EX2") ? atoi(getenv("NSA_B200_POLY_EX2")) : kScPolyDefault;
  const int pe = MT == 4 ? pe_env : 0;
  const int grid = dm.B * dm.G * ceil_div(dm.S, MT * TOK);
  const int threads = 32 * (4 * MT + 2);
#define NSA_SCORE_LAUNCH(PEV)                                                                                           \
  do {                                                                                                                  \
    auto kern = score_tc_kernel<T, MT, 4, PEV>;                                                                         \
    static std::atomic<unsigned long long> attr_done{0};                                                               \
    if (int rc = ensure_smem_attr(kern, ScSmem<MT>::total, attr_done, "score tc")) return rc;                         \
    kern<<<grid, threads, ScSmem<MT>::total, stream>>>(tmQ, tmK, dm, S_sel, p_grp, TOK, sel_only ? 1 : 0, stats_out);  \
  } while (0)
  if (MT == 4 && pe == 8) NSA_SCORE_LAUNCH(8);
  else if (MT == 4 && pe == 6) NSA_SCORE_LAUNCH(6);
  else if (MT == 4 && pe == 5) NSA_SCORE_LAUNCH(5);
  else if (MT == 4 && pe == 4) NSA_SCORE_LAUNCH(4);
  else NSA_SCORE_LAUNCH(0);
#undef NSA_SCORE_LAUNCH
  return check_launch("score_tc_kernel");
}

// Pass 1 alone (4 M-tiles per CTA): per (token, head) row the pair (m*c + log2 l, causal max * c) that tc_score_cmp.cu consumes.
int launch_score_stats_tc(const nsa_dims_t& dm, const void* Q, const void* Kc, float* stats, cudaStream_t stream) {
  if (dm.B * dm.S * dm.G == 0) return NSA_OK;
  if (dm.dtype == NSA_BF16) return launch_score_t<__nv_bfloat16, 4>(dm, Q, Kc, 1, nullptr, false, stream, reinterpret_cast<float2*>(stats));
  return launch_score_t<__half, 4>(dm, Q, Kc, 1, nullptr, false, stream, reinterpret_cast<float2*>(stats));
}

int launch_score_tc(const nsa_dims_t& dm, const void* Q, const void* Kc, int S_sel, int S_total, int sel_mode, int Kr,
                    float* p_grp, int32_t* ranges, void* workspace, cudaStream_t stream) {
  static_assert(sizeof(ScMisc<4>) <= 256, "ScMisc must fit its slot");
  if (dm.B * dm.S * dm.G == 0) return NSA_OK;
  float* pg = p_grp;
  if (!pg) {
    NSA_REQUIRE(workspace, "score_select(tc): needs a workspace of nsa_workspace_bytes(NSA_WS_SCORE_SELECT) bytes");
    NSA_REQUIRE((int64_t)dm.B * dm.S * dm.G * S_sel * 4 <= tc_score_workspace(dm), "score_select(tc): S_sel=%d exceeds the workspace", S_sel);
    pg = reinterpret_cast<float*>(workspace);
  }
  // 4 M-tiles per CTA (16 softmax warps) when there are enough rows to fill the machine, 2 otherwise;
  // NSA_B200_SCORE_MT=2|4 forces one (benchmarks / tests)
  static const int mt_env = getenv("NSA_B200_SCORE_MT") ? atoi(getenv("NSA_B200_SCORE_MT")) : 0;
  const bool big = mt_env == 4 || (mt_env != 2 && (long long)dm.B * dm.G * dm.S >= 4LL * (128 / dm.h) * 148);
  // the staged p_grp only feeds the selection: pass 2 may stop at each CTA's causal limit (NSA_B200_SCORE_CAUSAL_P2=0 keeps
  // every column, for A/B runs)
  static const bool p2_env = !(getenv("NSA_B200_SCORE_CAUSAL_P2") && atoi(getenv("NSA_B200_SCORE_CAUSAL_P2")) == 0);
  const bool so = !p_grp && ranges && p2_env;
  int rc;
  if (dm.dtype == NSA_BF16)
    rc = big ? launch_score_t<__nv_bfloat16, 4>(dm, Q, Kc, S_sel, pg, so, stream) : launch_score_t<__nv_bfloat16, 2>(dm, Q, Kc, S_sel, pg, so, stream);
  else
    rc = big ? launch_score_t<__half, 4>(dm, Q, Kc, S_sel, pg, so, stream) : launch_score_t<__half, 2>(dm, Q, Kc, S_sel, pg, so, stream);
  if (rc) return rc;
  if (!ranges) return NSA_OK;
  const int nf = forced_code_default(sel_mode, S_total, dm.l_sel);
  return launch_select(pg, dm.B * dm.S * dm.G, dm.S, dm.G, S_sel, dm.l_sel, dm.n_sel, sel_mode, nf, Kr, dm.t0, ranges, stream);
}

}  // namespace nsa
