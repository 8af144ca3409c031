// Block-sparse gather attention on tcgen05 tensor cores: the selected branch of prefill and all three branches of a
// decode step.  A work item is one (b, s, g, branch): <= 1024 keys described as up to 16 blocks of 64 cache rows
//   sel: the selected ranges cut into 64-key blocks        (grouped_selection_attention_masked, attention_kernels.py:705-772)
//   win: cache rows of tokens [max(0, t-w+1), t]           (sliding_window_attention, attention_kernels.py:146-178)
//   cmp: compressed tokens [0, num_cmp(t))                 (attention_kernels.py:106-143 with the mask it builds)
// All h <= 8 heads of the KV group share every block (GQA).  Swapped operands make h rows of work fill a tensor tile:
//   S^T[128 keys x 16] = K_pair[128 x 64] . Q^T[64 x 16]      (M=128, N=16, K=16 x4; keys on TMEM lanes)
//   exact softmax over the item's keys (every S^T tile of the item stays in TMEM: 8 pairs x 16 columns)
//   O^T[64 dv x 8]    += V_pair^T[64 x 128] . P^T[128 x 8]     (M=64, N=8, K=16 x8; V is the MN-major A operand)
// The kernel is a stream: it is bound by how fast K/V blocks arrive (L2 in prefill, HBM in decode), so the CTA is
// warp-specialised and pipelined ACROSS items: the producer warp builds the next item's block list and keeps a TMA ring
// full (K pairs of item i+1 are in flight while item i is in its softmax / P.V), one thread issues the MMAs, four warps do
// softmax + epilogue.  With gates given, the three branch items of a token are combined in registers and only the gated
// O is written (decode: branch outputs never reach HBM).
#include "tc_common.cuh"
#include <stdlib.h>
#include <string.h>

#include "launchers.h"
#include "select.cuh"
#include "gate.cuh"

namespace nsa {
using namespace tc;

constexpr int kGN = 8;            // head slots (N of the P.V MMA)
constexpr int kGNQ = 16;          // N of the Q.K^T MMA (M=128 needs N % 16 == 0); slots 8..15 are zero
constexpr int kGStages = 4;       // TMA ring depth in 128-key pair stages (16 KB each)
constexpr int kGMaxBlk = 16;      // <= 1024 keys per item
constexpr int kGMaxPairs = kGMaxBlk / 2;
constexpr int kGPair = 128 * 128; // bytes of one pair stage
constexpr int kGTmemCols = 256;   // S^T: 8 pairs x 16 columns; O^T: 2 x 16 columns
constexpr int kGSlots = 4;        // block-list / Q^T slots: how far the producer may run ahead of the epilogue

struct GSmem {
  static constexpr int ring = 0;
  static constexpr int P = ring + kGStages * kGPair;               // [2 slots][8 pairs][128 keys][8 heads] 16-bit
  static constexpr int Qt = P + 2 * kGMaxPairs * 128 * kGN * 2;    // [4 slots][2 head groups][8 k-chunks][8 heads][8] 16-bit
  static constexpr int misc = Qt + kGSlots * 2048;
  static constexpr int fuse = misc + 1280;                          // fused decode scoring scratch (GFuse)
  static constexpr int total = fuse + 6784 + 1024;                  // 2 CTAs per SM: 2 x (total + 1 KB) <= 228 KB
};

constexpr int kGGateSlots = 8;
constexpr int kGMaxSel = 320;  // selection blocks the fused decode scorer can rank (1024 compressed keys -> <= 257)

// scratch of the fused decode step: p_cmp summed over heads per compressed key, p_grp, the selected ranges, gate MLP
struct GFuse {
  float pkey[kGMaxPairs * 128];
  float pg[kGMaxSel];
  int32_t ranges[2][64];
  float gate3[kGGateSlots][4];  // per token of this CTA (slot = token % kGGateSlots)
  float qgp[64];
  float xs[128];
};
static_assert(sizeof(GFuse) <= 6784, "GFuse must fit its slot");


struct GMisc {
  uint64_t full[kGStages], empty[kGStages];
  uint64_t list_full[kGSlots], row_free[kGSlots], p_ready[2], o_done[2], o_free[2], s_done, s_free, sel_ready[2];
  uint32_t tmem_base;
  int nblk[kGSlots];
  int blk_row[kGSlots][kGMaxBlk], blk_valid[kGSlots][kGMaxBlk];
  float red_max[2][4][kGN], red_sum[2][4][kGN];
};

struct GatherArgs {
  const void* Q;
  const int32_t* ranges;
  void* O_br[3];        // per-branch outputs [rows][h][64] (may be NULL)
  float* lse[3];        // per-branch [rows][h] (may be NULL)
  const float* gates;   // [rows][3] (cmp, sel, win); with O != NULL the branches are combined in registers
  void* O;              // gated output [rows][h][64] (may be NULL)
  int branch_mask;
  // fused decode step (branch_mask == 7, S == 1): score the compressed keys from the cmp item's softmax, select with the
  // decode rule, evaluate the gate MLP -- no other kernel runs
  int fuse;
  int S_sel;
  nsa_gate_params_t gp;
  int32_t* ranges_out;  // [rows][n_sel][2] (may be NULL)
  const nsa_decode_state_t* state;  // device-stepped decode: t0 / rows present come from here (CUDA-graph replay), else NULL
  long long* dbg;       // debug timeline of CTA 0 (NSA_B200_GATHER_DBG=1), else NULL
};

#ifdef NSA_GATHER_DBG  // compile-time (NSA_B200_NVCC_FLAGS=-DNSA_GATHER_DBG): timeline of CTA 0 for tools/dbg_gather.py
#define GDBG(tag, it)                                                                 \
  do {                                                                                \
    if (a.dbg && blockIdx.x == 0) {                                                   \
      const unsigned long long i_ = atomicAdd((unsigned long long*)a.dbg, 1ull);      \
      if (i_ < 4000) { a.dbg[1 + 2 * i_] = ((long long)(tag) << 32) | (unsigned)(it); a.dbg[2 + 2 * i_] = clock64(); } \
    }                                                                                 \
  } while (0)
#else
#define GDBG(tag, it) do { } while (0)
#endif

__device__ __forceinline__ float g_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void g_tmem_ld8f(uint32_t taddr, float (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void g_named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <typename T>
__global__ void __launch_bounds__(192, 2)
gather_attn_tc_kernel(const __grid_constant__ CUtensorMap tmK0, const __grid_constant__ CUtensorMap tmV0,
                      const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV1,
                      const __grid_constant__ CUtensorMap tmK2, const __grid_constant__ CUtensorMap tmV2, nsa_dims_t dm,
                      GatherArgs a) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  GMisc* ms = reinterpret_cast<GMisc*>(smem + GSmem::misc);
  uint8_t* ring = smem + GSmem::ring;
  // position / rows present: kernel parameters, or (device-stepped decode, CUDA-graph replay) the device record after this step's
  // produce + emit.  Kept in scalars: writing into the by-value `dm` would move the whole struct to local memory.
  int k_t0 = dm.t0, k_S_sel_kv = dm.S_sel_kv, k_S_win_kv = dm.S_win_kv, k_win_off = dm.win_off, k_S_cmp = dm.S_cmp, k_S_sel = a.S_sel;
  if (a.state) {
    const int t = a.state->t, S_raw = a.state->row_raw + 1;
    k_t0 = t;
    k_S_sel_kv = t + 1;
    k_S_win_kv = a.state->row_win + 1;
    k_win_off = t + 1 - k_S_win_kv;
    k_S_cmp = a.state->S_cmp + ((S_raw >= dm.l && (S_raw - dm.l) % dm.d == 0) ? 1 : 0);
    const int cover = t + 1 > dm.l_sel ? t + 1 : dm.l_sel;  // meta covers max(t+1, l_sel) tokens (nsa_attention.py:609-632)
    k_S_sel = ceil_div(cover, dm.l_sel);
  }
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = dm.h;

  // branches present, in order
  int brs[3], nbr = 0;  // cmp, win, sel: the selected branch last, so a fused decode step can pick its blocks first
  if (a.branch_mask & 1) brs[nbr++] = 0;
  if (a.branch_mask & 4) brs[nbr++] = 2;
  if (a.branch_mask & 2) brs[nbr++] = 1;
  GFuse* fz = reinterpret_cast<GFuse*>(smem + GSmem::fuse);
  // balanced partition of the tokens (b, g, s) over the grid; a token's branch items stay in one CTA
  const long long tokens = (long long)dm.B * dm.G * dm.S;  // < 2^31 (checked on the host)
  const int tok_begin = (int)((long long)blockIdx.x * tokens / gridDim.x);
  const int tok_end = (int)((long long)(blockIdx.x + 1) * tokens / gridDim.x);
  const int n_it = (int)(tok_end - tok_begin) * nbr;

  // item -> (row, branch, t, token parity).  Plain order is (b, g, s, branch), so a CTA walks consecutive tokens of one
  // (b, g).  A fused decode step interleaves tokens -- c0 w0 c1 s0 w1 c2 s1 w2 ... -- so the ranges a token selects
  // (during its cmp item) are three items old when its sel item needs them.
  const int n_tok = n_it / (nbr > 0 ? nbr : 1);
  auto decode_item = [&](int it, size_t& row, int& bg, int& br, int& t, int& tp) {
    int tk, b3;
    if (a.fuse) {
      if (it < 2) { tk = 0; b3 = it; }
      else if (it == n_it - 1) { tk = n_tok - 1; b3 = 2; }
      else {
        const int k = (it + 1) / 3, r = (it + 1) % 3;
        if (r == 0) { tk = k; b3 = 0; } else if (r == 1) { tk = k - 1; b3 = 2; } else { tk = k; b3 = 1; }
      }
    } else {
      tk = it / nbr;
      b3 = it % nbr;
    }
    br = brs[b3];
    const int tok = tok_begin + tk;
    bg = tok / dm.S;
    const int s = tok - bg * dm.S;
    const int g = bg % dm.G, b = bg / dm.G;
    row = ((size_t)b * dm.S + s) * dm.G + g;
    t = k_t0 + s;
    tp = tk;
  };

  // ---- one-time setup ---------------------------------------------------------------------------------
  {  // a never-loaded half of a pair stage must hold finite data (it is masked, but 0 * NaN = NaN)
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < kGStages * kGPair / 16; i += blockDim.x) reinterpret_cast<uint4*>(ring)[i] = z;
  }
  if (tid == 0) {
    for (int i = 0; i < kGStages; ++i) { mbar_init(&ms->full[i], 1); mbar_init(&ms->empty[i], 1); }
    for (int i = 0; i < kGSlots; ++i) {
      mbar_init(&ms->list_full[i], 1);
      mbar_init(&ms->row_free[i], 4);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ms->o_free[i], 4);
      mbar_init(&ms->p_ready[i], 4);
      mbar_init(&ms->o_done[i], 1);
      mbar_init(&ms->sel_ready[i], 1);
    }
    mbar_init(&ms->s_done, 1);
    mbar_init(&ms->s_free, 4);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(&ms->tmem_base, kGTmemCols);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ms->tmem_base;
  const uint32_t tmem_S = tmem;                       // pair j -> columns [j*16, j*16+16), first 8 used
  const uint32_t tmem_O = tmem + kGMaxPairs * kGNQ;   // slot rs -> columns [128 + rs*16, +8), M=64 layout

  if (warp == 4) {
    // ===== producer: block lists, Q^T, TMA ring ===========================================================
    if (lane == 0) {
      if (a.branch_mask & 1) { tma_prefetch_desc(&tmK0); tma_prefetch_desc(&tmV0); }
      if (a.branch_mask & 2) { tma_prefetch_desc(&tmK1); tma_prefetch_desc(&tmV1); }
      if (a.branch_mask & 4) { tma_prefetch_desc(&tmK2); tma_prefetch_desc(&tmV2); }
    }
    uint32_t n_loads = 0;
    int prev_bg = 0, prev_br = 0;
    int tok_sel_count[2] = {0, 0};
    auto issue_pairs = [&](int rs, int bg, int br, bool is_v) {  // lane 0 only; rs = list slot
      const CUtensorMap* tm = br == 0 ? (is_v ? &tmV0 : &tmK0) : (br == 1 ? (is_v ? &tmV1 : &tmK1) : (is_v ? &tmV2 : &tmK2));
      const int nblk = ms->nblk[rs];
      const int np = (nblk + 1) >> 1;
      for (int j = 0; j < np; ++j) {
        const uint32_t st = n_loads % kGStages;
        mbar_wait(&ms->empty[st], ((n_loads / kGStages) & 1) ^ 1);
        const int nb = (2 * j + 1 < nblk) ? 2 : 1;
        mbar_expect_tx(&ms->full[st], nb * (kGPair / 2));
        tma_load_3d(ring + st * kGPair, tm, &ms->full[st], 0, ms->blk_row[rs][2 * j], bg);
        if (nb == 2) tma_load_3d(ring + st * kGPair + kGPair / 2, tm, &ms->full[st], 0, ms->blk_row[rs][2 * j + 1], bg);
        ++n_loads;
      }
    };
    // per-item inputs (range piece of this lane, 4 x 16 B of Q) are fetched one item ahead so their global latency
    // hides behind the TMA issue loop
    struct Pre { int a0, a1; uint4 q[4]; };
    auto prefetch = [&](int it, Pre& pr) {
      size_t row; int bg, br, t, tp;
      decode_item(it, row, bg, br, t, tp);
      pr.a0 = 0; pr.a1 = 0;
      if (br == 1) {
        if (lane < dm.n_ranges && !a.fuse) {
          const int2 rr = *reinterpret_cast<const int2*>(a.ranges + (row * dm.n_ranges + lane) * 2);
          pr.a0 = rr.x < 0 ? 0 : rr.x;
          pr.a1 = rr.y > k_S_sel_kv ? k_S_sel_kv : rr.y;
        }
      } else if (lane == 0) {
        if (br == 0) {
          pr.a1 = num_cmp_at(t, dm.l, dm.d, k_S_cmp);
        } else {
          int lo = t - dm.w + 1;
          if (lo < k_win_off) lo = k_win_off;
          if (lo < 0) lo = 0;
          pr.a0 = lo - k_win_off;
          pr.a1 = t + 1 - k_win_off;
          if (pr.a1 > k_S_win_kv) pr.a1 = k_S_win_kv;
          if (dm.w <= 0) pr.a1 = pr.a0;
        }
      }
      const T* qrow = reinterpret_cast<const T*>(a.Q) + row * h * 64;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int i = lane + 32 * c, head = i >> 3, kc = i & 7;
        pr.q[c] = head < h ? *reinterpret_cast<const uint4*>(qrow + head * 64 + kc * 8) : make_uint4(0, 0, 0, 0);
      }
    };
    Pre cur, nxt;
    if (n_it > 0) prefetch(0, cur);
    for (int it = 0; it < n_it; ++it) {
      const int rs = it & (kGSlots - 1);  // list / Q^T slot
      size_t row; int bg, br, t, tp;
      decode_item(it, row, bg, br, t, tp);
      if (it + 1 < n_it) prefetch(it + 1, nxt);
      if (lane == 0) GDBG(1, it);
      mbar_wait(&ms->row_free[rs], ((it / kGSlots) & 1) ^ 1);
      if (lane == 0) GDBG(2, it);
      // ---- block list: each lane owns one [a0, a1) piece, a warp prefix sum places its 64-key blocks ----
      int a0 = cur.a0, a1 = cur.a1;
      if (a.fuse && br == 1) {  // ranges this CTA selected while the token's cmp item was in its softmax
        mbar_wait(&ms->sel_ready[tp & 1], ((tok_sel_count[tp & 1]++) & 1));
        if (lane < dm.n_ranges) {
          a0 = fz->ranges[tp & 1][2 * lane] < 0 ? 0 : fz->ranges[tp & 1][2 * lane];
          a1 = fz->ranges[tp & 1][2 * lane + 1] > k_S_sel_kv ? k_S_sel_kv : fz->ranges[tp & 1][2 * lane + 1];
        }
      }
      const int nb = a1 > a0 ? (a1 - a0 + 63) >> 6 : 0;
      int incl = nb;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const int off = incl - nb;
      for (int k = 0; k < nb; ++k) {
        const int idx = off + k;
        if (idx < kGMaxBlk) {
          ms->blk_row[rs][idx] = a0 + 64 * k;
          ms->blk_valid[rs][idx] = a1 - (a0 + 64 * k) < 64 ? a1 - (a0 + 64 * k) : 64;
        }
      }
      const int tot = __shfl_sync(0xffffffffu, incl, 31);
      if (lane == 0) ms->nblk[rs] = tot < kGMaxBlk ? tot : kGMaxBlk;
      // ---- Q^T into the no-swizzle K-major core-matrix layout: (head, k) -> (head/8)*1024 + (k/8)*128 + (head%8)*16 ----
      {
        uint8_t* qb = smem + GSmem::Qt + rs * 2048;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int i = lane + 32 * c, head = i >> 3, kc = i & 7;
          *reinterpret_cast<uint4*>(qb + (head >> 3) * 1024 + kc * 128 + (head & 7) * 16) = cur.q[c];
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ms->list_full[rs]);
        GDBG(3, it);
        issue_pairs(rs, bg, br, false);                       // K pairs of item it
        GDBG(4, it);
        if (it > 0) issue_pairs((it - 1) & (kGSlots - 1), prev_bg, prev_br, true);  // V pairs of item it-1
        GDBG(5, it);
      }
      prev_bg = bg; prev_br = br;
      cur = nxt;
      __syncwarp();
    }
    if (lane == 0 && n_it > 0) issue_pairs((n_it - 1) & (kGSlots - 1), prev_bg, prev_br, true);
  } else if (warp == 5) {
    // ===== MMA issuer ======================================================================================
    if (lane == 0) {
      constexpr uint32_t idesc_qk = make_idesc_f16(128, kGNQ, TcType<T>::fmt, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_f16(64, kGN, TcType<T>::fmt, 1, 1);
      uint32_t n_cons = 0;
      auto pv = [&](int i) {
        const int rs = i & 1, ls = i & (kGSlots - 1);
        const int nblk = ms->nblk[ls];
        const int np = (nblk + 1) >> 1;
        GDBG(20, i);
        mbar_wait(&ms->p_ready[rs], (i >> 1) & 1);
        GDBG(21, i);
        mbar_wait(&ms->o_free[rs], ((i >> 1) & 1) ^ 1);  // the epilogue of item i-2 has read this O^T slot
        tc_fence_after();
        bool first_mma = true;
        for (int j = 0; j < np; ++j) {
          const uint32_t st = n_cons % kGStages;
          mbar_wait(&ms->full[st], (n_cons / kGStages) & 1);
          tc_fence_after();
          const uint32_t v_base = smem_u32(ring + st * kGPair);
          const uint32_t p_base = smem_u32(smem + GSmem::P + (rs * kGMaxPairs + j) * 128 * 16);
          int keys = ms->blk_valid[ls][2 * j];
          if (2 * j + 1 < nblk) keys = 64 + ms->blk_valid[ls][2 * j + 1];
          const int ksteps = (keys + 15) >> 4;  // skip k-steps made only of masked keys
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t ad = make_smem_desc(v_base + k * 2048, 8192, 1024, kSwizzle128B);  // V^T: MN-major A
            const uint64_t bd = make_smem_desc(p_base + k * 256, 128, 2048, kSwizzleNone);    // P^T: MN-major B
            umma_f16(tmem_O + rs * 16, ad, bd, idesc_pv, first_mma ? 0u : 1u);
            first_mma = false;
          }
          umma_commit(&ms->empty[st]);
          ++n_cons;
        }
        umma_commit(&ms->o_done[rs]);
        GDBG(22, i);
      };
      for (int it = 0; it < n_it; ++it) {
        const int ls = it & (kGSlots - 1);
        GDBG(10, it);
        mbar_wait(&ms->list_full[ls], (it / kGSlots) & 1);
        const int nblk = ms->nblk[ls];
        const int np = (nblk + 1) >> 1;
        GDBG(11, it);
        mbar_wait(&ms->s_free, (it & 1) ^ 1);
        GDBG(12, it);  // the softmax warps hold S^T of item it-1 in registers
        tc_fence_after();
        const uint32_t q_base = smem_u32(smem + GSmem::Qt + ls * 2048);
        for (int j = 0; j < np; ++j) {
          const uint32_t st = n_cons % kGStages;
          mbar_wait(&ms->full[st], (n_cons / kGStages) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(ring + st * kGPair);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ad = make_smem_desc(a_base + k * 32, 16, 1024, kSwizzle128B);
            const uint64_t bd = make_smem_desc(q_base + k * 256, 128, 1024, kSwizzleNone);
            umma_f16(tmem_S + j * kGNQ, ad, bd, idesc_qk, k > 0);
          }
          umma_commit(&ms->empty[st]);
          ++n_cons;
        }
        umma_commit(&ms->s_done);
        GDBG(13, it);
        if (it > 0) pv(it - 1);
      }
      if (n_it > 0) pv(n_it - 1);
    }
  } else {
    // ===== softmax + epilogue warps: thread = key lane of every pair ========================================
    // Per item: A = pull S^T out of TMEM (frees it for the next item's Q.K^T), B = softmax + P^T (+ fused decode
    // scoring), C = epilogue once P.V is done.  A of item it+1 runs BEFORE C of item it, so the MMA thread never waits
    // for an epilogue and the TMA ring keeps moving across item boundaries.
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const float sl2 = dm.scale * kLog2e;
    const bool combine = (a.gates != nullptr || a.fuse) && a.O != nullptr;
    float comb0[kGN], comb1[kGN];  // gated sums of the (at most two) tokens in flight
#pragma unroll
    for (int e = 0; e < kGN; ++e) { comb0[e] = 0.f; comb1[e] = 0.f; }
    float sc[kGMaxPairs][kGN];

    // gate MLP on q_gp = mean over heads (nsa_attention.py:32-82, :908-912) by one warp -> fz->gate3[token slot]
    auto gate_of_token = [&](size_t row, int tk, float* qgp, float* xs) {
      const T* qrow = reinterpret_cast<const T*>(a.Q) + row * h * 64;
      for (int k = lane; k < 64; k += 32) {
        float m = 0.f;
        for (int hh = 0; hh < h; ++hh) m += (float)qrow[hh * 64 + k];
        qgp[k] = m / (float)h;
      }
      __syncwarp();
      Gate3 gt = gate_forward_warp(qgp, xs, nullptr, a.gp, 64, dm.gate_hidden, dm.gate_tau, dm.gate_mode, nullptr);
      if (lane == 0) { fz->gate3[tk % kGGateSlots][0] = gt.c; fz->gate3[tk % kGGateSlots][1] = gt.s; fz->gate3[tk % kGGateSlots][2] = gt.w; }
      __syncwarp();
    };
    if (a.fuse) {
      // prologue: the gates of this CTA's first tokens, one warp per token, while the first K blocks are in flight.
      // Scratch lives in the (still unused) P buffer.
      float* scratch = reinterpret_cast<float*>(smem + GSmem::P) + warp * 256;
      for (int tk = warp; tk < n_tok && tk < kGGateSlots; tk += 4) {
        const int tok = tok_begin + tk;
        const int bg_ = tok / dm.S, s_ = tok - bg_ * dm.S;
        const size_t row_ = ((size_t)(bg_ / dm.G) * dm.S + s_) * dm.G + (bg_ % dm.G);
        gate_of_token(row_, tk, scratch, scratch + 64);
      }
      g_named_bar(1, 128);
    }

    auto stage_a = [&](int it) {  // S^T -> registers (masked), arrive s_free
      const int ls = it & (kGSlots - 1);
      mbar_wait(&ms->list_full[ls], (it / kGSlots) & 1);
      const int nblk = ms->nblk[ls];
      const int np = (nblk + 1) >> 1;
      if (tid == 0) GDBG(30, it);
      mbar_wait(&ms->s_done, it & 1);
      if (tid == 0) GDBG(31, it);
      tc_fence_after();
#pragma unroll
      for (int j = 0; j < kGMaxPairs; ++j)
        if (j < np) g_tmem_ld8f(tmem_S + lane_base + j * kGNQ, sc[j]);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < kGMaxPairs; ++j)  // no use of the loaded registers may be scheduled above the wait
        asm volatile("" : "+f"(sc[j][0]), "+f"(sc[j][1]), "+f"(sc[j][2]), "+f"(sc[j][3]), "+f"(sc[j][4]), "+f"(sc[j][5]),
                          "+f"(sc[j][6]), "+f"(sc[j][7]));
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->s_free);
#pragma unroll
      for (int j = 0; j < kGMaxPairs; ++j) {
        const int blk = 2 * j + (tid >> 6);
        const bool ok = j < np && blk < nblk && (tid & 63) < ms->blk_valid[ls][blk];
#pragma unroll
        for (int e = 0; e < kGN; ++e) sc[j][e] = ok ? sc[j][e] : -INFINITY;
      }
    };

    if (n_it > 0) stage_a(0);
    for (int it = 0; it < n_it; ++it) {
      const int rs = it & 1, ls = it & (kGSlots - 1);
      size_t row; int bg, br, t, tp;
      decode_item(it, row, bg, br, t, tp);
      const int nblk = ms->nblk[ls];
      const int np = (nblk + 1) >> 1;
      // ---- B: exact softmax over the item's keys ------------------------------------------------------------
      float mx[kGN];
#pragma unroll
      for (int e = 0; e < kGN; ++e) mx[e] = -INFINITY;
#pragma unroll
      for (int j = 0; j < kGMaxPairs; ++j)
        if (j < np) {
#pragma unroll
          for (int e = 0; e < kGN; ++e) mx[e] = fmaxf(mx[e], sc[j][e]);
        }
#pragma unroll
      for (int e = 0; e < kGN; ++e) mx[e] = warp_max(mx[e]);
      if (lane == 0) {
#pragma unroll
        for (int e = 0; e < kGN; ++e) ms->red_max[rs][warp][e] = mx[e];
      }
      g_named_bar(1, 128);
      float sum[kGN];
#pragma unroll
      for (int e = 0; e < kGN; ++e) {
        mx[e] = fmaxf(fmaxf(ms->red_max[rs][0][e], ms->red_max[rs][1][e]), fmaxf(ms->red_max[rs][2][e], ms->red_max[rs][3][e]));
        mx[e] = mx[e] > -INFINITY ? mx[e] * sl2 : 0.f;  // log2-domain offset; all-masked head slots stay finite
        sum[e] = 0.f;
      }
      uint8_t* Pbuf = smem + GSmem::P + rs * kGMaxPairs * 128 * 16;
#pragma unroll
      for (int j = 0; j < kGMaxPairs; ++j) {
        if (j < np) {
          uint32_t pk[4];
#pragma unroll
          for (int e = 0; e < kGN; e += 2) {
            const float p0 = g_ex2(fmaf(sc[j][e], sl2, -mx[e]));  // masked keys: ex2(-inf) = 0
            const float p1 = g_ex2(fmaf(sc[j][e + 1], sl2, -mx[e + 1]));
            sum[e] += p0;
            sum[e + 1] += p1;
            sc[j][e] = p0;
            sc[j][e + 1] = p1;
            pk[e >> 1] = pack2(T(), p0, p1);
          }
          // P^T for the MN-major no-swizzle B operand: [pair][key][8 heads] -> one 16-byte row per key
          *reinterpret_cast<uint4*>(Pbuf + (j * 128 + tid) * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        }
      }
#pragma unroll
      for (int e = 0; e < kGN; ++e) sum[e] = warp_sum(sum[e]);
      if (lane == 0) {
#pragma unroll
        for (int e = 0; e < kGN; ++e) ms->red_sum[rs][warp][e] = sum[e];
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->p_ready[rs]);
      g_named_bar(1, 128);
      if (tid == 0) GDBG(32, it);

      if (a.fuse && br == 0) {
        // ---- fused decode scoring: p_cmp (this item's softmax) -> Eq.9 -> Eq.10 -> top-n ranges; gate MLP ----
        // (compute_pcmp_all / map_pcmp_to_pslc / select_topn_ranges, selection_scorer.py:42-61, :64-86, :124-249)
        float invl[kGN];
#pragma unroll
        for (int e = 0; e < kGN; ++e) {
          const float l = ms->red_sum[rs][0][e] + ms->red_sum[rs][1][e] + ms->red_sum[rs][2][e] + ms->red_sum[rs][3][e];
          invl[e] = (e < h && l > 0.f) ? 1.0f / l : 0.f;
        }
#pragma unroll
        for (int j = 0; j < kGMaxPairs; ++j) {
          if (j < np) {
            float ph = 0.f;  // Eq.10 first: probabilities of the h heads of the group, summed per compressed key
#pragma unroll
            for (int e = 0; e < kGN; ++e) ph = fmaf(sc[j][e], invl[e], ph);
            fz->pkey[j * 128 + tid] = ph;
          }
        }
      }

      // ---- A of the next item ----------------------------------------------------------------------------------
      if (it + 1 < n_it) stage_a(it + 1);

      if (a.fuse && br == 0) {
        g_named_bar(1, 128);
        const int nkeys = nblk > 0 ? (nblk - 1) * 64 + ms->blk_valid[ls][nblk - 1] : 0;
        for (int blk = tid; blk < k_S_sel; blk += 128) {  // Eq.9 with l = 2d, l_sel = 4d, ascending compressed index
          const int i0 = 4 * blk;
          float acc = (i0 - 1 >= 0 && i0 - 1 < nkeys) ? 0.5f * fz->pkey[i0 - 1] : 0.f;
#pragma unroll
          for (int q = 0; q < 3; ++q)
            if (i0 + q < nkeys) acc += fz->pkey[i0 + q];
          if (i0 + 3 < nkeys) acc += 0.5f * fz->pkey[i0 + 3];
          fz->pg[blk] = acc;
        }
        if (warp == 1 && tp >= kGGateSlots) gate_of_token(row, tp, fz->qgp, fz->xs);  // tokens the prologue did not cover
        g_named_bar(1, 128);
        if (warp == 0) {
          select_row_warp(fz->pg, k_S_sel, dm.l_sel, dm.n_sel, 1, kForcedDecodeDefault, dm.n_sel, t, fz->ranges[tp & 1]);
          __syncwarp();
          if (a.ranges_out && lane < dm.n_sel)
            *reinterpret_cast<int2*>(a.ranges_out + (row * dm.n_sel + lane) * 2) =
                make_int2(fz->ranges[tp & 1][2 * lane], fz->ranges[tp & 1][2 * lane + 1]);
          __syncwarp();
          if (lane == 0) mbar_arrive(&ms->sel_ready[tp & 1]);
        }
      }


      // ---- C: epilogue: O^T (64 dv x 8 heads, M=64 layout: dv row r on lane 32*(r/16) + r%16) ---------------------
      if (tid == 0) GDBG(33, it);
      mbar_wait(&ms->o_done[rs], (it >> 1) & 1);
      if (tid == 0) GDBG(34, it);
      tc_fence_after();
      uint32_t r[8];
      if (nblk > 0) {
        tmem_ld8(tmem_O + rs * 16 + lane_base, r);
        tmem_ld_wait();
        asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]));
      } else {
#pragma unroll
        for (int e = 0; e < kGN; ++e) r[e] = 0u;  // empty item -> zeros, lse = -inf (attention_kernels.py:769-771)
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->o_free[rs]);
      float gate = 1.f;
      if (a.fuse) gate = fz->gate3[tp % kGGateSlots][br];
      else if (combine) gate = a.gates[row * 3 + br];
      if (lane < 16) {
        const int dv = warp * 16 + lane;
        T* ob = a.O_br[br] ? reinterpret_cast<T*>(a.O_br[br]) : nullptr;
        const bool last = combine && br == brs[nbr - 1];  // last branch item of this token: write the gated output
        T* of = reinterpret_cast<T*>(a.O);
#pragma unroll
        for (int e = 0; e < kGN; ++e) {
          if (e < h) {
            const float l = ms->red_sum[rs][0][e] + ms->red_sum[rs][1][e] + ms->red_sum[rs][2][e] + ms->red_sum[rs][3][e];
            const float o = l > 0.f ? __fdividef(__uint_as_float(r[e]), l) : 0.f;
            if (ob) ob[(row * h + e) * 64 + dv] = T(o);
            if (combine) {
              const float c = fmaf(gate, o, (tp & 1) ? comb1[e] : comb0[e]);
              if (last) of[(row * h + e) * 64 + dv] = T(c);
              if (tp & 1) comb1[e] = last ? 0.f : c;
              else comb0[e] = last ? 0.f : c;
            }
          }
        }
      }
      if (a.lse[br] && tid < h) {
        const float l = ms->red_sum[rs][0][tid] + ms->red_sum[rs][1][tid] + ms->red_sum[rs][2][tid] + ms->red_sum[rs][3][tid];
        const float m = fmaxf(fmaxf(ms->red_max[rs][0][tid], ms->red_max[rs][1][tid]), fmaxf(ms->red_max[rs][2][tid], ms->red_max[rs][3][tid]));
        a.lse[br][row * h + tid] = l > 0.f ? m * dm.scale + logf(l) : -INFINITY;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->row_free[ls]);
      if (tid == 0) GDBG(35, it);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kGTmemCols);
}

// ---- host ------------------------------------------------------------------------------------------------------
bool tc_gather_supported(const nsa_dims_t& dm, int branch_mask) {
  if (dm.impl == NSA_IMPL_SIMT) return false;
  if (!((dm.dtype == NSA_BF16 || dm.dtype == NSA_F16) && dm.Dk == 64 && dm.Dv == 64 && dm.h >= 1 && dm.h <= kGN)) return false;
  if (branch_mask & 2) {
    if (dm.l_sel % 64 != 0 || (long long)dm.n_sel * dm.l_sel > 64 * kGMaxBlk || dm.n_ranges > 32 || dm.S_sel_kv < 1) return false;
  }
  if (branch_mask & 1) {  // every row's compressed prefix must fit 16 blocks
    if (dm.S_cmp < 1 || num_cmp_at(dm.t0 + dm.S - 1, dm.l, dm.d, dm.S_cmp) > 64 * kGMaxBlk) return false;
  }
  if (branch_mask & 4) {
    if (dm.S_win_kv < 1 || dm.w < 1 || dm.w > 64 * (kGMaxBlk - 1)) return false;
  }
  return true;
}

struct GatherPtrs {
  const void *K[3], *V[3];
};

template <typename T>
static int launch_gather_t(const nsa_dims_t& dm, const GatherPtrs& kv, GatherArgs a, cudaStream_t stream) {
  CUtensorMap tm[6];
  memset(tm, 0, sizeof(tm));
  const int slabs = dm.B * dm.G;
  // stepped decode: the maps are baked into a CUDA graph, so they bound the slab capacity; rows past the ones present are
  // masked by their counts inside the kernel (slabs are zero-initialised / hold finite stale rows)
  const int rows[3] = {a.state ? dm.cap_cmp : dm.S_cmp, a.state ? dm.cap_sel : dm.S_sel_kv, a.state ? dm.cap_win : dm.S_win_kv};
  const long long caps[3] = {dm.cap_cmp, dm.cap_sel, dm.cap_win};
  for (int br = 0; br < 3; ++br) {
    if (!(a.branch_mask & (1 << br))) continue;
    if (int rc = make_tmap_rows(&tm[2 * br], kv.K[br], dm.dtype, 64, rows[br], 64, caps[br] * 64, slabs, 64)) return rc;
    if (int rc = make_tmap_rows(&tm[2 * br + 1], kv.V[br], dm.dtype, 64, rows[br], 64, caps[br] * 64, slabs, 64)) return rc;
  }
  int nbr = 0;
  for (int br = 0; br < 3; ++br) nbr += (a.branch_mask >> br) & 1;
  // prefill: ~16 consecutive tokens per CTA; short launches (decode): two CTAs per SM, tokens split evenly
  const long long tokens = (long long)slabs * dm.S;
  long long grid_ll = tokens >= 16LL * 296 ? (tokens + 15) / 16 : (tokens < 296 ? tokens : 296);
  const int grid = (int)grid_ll;
  auto kern = gather_attn_tc_kernel<T>;
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_smem_attr(kern, GSmem::total, attr_done, "gather tc", /*two CTAs per SM*/ true)) return rc;
#ifdef NSA_GATHER_DBG
  static const bool dbg_on = getenv("NSA_B200_GATHER_DBG") != nullptr;
  static long long* dbg_buf = nullptr;
  if (dbg_on) {
    if (!dbg_buf) cudaMalloc(&dbg_buf, 8008 * sizeof(long long));
    cudaMemsetAsync(dbg_buf, 0, 8008 * sizeof(long long), stream);
    a.dbg = dbg_buf;
  }
#endif
  kern<<<grid, 192, GSmem::total, stream>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], dm, a);
#ifdef NSA_GATHER_DBG
  if (dbg_on) {  // debug only: dump the timeline of CTA 0 (tag, item, clock)
    static int dumps = 0;
    cudaStreamSynchronize(stream);
    if (dumps++ == 3) {
      static long long host[8008];
      cudaMemcpy(host, dbg_buf, sizeof(host), cudaMemcpyDeviceToHost);
      long long n = host[0] < 4000 ? host[0] : 4000;
      for (long long i = 0; i < n; ++i)
        fprintf(stderr, "GDBG %lld %lld %lld\n", host[1 + 2 * i] >> 32, host[1 + 2 * i] & 0xffffffff, host[2 + 2 * i]);
    }
  }
#endif
  return check_launch("gather_attn_tc_kernel");
}

// branch_mask selects the branches; K[br]/V[br] are read only for selected branches.
bool tc_gather_fuse_supported(const nsa_dims_t& dm, int S_sel) {
  return dm.S == 1 && dm.l == 2 * dm.d && dm.l_sel == 4 * dm.d && dm.gate_hidden <= 128 && S_sel >= 1 && S_sel <= kGMaxSel &&
         dm.n_sel <= 32 && tc_gather_supported(dm, 7);
}

// gp_fuse != NULL: fused decode step (scoring + selection + gate inside the kernel; `ranges`/`gates` are ignored).
int launch_gather_tc(const nsa_dims_t& dm, int branch_mask, const void* Q, const void* const* K, const void* const* V,
                     const int32_t* ranges, void* const* O_br, float* const* lse, const float* gates, void* O,
                     const nsa_gate_params_t* gp_fuse, int S_sel, int32_t* ranges_out, cudaStream_t stream,
                     const nsa_decode_state_t* state) {
  static_assert(sizeof(GMisc) <= 1280, "GMisc must fit its slot");
  NSA_REQUIRE(forced_code_default(1, 0, dm.l_sel) == kForcedDecodeDefault, "gather(tc): forced-block code changed");
  if (dm.B * dm.S * dm.G == 0 || branch_mask == 0) return NSA_OK;
  NSA_REQUIRE((long long)dm.B * dm.S * dm.G < (1LL << 31), "gather(tc): B*S*G must be below 2^31");
  GatherArgs a;
  memset(&a, 0, sizeof(a));
  GatherPtrs kv;
  a.Q = Q;
  a.ranges = ranges;
  a.gates = gates;
  a.O = O;
  a.branch_mask = branch_mask;
  if (gp_fuse) {
    NSA_REQUIRE(branch_mask == 7 && (state || tc_gather_fuse_supported(dm, S_sel)) && O, "gather(tc): fused decode step not available for this shape");
    a.state = state;
    a.fuse = 1;
    a.S_sel = S_sel;
    a.gp = *gp_fuse;
    a.ranges_out = ranges_out;
  }
  for (int br = 0; br < 3; ++br) {
    kv.K[br] = K ? K[br] : nullptr;
    kv.V[br] = V ? V[br] : nullptr;
    a.O_br[br] = O_br ? O_br[br] : nullptr;
    a.lse[br] = lse ? lse[br] : nullptr;
  }
  if (dm.dtype == NSA_BF16) return launch_gather_t<__nv_bfloat16>(dm, kv, a, stream);
  return launch_gather_t<__half>(dm, kv, a, stream);
}

}  // namespace nsa
