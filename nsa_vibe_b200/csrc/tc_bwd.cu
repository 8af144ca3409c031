// Analytical backward of the three NSA branches on tcgen05 tensor cores (replaces _selection_attention_backward,
// kernels/triton_sel_kernel/__init__.py:163-231, without its first-key-only line, and the autograd of
// sliding_window_attention / batched_causal_attention_compressed, attention_kernels.py:106-178):
//   P = exp(S.scale - lse)            S = Q.K^T
//   dV += P^T.dO                      dP = dO.V^T
//   dS = P o (dP - D).scale           D  = rowsum(dO o O)
//   dQ += dS.K                        dK += dS^T.Q
// KV-tile-major, like the block-major forward (tc_sel2.cu): one CTA owns a 64-key tile of one (b, g) slab and walks the
// M-tiles of query rows that attend it, so dK/dV of the tile accumulate in TMEM over the whole walk (the h heads of a
// token and all tokens of the walk are summed by the MMA itself) and only dQ needs atomics.
//   cmp : tile j = compressed rows [64j, 64j+64)   queries t >= 64j.d + l - 1, contiguous, masked to col < num_cmp(t)
//   win : tile j = cache rows [64j, 64j+64)        queries in [first key, last key + w - 1], masked to [t-w+1, t]
//   sel : tile j = selection block j               queries from the device-built inverted index (tc_sel2.cuh), masked to the
//                                                  clamped block length
// An M-tile is 128 rows = TOK queries x h heads; its Q and dO rows arrive by TMA with hardware swizzle: one box per tile
// for cmp / win (consecutive tokens), one box per query for sel.
// Per (M-tile, key tile) pair:  S, dP (M=128, N=64) -> softmax warps form P~ = g.P and dS~ = g.scale.P o (dP - D) as 16-bit
// swizzled tiles -> dV += P~^T.dO, dK += dS~^T.Q (M=64, N=64, K=128; A and B both MN-major), dQ_t = dS~.K (M=128, N=64)
// -> four drain warps stage dQ_t as fp32 rows in shared memory and add each row into the fp32 dQ with one bulk
// reduction (cp.reduce.async.bulk .add.f32: the TMA unit streams the row to the L2 atomic units; per-lane REDG tops out
// near one fp32 per clock per SM, which made the first version of this kernel atomic-bound at ~3 us per pair).
// S/dP, P~/dS~ and dQ_t are double-buffered, the Q/dO tiles triple-buffered: the tensor core works on pair i+1 while the
// softmax warps are on pair i and the drain warps on i-1.  The binding resource is shared-memory bandwidth (136 KB of MMA
// operand reads + 128 KB of tile writes/reads per pair).
// g = gate weight of the branch for the row (the gated combine O = sum_b g_b O_b is folded in: dO_b = g_b dO).
// Warp roles: 0-7 softmax (two per scheduler, each half of the key columns), 8-11 dQ drain, 12 TMA producer, 13 MMA issuer;
// the epilogue (dK, dV -> global) uses warps 0-7.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"
#include "launchers.h"
#include "tc_sel2.cuh"

namespace nsa {
using namespace tc;

constexpr int kBwTile = 128 * 128;  // bytes: 128 rows x 64 x 2 B

constexpr int kBwQS = 3;            // Q/dO stages
constexpr int kBwDqRow = 256 + 16;  // bytes per staged dQ row: 64 fp32 + 16 B so that 16-byte stores of a warp spread over banks

constexpr int kBwThreads = 14 * 32;  // warps 0-7 softmax, 8-11 dQ drain, 12 TMA producer, 13 MMA issuer

struct BwSmem {
  static constexpr int k = 0;                       // K tile 8 KB
  static constexpr int v = 8192;                    // V tile 8 KB
  static constexpr int q = 16384;                   // [kBwQS] x 16 KB
  static constexpr int dO = q + kBwQS * kBwTile;    // [kBwQS] x 16 KB
  static constexpr int p = dO + kBwQS * kBwTile;    // [2] x 16 KB   P~  (128 rows x 64 keys)
  static constexpr int ds = p + 2 * kBwTile;        // [2] x 16 KB   dS~
  static constexpr int dq = ds + 2 * kBwTile;       // 128 rows x kBwDqRow: dQ_t staged for the bulk reductions
  static constexpr int misc = dq + 128 * kBwDqRow;
  static constexpr int total = misc + 512 + 1024;
};

struct BwMisc {
  uint64_t kv_full, fin;
  uint64_t qdo_full[kBwQS], qdo_empty[kBwQS];
  uint64_t s_full[2], s_empty[2], pds_full[2], pds_empty[2], dq_full[2], dq_empty[2];
  uint32_t tmem_base;
};

struct BwKArgs {
  const float* lse;    // [rows_h] natural-log normalisers of this branch
  const float* delta;  // [rows_h] rowsum(dO o O_b) (ungated)
  const float* gates;  // [rows][3] or NULL (weight 1)
  float *dQ, *dK, *dV;
  const S2Run* runs;   // sel only
  const int* n_runs;
  const int* tok;
  const int* hi;
  int NB;              // sel: 64-key blocks per slab
  int TOK, R;          // queries per M-tile; M-tiles per CTA (cmp / win)
  int rows_present, cap;
  long long* dbg;      // -DNSA_BWD_DBG: timeline of one CTA (tag, tile, clock) for tools/dbg_bwd.py
};

__device__ __forceinline__ float bw_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void bw_ld_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void bw_red4(float* p, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(__uint_as_float(a)), "f"(__uint_as_float(b)),
               "f"(__uint_as_float(c)), "f"(__uint_as_float(d))
               : "memory");
}

// global[dst .. dst+bytes) += shared[src ..) as fp32, done by the TMA unit (bulk async-group completion)
__device__ __forceinline__ void bw_bulk_red_add_f32(float* dst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(smem_src)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bw_bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bw_bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bw_bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// key range of one query row in cache-row coordinates (cmp: packing.py:15-23; win: attention_kernels.py:146-178)
__device__ __forceinline__ void bw_row_range(const nsa_dims_t& dm, int branch, int t, int& lo, int& hi) {
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

#ifdef NSA_BWD_DBG  // one slot per (tag, tile): plain stores, no atomics, so the probes cost a clock read each
#define BDBG(tag, it)                                                                                  \
  do {                                                                                                 \
    if (a.dbg && blockIdx.x == 0 && blockIdx.y == gridDim.y / 2 && blockIdx.z == 0 && (it) < 100)     \
      a.dbg[(tag) * 100 + (it)] = clock64();                                                           \
  } while (0)
#else
#define BDBG(tag, it) do { } while (0)
#endif

struct BwRow {
  int klo, khi;        // valid key columns of the tile for this row ([0,0) = row contributes nothing)
  float lse, dl, gt;   // natural-log normaliser, D, gate weight
};

template <typename T, int BR>
__global__ void __launch_bounds__(kBwThreads, 1)
bwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
              const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, nsa_dims_t dm, BwKArgs a) {
  // ---- work item: (slab bg, key tile j, M-tiles [tile0, tile0 + n)) -----------------------------------------------
  int bg, j, tile0, n, s_lo = 0;
  const int TOK = a.TOK, h = dm.h;
  if (BR == 1) {
    if ((int)blockIdx.x >= *a.n_runs) return;
    const S2Run run = a.runs[blockIdx.x];
    bg = run.list / a.NB;
    j = run.list % a.NB;
    tile0 = run.tile0;
    n = run.ntiles;
  } else {
    bg = blockIdx.z;
    j = blockIdx.y;
    int s_hi;
    if (BR == 0) {
      s_lo = 64 * j * dm.d + dm.l - 1 - dm.t0;
      s_hi = dm.S;
    } else {
      s_lo = dm.win_off + 64 * j - dm.t0;
      s_hi = dm.win_off + 64 * j + 63 + dm.w - dm.t0;  // exclusive
      if (s_hi > dm.S) s_hi = dm.S;
    }
    if (s_lo < 0) s_lo = 0;
    const int nt = s_hi > s_lo ? ceil_div(s_hi - s_lo, TOK) : 0;
    tile0 = blockIdx.x * a.R;
    n = nt - tile0 < a.R ? nt - tile0 : a.R;
  }
  if (n <= 0) return;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  BwMisc* ms = reinterpret_cast<BwMisc*>(smem + BwSmem::misc);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = bg % dm.G, b = bg / dm.G;

  // query of (tile i, position tok_l): global token index b*S + s, or -1 (padding)
  auto get_tk = [&](int i, int tok_l) -> int {
    if (i >= n || tok_l >= TOK) return -1;
    if (BR == 1) return a.tok[(size_t)(tile0 + i) * TOK + tok_l];
    const int s = s_lo + (tile0 + i) * TOK + tok_l;
    return s < dm.S ? b * dm.S + s : -1;
  };

  // ---- setup ---------------------------------------------------------------------------------------------------
  {  // rows the TMA never writes (padding rows, padded queries) must hold finite data
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < 2 * kBwQS * kBwTile / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + BwSmem::q)[i] = z;
  }
  if (tid == 0) {
    mbar_init(&ms->kv_full, 1);
    mbar_init(&ms->fin, 1);
    for (int s = 0; s < kBwQS; ++s) {
      mbar_init(&ms->qdo_full[s], 1);
      mbar_init(&ms->qdo_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&ms->s_full[s], 1);
      mbar_init(&ms->s_empty[s], 8);
      mbar_init(&ms->pds_full[s], 8);
      mbar_init(&ms->pds_empty[s], 1);
      mbar_init(&ms->dq_full[s], 1);
      mbar_init(&ms->dq_empty[s], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmdO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
  }
  if (warp == 0) tmem_alloc(&ms->tmem_base, 512);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ms->tmem_base;
  // TMEM columns: S[s] = s*64, dP[s] = 128 + s*64, dQ_t[s] = 256 + s*64, dK = 384 (M=64), dV = 448 (M=64)
  constexpr uint32_t kColS = 0, kColDP = 128, kColDQ = 256, kColDK = 384, kColDV = 448;

  if (warp == 12) {
    // ===== TMA producer: the key tile once, then the Q and dO rows of every M-tile.  cmp / win: the queries of a tile are
    // consecutive tokens -> one box (64, h, 1, TOK) per tensor; sel: one box (64, h, 1, 1) per query and tensor (small boxes
    // cost ~70 cycles each in the TMA unit, which bounds the sel walk; a 16-byte cp.async gather was 4x slower still). ====
    if (lane == 0) {
      mbar_expect_tx(&ms->kv_full, 2 * 8192);
      tma_load_3d(smem + BwSmem::k, &tmK, &ms->kv_full, 0, j * 64, bg);
      tma_load_3d(smem + BwSmem::v, &tmV, &ms->kv_full, 0, j * 64, bg);
    }
    if (BR != 1) {
      if (lane == 0) {
        for (int i = 0; i < n; ++i) {
          const int st = i % kBwQS, k = i / kBwQS;
          BDBG(1, i);
          mbar_wait(&ms->qdo_empty[st], (k & 1) ^ 1);
          BDBG(2, i);
          mbar_expect_tx(&ms->qdo_full[st], TOK * h * 256);
          const int tok0 = b * dm.S + s_lo + (tile0 + i) * TOK;  // rows past the sequence are loaded (or zero-filled) and masked
          tma_load_4d(smem + BwSmem::q + st * kBwTile, &tmQ, &ms->qdo_full[st], 0, 0, g, tok0);
          tma_load_4d(smem + BwSmem::dO + st * kBwTile, &tmdO, &ms->qdo_full[st], 0, 0, g, tok0);
        }
      }
    } else {
      int tk = get_tk(0, lane);
      for (int i = 0; i < n; ++i) {
        const int st = i % kBwQS, k = i / kBwQS;
        const int tk_next = get_tk(i + 1, lane);
        const unsigned have = __ballot_sync(0xffffffffu, tk >= 0);
        if (lane == 0) {
          mbar_wait(&ms->qdo_empty[st], (k & 1) ^ 1);
          mbar_expect_tx(&ms->qdo_full[st], __popc(have) * h * 256);
        }
        __syncwarp();
        if (tk >= 0) {
          tma_load_4d(smem + BwSmem::q + st * kBwTile + lane * h * 128, &tmQ, &ms->qdo_full[st], 0, 0, g, tk);
          tma_load_4d(smem + BwSmem::dO + st * kBwTile + lane * h * 128, &tmdO, &ms->qdo_full[st], 0, 0, g, tk);
        }
        __syncwarp();
        tk = tk_next;
      }
    }
  } else if (warp == 13) {
    // ===== MMA issuer: the whole warp runs this code (warp-uniform control flow and operands, so the descriptors live in
    // uniform registers and each tcgen05.mma costs a couple of uniform adds); one elected lane issues.  Building every
    // descriptor from scratch under `if (lane == 0)` cost ~80 cycles per MMA and bounded the walk at ~3000 cycles per pair.
    constexpr uint32_t idesc_s = make_idesc_f16(128, 64, TcType<T>::fmt, 0, 0);    // S = Q.K^T, dP = dO.V^T
    constexpr uint32_t idesc_dq = make_idesc_f16(128, 64, TcType<T>::fmt, 0, 1);   // dQ_t = dS~.K   (K: MN-major B)
    constexpr uint32_t idesc_kv = make_idesc_f16(64, 64, TcType<T>::fmt, 1, 1);    // dV += P~^T.dO, dK += dS~^T.Q
    // descriptor words: lo = addr>>4 | (LBO>>4)<<16, hi = SBO>>4 | version | swizzle; K-major tiles use LBO 16 and advance
    // 32 B per 16-element k-step, MN-major tiles use LBO 8192 and advance 2048 B (16 rows) per k-step
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | ((uint32_t)kSwizzle128B << 29);
    constexpr uint32_t kLoK = (16u >> 4) << 16, kLoMN = (8192u >> 4) << 16;
    constexpr uint32_t kStepK = 32u >> 4, kStepMN = 2048u >> 4;
    const uint32_t smem0 = smem_u32(smem) >> 4;
    const uint32_t k_k = (smem0 + (BwSmem::k >> 4)) | kLoK, k_mn = (smem0 + (BwSmem::k >> 4)) | kLoMN;
    const uint32_t v_k = (smem0 + (BwSmem::v >> 4)) | kLoK;
    auto issue_sdp = [&](int i) {
      const int s = i & 1, k = i >> 1, st = i % kBwQS;
      if (lane == 0) BDBG(10, i);
      mbar_wait(&ms->qdo_full[st], (i / kBwQS) & 1);
      mbar_wait(&ms->s_empty[s], (k & 1) ^ 1);
      if (lane == 0) BDBG(12, i);
      tc_fence_after();
      const uint32_t q_k = (smem0 + ((BwSmem::q + st * kBwTile) >> 4)) | kLoK;
      const uint32_t o_k = (smem0 + ((BwSmem::dO + st * kBwTile) >> 4)) | kLoK;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_f16_elect(tmem + kColS + s * 64, q_k + kk * kStepK, kHi, k_k + kk * kStepK, kHi, idesc_s, kk > 0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
        umma_f16_elect(tmem + kColDP + s * 64, o_k + kk * kStepK, kHi, v_k + kk * kStepK, kHi, idesc_s, kk > 0);
      umma_commit_elect(&ms->s_full[s]);
    };
    auto issue_grads = [&](int i) {
      const int s = i & 1, k = i >> 1, st = i % kBwQS;
      if (lane == 0) BDBG(13, i);
      mbar_wait(&ms->pds_full[s], k & 1);
      if (lane == 0) BDBG(14, i);
      mbar_wait(&ms->dq_empty[s], (k & 1) ^ 1);
      if (lane == 0) BDBG(15, i);
      tc_fence_after();
      const uint32_t q_mn = (smem0 + ((BwSmem::q + st * kBwTile) >> 4)) | kLoMN;
      const uint32_t o_mn = (smem0 + ((BwSmem::dO + st * kBwTile) >> 4)) | kLoMN;
      const uint32_t p_mn = (smem0 + ((BwSmem::p + s * kBwTile) >> 4)) | kLoMN;
      const uint32_t s_mn = (smem0 + ((BwSmem::ds + s * kBwTile) >> 4)) | kLoMN;
      const uint32_t s_k = (smem0 + ((BwSmem::ds + s * kBwTile) >> 4)) | kLoK;
      const uint32_t acc0 = i > 0 ? 1u : 0u;
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)  // contraction over the 128 rows, 16 per step
        umma_f16_elect(tmem + kColDV, p_mn + kk * kStepMN, kHi, o_mn + kk * kStepMN, kHi, idesc_kv, kk > 0 ? 1u : acc0);
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        umma_f16_elect(tmem + kColDK, s_mn + kk * kStepMN, kHi, q_mn + kk * kStepMN, kHi, idesc_kv, kk > 0 ? 1u : acc0);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)  // contraction over the 64 keys
        umma_f16_elect(tmem + kColDQ + s * 64, s_k + kk * kStepK, kHi, k_mn + kk * kStepMN, kHi, idesc_dq, kk > 0);
      umma_commit_elect(&ms->dq_full[s]);
      umma_commit_elect(&ms->pds_empty[s]);
      umma_commit_elect(&ms->qdo_empty[st]);
      if (lane == 0) BDBG(16, i);
    };
    mbar_wait(&ms->kv_full, 0);
    issue_sdp(0);
    if (n > 1) issue_sdp(1);
    for (int i = 0; i < n; ++i) {
      issue_grads(i);
      if (i + 2 < n) issue_sdp(i + 2);
    }
    umma_commit_elect(&ms->fin);
  } else if (warp < 8) {
    // ===== softmax warps: thread = TMEM lane = row (query, head); warps 0-3 take key columns 0-31, warps 4-7 columns 32-63
    // (two warps per scheduler: one warp alone issues ~1 instruction per 2-3 cycles, which bounded the first version) =====
    const int r = tid & 127, hf = warp >> 2;
    const int tok_l = r / h, head = r - tok_l * h;
    const float c = dm.scale * kLog2e;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int sw = r & 7;
    auto get_hi = [&](int i) -> int {
      if (BR != 1 || i >= n || tok_l >= TOK) return 0;
      return a.hi[(size_t)(tile0 + i) * TOK + tok_l];
    };
    // Row data as loaded; nothing here is looked at until the tile is processed, so the loads of tile i+1 stay in flight
    // while tile i is computed (two-level prefetch: the query of tile i+2 is requested at the same time).
    auto meta = [&](int tk, int hi_blk) -> BwRow {
      BwRow m;
      m.klo = 0; m.khi = 0; m.lse = -INFINITY; m.dl = 0.f; m.gt = 0.f;
      if (tk >= 0) {
        const size_t grow = ((size_t)tk * dm.G + g) * h + head;
        m.lse = a.lse[grow];
        m.dl = a.delta[grow];
        m.gt = a.gates ? a.gates[((size_t)tk * dm.G + g) * 3 + BR] : 1.0f;
        if (BR == 1) {
          m.khi = hi_blk;
        } else {
          int lo, hi;
          bw_row_range(dm, BR, dm.t0 + tk - b * dm.S, lo, hi);
          lo -= 64 * j;
          hi -= 64 * j;
          m.klo = lo < 0 ? 0 : (lo > 64 ? 64 : lo);
          m.khi = hi < 0 ? 0 : (hi > 64 ? 64 : hi);
        }
      }
      return m;
    };
    BwRow cur = meta(get_tk(0, tok_l), get_hi(0));
    int tk_n = get_tk(1, tok_l), hi_n = get_hi(1);
    for (int i = 0; i < n; ++i) {
      const int s = i & 1, k = i >> 1;
      const BwRow nxt = meta(tk_n, hi_n);
      tk_n = get_tk(i + 2, tok_l);
      hi_n = get_hi(i + 2);
      // P~ = g.exp2(S.c - lse2) = exp2(S.c - (lse2 - log2 g));  dS~ = P~.(dP.scale - D.scale)
      const bool live = cur.lse > -INFINITY && cur.gt > 0.f && cur.khi > cur.klo;
      const int klo = live ? cur.klo : 0, khi = live ? cur.khi : 0;
      const float e0 = live ? cur.lse * kLog2e - __log2f(cur.gt) : 0.f;
      const float dls = cur.dl * dm.scale;
      const bool full = __all_sync(0xffffffffu, klo == 0 && khi == 64);
      uint8_t* prow = smem + BwSmem::p + s * kBwTile + r * 128;
      uint8_t* srow = smem + BwSmem::ds + s * kBwTile + r * 128;
      if (tid == 0) BDBG(20, i);
      mbar_wait(&ms->s_full[s], k & 1);
      if (tid == 0) BDBG(21, i);
      mbar_wait(&ms->pds_empty[s], (k & 1) ^ 1);  // the gradient MMAs of tile i-2 have read these P~/dS~ buffers
      if (tid == 0) BDBG(22, i);
      tc_fence_after();
      uint32_t sa[32], da[32];
      tmem_ld32(tmem + lane_off + kColS + s * 64 + hf * 32, sa);
      tmem_ld32(tmem + lane_off + kColDP + s * 64 + hf * 32, da);
      bw_ld_wait32(sa);
      bw_ld_wait32(da);
      if (full) {
#pragma unroll
        for (int e = 0; e < 32; ++e) sa[e] = __float_as_uint(bw_ex2(fmaf(__uint_as_float(sa[e]), c, -e0)));
#pragma unroll
        for (int e = 0; e < 32; ++e)
          da[e] = __float_as_uint(__uint_as_float(sa[e]) * fmaf(__uint_as_float(da[e]), dm.scale, -dls));
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col = hf * 32 + e;
          const bool ok = col >= klo && col < khi;
          const float p = ok ? bw_ex2(fmaf(__uint_as_float(sa[e]), c, -e0)) : 0.f;
          da[e] = ok ? __float_as_uint(p * fmaf(__uint_as_float(da[e]), dm.scale, -dls)) : 0u;
          sa[e] = __float_as_uint(p);
        }
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // 8 keys = one 16-byte chunk; chunk index hf*4 + q, swizzled by the row
        uint4 u, w;
        u.x = pack2(T(), __uint_as_float(sa[q * 8 + 0]), __uint_as_float(sa[q * 8 + 1]));
        u.y = pack2(T(), __uint_as_float(sa[q * 8 + 2]), __uint_as_float(sa[q * 8 + 3]));
        u.z = pack2(T(), __uint_as_float(sa[q * 8 + 4]), __uint_as_float(sa[q * 8 + 5]));
        u.w = pack2(T(), __uint_as_float(sa[q * 8 + 6]), __uint_as_float(sa[q * 8 + 7]));
        w.x = pack2(T(), __uint_as_float(da[q * 8 + 0]), __uint_as_float(da[q * 8 + 1]));
        w.y = pack2(T(), __uint_as_float(da[q * 8 + 2]), __uint_as_float(da[q * 8 + 3]));
        w.z = pack2(T(), __uint_as_float(da[q * 8 + 4]), __uint_as_float(da[q * 8 + 5]));
        w.w = pack2(T(), __uint_as_float(da[q * 8 + 6]), __uint_as_float(da[q * 8 + 7]));
        *reinterpret_cast<uint4*>(prow + (((hf * 4 + q) ^ sw) << 4)) = u;
        *reinterpret_cast<uint4*>(srow + (((hf * 4 + q) ^ sw) << 4)) = w;
      }
      tc_fence_before();
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ms->s_empty[s]);
        mbar_arrive(&ms->pds_full[s]);
      }
      if (tid == 0) BDBG(23, i);
      cur = nxt;
    }
  } else {
    // ===== dQ drain warps: dQ_t (TMEM) -> fp32 dQ by vector reductions ============================================
    const int r = tid - 256;
    const int tok_l = r / h, head = r - tok_l * h;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    int tk = get_tk(0, tok_l);
    for (int i = 0; i < n; ++i) {
      const int s = i & 1, k = i >> 1;
      const int tk_next = get_tk(i + 1, tok_l);
      if (tid == 256) BDBG(30, i);
      mbar_wait(&ms->dq_full[s], k & 1);
      if (tid == 256) BDBG(31, i);
      tc_fence_after();
      uint32_t va[32], vb2[32];
      tmem_ld32(tmem + lane_off + kColDQ + s * 64, va);
      tmem_ld32(tmem + lane_off + kColDQ + s * 64 + 32, vb2);
      bw_ld_wait32(va);
      bw_ld_wait32(vb2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->dq_empty[s]);
      if (tk >= 0 && a.dQ) {  // this thread's staging row is its own: no cross-thread synchronisation
        bw_bulk_wait_read0();  // the previous tile's reduction has read the row
        uint4* stage = reinterpret_cast<uint4*>(smem + BwSmem::dq + r * kBwDqRow);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          stage[q] = make_uint4(va[q * 4], va[q * 4 + 1], va[q * 4 + 2], va[q * 4 + 3]);
          stage[8 + q] = make_uint4(vb2[q * 4], vb2[q * 4 + 1], vb2[q * 4 + 2], vb2[q * 4 + 3]);
        }
        fence_proxy_async();
        bw_bulk_red_add_f32(a.dQ + (((size_t)tk * dm.G + g) * h + head) * 64, stage, 256);
        bw_bulk_commit();
      }
      if (tid == 256) BDBG(32, i);
      tk = tk_next;
    }
    bw_bulk_wait0();
  }

  // ---- epilogue: dK (warps 0-3) and dV (warps 4-7) of the tile -> global, M=64 layout: key r on lane 32*(r/16) + r%16 ----
  if (warp < 8) {
    mbar_wait(&ms->fin, 0);
    tc_fence_after();
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t col = warp < 4 ? kColDK : kColDV;
    float* base = warp < 4 ? a.dK : a.dV;
    uint32_t va[32], vb2[32];
    tmem_ld32(tmem + lane_off + col, va);
    tmem_ld32(tmem + lane_off + col + 32, vb2);
    bw_ld_wait32(va);
    bw_ld_wait32(vb2);
    const int key = 64 * j + 16 * (warp & 3) + lane;
    if (lane < 16 && key < a.rows_present) {
      float* dst = base + ((size_t)bg * a.cap + key) * 64;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        bw_red4(dst + q * 4, va[q * 4], va[q * 4 + 1], va[q * 4 + 2], va[q * 4 + 3]);
        bw_red4(dst + 32 + q * 4, vb2[q * 4], vb2[q * 4 + 1], vb2[q * 4 + 2], vb2[q * 4 + 3]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------------------------------------------------
// D = rowsum(dO o O_b) per (row, head, branch) and d gate_b = sum_{h,dv} dO o O_b (nsa_attention.py:1393-1398).
// One warp per (b, s, g) row; HBM-bound: (1 + branches) * h * Dv elements per row.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kBwDeltaWarps = 8;

__global__ void __launch_bounds__(kBwDeltaWarps * 32)
bwd_delta_kernel(nsa_dims_t dm, const void* __restrict__ dO, const void* __restrict__ O_br, int branch_mask,
                 float* __restrict__ delta, float* __restrict__ dgates) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rows = dm.B * dm.S * dm.G;
  const size_t rows_h = (size_t)n_rows * dm.h;
  for (int row = blockIdx.x * kBwDeltaWarps + warp; row < n_rows; row += gridDim.x * kBwDeltaWarps) {
    for (int br = 0; br < 3; ++br) {
      if (!(branch_mask & (1 << br))) continue;
      float dg = 0.f;
      for (int hh = 0; hh < dm.h; ++hh) {
        const size_t e0 = ((size_t)row * dm.h + hh) * dm.Dv;
        float part = 0.f;
        for (int k = lane; k < dm.Dv; k += 32)
          part = fmaf(ld_elt(dO, e0 + k, dm.dtype), ld_elt(O_br, br * rows_h * dm.Dv + e0 + k, dm.dtype), part);
        part = warp_sum(part);
        dg += part;
        if (lane == 0) delta[br * rows_h + (size_t)row * dm.h + hh] = part;
      }
      if (dgates && lane == 0) dgates[(size_t)row * 3 + br] = dg;
    }
  }
}

// 16-bit, Dv = 64: eight lanes per (b,s,g) row, 16-byte loads (a 64-element head row = 8 lanes x 8 elements), a 3-step shuffle
// per head and branch; a warp works on four rows at once.  (One warp per head row with 4-byte loads and a 5-step warp sum per
// head and branch was instruction-bound: 59 % issue-active at 2.6 TB/s.)
template <typename T>
__global__ void __launch_bounds__(kBwDeltaWarps * 32)
bwd_delta16_kernel(nsa_dims_t dm, const T* __restrict__ dO, const T* __restrict__ O_br, int branch_mask, float* __restrict__ delta,
                   float* __restrict__ dgates) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane & 7, grp = lane >> 3;
  const int n_rows = dm.B * dm.S * dm.G, h = dm.h;
  const size_t rows_h = (size_t)n_rows * h;
  auto dot8 = [](const uint4& a, const uint4& b) {
    const uint32_t ua[4] = {a.x, a.y, a.z, a.w}, ub[4] = {b.x, b.y, b.z, b.w};
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      T a0, a1, b0, b1;
      memcpy(&a0, &ua[k], 2);
      memcpy(&a1, reinterpret_cast<const char*>(&ua[k]) + 2, 2);
      memcpy(&b0, &ub[k], 2);
      memcpy(&b1, reinterpret_cast<const char*>(&ub[k]) + 2, 2);
      acc = fmaf((float)a0, (float)b0, acc);
      acc = fmaf((float)a1, (float)b1, acc);
    }
    return acc;
  };
  const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
  for (int base = (blockIdx.x * kBwDeltaWarps + warp) * 4; base < n_rows; base += gridDim.x * kBwDeltaWarps * 4) {  // warp-uniform
    const int row = base + grp;
    const bool valid = row < n_rows;
    float dg[3] = {0.f, 0.f, 0.f};
    for (int hh = 0; hh < h; ++hh) {
      const size_t e0 = ((size_t)(valid ? row : 0) * h + hh) * 64 + sub * 8;
      const uint4 ud = valid ? *reinterpret_cast<const uint4*>(dO + e0) : zero;
      uint4 uo[3];
#pragma unroll
      for (int br = 0; br < 3; ++br)
        uo[br] = (valid && (branch_mask & (1 << br))) ? *reinterpret_cast<const uint4*>(O_br + br * rows_h * 64 + e0) : zero;
#pragma unroll
      for (int br = 0; br < 3; ++br) {
        if (!(branch_mask & (1 << br))) continue;  // kernel-uniform
        float part = dot8(ud, uo[br]);
        part += __shfl_xor_sync(0xffffffffu, part, 4);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        dg[br] += part;
        if (valid && sub == 0) delta[br * rows_h + (size_t)row * h + hh] = part;
      }
    }
    if (dgates && valid && sub < 3 && (branch_mask & (1 << sub))) dgates[(size_t)row * 3 + sub] = sub == 0 ? dg[0] : (sub == 1 ? dg[1] : dg[2]);
  }
}

// ---- host ------------------------------------------------------------------------------------------------------
int make_tmap_q_heads(CUtensorMap* out, const void* base, int dtype, int D, int h, int G, long long n_tokens, int box_tokens);

static int bw_tokp(const nsa_dims_t& dm) { return 128 / dm.h > 32 ? 32 : 128 / dm.h; }  // one producer lane per query

bool tc_bwd_supported(const nsa_dims_t& dm, int branch) {
  if (dm.impl == NSA_IMPL_SIMT) return false;
  if (!((dm.dtype == NSA_BF16 || dm.dtype == NSA_F16) && dm.Dk == 64 && dm.Dv == 64 && dm.h >= 1 && dm.h <= 128 && dm.S >= 1))
    return false;
  if (branch == 0) return dm.S_cmp >= 1;
  if (branch == 2) return dm.S_win_kv >= 1 && dm.w >= 1;
  return sel2_index_supported(dm);
}

static size_t bw_delta_bytes(const nsa_dims_t& dm) {
  return (((size_t)3 * dm.B * dm.S * dm.G * dm.h * sizeof(float)) + 255) & ~(size_t)255;
}

int64_t tc_bwd_workspace(const nsa_dims_t& dm) {
  bool any = false;
  for (int br = 0; br < 3; ++br) any = any || tc_bwd_supported(dm, br);
  if (!any) return 0;
  return (int64_t)bw_delta_bytes(dm) + (tc_bwd_supported(dm, 1) ? sel2_index_workspace(dm) : 0);
}

template <typename T, int BR>
static int launch_bwd_branch(const nsa_dims_t& dm, const BwdArgs& a, const float* delta, const S2Index* idx, cudaStream_t stream) {
  const size_t rows_h = (size_t)dm.B * dm.S * dm.G * dm.h;
  const int rows = BR == 0 ? dm.S_cmp : (BR == 1 ? dm.S_sel_kv : dm.S_win_kv);
  const int cap = BR == 0 ? dm.cap_cmp : (BR == 1 ? dm.cap_sel : dm.cap_win);
  const int slabs = dm.B * dm.G;
  CUtensorMap tmQ, tmdO, tmK, tmV;
  const int box_tok = BR == 1 ? 1 : bw_tokp(dm);  // cmp / win tiles are runs of consecutive tokens: one box per tile
  if (int rc = make_tmap_q_heads(&tmQ, a.Q, dm.dtype, 64, dm.h, dm.G, (long long)dm.B * dm.S, box_tok)) return rc;
  if (int rc = make_tmap_q_heads(&tmdO, a.dO, dm.dtype, 64, dm.h, dm.G, (long long)dm.B * dm.S, box_tok)) return rc;
  if (int rc = make_tmap_rows(&tmK, a.K[BR], dm.dtype, 64, rows, 64, (long long)cap * 64, slabs, 64)) return rc;
  if (int rc = make_tmap_rows(&tmV, a.V[BR], dm.dtype, 64, rows, 64, (long long)cap * 64, slabs, 64)) return rc;
  BwKArgs k;
  memset(&k, 0, sizeof(k));
  k.lse = a.lse + BR * rows_h;
  k.delta = delta + BR * rows_h;
  k.gates = a.gates;
  static const bool no_dq = getenv("NSA_B200_BWD_NODQ") != nullptr;  // A/B switch (benchmarks only): skip the dQ reductions
  k.dQ = no_dq ? nullptr : a.dQ;
  k.dK = a.dK[BR];
  k.dV = a.dV[BR];
  k.TOK = bw_tokp(dm);
  k.rows_present = rows;
  k.cap = cap;
  dim3 grid;
  if (BR == 1) {
    k.runs = idx->runs;
    k.n_runs = idx->n_runs;
    k.tok = idx->tok;
    k.hi = idx->hi;
    k.NB = idx->gm.NB;
    k.R = kS2Run;
    grid = dim3(idx->max_runs, 1, 1);
  } else {
    const int n_kt = ceil_div(rows, 64);
    // M-tiles a key tile can have; split into chunks of R so that the grid fills the machine at least twice
    const int nt_max = BR == 0 ? ceil_div(dm.S, k.TOK) : ceil_div(64 + dm.w - 1, k.TOK) + 1;
    const int want = ceil_div(2 * 148, slabs * n_kt);
    int R = ceil_div(nt_max, want);
    if (R < 4) R = 4;
    if (R > 32) R = 32;
    k.R = R;
    grid = dim3(ceil_div(nt_max, R), n_kt, slabs);
    NSA_REQUIRE(n_kt <= 65535 && slabs <= 65535, "bwd(tc): grid too large (key tiles %d, slabs %d)", n_kt, slabs);
  }
  auto kern = bwd_tc_kernel<T, BR>;
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_smem_attr(kern, BwSmem::total, attr_done, "bwd tc")) return rc;
#ifdef NSA_BWD_DBG
  static long long* dbg_buf = nullptr;
  if (!dbg_buf) cudaMalloc(&dbg_buf, 8008 * sizeof(long long));
  cudaMemsetAsync(dbg_buf, 0, 8008 * sizeof(long long), stream);
  k.dbg = BR == 2 ? dbg_buf : nullptr;
#endif
  kern<<<grid, kBwThreads, BwSmem::total, stream>>>(tmQ, tmdO, tmK, tmV, dm, k);
#ifdef NSA_BWD_DBG
  if (BR == 2) {  // debug only: timeline of one CTA of the win walk (tag, tile, clock)
    static int dumps = 0;
    cudaStreamSynchronize(stream);
    if (dumps++ == 2) {
      static long long host[8008];
      cudaMemcpy(host, dbg_buf, sizeof(host), cudaMemcpyDeviceToHost);
      for (int tag = 0; tag < 40; ++tag)
        for (int it = 0; it < 100; ++it)
          if (host[tag * 100 + it]) fprintf(stderr, "BDBG %d %d %lld\n", tag, it, host[tag * 100 + it]);
    }
  }
#endif
  return check_launch("bwd_tc_kernel");
}

template <typename T>
static int launch_bwd_tc_t(const nsa_dims_t& dm, const BwdArgs& a, int tc_mask, void* workspace, cudaStream_t stream) {
  char* ws = reinterpret_cast<char*>(workspace);
  float* delta = reinterpret_cast<float*>(ws);
  const int n_rows = dm.B * dm.S * dm.G;
  int blocks = ceil_div(n_rows, kBwDeltaWarps * (dm.Dv == 64 ? 4 : 1));
  if (blocks > 148 * 32) blocks = 148 * 32;
  if (dm.Dv == 64)
    bwd_delta16_kernel<T><<<blocks, kBwDeltaWarps * 32, 0, stream>>>(dm, (const T*)a.dO, (const T*)a.O_br, a.branch_mask, delta, a.dgates);
  else
    bwd_delta_kernel<<<blocks, kBwDeltaWarps * 32, 0, stream>>>(dm, a.dO, a.O_br, a.branch_mask, delta, a.dgates);
  if (int rc = check_launch("bwd_delta_kernel")) return rc;
  if (tc_mask & 1)
    if (int rc = launch_bwd_branch<T, 0>(dm, a, delta, nullptr, stream)) return rc;
  if (tc_mask & 4)
    if (int rc = launch_bwd_branch<T, 2>(dm, a, delta, nullptr, stream)) return rc;
  if (tc_mask & 2) {
    S2Index idx;
    if (int rc = sel2_build_index(dm, a.ranges, ws + bw_delta_bytes(dm), stream, &idx)) return rc;
    if (int rc = launch_bwd_branch<T, 1>(dm, a, delta, &idx, stream)) return rc;
  }
  return NSA_OK;
}

// Backward of the branches in a.branch_mask: tensor-core kernels for the branches they serve, the SIMT kernel for the rest.
int launch_bwd_tc(const nsa_dims_t& dm, const BwdArgs& a, void* workspace, cudaStream_t stream) {
  static_assert(sizeof(BwMisc) <= 512, "BwMisc must fit its slot");
  if (dm.B * dm.S * dm.G == 0) return NSA_OK;
  int tc_mask = 0;
  for (int br = 0; br < 3; ++br)
    if ((a.branch_mask & (1 << br)) && tc_bwd_supported(dm, br) && (br != 1 || a.ranges)) tc_mask |= 1 << br;
  if (!workspace || tc_mask == 0) {
    NSA_REQUIRE(dm.impl != NSA_IMPL_TC, "attention bwd: no tcgen05 kernel for this shape or no workspace (impl=TC was forced)");
    return launch_bwd_generic(dm, a, stream);
  }
  int rc = dm.dtype == NSA_BF16 ? launch_bwd_tc_t<__nv_bfloat16>(dm, a, tc_mask, workspace, stream)
                                : launch_bwd_tc_t<__half>(dm, a, tc_mask, workspace, stream);
  if (rc) return rc;
  const int rest = a.branch_mask & ~tc_mask;
  if (rest) {
    BwdArgs b = a;
    b.branch_mask = rest;
    b.dgates = nullptr;  // already written by the delta kernel
    return launch_bwd_generic(dm, b, stream);
  }
  return NSA_OK;
}

}  // namespace nsa
