// Selected-branch attention, KV-block-major ("inverted index") -- the prefill path for long sequences.
//
// The query-major gather kernel (tc_gather.cu) re-reads 16 x 16 KB of K/V for every (b, t, g) row: 34 GB per 64k
// sequence, bound by L2 bandwidth.  Here the roles are swapped: for every 64-key block the queries that selected it are
// listed (a CSR built on the device from the ranges), the block's K/V tile is loaded ONCE per run of those queries, and
// each (query, block) pair yields a partial (O, lse) that a merge kernel folds per query -- exactly the split-softmax
// identity, so the result equals grouped_selection_attention_masked (attention_kernels.py:705-772) like the gather kernel.
//   1. sel2_count / sel2_scan / sel2_fill : ranges -> per-block pair lists (padded to whole M-tiles), pair_of[row][slot]
//   2. sel2_attn_kernel (tcgen05)         : per block, M-tiles of TOK queries x h heads against the block's 64 keys:
//        S = Q_g.K_j^T (M=128, N=64), exact softmax of the tile, O_p = P.V_j (M=128, N=64, K=64); Q rows arrive by one TMA
//        box per query (hardware swizzle), two M-tile slots ping-pong on the tensor core
//   3. sel2_merge                         : O[row] = sum_k w_k O_k, w_k = exp(lse_k - LSE), LSE = logsumexp_k lse_k
// Traffic per 64k sequence: Q rows 1.6 GB (L2-resident source) + partials 1.7 GB written + read, instead of 34 GB.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"
#include "launchers.h"
#include "tc_sel2.cuh"

namespace nsa {
using namespace tc;

constexpr int kS2Tile = 128 * 128;  // bytes of one 128-row Q tile / P tile half

// ---------------------------------------------------------------------------------------------------------------------
// 1. index build.  One thread per (b, s, g) row; CTA-level shared-memory aggregation keeps the hot counters (block 0 and
//    the local blocks are picked by every row) off the global atomics.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int s2_row_blocks(const S2Geom& gm, const int32_t* __restrict__ rr, int* blk, int* valid) {
  int n = 0;
  for (int i = 0; i < gm.n_ranges; ++i) {
    int a0 = rr[2 * i], a1 = rr[2 * i + 1];
    if (a0 < 0) a0 = 0;
    if (a1 > gm.S_kv) a1 = gm.S_kv;
    for (int p = a0; p < a1 && n < kS2MaxSlots; p += 64) {
      blk[n] = p >> 6;
      valid[n] = a1 - p < 64 ? a1 - p : 64;
      ++n;
    }
  }
  return n;
}

constexpr int kS2IdxThreads = 256;

// counts[bg * NB + j] += number of rows that selected block j
__global__ void __launch_bounds__(kS2IdxThreads)
sel2_count_kernel(S2Geom gm, const int32_t* __restrict__ ranges, int* __restrict__ counts) {
  extern __shared__ int hist[];  // [G * NB] (rows of one CTA are consecutive (b, s, g): one b, every g)
  const int n_rows = gm.B * gm.S * gm.G;
  const int row0 = blockIdx.x * kS2IdxThreads;
  const int b0 = row0 / (gm.S * gm.G);
  for (int i = threadIdx.x; i < 2 * gm.G * gm.NB; i += blockDim.x) hist[i] = 0;  // two batches may meet in one CTA
  __syncthreads();
  const int row = row0 + threadIdx.x;
  if (row < n_rows) {
    const int g = row % gm.G, b = row / (gm.S * gm.G);
    int blk[kS2MaxSlots], valid[kS2MaxSlots];
    const int n = s2_row_blocks(gm, ranges + (size_t)row * gm.n_ranges * 2, blk, valid);
    for (int k = 0; k < n; ++k) atomicAdd(&hist[((b - b0) * gm.G + g) * gm.NB + blk[k]], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * gm.G * gm.NB; i += blockDim.x) {
    const int v = hist[i];
    if (v) {
      const int bb = b0 + i / (gm.G * gm.NB);
      if (bb < gm.B) atomicAdd(&counts[(size_t)bb * gm.G * gm.NB + i % (gm.G * gm.NB)], v);
    }
  }
}

// Offsets of the block lists (each padded to a whole number of M-tiles) and the run table of the attention kernel:
// run = (list, first M-tile, number of M-tiles <= kS2Run), every run inside one list so its CTA loads one K/V tile.
// One CTA, sequential over chunks of 1024 lists (2 x 1024 lists at 64k: microseconds).
__device__ __forceinline__ int s2_block_scan(int v, int* buf) {  // inclusive scan over 1024 threads
  buf[threadIdx.x] = v;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int a = threadIdx.x >= o ? buf[threadIdx.x - o] : 0;
    __syncthreads();
    buf[threadIdx.x] += a;
    __syncthreads();
  }
  const int r = buf[threadIdx.x];
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(1024)
sel2_scan_kernel(S2Geom gm, const int* __restrict__ counts, int* __restrict__ offs, int* __restrict__ cursors,
                 S2Run* __restrict__ runs, int* __restrict__ n_runs) {
  __shared__ int buf[1024];
  __shared__ int tile_base, run_base;
  if (threadIdx.x == 0) { tile_base = 0; run_base = 0; }
  __syncthreads();
  const int nlists = gm.B * gm.G * gm.NB;
  for (int l0 = 0; l0 < nlists; l0 += 1024) {
    const int l = l0 + threadIdx.x;
    const int c = l < nlists ? counts[l] : 0;
    const int tiles = (c + gm.tokp - 1) / gm.tokp;
    const int nr = (tiles + kS2Run - 1) / kS2Run;
    const int t_incl = s2_block_scan(tiles, buf);
    const int r_incl = s2_block_scan(nr, buf);
    const int tb = tile_base, rb = run_base;
    if (l < nlists) {
      const int t0 = tb + t_incl - tiles;
      offs[l] = t0;
      cursors[l] = 0;
      for (int r = 0; r < nr; ++r) {
        S2Run run;
        run.list = l;
        run.tile0 = t0 + r * kS2Run;
        run.ntiles = tiles - r * kS2Run < kS2Run ? tiles - r * kS2Run : kS2Run;
        run.pad = 0;
        runs[rb + r_incl - nr + r] = run;
      }
    }
    __syncthreads();
    if (threadIdx.x == 1023) { tile_base = tb + t_incl; run_base = rb + r_incl; }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    offs[nlists] = tile_base;  // sentinel = total M-tiles
    *n_runs = run_base;
  }
}

// pair p (= tile * tokp + position) gets: tok[p] = b*S + s (query), hi[p] = valid keys of the block for that row;
// pair_of[row][slot] = p.  Order inside a block list is arbitrary: partials are merged per (row, slot), so the result does
// not depend on it.
__global__ void __launch_bounds__(kS2IdxThreads)
sel2_fill_kernel(S2Geom gm, const int32_t* __restrict__ ranges, const int* __restrict__ offs, int* __restrict__ cursors,
                 int* __restrict__ tok, int* __restrict__ hi, int* __restrict__ pair_of) {
  extern __shared__ int sm[];  // hist[2*G*NB] then base[2*G*NB]
  const int nb2 = 2 * gm.G * gm.NB;
  int* hist = sm;
  int* basep = sm + nb2;
  const int n_rows = gm.B * gm.S * gm.G;
  const int row0 = blockIdx.x * kS2IdxThreads;
  const int b0 = row0 / (gm.S * gm.G);
  for (int i = threadIdx.x; i < nb2; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  const int row = row0 + threadIdx.x;
  int blk[kS2MaxSlots], valid[kS2MaxSlots], rank[kS2MaxSlots];
  int n = 0, g = 0, b = 0;
  if (row < n_rows) {
    g = row % gm.G;
    b = row / (gm.S * gm.G);
    n = s2_row_blocks(gm, ranges + (size_t)row * gm.n_ranges * 2, blk, valid);
    for (int k = 0; k < n; ++k) rank[k] = atomicAdd(&hist[((b - b0) * gm.G + g) * gm.NB + blk[k]], 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nb2; i += blockDim.x) {
    const int v = hist[i];
    const int bb = b0 + i / (gm.G * gm.NB);
    basep[i] = (v && bb < gm.B) ? atomicAdd(&cursors[(size_t)bb * gm.G * gm.NB + i % (gm.G * gm.NB)], v) : 0;
  }
  __syncthreads();
  if (row < n_rows) {
    const int s = (row / gm.G) % gm.S;
    for (int k = 0; k < kS2MaxSlots; ++k) {
      int p = -1;
      if (k < n) {
        const int li = ((b - b0) * gm.G + g) * gm.NB + blk[k];
        p = offs[(size_t)(b * gm.G + g) * gm.NB + blk[k]] * gm.tokp + basep[li] + rank[k];
        tok[p] = b * gm.S + s;
        hi[p] = valid[k];
      }
      if (pair_of) pair_of[(size_t)row * kS2MaxSlots + k] = p;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 2. block-major attention on tcgen05
//
// One CTA loads the K/V tile of its block once and walks its M-tiles as a three-stage pipeline over specialised warps:
//   warps 0-3  softmax : S(i) from TMEM -> exact softmax of the 64 keys -> P(i) (16-bit, swizzled) + the row's (l, m)
//   warps 4-7  epilogue: O(i) from TMEM -> / l -> 16-bit rows staged in the tile's own P buffer -> one TMA tensor store per warp
//   warp 8 producer (one TMA box per query into a ring of three Q buffers), warp 9 MMA issuer (QK(i) and PV(j) as they arrive).
// S, O and P are double buffered, so softmax(i+1) runs while PV(i) and the epilogue of tile i are in flight.  The first version
// gave each M-tile to ONE group of four warps that did softmax and epilogue back to back: its timeline (-DNSA_SEL2_DBG,
// profiles/r2_sel2_timeline.log) showed ~6100 cycles per tile and slot, of which ~1900 in an epilogue whose un-swizzled staging
// stores were 8-way bank conflicted, with the tensor pipe and MUFU idle 3/4 of the time.  The second staged O(i) in the tile's Q
// buffer: a Q buffer then lived from the issue of its loads to the store of its O (~7000 cycles: 3500 of load latency, QK,
// softmax, PV, epilogue; profiles/r2_sel2_timeline_v2.log), and three of them bounded the CTA at ~2850 cycles per tile with every
// unit under 50 %.  The P buffer of a tile is free exactly when O(i) is there (P.V(i) complete), so it is the staging area now:
// a Q buffer is released by the commit of QK(i) (q_free), and softmax(i+2) waits for the store of O(i) to leave P (ps_free).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kS2QRing = 3;  // Q buffers in flight: tile i loads into i % 3 while tiles i-1, i-2 are in softmax / epilogue

struct S2Smem {
  static constexpr int kv = 0;                              // K tile 8 KB, V tile 8 KB (64 keys)
  static constexpr int q = kv + 2 * 8192;                   // ring of kS2QRing x 16 KB (Q rows of a tile; later its O staging)
  static constexpr int p = q + kS2QRing * kS2Tile;          // [2] x 16 KB  (128 rows x 64 keys, 16-bit, 128B-swizzled)
  static constexpr int misc = p + 2 * kS2Tile;
  // Two CTAs must fit one SM (2 x (total + 1 KB reserved) <= 228 KB), so there is no alignment slack: the kernel requires the
  // dynamic shared-memory window to start 1024-byte aligned (it does on sm_100: it follows the 1 KB reserved region) and
  // traps otherwise.
  static constexpr int stat = misc + 256;                   // [kS2QRing][128] float2 (l, m) of a tile's rows: softmax -> epilogue
  static constexpr int total = stat + kS2QRing * 128 * 8;
};

struct S2Misc {
  uint64_t kv_full;
  uint64_t q_full[kS2QRing], q_free[kS2QRing], st_full[kS2QRing];
  uint64_t s_full[2], s_empty[2], p_full[2], p_empty[2], o_full[2], o_empty[2], ps_free[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ float s2_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void s2_ld_wait32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

#ifdef NSA_SEL2_DBG  // timeline of one CTA: one slot per (tag, tile), plain stores (tools/dbg_bwd.py prints it)
#define S2DBG(tag, it)                                                                      \
  do {                                                                                      \
    if (dbg && blockIdx.x == 2000 && (it) < 100) dbg[(tag) * 100 + (it)] = clock64();       \
  } while (0)
#else
#define S2DBG(tag, it) do { } while (0)
#endif

// grid: an upper bound on the number of runs; CTAs beyond *n_runs exit.
// O_p [tiles][TOK*h][64] (dtype T, normalised by the tile's own l; written through tmO), lse_p [pairs][h] fp32 (natural log).
template <typename T>
__global__ void __launch_bounds__(320, 2)
sel2_attn_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                 const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, nsa_dims_t dm, S2Geom gm,
                 const S2Run* __restrict__ runs, const int* __restrict__ n_runs, const int* __restrict__ tok,
                 const int* __restrict__ hi, float* __restrict__ lse_p, long long* dbg) {
  if ((int)blockIdx.x >= *n_runs) return;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  if (smem_u32(smem_raw) & 1023u) __trap();  // see S2Smem
  uint8_t* smem = smem_raw;
  S2Misc* ms = reinterpret_cast<S2Misc*>(smem + S2Smem::misc);
  float2* stat = reinterpret_cast<float2*>(smem + S2Smem::stat);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int h = dm.h, TOK = gm.tokp;
  const S2Run run = runs[blockIdx.x];
  const int bg = run.list / gm.NB, blk = run.list % gm.NB;
  const int g = bg % gm.G;
  const int n = run.ntiles;  // M-tiles of this CTA; tile i uses Q buffer i % kS2QRing and S / O / P buffer i & 1

  // ---- setup ---------------------------------------------------------------------------------------------
  {  // rows the TMA never writes (padding rows >= TOK*h, padded queries) must hold finite data
    uint4 z = make_uint4(0, 0, 0, 0);
    for (int i = tid; i < kS2QRing * kS2Tile / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem + S2Smem::q)[i] = z;
  }
  if (tid == 0) {
    mbar_init(&ms->kv_full, 1);
    for (int i = 0; i < kS2QRing; ++i) { mbar_init(&ms->q_full[i], 1); mbar_init(&ms->q_free[i], 1); mbar_init(&ms->st_full[i], 4); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&ms->s_full[i], 1);
      mbar_init(&ms->s_empty[i], 4);
      mbar_init(&ms->p_full[i], 4);
      mbar_init(&ms->p_empty[i], 1);
      mbar_init(&ms->o_full[i], 1);
      mbar_init(&ms->o_empty[i], 4);
      mbar_init(&ms->ps_free[i], 4);
    }
    fence_barrier_init();
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    tma_prefetch_desc(&tmO);
  }
  if (warp == 0) tmem_alloc(&ms->tmem_base, 256);
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = ms->tmem_base;
  // TMEM columns: S[b] at b*64 ; O[b] at 128 + b*64

  if (warp == 8) {
    // ===== TMA producer: K/V tile once, then one box per query of every M-tile =================================
    // Issuing a box costs this warp ~110 cycles whichever lane does it (profiles/r2_sel2_timeline_v3.log: ~2300 cycles for the 21
    // boxes of a tile), which is the period of the whole CTA now that the Q buffers turn over fast enough.  Tried and measured
    // slower at 64k: 16-byte cp.async copies from this warp instead of boxes (~20 instructions per chunk on one warp, 1.19 ms
    // for the branch against 1.11), both engines side by side (the warp issues them one after the other: 1.22), a second
    // producer warp (352 threads leave 80 registers: spills in the softmax warps, 1.20), one lane issuing all boxes (1.41), the
    // four epilogue warps issuing a quarter of every tile's boxes each (1.06 against 1.02: the limit is the TMA unit's box rate
    // for the SM, ~55 cycles per box with two CTAs resident, not the issuing warp).
    if (lane == 0) {
      mbar_expect_tx(&ms->kv_full, 2 * 8192);
      tma_load_3d(smem + S2Smem::kv, &tmK, &ms->kv_full, 0, blk * 64, bg);
      tma_load_3d(smem + S2Smem::kv + 8192, &tmV, &ms->kv_full, 0, blk * 64, bg);
    }
    int tk_next = (n > 0 && lane < TOK) ? tok[run.tile0 * TOK + lane] : -1;
    for (int i = 0; i < n; ++i) {
      const int qb = i % kS2QRing;
      const int tk = tk_next;                                         // query (b*S + s) of pair p0 + lane, -1 = padding
      // the next tile's queries are fetched while this tile's boxes are issued: no global-load latency at the head of a tile
      tk_next = (i + 1 < n && lane < TOK) ? tok[(run.tile0 + i + 1) * TOK + lane] : -1;
      const unsigned have = __ballot_sync(0xffffffffu, tk >= 0);
      if (lane == 0) {
        S2DBG(1, i);
        mbar_wait(&ms->q_free[qb], ((i / kS2QRing) & 1) ^ 1);         // QK of tile i - kS2QRing has read this buffer
        S2DBG(2, i);
        mbar_expect_tx(&ms->q_full[qb], __popc(have) * h * 128);
      }
      __syncwarp();
      if (tk >= 0)  // one box (64 x h x 1 x 1) = the h head rows of one query, swizzled by the TMA unit
        tma_load_4d(smem + S2Smem::q + qb * kS2Tile + lane * h * 128, &tmQ, &ms->q_full[qb], 0, 0, g, tk);
      __syncwarp();
    }
  } else if (warp == 9) {
    // ===== MMA issuer: whole warp, warp-uniform operands (descriptors in uniform registers), one elected lane issues ======
    constexpr uint32_t idesc_qk = make_idesc_f16(128, 64, TcType<T>::fmt, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_f16(128, 64, TcType<T>::fmt, 0, 1);
    constexpr uint32_t kHi = (1024u >> 4) | (1u << 14) | ((uint32_t)kSwizzle128B << 29);
    constexpr uint32_t kLoK = (16u >> 4) << 16, kLoMN = (8192u >> 4) << 16;
    const uint32_t smem0 = smem_u32(smem) >> 4;
    const uint32_t k_lo = (smem0 + (S2Smem::kv >> 4)) | kLoK, v_lo = (smem0 + ((S2Smem::kv + 8192) >> 4)) | kLoMN;
    mbar_wait(&ms->kv_full, 0);
    // QK(i) and PV(j) are issued in whichever order their inputs arrive (lane 0 polls, the warp follows): with a fixed order a
    // late Q tile held back the P.V of the tile before it, which held back that tile's epilogue, which held back the producer
    int qi = 0, pj = 0;  // next S = Q.K^T / next O = P.V to issue
    while (pj < n) {
      if (pj < qi) {
        const int b = pj & 1;
        bool ok = false;
        if (lane == 0) ok = mbar_test_wait(&ms->p_full[b], (pj >> 1) & 1) && mbar_test_wait(&ms->o_empty[b], ((pj >> 1) & 1) ^ 1);
        if (__shfl_sync(0xffffffffu, ok ? 1 : 0, 0)) {
          if (lane == 0) S2DBG(15, pj);
          tc_fence_after();
          const uint32_t p_lo = (smem0 + ((S2Smem::p + b * kS2Tile) >> 4)) | kLoK;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16_elect(tmem + 128 + b * 64, p_lo + kk * 2, kHi, v_lo + kk * (2048 >> 4), kHi, idesc_pv, kk > 0);
          umma_commit_elect(&ms->o_full[b]);
          umma_commit_elect(&ms->p_empty[b]);
          ++pj;
        }
      }
      if (qi < n) {
        const int b = qi & 1, qb = qi % kS2QRing;
        bool ok = false;
        if (lane == 0) ok = mbar_test_wait(&ms->q_full[qb], (qi / kS2QRing) & 1) && mbar_test_wait(&ms->s_empty[b], ((qi >> 1) & 1) ^ 1);
        if (__shfl_sync(0xffffffffu, ok ? 1 : 0, 0)) {
          if (lane == 0) S2DBG(12, qi);
          tc_fence_after();
          const uint32_t q_lo = (smem0 + ((S2Smem::q + qb * kS2Tile) >> 4)) | kLoK;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_f16_elect(tmem + b * 64, q_lo + kk * 2, kHi, k_lo + kk * 2, kHi, idesc_qk, kk > 0);
          umma_commit_elect(&ms->s_full[b]);
          umma_commit_elect(&ms->q_free[qb]);  // S(qi) complete = the Q buffer has been read: the producer may refill it
          ++qi;
        }
      }
    }
  } else if (warp < 4) {
    // ===== softmax warps: thread = TMEM lane = row (query, head) ==============================================
    const int r = tid;  // 0..127
    const int tok_l = r / h;
    const float c = dm.scale * kLog2e;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    const int sw = r & 7;
    // the pair's valid-key count of the NEXT tile is fetched while the current one is processed
    int tok_n = -1, hi_n = 64;
    if (n > 0 && tok_l < TOK) {
      const int p0 = run.tile0 * TOK + tok_l;
      tok_n = tok[p0];
      hi_n = hi[p0];
    }
    for (int i = 0; i < n; ++i) {
      const int b = i & 1;
      const int p = (run.tile0 + i) * TOK + tok_l;
      const bool row_ok = tok_l < TOK && tok_n >= 0;
      // rows that are not stored see the whole tile, so they never push their warp onto the masked path
      const int nk = row_ok ? hi_n : 64;
      if (i + 1 < n && tok_l < TOK) {
        tok_n = tok[p + TOK];
        hi_n = hi[p + TOK];
      }
      if (tid == 0) S2DBG(20, i);
      mbar_wait(&ms->s_full[b], (i >> 1) & 1);
      if (tid == 0) S2DBG(21, i);
      tc_fence_after();
      uint32_t va[32], vb2[32];
      tmem_ld32(tmem + lane_off + b * 64, va);
      tmem_ld32(tmem + lane_off + b * 64 + 32, vb2);
      s2_ld_wait32(va);
      s2_ld_wait32(vb2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->s_empty[b]);
      // exact softmax of the tile (64 keys): max, exponentials, sum
      const bool full = __all_sync(0xffffffffu, nk >= 64);
      float m = -INFINITY;
      if (full) {
#pragma unroll
        for (int e = 0; e < 32; ++e) m = fmaxf(m, fmaxf(__uint_as_float(va[e]), __uint_as_float(vb2[e])));
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          if (e < nk) m = fmaxf(m, __uint_as_float(va[e]));
          if (32 + e < nk) m = fmaxf(m, __uint_as_float(vb2[e]));
        }
      }
      const bool any = m > -INFINITY;
      const float mc = any ? m * c : 0.f;
      if (full) {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          va[e] = __float_as_uint(s2_ex2(fmaf(__uint_as_float(va[e]), c, -mc)));
          vb2[e] = __float_as_uint(s2_ex2(fmaf(__uint_as_float(vb2[e]), c, -mc)));
        }
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          va[e] = (any && e < nk) ? __float_as_uint(s2_ex2(fmaf(__uint_as_float(va[e]), c, -mc))) : 0u;
          vb2[e] = (any && 32 + e < nk) ? __float_as_uint(s2_ex2(fmaf(__uint_as_float(vb2[e]), c, -mc))) : 0u;
        }
      }
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
      for (int e = 0; e < 32; e += 2) {
        l0 += __uint_as_float(va[e]);
        l1 += __uint_as_float(va[e + 1]);
        l2 += __uint_as_float(vb2[e]);
        l3 += __uint_as_float(vb2[e + 1]);
      }
      const float l = (l0 + l1) + (l2 + l3);
      stat[(i % kS2QRing) * 128 + r] = make_float2(l, m);  // softmax(i + 2) waits for the epilogue of tile i (ps_free): no overrun
      if (tid == 0) S2DBG(22, i);
      mbar_wait(&ms->p_empty[b], ((i >> 1) & 1) ^ 1);  // P.V of tile i - 2 has read this P buffer ...
      mbar_wait(&ms->ps_free[b], ((i >> 1) & 1) ^ 1);  // ... and the store of O(i - 2), staged in it afterwards, has left it
      uint8_t* prow = smem + S2Smem::p + b * kS2Tile + r * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // keys 0..31: chunks 0..3 ; keys 32..63: chunks 4..7 (16 B = 8 keys each)
        uint4 u, w;
        u.x = pack2(T(), __uint_as_float(va[q * 8 + 0]), __uint_as_float(va[q * 8 + 1]));
        u.y = pack2(T(), __uint_as_float(va[q * 8 + 2]), __uint_as_float(va[q * 8 + 3]));
        u.z = pack2(T(), __uint_as_float(va[q * 8 + 4]), __uint_as_float(va[q * 8 + 5]));
        u.w = pack2(T(), __uint_as_float(va[q * 8 + 6]), __uint_as_float(va[q * 8 + 7]));
        w.x = pack2(T(), __uint_as_float(vb2[q * 8 + 0]), __uint_as_float(vb2[q * 8 + 1]));
        w.y = pack2(T(), __uint_as_float(vb2[q * 8 + 2]), __uint_as_float(vb2[q * 8 + 3]));
        w.z = pack2(T(), __uint_as_float(vb2[q * 8 + 4]), __uint_as_float(vb2[q * 8 + 5]));
        w.w = pack2(T(), __uint_as_float(vb2[q * 8 + 6]), __uint_as_float(vb2[q * 8 + 7]));
        *reinterpret_cast<uint4*>(prow + ((q ^ sw) << 4)) = u;
        *reinterpret_cast<uint4*>(prow + (((4 + q) ^ sw) << 4)) = w;
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&ms->p_full[b]);
        mbar_arrive(&ms->st_full[i % kS2QRing]);
      }
      if (tid == 0) S2DBG(23, i);
    }
  } else {
    // ===== epilogue warps: thread = TMEM lane = row; partial O of every (query, block) pair -> global ========================
    const int ew = warp - 4;                 // TMEM lane quarter
    const int r = ew * 32 + lane;
    const int tok_l = r / h, head = r - tok_l * h;
    const uint32_t lane_off = (uint32_t)(ew * 32) << 16;
    const int sw = r & 7;
    const int row0 = ew * 32;
    const bool warp_has_rows = row0 < TOK * h;
    int tok_n = -1;
    if (n > 0 && tok_l < TOK) tok_n = tok[run.tile0 * TOK + tok_l];
    for (int i = 0; i < n; ++i) {
      const int b = i & 1, qb = i % kS2QRing;
      const int p = (run.tile0 + i) * TOK + tok_l;
      const bool row_ok = tok_l < TOK && tok_n >= 0;
      if (i + 1 < n && tok_l < TOK) tok_n = tok[p + TOK];
      if (i > 0 && lane == 0) {  // the store of tile i - 1 has read its staging rows: softmax may write that P buffer again
        bulk_wait_read0();
        mbar_arrive(&ms->ps_free[(i - 1) & 1]);
      }
      if (tid == 128) S2DBG(30, i);
      mbar_wait(&ms->o_full[b], (i >> 1) & 1);
      if (tid == 128) S2DBG(31, i);
      tc_fence_after();
      uint32_t va[32], vb2[32];
      tmem_ld32(tmem + lane_off + 128 + b * 64, va);
      tmem_ld32(tmem + lane_off + 128 + b * 64 + 32, vb2);
      s2_ld_wait32(va);
      s2_ld_wait32(vb2);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ms->o_empty[b]);
      mbar_wait(&ms->st_full[qb], (i / kS2QRing) & 1);
      const float2 lm = stat[qb * 128 + r];
      const float inv = lm.x > 0.f ? 1.0f / lm.x : 0.f;
      // staging = this tile's P buffer (P.V(i) has completed: O(i) is here), rows in the 128B-swizzle pattern the tensor store undoes
      uint8_t* orow = smem + S2Smem::p + b * kS2Tile + r * 128;
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 u, w;
          u.x = pack2(T(), __uint_as_float(va[q * 8 + 0]) * inv, __uint_as_float(va[q * 8 + 1]) * inv);
          u.y = pack2(T(), __uint_as_float(va[q * 8 + 2]) * inv, __uint_as_float(va[q * 8 + 3]) * inv);
          u.z = pack2(T(), __uint_as_float(va[q * 8 + 4]) * inv, __uint_as_float(va[q * 8 + 5]) * inv);
          u.w = pack2(T(), __uint_as_float(va[q * 8 + 6]) * inv, __uint_as_float(va[q * 8 + 7]) * inv);
          w.x = pack2(T(), __uint_as_float(vb2[q * 8 + 0]) * inv, __uint_as_float(vb2[q * 8 + 1]) * inv);
          w.y = pack2(T(), __uint_as_float(vb2[q * 8 + 2]) * inv, __uint_as_float(vb2[q * 8 + 3]) * inv);
          w.z = pack2(T(), __uint_as_float(vb2[q * 8 + 4]) * inv, __uint_as_float(vb2[q * 8 + 5]) * inv);
          w.w = pack2(T(), __uint_as_float(vb2[q * 8 + 6]) * inv, __uint_as_float(vb2[q * 8 + 7]) * inv);
          *reinterpret_cast<uint4*>(orow + ((q ^ sw) << 4)) = u;
          *reinterpret_cast<uint4*>(orow + (((4 + q) ^ sw) << 4)) = w;
        }
        lse_p[(size_t)p * h + head] = lm.x > 0.f ? lm.y * dm.scale + logf(lm.x) : -INFINITY;
      }
      fence_proxy_async();
      __syncwarp();
      // rows of padding pairs carry stale bytes (the merge never reads them); rows >= TOK*h of the box are clipped by the TMA unit
      if (lane == 0) {
        if (warp_has_rows) tma_store_3d(&tmO, smem + S2Smem::p + b * kS2Tile + row0 * 128, 0, row0, run.tile0 + i);
        bulk_commit();
      }
      if (tid == 128) S2DBG(32, i);
    }
    if (lane == 0) bulk_wait0();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

// ---------------------------------------------------------------------------------------------------------------------
// 3. merge: one warp per (b, s, g) row.  O[row][head] = sum_k exp(lse_k - LSE) O_k, LSE = logsumexp_k lse_k; empty row ->
//    zeros and lse = -inf (attention_kernels.py:769-771).  Slots are folded in slot order, so the result is deterministic.
// ---------------------------------------------------------------------------------------------------------------------
// One thread per 16-byte chunk of a row (h*8 chunks): its 16 lse values and 16 partial chunks are all requested before any
// is used, and the slot table of the thread's next row is already on its way: a row costs one memory round trip.
constexpr int kS2MergeRows = 4;  // rows per CTA

template <typename T>
__global__ void __launch_bounds__(kS2MergeRows * 64)
sel2_merge_kernel(int n_rows, int h, const int* __restrict__ pair_of, const T* __restrict__ O_p, const float* __restrict__ lse_p,
                  T* __restrict__ O, float* __restrict__ lse) {
  const int chunks = h * 8;                   // 16-byte chunks per row (h heads x 64 x 2 B), <= 64 threads per row
  const int rl = threadIdx.x / 64, cidx = threadIdx.x % 64;
  if (cidx >= chunks) return;
  const int head = cidx >> 3;
  const int stride = gridDim.x * kS2MergeRows;
  // the NEXT row's slot table is requested before this row's partials are consumed: one memory round trip per row, not two
  int4 nx[kS2MaxSlots / 4];
  int row = blockIdx.x * kS2MergeRows + rl;
  if (row < n_rows) {
    const int4* pp = reinterpret_cast<const int4*>(pair_of + (size_t)row * kS2MaxSlots);
#pragma unroll
    for (int q = 0; q < kS2MaxSlots / 4; ++q) nx[q] = pp[q];
  }
  for (; row < n_rows; row += stride) {
    int pk[kS2MaxSlots];
#pragma unroll
    for (int q = 0; q < kS2MaxSlots / 4; ++q) {
      pk[4 * q] = nx[q].x; pk[4 * q + 1] = nx[q].y; pk[4 * q + 2] = nx[q].z; pk[4 * q + 3] = nx[q].w;
    }
    if (row + stride < n_rows) {
      const int4* pp = reinterpret_cast<const int4*>(pair_of + (size_t)(row + stride) * kS2MaxSlots);
#pragma unroll
      for (int q = 0; q < kS2MaxSlots / 4; ++q) nx[q] = pp[q];
    }
    float ls[kS2MaxSlots];
    uint4 v[kS2MaxSlots];
#pragma unroll
    for (int k = 0; k < kS2MaxSlots; ++k) {
      const int p = pk[k];
      ls[k] = p >= 0 ? lse_p[(size_t)p * h + head] : -INFINITY;
      v[k] = p >= 0 ? *reinterpret_cast<const uint4*>(O_p + (size_t)p * h * 64 + cidx * 8) : make_uint4(0, 0, 0, 0);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kS2MaxSlots; ++k) mx = fmaxf(mx, ls[k]);
    float acc[8], den = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k = 0; k < kS2MaxSlots; ++k) {
      if (ls[k] > -INFINITY) {
        const float w = __expf(ls[k] - mx);
        den += w;
        const T* pv = reinterpret_cast<const T*>(&v[k]);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, (float)pv[e], acc[e]);
      }
    }
    const float inv = den > 0.f ? 1.0f / den : 0.f;
    uint4 o;
    T* po = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int e = 0; e < 8; ++e) po[e] = T(acc[e] * inv);
    *reinterpret_cast<uint4*>(O + (size_t)row * h * 64 + cidx * 8) = o;
    if (lse && (cidx & 7) == 0) lse[(size_t)row * h + head] = den > 0.f ? mx + logf(den) : -INFINITY;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 3a. merge + blend (long no-grad prefill): the merge kernel above with the other two branches' chunks and the row's gates (already
//     evaluated, launch_gate_fast on a side stream) folded in: O[row] = g_cmp O_cmp + (g_sel / den) sum_k w_k O_k + g_win O_win.  The
//     selected branch's output never reaches HBM and the separate combine pass (405 MB read + 88 MB written at 64k) disappears.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kS2MergeRows * 64, 2)
sel2_merge_blend_kernel(int n_rows, int h, const int* __restrict__ pair_of, const T* __restrict__ O_p, const float* __restrict__ lse_p,
                        const T* __restrict__ O_cmp, const T* __restrict__ O_win, const float* __restrict__ gates, T* __restrict__ O) {
  const int chunks = h * 8;
  const int rl = threadIdx.x / 64, cidx = threadIdx.x % 64;
  if (cidx >= chunks) return;
  const int head = cidx >> 3;
  const int stride = gridDim.x * kS2MergeRows;
  int4 nx[kS2MaxSlots / 4];
  int row = blockIdx.x * kS2MergeRows + rl;
  if (row < n_rows) {
    const int4* pp = reinterpret_cast<const int4*>(pair_of + (size_t)row * kS2MaxSlots);
#pragma unroll
    for (int q = 0; q < kS2MaxSlots / 4; ++q) nx[q] = pp[q];
  }
  for (; row < n_rows; row += stride) {
    int pk[kS2MaxSlots];
#pragma unroll
    for (int q = 0; q < kS2MaxSlots / 4; ++q) {
      pk[4 * q] = nx[q].x; pk[4 * q + 1] = nx[q].y; pk[4 * q + 2] = nx[q].z; pk[4 * q + 3] = nx[q].w;
    }
    if (row + stride < n_rows) {
      const int4* pp = reinterpret_cast<const int4*>(pair_of + (size_t)(row + stride) * kS2MaxSlots);
#pragma unroll
      for (int q = 0; q < kS2MaxSlots / 4; ++q) nx[q] = pp[q];
    }
    const size_t ro = (size_t)row * h * 64 + cidx * 8;
    const uint4 oc = *reinterpret_cast<const uint4*>(O_cmp + ro);
    const uint4 ow = *reinterpret_cast<const uint4*>(O_win + ro);
    const float g0 = gates[(size_t)row * 3], g1 = gates[(size_t)row * 3 + 1], g2 = gates[(size_t)row * 3 + 2];
    float ls[kS2MaxSlots];
    uint4 v[kS2MaxSlots];
#pragma unroll
    for (int k = 0; k < kS2MaxSlots; ++k) {
      const int p = pk[k];
      ls[k] = p >= 0 ? lse_p[(size_t)p * h + head] : -INFINITY;
      v[k] = p >= 0 ? *reinterpret_cast<const uint4*>(O_p + (size_t)p * h * 64 + cidx * 8) : make_uint4(0, 0, 0, 0);
    }
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kS2MaxSlots; ++k) mx = fmaxf(mx, ls[k]);
    float acc[8], den = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k = 0; k < kS2MaxSlots; ++k) {
      if (ls[k] > -INFINITY) {
        const float w = __expf(ls[k] - mx);
        den += w;
        const T* pv = reinterpret_cast<const T*>(&v[k]);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, (float)pv[e], acc[e]);
      }
    }
    const float sc = den > 0.f ? g1 / den : 0.f;  // empty row -> the selected branch contributes zeros (attention_kernels.py:769-771)
    const T* ec = reinterpret_cast<const T*>(&oc);
    const T* ew = reinterpret_cast<const T*>(&ow);
    uint4 o;
    T* po = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int e = 0; e < 8; ++e) po[e] = T(g0 * (float)ec[e] + sc * acc[e] + g2 * (float)ew[e]);
    *reinterpret_cast<uint4*>(O + ro) = o;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// 3b. merge + gate + combine in one pass (no-grad prefill): O[row] = g_cmp O_cmp + g_sel merge_k(O_k, lse_k) + g_win O_win with the
//     gates of GateMLP(mean_h Q[row]) (nsa_attention.py:32-82, :1356-1398) -- the selected branch's output and the gates never
//     reach HBM, and the separate combine pass (405 MB read + 88 MB written at 64k) disappears.  64 threads per row (one per
//     16-byte chunk, as in the merge kernel); the first warp of a row evaluates the gate from weights staged in shared memory
//     while the row's partial loads are in flight.
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kS2MergeRows * 64)
sel2_merge_combine_kernel(nsa_dims_t dm, const T* __restrict__ Q, nsa_gate_params_t gp, const int* __restrict__ pair_of,
                          const T* __restrict__ O_p, const float* __restrict__ lse_p, const T* __restrict__ O_cmp,
                          const T* __restrict__ O_win, T* __restrict__ O, float* __restrict__ gates) {
  extern __shared__ float smf[];
  const int Dk = dm.Dk, H = dm.gate_hidden, h = dm.h;
  float* w1t = smf;                 // [Dk][H]
  float* b1 = w1t + Dk * H;         // [H]
  float* w2 = b1 + H;               // [3][H]
  float* b2 = w2 + 3 * H;           // [4]
  float* qg = b2 + 4;               // [warps][Dk]
  const bool mlp = dm.gate_mode == NSA_GATE_MLP;
  if (mlp) {
    for (int i = threadIdx.x; i < Dk * H; i += blockDim.x) w1t[(i % Dk) * H + i / Dk] = gp.fc1_w[i];
    for (int i = threadIdx.x; i < H; i += blockDim.x) b1[i] = gp.fc1_b ? gp.fc1_b[i] : 0.f;
    for (int i = threadIdx.x; i < 3 * H; i += blockDim.x) w2[i] = gp.fc2_w[i];
    if (threadIdx.x < 3) b2[threadIdx.x] = gp.fc2_b ? gp.fc2_b[threadIdx.x] : 0.f;
  }
  __syncthreads();
  const int n_rows = dm.B * dm.S * dm.G;
  const int chunks = h * 8;                   // 16-byte chunks per row (Dv = 64), <= 64 threads per row
  const int rl = threadIdx.x / 64, cidx = threadIdx.x % 64, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* qgp = qg + warp * Dk;
  const float inv_tau = 1.0f / fmaxf(dm.gate_tau, 1e-6f);
  for (int row = blockIdx.x * kS2MergeRows + rl; row < n_rows; row += gridDim.x * kS2MergeRows) {  // warp-uniform trip count
    const bool act = cidx < chunks;
    const int head = cidx >> 3;
    // ---- requests that depend on nothing: slot table, the other two branches' chunks ----
    int pk[kS2MaxSlots];
    uint4 oc = make_uint4(0, 0, 0, 0), ow = make_uint4(0, 0, 0, 0);
    if (act) {
      const int4* pp = reinterpret_cast<const int4*>(pair_of + (size_t)row * kS2MaxSlots);
#pragma unroll
      for (int q = 0; q < kS2MaxSlots / 4; ++q) {
        const int4 t = pp[q];
        pk[4 * q] = t.x; pk[4 * q + 1] = t.y; pk[4 * q + 2] = t.z; pk[4 * q + 3] = t.w;
      }
      oc = *reinterpret_cast<const uint4*>(O_cmp + (size_t)row * h * 64 + cidx * 8);
      ow = *reinterpret_cast<const uint4*>(O_win + (size_t)row * h * 64 + cidx * 8);
    } else {
#pragma unroll
      for (int k = 0; k < kS2MaxSlots; ++k) pk[k] = -1;
    }
    // ---- q_gp = mean over heads (nsa_attention.py:1357): each warp of the row builds its own copy (lane owns dims 2*lane, +1) ----
    if (mlp) {
      const T* qrow = Q + (size_t)row * h * Dk;
      for (int k = 2 * lane; k < Dk; k += 64) {
        float m0 = 0.f, m1 = 0.f;
        for (int h0 = 0; h0 < h; h0 += 8) {
          uint32_t v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = h0 + u < h ? *reinterpret_cast<const uint32_t*>(qrow + (h0 + u) * Dk + k) : 0u;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            m0 += (float)reinterpret_cast<const T*>(&v[u])[0];
            m1 += (float)reinterpret_cast<const T*>(&v[u])[1];
          }
        }
        qgp[k] = m0 / (float)h;
        qgp[k + 1] = m1 / (float)h;
      }
      __syncwarp();
    }
    // ---- second round trip: the row's lse values and the first eight partial chunks (the slot table has arrived by now) ----
    float ls[kS2MaxSlots];
#pragma unroll
    for (int k = 0; k < kS2MaxSlots; ++k) ls[k] = pk[k] >= 0 ? lse_p[(size_t)pk[k] * h + head] : -INFINITY;
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k)
      v[k] = pk[k] >= 0 ? *reinterpret_cast<const uint4*>(O_p + (size_t)pk[k] * h * 64 + cidx * 8) : make_uint4(0, 0, 0, 0);
    // ---- gate MLP while those loads fly (both warps of a row evaluate it: no barrier between them) ----
    float g0 = 1.0f / 3.0f, g1 = 1.0f / 3.0f, g2 = 1.0f / 3.0f;
    if (dm.gate_mode == NSA_GATE_CMP) { g0 = 1.f; g1 = 0.f; g2 = 0.f; }
    else if (dm.gate_mode == NSA_GATE_SEL) { g0 = 0.f; g1 = 1.f; g2 = 0.f; }
    else if (dm.gate_mode == NSA_GATE_WIN) { g0 = 0.f; g1 = 0.f; g2 = 1.f; }
    else if (mlp) {
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
      const float mxg = fmaxf(l0, fmaxf(l1, l2));
      const int am = l0 >= l1 ? (l0 >= l2 ? 0 : 2) : (l1 >= l2 ? 1 : 2);  // first maximum
      const float second = am == 0 ? fmaxf(l1, l2) : (am == 1 ? fmaxf(l0, l2) : fmaxf(l0, l1));
      if (mxg - second > 50.0f) {  // hard one-hot (nsa_attention.py:74-81)
        g0 = am == 0 ? 1.f : 0.f; g1 = am == 1 ? 1.f : 0.f; g2 = am == 2 ? 1.f : 0.f;
      } else {
        const float e0 = expf(l0 - mxg), e1 = expf(l1 - mxg), e2 = expf(l2 - mxg);
        const float inv = 1.0f / (e0 + e1 + e2);
        g0 = e0 * inv; g1 = e1 * inv; g2 = e2 * inv;
      }
      __syncwarp();  // qgp is rewritten for the next row
    }
    if (gates && cidx == 0) {
      gates[(size_t)row * 3] = g0;
      gates[(size_t)row * 3 + 1] = g1;
      gates[(size_t)row * 3 + 2] = g2;
    }
    if (!act) continue;
    // ---- merge the row's partials in slot order (deterministic) ----
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < kS2MaxSlots; ++k) mx = fmaxf(mx, ls[k]);
    float acc[8], den = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
    for (int k0 = 0; k0 < kS2MaxSlots; k0 += 8) {
      if (k0 > 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k)
          v[k] = pk[k0 + k] >= 0 ? *reinterpret_cast<const uint4*>(O_p + (size_t)pk[k0 + k] * h * 64 + cidx * 8) : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (ls[k0 + k] > -INFINITY) {
          const float w = __expf(ls[k0 + k] - mx);
          den += w;
          const T* pv = reinterpret_cast<const T*>(&v[k]);
#pragma unroll
          for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, (float)pv[e], acc[e]);
        }
      }
    }
    const float sc = den > 0.f ? g1 / den : 0.f;  // empty row -> the selected branch contributes zeros (attention_kernels.py:769-771)
    const T* ec = reinterpret_cast<const T*>(&oc);
    const T* ew = reinterpret_cast<const T*>(&ow);
    uint4 o;
    T* po = reinterpret_cast<T*>(&o);
#pragma unroll
    for (int e = 0; e < 8; ++e) po[e] = T(g0 * (float)ec[e] + sc * acc[e] + g2 * (float)ew[e]);
    *reinterpret_cast<uint4*>(O + (size_t)row * h * 64 + cidx * 8) = o;
  }
}

// ---- host ------------------------------------------------------------------------------------------------------
int make_tmap_q_heads(CUtensorMap* out, const void* base, int dtype, int D, int h, int G, long long n_tokens, int box_tokens);

static S2Geom s2_geom(const nsa_dims_t& dm) {
  S2Geom gm;
  gm.B = dm.B; gm.S = dm.S; gm.G = dm.G; gm.n_ranges = dm.n_ranges; gm.S_kv = dm.S_sel_kv;
  gm.NB = (dm.S_sel_kv + 63) / 64;
  gm.t0 = dm.t0;
  gm.tokp = 128 / dm.h > 32 ? 32 : 128 / dm.h;  // one producer lane per query
  return gm;
}

// workspace carve-up (bytes, all 256-aligned); the index comes first so the backward can use it without the partials
struct S2Ws {
  size_t counts, offs, cursors, runs, n_runs, tok, hi, index_total, pair_of, lse_p, O_p, total;
  int max_pairs, max_runs;
};

static S2Ws s2_ws(const nsa_dims_t& dm) {
  const S2Geom gm = s2_geom(dm);
  S2Ws w;
  const long long rows = (long long)dm.B * dm.S * dm.G;
  const long long nlists = (long long)dm.B * dm.G * gm.NB;
  const long long max_tiles = (rows * kS2MaxSlots + gm.tokp - 1) / gm.tokp + nlists;  // every list padded by < one tile
  w.max_pairs = (int)(max_tiles * gm.tokp);
  w.max_runs = (int)(max_tiles / kS2Run + nlists + 1);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t at = o; o += (bytes + 255) & ~(size_t)255; return at; };
  w.counts = take(nlists * 4);
  w.offs = take((nlists + 1) * 4);
  w.cursors = take(nlists * 4);
  w.runs = take((size_t)w.max_runs * sizeof(S2Run));
  w.n_runs = take(4);
  w.tok = take((size_t)w.max_pairs * 4);
  w.hi = take((size_t)w.max_pairs * 4);
  w.index_total = o;
  w.pair_of = take((size_t)rows * kS2MaxSlots * 4);
  w.lse_p = take((size_t)w.max_pairs * dm.h * 4);
  w.O_p = take((size_t)w.max_pairs * dm.h * 64 * 2);
  w.total = o;
  return w;
}

bool sel2_index_supported(const nsa_dims_t& dm) {
  if (dm.h < 1 || dm.h > 128) return false;
  if (dm.l_sel % 64 != 0 || (long long)dm.n_sel * dm.l_sel > 64 * kS2MaxSlots || dm.n_ranges < 1 || dm.S_sel_kv < 1) return false;
  const long long rows = (long long)dm.B * dm.S * dm.G;
  const long long nb = (dm.S_sel_kv + 63) / 64;
  if (dm.S * dm.G < kS2IdxThreads) return false;            // an index CTA (256 rows) may touch at most two batches
  if (2 * dm.G * nb * 2 * 4 > 200 * 1024) return false;     // shared-memory histograms of the index kernels
  if (rows * kS2MaxSlots + (long long)dm.B * dm.G * nb * 64 >= (1LL << 31)) return false;
  return true;
}

bool tc_sel2_supported(const nsa_dims_t& dm) {
  if (dm.impl == NSA_IMPL_SIMT) return false;
  if (!((dm.dtype == NSA_BF16 || dm.dtype == NSA_F16) && dm.Dk == 64 && dm.Dv == 64 && dm.h >= 1 && dm.h <= 8)) return false;  // merge: h*8 <= 64 threads per row
  return sel2_index_supported(dm);
}

int64_t tc_sel2_workspace(const nsa_dims_t& dm) { return tc_sel2_supported(dm) ? (int64_t)s2_ws(dm).total : 0; }
int64_t sel2_index_workspace(const nsa_dims_t& dm) { return sel2_index_supported(dm) ? (int64_t)s2_ws(dm).index_total : 0; }

// ranges -> block lists + run table.  pair_of [rows][16] is written only when the caller merges partials (forward).
static int s2_build(const nsa_dims_t& dm, const int32_t* ranges, char* ws, const S2Ws& w, bool with_pair_of, cudaStream_t stream) {
  const S2Geom gm = s2_geom(dm);
  int* counts = reinterpret_cast<int*>(ws + w.counts);
  int* offs = reinterpret_cast<int*>(ws + w.offs);
  int* cursors = reinterpret_cast<int*>(ws + w.cursors);
  S2Run* runs = reinterpret_cast<S2Run*>(ws + w.runs);
  int* n_runs = reinterpret_cast<int*>(ws + w.n_runs);
  int* tok = reinterpret_cast<int*>(ws + w.tok);
  int* hi = reinterpret_cast<int*>(ws + w.hi);
  int* pair_of = with_pair_of ? reinterpret_cast<int*>(ws + w.pair_of) : nullptr;
  const int n_rows = dm.B * dm.S * dm.G;
  const int nlists = dm.B * dm.G * gm.NB;
  cudaError_t me = cudaMemsetAsync(counts, 0, (size_t)nlists * 4, stream);
  if (me == cudaSuccess) me = cudaMemsetAsync(tok, 0xff, (size_t)w.max_pairs * 4, stream);  // -1 = padding pair
  if (me != cudaSuccess) { set_error("sel2 index: cudaMemsetAsync: %s", cudaGetErrorString(me)); return NSA_ERR_CUDA; }
  const int idx_blocks = ceil_div(n_rows, kS2IdxThreads);
  const size_t hist_bytes = (size_t)2 * gm.G * gm.NB * 4;
  static std::atomic<unsigned long long> cnt_done{0}, fill_done{0};
  if (int rc = ensure_smem_attr(sel2_count_kernel, 200 * 1024, cnt_done, "sel2 count")) return rc;
  if (int rc = ensure_smem_attr(sel2_fill_kernel, 200 * 1024, fill_done, "sel2 fill")) return rc;
  sel2_count_kernel<<<idx_blocks, kS2IdxThreads, hist_bytes, stream>>>(gm, ranges, counts);
  if (int rc = check_launch("sel2_count_kernel")) return rc;
  sel2_scan_kernel<<<1, 1024, 0, stream>>>(gm, counts, offs, cursors, runs, n_runs);
  if (int rc = check_launch("sel2_scan_kernel")) return rc;
  sel2_fill_kernel<<<idx_blocks, kS2IdxThreads, 2 * hist_bytes, stream>>>(gm, ranges, offs, cursors, tok, hi, pair_of);
  return check_launch("sel2_fill_kernel");
}

int sel2_build_index(const nsa_dims_t& dm, const int32_t* ranges, void* workspace, cudaStream_t stream, S2Index* out) {
  NSA_REQUIRE(sel2_index_supported(dm), "sel2 index: shape not supported");
  const S2Ws w = s2_ws(dm);
  char* ws = reinterpret_cast<char*>(workspace);
  if (int rc = s2_build(dm, ranges, ws, w, false, stream)) return rc;
  out->gm = s2_geom(dm);
  out->runs = reinterpret_cast<const S2Run*>(ws + w.runs);
  out->n_runs = reinterpret_cast<const int*>(ws + w.n_runs);
  out->tok = reinterpret_cast<const int*>(ws + w.tok);
  out->hi = reinterpret_cast<const int*>(ws + w.hi);
  out->max_runs = w.max_runs;
  return NSA_OK;
}

template <typename T>
static int launch_sel2_t(const nsa_dims_t& dm, const void* Q, const void* K, const void* V, const int32_t* ranges, void* O,
                         float* lse, void* workspace, cudaStream_t stream, const Sel2Fuse* fuse) {
  const S2Geom gm = s2_geom(dm);
  const S2Ws w = s2_ws(dm);
  char* ws = reinterpret_cast<char*>(workspace);
  S2Run* runs = reinterpret_cast<S2Run*>(ws + w.runs);
  int* n_runs = reinterpret_cast<int*>(ws + w.n_runs);
  int* tok = reinterpret_cast<int*>(ws + w.tok);
  int* hi = reinterpret_cast<int*>(ws + w.hi);
  int* pair_of = reinterpret_cast<int*>(ws + w.pair_of);
  float* lse_p = reinterpret_cast<float*>(ws + w.lse_p);
  T* O_p = reinterpret_cast<T*>(ws + w.O_p);
  const int n_rows = dm.B * dm.S * dm.G;
  if (int rc = s2_build(dm, ranges, ws, w, true, stream)) return rc;

  CUtensorMap tmQ, tmK, tmV, tmO;
  const int slabs = dm.B * dm.G;
  if (int rc = make_tmap_q_heads(&tmQ, Q, dm.dtype, 64, dm.h, dm.G, (long long)dm.B * dm.S, 1)) return rc;
  if (int rc = make_tmap_rows(&tmK, K, dm.dtype, 64, dm.S_sel_kv, 64, (long long)dm.cap_sel * 64, slabs, 64)) return rc;
  if (int rc = make_tmap_rows(&tmV, V, dm.dtype, 64, dm.S_sel_kv, 64, (long long)dm.cap_sel * 64, slabs, 64)) return rc;
  if (int rc = make_tmap_tiles(&tmO, O_p, dm.dtype, 64, gm.tokp * dm.h, w.max_pairs / gm.tokp, 32)) return rc;
  auto kern = sel2_attn_kernel<T>;
  // two CTAs per SM (S2Smem::total): needs the largest carve-out -- with the driver's default this kernel ran one CTA per SM
  static std::atomic<unsigned long long> attr_done{0};
  if (int rc = ensure_smem_attr(kern, S2Smem::total, attr_done, "sel2", true)) return rc;
  long long* dbg_buf = nullptr;
#ifdef NSA_SEL2_DBG
  static long long* dbg_static = nullptr;
  if (!dbg_static) cudaMalloc(&dbg_static, 4000 * sizeof(long long));
  cudaMemsetAsync(dbg_static, 0, 4000 * sizeof(long long), stream);
  dbg_buf = dbg_static;
#endif
  kern<<<w.max_runs, 320, S2Smem::total, stream>>>(tmQ, tmK, tmV, tmO, dm, gm, runs, n_runs, tok, hi, lse_p, dbg_buf);
  if (int rc = check_launch("sel2_attn_kernel")) return rc;
#ifdef NSA_SEL2_DBG
  {
    static int dumps = 0;
    cudaStreamSynchronize(stream);
    if (dumps++ == 2) {
      static long long host[4000];
      cudaMemcpy(host, dbg_buf, sizeof(host), cudaMemcpyDeviceToHost);
      for (int tag = 0; tag < 40; ++tag)
        for (int it = 0; it < 100; ++it)
          if (host[tag * 100 + it]) fprintf(stderr, "BDBG %d %d %lld\n", tag, it, host[tag * 100 + it]);
    }
  }
#endif
  int mblocks = ceil_div(n_rows, kS2MergeRows);
  if (mblocks > 148 * 8) mblocks = 148 * 8;  // two resident CTAs per SM, several rows each: the slot-table prefetch needs a next row
  if (fuse && fuse->gates_in) {  // gates already evaluated: merge + blend
    sel2_merge_blend_kernel<T><<<mblocks, kS2MergeRows * 64, 0, stream>>>(n_rows, dm.h, pair_of, O_p, lse_p, (const T*)fuse->O_cmp,
                                                                        (const T*)fuse->O_win, fuse->gates_in, (T*)fuse->O);
    return check_launch("sel2_merge_blend_kernel");
  }
  if (fuse) {  // no-grad prefill: merge + gate + combine in one pass, the selected branch's output never reaches HBM
    const int Hh = dm.gate_mode == NSA_GATE_MLP ? dm.gate_hidden : 0;
    const size_t smem = ((size_t)dm.Dk * Hh + Hh + 3 * Hh + 4 + (size_t)kS2MergeRows * 2 * dm.Dk) * sizeof(float);
    sel2_merge_combine_kernel<T><<<mblocks, kS2MergeRows * 64, smem, stream>>>(dm, (const T*)Q, *fuse->gp, pair_of, O_p, lse_p,
                                                                             (const T*)fuse->O_cmp, (const T*)fuse->O_win, (T*)fuse->O,
                                                                             fuse->gates);
    return check_launch("sel2_merge_combine_kernel");
  }
  sel2_merge_kernel<T><<<mblocks, kS2MergeRows * 64, 0, stream>>>(n_rows, dm.h, pair_of, O_p, lse_p, (T*)O, lse);
  return check_launch("sel2_merge_kernel");
}

bool sel2_fuse_supported(const nsa_dims_t& dm) {
  const int Hh = dm.gate_mode == NSA_GATE_MLP ? dm.gate_hidden : 0;
  return tc_sel2_supported(dm) && dm.Dk == 64 && dm.Dk % 4 == 0 &&
         ((size_t)dm.Dk * Hh + 4 * Hh + 4 + (size_t)kS2MergeRows * 2 * dm.Dk) * sizeof(float) <= 48 * 1024 && ((uintptr_t)0 == 0);
}

int launch_sel2_tc(const nsa_dims_t& dm, const void* Q, const void* K, const void* V, const int32_t* ranges, void* O, float* lse,
                   void* workspace, cudaStream_t stream, const Sel2Fuse* fuse) {
  static_assert(sizeof(S2Misc) <= 256, "S2Misc must fit its slot");
  static_assert(2 * (S2Smem::total + 1024) <= 228 * 1024, "two sel2 CTAs must fit one SM");
  if (dm.B * dm.S * dm.G == 0) return NSA_OK;
  NSA_REQUIRE(workspace, "sel2: needs a workspace of nsa_workspace_bytes(NSA_WS_SEL) bytes");
  NSA_REQUIRE(!fuse || sel2_fuse_supported(dm), "sel2: fused merge + combine does not serve this shape");
  if (dm.dtype == NSA_BF16) return launch_sel2_t<__nv_bfloat16>(dm, Q, K, V, ranges, O, lse, workspace, stream, fuse);
  return launch_sel2_t<__half>(dm, Q, K, V, ranges, O, lse, workspace, stream, fuse);
}

}  // namespace nsa
