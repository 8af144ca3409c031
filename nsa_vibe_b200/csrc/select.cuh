// Warp-level top-n block selection and range merging (NSA Eq.11-12 as the reference implements it).
//
// One warp owns one (b, t, g) row of p_grp.  Restates, for one row:
//   prefill: select_topn_ranges_batched + convert_indices_to_ranges_batched_v2
//            (nsa/core/selection_scorer.py:255-362, :434-605)
//   decode : select_topn_ranges (selection_scorer.py:124-249)
// The selected block ids are kept as a bitmap spread over the warp, so "sort, drop duplicates, merge
// adjacent blocks" is simply "find runs of set bits".
#pragma once
#include "common.cuh"

namespace nsa {

constexpr int kSelMaxWords = 4;  // bitmap words per lane -> S_sel <= 4096 blocks (262144 tokens at l_sel=64)

__host__ __device__ inline int prefill_forced_cols(int S_total, int l_sel) {
  // the reference drops duplicate forced COLUMNS (unique_consecutive(dim=-1), selection_scorer.py:299-300)
  return S_total <= l_sel ? 1 : (S_total <= 2 * l_sel ? 2 : 3);
}

__host__ __device__ inline int prefill_range_cols(int S_total, int l_sel, int n_sel) {
  int S_sel = S_total <= 0 ? 0 : (S_total + l_sel - 1) / l_sel;
  if (n_sel >= S_sel) return S_sel;
  int nf = prefill_forced_cols(S_total, l_sel);
  int k_rest = n_sel - nf > 0 ? n_sel - nf : 0;
  if (k_rest > 0) return nf + (k_rest < S_sel ? k_rest : S_sel);
  return nf < n_sel ? nf : n_sel;
}

// Forced blocks of a row as a code: bits [0,2) = count n <= 3, slot i at bits [2+4i, 6+4i): 15 = block 0 (force_init), k = the
// local block max(cb - k, 0) with cb = t / l_sel (force_local).  Slots are in ascending block order.
//   decode rule (select_topn_ranges, selection_scorer.py:159-170): [0 if force_init] + [cb - k for k < force_local], no dedup;
//   prefill rule (select_topn_ranges_batched, :283-300): the same columns sorted per row, then COLUMNS that are identical for
//   every row of the S_total-row tensor collapse (unique_consecutive(dim=-1)): a column is all-zero iff max_cb <= k.
__host__ __device__ inline int forced_code(int mode, int S_total, int l_sel, int force_init, int force_local) {
  int codes[3] = {0, 0, 0};
  int n = 0;
  if (mode == 1) {
    if (force_init) codes[n++] = 15;
    for (int k = 0; k < force_local && n < 3; ++k) codes[n++] = k;
  } else {
    const int max_cb = S_total > 0 ? (S_total - 1) / l_sel : 0;
    int nz = 0;
    for (int k = 0; k < force_local; ++k) nz += max_cb > k ? 1 : 0;
    if ((force_init ? 1 : 0) + (force_local - nz) > 0) codes[n++] = 15;
    for (int k = force_local - 1; k >= 0; --k)
      if (max_cb > k && n < 3) codes[n++] = k;
  }
  int c = n;
  for (int i = 0; i < n; ++i) c |= codes[i] << (2 + 4 * i);
  return c;
}
__host__ __device__ inline int forced_code_default(int mode, int S_total, int l_sel) { return forced_code(mode, S_total, l_sel, 1, 2); }
// forced_code(1, *, *, 1, 2) = {block 0, cb, cb - 1}: the decode rule's default, as a constant for kernels that select in place
constexpr int kForcedDecodeDefault = 3 | (15 << 2) | (0 << 6) | (1 << 10);

__host__ __device__ inline int prefill_range_cols_ex(int S_total, int l_sel, int n_sel, int force_init, int force_local) {
  int S_sel = S_total <= 0 ? 0 : (S_total + l_sel - 1) / l_sel;
  if (n_sel >= S_sel) return S_sel;
  int nf = forced_code(0, S_total, l_sel, force_init, force_local) & 3;
  int k_rest = n_sel - nf > 0 ? n_sel - nf : 0;
  if (k_rest > 0) return nf + (k_rest < S_sel ? k_rest : S_sel);
  return nf < n_sel ? nf : n_sel;
}

__device__ __forceinline__ int forced_decode(int fcode, int cb, int (&forced)[3]) {
  const int n = fcode & 3;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int code = (fcode >> (2 + 4 * i)) & 15;
    int v = code == 15 ? 0 : cb - code;
    if (v < 0) v = 0;
    forced[i] = i < n ? v : -1;
  }
  return n;
}

__device__ __forceinline__ void bitmap_set(uint32_t (&bm)[kSelMaxWords], int lane, int j) {
  int word = j >> 5;
  if ((word & 31) == lane) bm[word >> 5] |= 1u << (j & 31);
}

// ---- runs of set bits -> [start,end) ranges (sort, dedup and merge of adjacent blocks in one step) ----------------
__device__ inline void sel_bitmap_to_ranges(const uint32_t (&bm)[kSelMaxWords], int S_sel, int l_sel, int K, int t,
                                            int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  int nstart_before = 0, nend_before = 0;
  const int nwords = (S_sel + 31) >> 5;
  uint32_t prev_hi = 0;  // bit 31 of lane 31's word in the previous slot
#pragma unroll
  for (int s = 0; s < kSelMaxWords; ++s) {
    if (s * 32 >= nwords) break;
    const uint32_t b = bm[s];
    uint32_t left = __shfl_up_sync(0xffffffffu, b, 1);
    uint32_t carry_in = (lane == 0 ? prev_hi : left) >> 31;
    uint32_t right = __shfl_down_sync(0xffffffffu, b, 1);
    uint32_t next_first = (s + 1 < kSelMaxWords) ? bm[s + 1] : 0u;
    next_first = __shfl_sync(0xffffffffu, next_first, 0);
    uint32_t carry_out = (lane == 31 ? next_first : right) & 1u;
    uint32_t starts = b & ~((b << 1) | carry_in);
    uint32_t ends = b & ~((b >> 1) | (carry_out << 31));
    int ns = __popc(starts), ne = __popc(ends);
    // exclusive prefix over lanes
    int ps = ns, pe = ne;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int a = __shfl_up_sync(0xffffffffu, ps, o);
      int c = __shfl_up_sync(0xffffffffu, pe, o);
      if (lane >= o) { ps += a; pe += c; }
    }
    int is = nstart_before + ps - ns, ie = nend_before + pe - ne;
    const int base_id = (s * 32 + lane) * 32;
    while (starts) {
      int bit = __ffs(starts) - 1;
      starts &= starts - 1;
      if (is < K) out[2 * is] = (base_id + bit) * l_sel;
      ++is;
    }
    while (ends) {
      int bit = __ffs(ends) - 1;
      ends &= ends - 1;
      int e = (base_id + bit + 1) * l_sel;
      if (e > t + 1) e = t + 1;  // clamp to the causal limit (:242-246, :567-573)
      if (ie < K) out[2 * ie + 1] = e;
      ++ie;
    }
    nstart_before += __shfl_sync(0xffffffffu, ps, 31);
    nend_before += __shfl_sync(0xffffffffu, pe, 31);
    prev_hi = __shfl_sync(0xffffffffu, b, 31);
  }
  for (int i = nstart_before + lane; i < K; i += 32) {
    out[2 * i] = 0;
    out[2 * i + 1] = 0;
  }
}

// `sc`: this warp's copy of the row (S_sel floats in shared memory, clobbered).
// mode 0 = prefill rule, 1 = decode rule; nf = forced_code(mode, S_total, l_sel, force_init, force_local).
// Writes out[K][2] (int32 token ranges, [0,0] padded).
__device__ inline void select_row_warp(float* sc, int S_sel, int l_sel, int n_sel, int mode, int nf, int K, int t,
                                       int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const float NEG = -INFINITY;
  uint32_t bm[kSelMaxWords];
#pragma unroll
  for (int i = 0; i < kSelMaxWords; ++i) bm[i] = 0u;

  int nvalid = (t + 1) / l_sel;
  if (nvalid > S_sel) nvalid = S_sel;
  const int cb = t / l_sel;

  if (mode == 0 && n_sel >= S_sel) {
    // selection_scorer.py:348-354: n_top >= S_sel -> exactly all complete blocks [0, nvalid).
    // bitmap word w (ids 32w..32w+31) is owned by lane w % 32, slot w / 32.
    for (int w = lane; w * 32 < nvalid; w += 32) {
      int rem = nvalid - w * 32;
      bm[w >> 5] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    }
  } else {
    // forced ids (sorted ascending): prefill keeps nf distinct columns, decode always three
    int forced[3];
    const int nfc = forced_decode(nf, cb, forced);  // nf = forced_code(...)
    const int k_rest = n_sel - nfc > 0 ? n_sel - nfc : 0;
    // forced members of the selection
    int take = nfc;
    if (mode == 0 && k_rest == 0) take = nfc < n_sel ? nfc : n_sel;  // forced[..., :n_top]
    for (int i = 0; i < take; ++i) {
      int j = forced[i];
      bool ok = mode == 0 ? (j < nvalid) : (j < S_sel);  // prefill drops incomplete blocks (:343-347)
      if (ok) bitmap_set(bm, lane, j);
    }
    if (k_rest > 0) {
      const int k_act = k_rest < S_sel ? k_rest : S_sel;
      // composite = fp32(p) - fp32(fp32(j) * 1e-8f): two separately rounded fp32 ops (:182-184, :312-318)
      float best = NEG;
      int best_j = 0x7fffffff;
      for (int j = lane; j < S_sel; j += 32) {
        float c = NEG;
        if (j < nvalid && j != forced[0] && j != forced[1] && j != forced[2])
          c = __fsub_rn(sc[j], __fmul_rn((float)j, 1e-8f));
        sc[j] = c;
        if (c > best) { best = c; best_j = j; }  // ascending j: first maximum = lowest index
      }
      if (S_sel <= 128) {
        // Few candidates (decode, short prefill): rank every candidate against all others instead of k rounds of warp
        // arg-max.  rank_j = #{i : c_i > c_j or (c_i == c_j and i < j)}; j is picked iff rank_j < k and c_j > -inf --
        // the same set the rounds below produce, with no serial dependence between picks.
        __syncwarp();
        for (int s = 0; s * 32 < S_sel; ++s) {
          const int j = s * 32 + lane;
          const float c = j < S_sel ? sc[j] : NEG;
          int rank = 0;
          for (int i = 0; i < S_sel; ++i) {
            const float ci = sc[i];
            rank += (ci > c || (ci == c && i < j)) ? 1 : 0;
          }
          const uint32_t word = __ballot_sync(0xffffffffu, c > NEG && rank < k_act);
          if ((s & 31) == lane) bm[s >> 5] |= word;
        }
      } else if (S_sel <= 1024) {
        // <= 32 candidates per lane in 4 groups of 8 (group = 8 consecutive strides): the group maxima live in registers, a
        // round is two warp reductions (redux max on an order-preserving key, redux min on the index for the lower-index
        // tie rule) and the owner lane rescans only the 8 entries of the group it took from.
        __syncwarp();
        float gv[4];
        int gj[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          gv[g] = NEG;
          gj[g] = 0x7fffffff;
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) {
            const int j = lane + 32 * (8 * g + kk);
            if (j < S_sel) {
              const float c = sc[j];
              if (c > gv[g]) { gv[g] = c; gj[g] = j; }
            }
          }
        }
        for (int it = 0; it < k_act; ++it) {
          float bv = gv[0];
          int bj = gj[0];
#pragma unroll
          for (int g = 1; g < 4; ++g)
            if (gv[g] > bv) { bv = gv[g]; bj = gj[g]; }  // ascending j across groups: strict > keeps the lower index
          const uint32_t bits = __float_as_uint(bv);
          const uint32_t key = bits ^ ((bits >> 31) ? 0xffffffffu : 0x80000000u);  // order-preserving
          const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
          if (kmax == (0xff800000u ^ 0xffffffffu)) break;  // only -inf left: invalid picks are dropped in both modes
          const int vj = (int)__reduce_min_sync(0xffffffffu, key == kmax ? (uint32_t)bj : 0x7fffffffu);
          bitmap_set(bm, lane, vj);
          if ((vj & 31) == lane) {
            sc[vj] = NEG;
            const int g = vj >> 8;  // (vj / 32) / 8
            float v = NEG;
            int vi = 0x7fffffff;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const int j = lane + 32 * (8 * g + kk);
              if (j < S_sel) {
                const float c = sc[j];
                if (c > v) { v = c; vi = j; }
              }
            }
#pragma unroll
            for (int g2 = 0; g2 < 4; ++g2)
              if (g2 == g) { gv[g2] = v; gj[g2] = vi; }
          }
        }
      } else
      for (int it = 0; it < k_act; ++it) {
        float v = best;
        int vj = best_j;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          float ov = __shfl_xor_sync(0xffffffffu, v, o);
          int oj = __shfl_xor_sync(0xffffffffu, vj, o);
          if (ov > v || (ov == v && oj < vj)) { v = ov; vj = oj; }
        }
        if (!(v > NEG)) break;  // only -inf left: invalid picks are dropped in both modes
        bitmap_set(bm, lane, vj);
        if ((vj & 31) == lane) {  // owner rescans its strided elements
          sc[vj] = NEG;
          best = NEG;
          best_j = 0x7fffffff;
          for (int j = lane; j < S_sel; j += 32) {
            float c = sc[j];
            if (c > best) { best = c; best_j = j; }
          }
        }
      }
    }
  }

  sel_bitmap_to_ranges(bm, S_sel, l_sel, K, t, out);
}

// Threshold form of the k_act picks among <= 1024 candidates held 32 per lane (raw[k] = p of block j = lane + 32 k; replaced by
// the composites).  T = the k_act-th largest LANE maximum (k_act rounds of one warp reduction, no rescans) is a lower bound of the
// k_act-th largest composite, so the picks are among the elements >= T -- about 17 of 1024 for k_act = 13, whatever the row
// length.  Those are compacted into shared memory, every candidate is ranked against the others under the rounds' order
// (larger composite first, lower block id on equal composites) and ranks < k_act are the picks: the same set as k_act rounds of
// arg-max.  Returns false (nothing written to bm) when the row has to take the rounds: T = -inf or more than 32 candidates.
// Needs 97 floats of `sc`.  The rounds cost ~2700 warp instructions per row and made the kernel issue-bound (82 % issue-active,
// 0.38 ms for the 131072 rows of a 64k sequence).
__device__ __forceinline__ bool select_threshold_1024(float (&raw)[32], float* sc, int nvalid, const int (&forced)[3], int k_act,
                                                      uint32_t (&bm)[kSelMaxWords]) {
  const int lane = threadIdx.x & 31;
  const float NEG = -INFINITY;
  // bit k of vm: block lane + 32 k is a candidate (complete and not forced)
  const int nv_l = nvalid > lane ? (nvalid - lane + 31) >> 5 : 0;
  uint32_t vm = nv_l >= 32 ? 0xffffffffu : ((1u << nv_l) - 1u);
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (forced[i] >= 0 && forced[i] < 1024 && (forced[i] & 31) == lane) vm &= ~(1u << (forced[i] >> 5));
  const float lanef = (float)lane;
  float lm = NEG;
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    // composite = fp32(p) - fp32(fp32(j) * 1e-8f): two separately rounded fp32 ops; fp32(j) = lane + 32 k exactly
    const float v = __fsub_rn(raw[k], __fmul_rn(lanef + 32.f * k, 1e-8f));
    raw[k] = ((vm >> k) & 1u) ? v : NEG;
    lm = fmaxf(lm, raw[k]);
  }
  const uint32_t lb = __float_as_uint(lm);
  uint32_t key = lb ^ ((lb >> 31) ? 0xffffffffu : 0x80000000u);  // order-preserving; -inf -> 0x007fffff
  uint32_t tkey = 0u;
  for (int it = 0; it < k_act; ++it) {  // equal lane maxima leave together: T only gets lower, never wrong
    tkey = __reduce_max_sync(0xffffffffu, key);
    if (key == tkey) key = 0u;
  }
  if (tkey <= 0x007fffffu) return false;  // fewer than k_act lanes with a candidate
  const float T = __uint_as_float(tkey ^ ((tkey >> 31) ? 0x80000000u : 0xffffffffu));
  float* cand_c = sc;
  int* cand_j = reinterpret_cast<int*>(sc + 32);
  uint32_t* sbm = reinterpret_cast<uint32_t*>(sc + 64);
  int* cnt = reinterpret_cast<int*>(sc + 96);
  sbm[lane] = bm[0];  // the forced members (ids < 1024: word = lane, slot 0)
  if (lane == 0) *cnt = 0;
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    if (raw[k] >= T) {
      const int pos = atomicAdd(cnt, 1);
      if (pos < 32) {
        cand_c[pos] = raw[k];
        cand_j[pos] = lane + 32 * k;
      }
    }
  }
  __syncwarp();
  const int n = *cnt;
  if (n > 32) return false;
  if (lane < n) {
    const float c = cand_c[lane];
    const int j = cand_j[lane];
    int rank = 0;
#pragma unroll 4
    for (int q = 0; q < n; ++q) {
      const float cq = cand_c[q];
      const int jq = cand_j[q];
      rank += (cq > c || (cq == c && jq < j)) ? 1 : 0;
    }
    if (rank < k_act) atomicOr(&sbm[j >> 5], 1u << (j & 31));
  }
  __syncwarp();
  bm[0] = sbm[lane];
  return true;
}

// Fast path of the standalone kernel for 128 < S_sel <= 1024 (long prefill): the row goes from global memory straight into
// registers (lane owns blocks j = lane + 32k, all loads in flight at once), composites are formed there and written to shared
// memory once (only the owner of a picked block ever re-reads them), and every (lane, group of 8 strides) bucket keeps its best
// AND second-best entry, so taking a bucket's maximum promotes the runner-up instead of rescanning; a rescan happens only
// when one bucket is picked twice.  Same comparisons, same tie rule (lower block id first) as select_row_warp: the picked
// set is identical.  The first version (row staged in shared memory, three passes over it, a rescan per pick) ran 3300
// instructions per row and was issue-bound (79 % issue-active, 0.49 ms for 131072 rows at 64k).
static __device__ __noinline__ void select_row_warp_1024_rounds(const float* __restrict__ src, float* sc, int S_sel, int l_sel, int n_sel,
                                                        int mode, int nf, int K, int t, int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const float NEG = -INFINITY;
  int nvalid = (t + 1) / l_sel;
  if (nvalid > S_sel) nvalid = S_sel;
  // only complete blocks (j < nvalid) can be candidates: buckets past them are neither loaded nor scanned (warp-uniform)
  float raw[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) {
    const int j = lane + 32 * k;
    raw[k] = (32 * (k & ~7) < nvalid && j < S_sel) ? __ldg(src + j) : 0.f;
  }
  uint32_t bm[kSelMaxWords];
#pragma unroll
  for (int i = 0; i < kSelMaxWords; ++i) bm[i] = 0u;
  const int cb = t / l_sel;
  if (mode == 0 && n_sel >= S_sel) {
    for (int w = lane; w * 32 < nvalid; w += 32) {
      int rem = nvalid - w * 32;
      bm[w >> 5] = rem >= 32 ? 0xffffffffu : ((1u << rem) - 1u);
    }
  } else {
    int forced[3];
    const int nfc = forced_decode(nf, cb, forced);  // nf = forced_code(...)
    const int k_rest = n_sel - nfc > 0 ? n_sel - nfc : 0;
    int take = nfc;
    if (mode == 0 && k_rest == 0) take = nfc < n_sel ? nfc : n_sel;
    for (int i = 0; i < take; ++i) {
      int j = forced[i];
      bool ok = mode == 0 ? (j < nvalid) : (j < S_sel);
      if (ok) bitmap_set(bm, lane, j);
    }
    if (k_rest > 0) {
      const int k_act = k_rest < S_sel ? k_rest : S_sel;
      // composites + best / second best of each bucket (ascending j inside a bucket: strict > keeps the lower id first)
      float v1[4], v2[4];
      int j1[4], j2[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        v1[g] = NEG; v2[g] = NEG; j1[g] = 0x7fffffff; j2[g] = 0x7fffffff;
        if (256 * g >= nvalid) continue;  // no candidate in this bucket for any lane; its sc[] entries are never read
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const int j = lane + 32 * (8 * g + kk);
          float c = NEG;
          if (j < nvalid && j != forced[0] && j != forced[1] && j != forced[2])
            c = __fsub_rn(raw[8 * g + kk], __fmul_rn((float)j, 1e-8f));  // two separately rounded fp32 ops (:182-184, :312-318)
          if (j < S_sel) sc[j] = c;
          if (c > v1[g]) { v2[g] = v1[g]; j2[g] = j1[g]; v1[g] = c; j1[g] = j; }
          else if (c > v2[g]) { v2[g] = c; j2[g] = j; }
        }
      }
      __syncwarp();
      // have2[g]: v2/j2 of bucket g are the true runner-up (false after a promotion until the bucket is rescanned)
      bool have2[4] = {true, true, true, true};
      for (int it = 0; it < k_act; ++it) {
        float bv = v1[0];
        int bj = j1[0];
#pragma unroll
        for (int g = 1; g < 4; ++g)
          if (v1[g] > bv) { bv = v1[g]; bj = j1[g]; }  // ascending j across buckets: strict > keeps the lower id
        const uint32_t bits = __float_as_uint(bv);
        const uint32_t key = bits ^ ((bits >> 31) ? 0xffffffffu : 0x80000000u);  // order-preserving
        const uint32_t kmax = __reduce_max_sync(0xffffffffu, key);
        if (kmax == (0xff800000u ^ 0xffffffffu)) break;  // only -inf left: invalid picks are dropped in both modes
        const int vj = (int)__reduce_min_sync(0xffffffffu, key == kmax ? (uint32_t)bj : 0x7fffffffu);
        bitmap_set(bm, lane, vj);
        if ((vj & 31) == lane) {
          sc[vj] = NEG;
          const int g = vj >> 8;  // (vj / 32) / 8
          bool promote = false;
#pragma unroll
          for (int g2 = 0; g2 < 4; ++g2)
            if (g2 == g) promote = have2[g2];
          if (promote) {
#pragma unroll
            for (int g2 = 0; g2 < 4; ++g2)
              if (g2 == g) { v1[g2] = v2[g2]; j1[g2] = j2[g2]; have2[g2] = false; }
          } else {  // second pick from this bucket since its last scan: rescan its 8 entries
            float a1 = NEG, a2 = NEG;
            int i1 = 0x7fffffff, i2 = 0x7fffffff;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              const int j = lane + 32 * (8 * g + kk);
              if (j < S_sel) {
                const float c = sc[j];
                if (c > a1) { a2 = a1; i2 = i1; a1 = c; i1 = j; }
                else if (c > a2) { a2 = c; i2 = j; }
              }
            }
#pragma unroll
            for (int g2 = 0; g2 < 4; ++g2)
              if (g2 == g) { v1[g2] = a1; j1[g2] = i1; v2[g2] = a2; j2[g2] = i2; have2[g2] = true; }
          }
        }
      }
    }
  }
  sel_bitmap_to_ranges(bm, S_sel, l_sel, K, t, out);
}

// The row of the standalone kernel for 128 < S_sel <= 1024: the threshold form when it applies (it does for every row with at
// least k_rest lanes holding a candidate), else the rounds (a call: the two never share registers).
__device__ inline void select_row_warp_1024(const float* __restrict__ src, float* sc, int S_sel, int l_sel, int n_sel, int mode,
                                            int nf, int K, int t, int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  int nvalid = (t + 1) / l_sel;
  if (nvalid > S_sel) nvalid = S_sel;
  int forced[3];
  const int nfc = forced_decode(nf, t / l_sel, forced);
  const int k_rest = n_sel - nfc;
  if (k_rest > 0 && k_rest <= 32 && n_sel < S_sel && nvalid >= k_rest) {  // warp-uniform
    float raw[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int j = lane + 32 * k;
      raw[k] = (32 * (k & ~7) < nvalid && j < S_sel) ? __ldg(src + j) : 0.f;
    }
    uint32_t bm[kSelMaxWords];
#pragma unroll
    for (int i = 0; i < kSelMaxWords; ++i) bm[i] = 0u;
    for (int i = 0; i < nfc; ++i) {
      const int j = forced[i];
      if (mode == 0 ? (j < nvalid) : (j < S_sel)) bitmap_set(bm, lane, j);  // prefill drops incomplete blocks
    }
    if (select_threshold_1024(raw, sc, nvalid, forced, k_rest, bm)) {
      sel_bitmap_to_ranges(bm, S_sel, l_sel, K, t, out);
      return;
    }
    __syncwarp();
  }
  select_row_warp_1024_rounds(src, sc, S_sel, l_sel, n_sel, mode, nf, K, t, out);
}

}  // namespace nsa
