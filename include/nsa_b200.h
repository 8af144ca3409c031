/*
 * nsa_b200.h -- C ABI of the B200-native NSA hot path (libnsa_b200.so).
 *
 * Drop-in boundary for seconds-0/nsa-vibe's attention math.  Every entry point
 *  - takes raw DEVICE pointers, plain ints/floats and a cudaStream_t passed as void*,
 *  - allocates nothing, never synchronises the host, is safe to call from several host
 *    threads on different streams,
 *  - returns 0 on success or a negative NSA_ERR_* code (never throws); nsa_last_error()
 *    gives the text of the last failure on the calling thread.
 * The caller owns all memory, including the workspaces sized by nsa_workspace_bytes().
 *
 * Tensor layouts (contiguous, row-major; "cap" = allocated token rows per (b,g) slab, so
 * pre-allocated caches can be passed without copies):
 *   Q      [B, S, G, h, Dk]      query rows, RoPE already applied      (nsa_attention.py:998-1009)
 *   K_*    [B, G, cap, Dk]       per-branch key caches  (NSA_KV, nsa/cache/kv_cache.py:8-26)
 *   V_*    [B, G, cap, Dv]
 *   O      [B, S, G, h, Dv]      what the reference hands to self.out   (nsa_attention.py:1401)
 *   ranges [B, S, G, K, 2] int32 [start,end) token ranges, [0,0] padded (selection_scorer.py:434-605)
 *   p_grp  [B, S, G, S_sel] fp32 Eq.10 group scores                     (nsa_attention.py:1091)
 *   gates  [B, S, G, 3] fp32     order (cmp, sel, win)                  (nsa_attention.py:1393-1396)
 *   lse    [3, B, S, G, h] fp32  natural-log softmax normalisers per branch (cmp, sel, win)
 * Row s of Q sits at absolute token position t = t0 + s.
 */
#ifndef NSA_B200_H_
#define NSA_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NSA_OK 0
#define NSA_ERR_INVALID_ARG (-1)   /* shape/dtype/alignment not supported: the wrapper raises */
#define NSA_ERR_CUDA (-2)          /* a CUDA runtime call or launch failed */
#define NSA_ERR_UNSUPPORTED (-3)   /* valid request with no kernel for it (no CPU fallback exists) */

enum { NSA_F32 = 0, NSA_BF16 = 1, NSA_F16 = 2 };                 /* element type of Q/K/V/O */
enum { NSA_NORM_FULL_ROW = 0, NSA_NORM_CAUSAL = 1 };              /* p_cmp normaliser */
enum { NSA_GATE_MLP = 0, NSA_GATE_UNIFORM = 1, NSA_GATE_CMP = 2, NSA_GATE_SEL = 3, NSA_GATE_WIN = 4 };
enum { NSA_IMPL_AUTO = 0, NSA_IMPL_SIMT = 1, NSA_IMPL_TC = 2 };   /* kernel family (both sm_100a CUDA) */

/* Geometry shared by the attention entry points. */
typedef struct nsa_dims {
  int32_t B, S, G, h, Dk, Dv;
  int32_t l, d, l_sel, n_sel, w;      /* NSA block parameters (d | l and d | l_sel) */
  int32_t t0;                         /* absolute position of query row 0 */
  int32_t S_sel_kv, cap_sel;          /* tokens present / rows allocated in K_sel,V_sel */
  int32_t S_win_kv, cap_win, win_off; /* rows present / allocated in K_win,V_win; absolute position of row 0 */
  int32_t S_cmp, cap_cmp;             /* compressed tokens present / rows allocated */
  int32_t n_ranges;                   /* K: columns of the ranges tensor */
  int32_t dtype;                      /* NSA_F32 | NSA_BF16 | NSA_F16 */
  int32_t gate_mode;                  /* NSA_GATE_* (NSA_FORCE_BRANCH / NSA_FORCE_UNIFORM_GATE) */
  int32_t gate_hidden;                /* rows of fc1 */
  int32_t norm_mode;                  /* NSA_NORM_* for the scorer */
  int32_t impl;                       /* NSA_IMPL_*; AUTO picks tcgen05 kernels when the shape allows */
  float   gate_tau;                   /* gate temperature */
  float   scale;                      /* softmax scale, 1/sqrt(Dk) */
} nsa_dims_t;

/* GateMLP parameters (nsa_attention.py:32-41), fp32, row-major like nn.Linear.weight. */
typedef struct nsa_gate_params {
  const float* fc1_w;  /* [hidden, Dk] */
  const float* fc1_b;  /* [hidden] */
  const float* fc2_w;  /* [3, hidden] */
  const float* fc2_b;  /* [3] */
} nsa_gate_params_t;

const char* nsa_version(void);
const char* nsa_last_error(void);
/* Number of kernels this process has launched through this library (bench.py reports it as gpu_launches). */
int64_t nsa_kernel_launches(void);

/* ---- (2) selection: p_grp -> ranges, bit-exact vs the reference ------------------------- */
/* Replaces select_topn_ranges_batched + convert_indices_to_ranges_batched_v2
 * (nsa/core/selection_scorer.py:255-362, :434-605).  S_total is the S argument of the reference
 * (decides the forced-column count); rows are t = t0 .. t0+S-1.  K must be
 * nsa_prefill_range_cols(S_total, l_sel, n_sel). */
int nsa_prefill_range_cols(int S_total, int l_sel, int n_sel);   /* force_init = 1, force_local = 2: the module's call */
int nsa_prefill_range_cols_ex(int S_total, int l_sel, int n_sel, int force_init, int force_local);
/* force_init / force_local: the reference's arguments of the same name (block 0 forced; the last force_local blocks forced);
 * force_init + force_local <= 3.  The module calls with (1, 2); the reference's tie-break tests with (0, 0). */
int nsa_select_ranges_prefill(const float* p_grp, int B, int S, int G, int S_sel, int l_sel, int n_sel,
                              int S_total, int t0, int K, int force_init, int force_local, int32_t* ranges, void* stream);
/* Replaces select_topn_ranges (selection_scorer.py:124-249): p_grp [B,G,S_sel], one position t,
 * ranges [B,G,n_sel,2].  Rows the reference leaves as end<=start garbage are written as [0,0]. */
int nsa_select_ranges_decode(const float* p_grp, int B, int G, int S_sel, int l_sel, int n_sel, int t,
                             int force_init, int force_local, int32_t* ranges, void* stream);
/* The stages of the scorer / selector as stand-alone functions, for callers of the reference's free functions (the fused hot
 * path never materialises these tensors):
 *   nsa_pcmp_all           compute_pcmp_all (selection_scorer.py:42-61): p_cmp [B,S,G,h,S_cmp] fp32 = softmax over ALL S_cmp keys;
 *   nsa_map_pcmp_to_pslc   map_pcmp_to_pslc(_batched) (:64-116): p_cmp [n_rows,S_cmp] -> p_slc [n_rows,S_sel], Eq.9, any d|l, d|l_sel;
 *   nsa_indices_to_ranges  convert_indices_to_ranges_batched(_v2) (:380-605): block ids [B,S,G,K] int32 (ascending per row,
 *                          negative = padding) -> ranges [B,S,G,K,2]: duplicates dropped, adjacent blocks merged, ends clamped to
 *                          t0 + s + 1, [0,0] padded. */
int nsa_pcmp_all(const nsa_dims_t* dm, const void* Q, const void* K_cmp, float* p_cmp, void* stream);
int nsa_map_pcmp_to_pslc(const float* p_cmp, int64_t n_rows, int S_cmp, int S_sel, int l, int d, int l_sel, float* p_slc,
                         void* stream);
int nsa_indices_to_ranges(const int32_t* indices, int B, int S, int G, int K, int S_sel, int l_sel, int t0, int32_t* ranges,
                          void* stream);

/* ---- (1) scoring: Q, K_cmp -> p_grp  (compute_pcmp_all + map_pcmp_to_pslc_batched + sum over h;
 * selection_scorer.py:42-61, :89-116, nsa_attention.py:1091).  p_grp [B,S,G,S_sel] fp32. */
int nsa_score(const nsa_dims_t* dm, const void* Q, const void* K_cmp, int S_sel, float* p_grp, void* stream);
/* Fused scoring + selection.  mode 0 = prefill rule, 1 = decode rule.  The SIMT scorer selects inside the scoring kernel
 * (nothing but the ranges leaves the SM); the tcgen05 scorer stages p_grp [B,S,G,S_sel] fp32 in the workspace for the
 * selection kernel that follows it.  workspace: nsa_workspace_bytes(dm, NSA_WS_SCORE_SELECT). */
int nsa_score_select(const nsa_dims_t* dm, const void* Q, const void* K_cmp, int S_sel, int S_total, int mode,
                     int32_t* ranges, void* workspace, void* stream);

/* ---- (3)(4) branch attention.  branch: 0 = cmp, 1 = sel, 2 = win.
 * Replace grouped_selection_attention* / selection_attention_{triton,cuda} (attention_kernels.py:181-772,
 * kernels/triton_sel_kernel, kernels/cuda_sel_kernel/sel_cuda.cpp:28-73), sliding_window_attention
 * (attention_kernels.py:146-178) and batched_causal_attention_compressed (:106-143) with softmax over
 * every allowed key.  O_b [B,S,G,h,Dv] in dm->dtype, lse_b [B,S,G,h] fp32 (either may be NULL). */
int nsa_branch_attn_fwd(const nsa_dims_t* dm, int branch, const void* Q, const void* K, const void* V,
                        const int32_t* ranges, void* O_b, float* lse_b, void* stream);
/* INVARIANT of the tcgen05 selected-branch kernels (16-bit, Dk = Dv = 64; forward, block-major forward and backward):
 * a row's ranges, each clamped to [0, S_sel_kv) and cut into 64-key blocks, give at most NSA_MAX_SEL_BLOCKS blocks.
 * Ranges produced by nsa_select_ranges_* / nsa_score_select always do (n_sel * l_sel <= 1024 keys in whole l_sel blocks).
 * For ranges from anywhere else (many short or unaligned ranges, > 1024 keys per row) call nsa_ranges_max_blocks first and
 * pass dims.impl = NSA_IMPL_SIMT when it reports more: the SIMT kernels take any ranges, like the reference's
 * grouped_selection_attention (attention_kernels.py:181-226).  max_blocks: ONE int32 in device memory. */
#define NSA_MAX_SEL_BLOCKS 16
int nsa_ranges_max_blocks(const int32_t* ranges, int64_t n_rows, int K, int S_kv, int32_t* max_blocks, void* stream);
/* Selected branch, KV-block-major: every 64-key block is read once per run of the queries that selected it and the
 * per-(query, block) partials are merged (same result as nsa_branch_attn_fwd(branch = 1); the form that scales to long
 * prefill, where the query-major gather is bound by L2 bandwidth).  workspace: nsa_workspace_bytes(dm, NSA_WS_SEL_BLOCKMAJOR).
 * Returns NSA_ERR_UNSUPPORTED when the shape has no block-major kernel (use nsa_branch_attn_fwd). */
int nsa_sel_attn_fwd_blockmajor(const nsa_dims_t* dm, const void* Q, const void* K_sel, const void* V_sel,
                                const int32_t* ranges, void* O_b, float* lse_b, void* workspace, void* stream);
/* Analytical backward of one branch (replaces _selection_attention_backward,
 * kernels/triton_sel_kernel/__init__.py:163-231, without its first-key-only line).
 * dO_b [B,S,G,h,Dv] in dm->dtype, O_b as saved, dQ/dK/dV fp32 accumulators (+=, caller zeroes).
 * workspace: nsa_workspace_bytes(dm, NSA_WS_BWD) bytes for the tcgen05 kernels (KV-tile-major, dK/dV accumulated in
 * TMEM, dQ by fp32 reductions); NULL or a shape they do not serve -> the SIMT kernel. */
int nsa_branch_attn_bwd(const nsa_dims_t* dm, int branch, const void* Q, const void* K, const void* V,
                        const int32_t* ranges, const void* O_b, const float* lse_b, const void* dO_b,
                        float* dQ, float* dK, float* dV, void* workspace, void* stream);

/* ---- (4) gate: GateMLP forward / backward (nsa_attention.py:32-82) on q_gp = mean_h(Q). */
int nsa_gate_fwd(const nsa_dims_t* dm, const void* Q, const nsa_gate_params_t* gp, float* gates, void* stream);
/* dgates [B,S,G,3] -> dQ (+= dq_gp / h per head, fp32) and fp32 parameter gradients (+=). */
int nsa_gate_bwd(const nsa_dims_t* dm, const void* Q, const nsa_gate_params_t* gp, const float* dgates,
                 float* dQ, float* d_fc1_w, float* d_fc1_b, float* d_fc2_w, float* d_fc2_b, void* stream);

/* ---- fused hot path ------------------------------------------------------------------------
 * nsa_prefill_fwd: the three branches + gate + combine in one pass; branch outputs stay on chip
 * unless O_branches (3 x [B,S,G,h,Dv], dm->dtype) is non-NULL (needed only to run backward).
 * workspace: nsa_workspace_bytes(dm, NSA_WS_PREFILL) bytes (0 when the single fused kernel serves the shape; otherwise
 * staging for the branch kernels that run separately plus the index / partials of the block-major selected branch).
 * Replaces nsa_attention.py:1137-1398. */
int nsa_prefill_fwd(const nsa_dims_t* dm, const void* Q,
                    const void* K_sel, const void* V_sel, const void* K_win, const void* V_win,
                    const void* K_cmp, const void* V_cmp, const int32_t* ranges,
                    const nsa_gate_params_t* gp, void* O, float* lse, float* gates, void* O_branches,
                    void* workspace, void* stream);
/* Scoring + selection + nsa_prefill_fwd in one call: ranges [B,S,G,K,2] is an OUTPUT (K as for nsa_score_select), everything else
 * as nsa_prefill_fwd.  For long 16-bit prefill this is more than the two calls back to back: the scorer's second pass and the
 * compressed branch run as ONE kernel (one exponential per (row, compressed key) feeds both p_grp and O_cmp).
 * For S >= 8192 the sliding branch and (when neither lse nor O_branches is requested) the GateMLP run on an internal side stream,
 * forked from and joined into `stream` inside the call (capture-safe; NSA_B200_WIN_SIDE=0 disables), and the selected branch's
 * merge then writes the gated output itself: O_sel and the gates' inputs never take an extra pass over HBM.
 * workspace: nsa_workspace_bytes(dm, NSA_WS_PREFILL_FULL).  Replaces nsa_attention.py:1066-1398. */
int nsa_prefill_full_fwd(const nsa_dims_t* dm, const void* Q,
                         const void* K_sel, const void* V_sel, const void* K_win, const void* V_win,
                         const void* K_cmp, const void* V_cmp, const nsa_gate_params_t* gp, int S_sel, int S_total,
                         int sel_mode, int32_t* ranges, void* O, float* lse, float* gates, void* O_branches,
                         void* workspace, void* stream);
/* The two kernels nsa_prefill_full_fwd runs for the scorer on long 16-bit prefill, on their own (profiling, tests):
 *   nsa_score_stats: pass 1 -- stats [B,S,G,h,2] fp32 = (m*c + log2 l of the p_cmp softmax, c * max over the row's causal keys),
 *                    c = scale * log2 e, normaliser per dims.norm_mode;
 *   nsa_score_cmp:   pass 2 fused with the compressed branch -- p_grp [B,S,G,S_sel] fp32 (written up to each CTA's selection limit,
 *                    as the staging of nsa_score_select is), O_cmp [B,S,G,h,Dv], lse_cmp [B,S,G,h] (may be NULL).
 * Both return NSA_ERR_UNSUPPORTED for shapes the fused path does not serve. */
int nsa_score_stats(const nsa_dims_t* dm, const void* Q, const void* K_cmp, float* stats, void* stream);
int nsa_score_cmp(const nsa_dims_t* dm, const void* Q, const void* K_cmp, const void* V_cmp, int S_sel, const float* stats,
                  float* p_grp, void* O_cmp, float* lse_cmp, void* stream);
/* Backward of nsa_prefill_fwd.  dQ [B,S,G,h,Dk], dK_x/dV_x like their caches but fp32 (+=, caller
 * zeroes), dgates [B,S,G,3] fp32 (written).  workspace: nsa_workspace_bytes(dm, NSA_WS_BWD) (see nsa_branch_attn_bwd). */
int nsa_prefill_bwd(const nsa_dims_t* dm, const void* Q,
                    const void* K_sel, const void* V_sel, const void* K_win, const void* V_win,
                    const void* K_cmp, const void* V_cmp, const int32_t* ranges,
                    const void* O_branches, const float* lse, const float* gates, const void* dO,
                    float* dQ, float* dK_sel, float* dV_sel, float* dK_win, float* dV_win,
                    float* dK_cmp, float* dV_cmp, float* dgates, void* workspace, void* stream);
/* nsa_decode_fwd: one decode step after the caches were appended (S must be 1; t = t0).  Scores the
 * emitted compressed keys, selects with the decode rule, attends over the three branches, gates and
 * combines (nsa_attention.py:648-971).  ranges_out [B,G,n_sel,2] may be NULL. */
int nsa_decode_fwd(const nsa_dims_t* dm, const void* Q,
                   const void* K_sel, const void* V_sel, const void* K_win, const void* V_win,
                   const void* K_cmp, const void* V_cmp, const nsa_gate_params_t* gp,
                   void* O, int32_t* ranges_out, void* workspace, void* stream);

/* Device-resident record of a decode loop.  With it the three decode entry points (nsa_decode_produce, nsa_decode_emit,
 * nsa_decode_fwd_stepped) read the position / row counts of the step from DEVICE memory and nsa_decode_advance moves them on, so
 * a caller can capture one decode step in a CUDA graph and replay it with no host-side patching of arguments (the reference's
 * decode loop is bench/bench_decode.py:123-136).  All caches must be slabs with spare capacity ([B,G,cap,D]). */
typedef struct nsa_decode_state {
  int32_t t;        /* position of the next token = rows present in K_sel / V_sel */
  int32_t row_win;  /* rows present in the K_win / V_win slabs (their whole history; the window is the last w of them) */
  int32_t row_raw;  /* rows present in the raw (pre-phi) K / V streams */
  int32_t S_cmp;    /* compressed tokens present in K_cmp / V_cmp */
  int32_t ctr_idx;  /* next column of the read-counter slab */
  int32_t pad_[3];
} nsa_decode_state_t;

/* nsa_decode_fwd with t0, S_sel_kv, S_win_kv, win_off and S_cmp taken from the device record (after this step's produce + emit:
 * t0 = state->t, S_sel_kv = t0 + 1, S_win_kv = row_win + 1, win_off = t0 + 1 - S_win_kv, S_cmp = state->S_cmp + emitted); the
 * dims only carry the static geometry and the slab capacities.  Served by the fused tcgen05 decode kernel only: returns
 * NSA_ERR_UNSUPPORTED when the shape / capacity has none (cap_sel <= 20480 tokens, compressed capacity <= 1024). */
int nsa_decode_stepped_supported(const nsa_dims_t* dm);  /* 1 when nsa_decode_fwd_stepped serves these dims / capacities */
int nsa_decode_fwd_stepped(const nsa_dims_t* dm, const void* Q,
                           const void* K_sel, const void* V_sel, const void* K_win, const void* V_win,
                           const void* K_cmp, const void* V_cmp, const nsa_gate_params_t* gp,
                           void* O, int32_t* ranges_out, const nsa_decode_state_t* state, void* stream);

/* ---- producers of the hot path's inputs ----------------------------------------------------------
 * RoPE (nsa/core/rope.py:16-51: interleaved pairs, angle = (pos / scale) * base^(-2i/dim) in fp32, sin/cos rounded to the
 * tensor's dtype) fused with the re-layout between the projection output [B,S,V,D] (layout 0) and the cache layout [B,V,S,D]
 * (layout 1; nsa_attention.py:403-405).  rot_dim = V*D rotates the V*D values of a token as ONE vector (what the reference
 * does to Q, nsa_attention.py:1002-1009), rot_dim = D rotates each D-vector, rot_dim = 0 only re-lays out.  Row s sits at
 * position t0 + s.  inverse = 1 applies the transposed rotation: the backward pass is the same call with the layouts swapped. */
int nsa_rope_shape(const void* x, void* y, int B, int S, int V, int D, int src_layout, int dst_layout, int rot_dim, int t0,
                   float base, float scale, int inverse, int dtype, void* stream);
/* phi (compress_pool.py:9-38): y[bg, c, :] = mean_{r<l} R(x[bg, c*d + r, :]) with R = RoPE (rope = 1, K) or identity (rope = 0,
 * V); x [BG,S,D] -> y [BG,(S-l)/d+1,D].  backward = 1 maps dy [BG,S_cmp,D] -> dx [BG,S,D] (written, not accumulated). */
int nsa_phi_avgpool(const void* x, void* y, int BG, int S, int D, int l, int d, int rope, int t0, float base, float scale,
                    int backward, int dtype, void* stream);

/* Learnable phi (phi = "mlp", nsa_attention.py:275-291, :1741-1777): a depthwise Conv1d over time, kernel l, stride d, no bias,
 * taps w [D][l] fp32 (nn.Conv1d.weight [D,1,l]), applied to RoPE(K_raw) (rope = 1) or V_raw (rope = 0); initialised to 1/l it is
 * the average pool.  mode 0: y [BG,S_cmp,D] = conv(x [BG,S,D]);  mode 1: x = dy [BG,S_cmp,D] -> y = dx [BG,S,D];
 * mode 2: x = the raw stream [BG,S,D], dy [BG,S_cmp,D] -> y = dw [D][l] fp32 (written). */
int nsa_phi_conv(const void* x, const float* w, void* y, const void* dy, int BG, int S, int D, int l, int d, int rope, int t0,
                 float base, float scale, int mode, int dtype, void* stream);

/* Projection-split producer: the fused projection output y [B, S, H*Dk + G*(3*Dk + 3*Dv)] =
 * (Q | K_sel | V_sel | K_win | V_win | K_raw | V_raw) of S tokens per sequence is rotated (Q as one H*Dk-wide vector, K_sel /
 * K_win per Dk-vector; row s at position t + s) and scattered: Q into q_out [B,S,H*Dk], the six streams into rows
 * row[i] .. row[i]+S-1 of slab[i] ([B,G,cap[i],D]) -- the seven rope / view / permute / torch.cat chains of the reference
 * (nsa_attention.py:545-586 decode, :998-1016 prefill, kv_cache.py:28-49) in one launch.  S = 1 with the cache slabs as
 * destinations is a decode step.  inverse = 1 is the backward pass: the seven gradients are read from q_out / slab and dy is
 * written to y.  counters (optional, [5][counters_cap] int64): the step's read counters (kv_cache.py:51-65) written at
 * counters_idx. */
typedef struct nsa_decode_produce {
  const void* y;
  void* q_out;
  void* slab[6];
  int32_t cap[6], row[6];
  int64_t* counters;
  int32_t counters_cap, counters_idx;
  int64_t counter_val[5];
  int32_t B, H, G, Dk, Dv, t;
  float base, scale;
  int32_t dtype;
  int32_t S, inverse;
  /* optional (S = 1): take t, row[] (= t, t, row_win, row_win, row_raw, row_raw), counters_idx and the counter values from this
   * device record instead of the fields above; l, d, l_sel, n_sel, w then give the counter formula (nsa_attention.py:634-638) */
  const nsa_decode_state_t* state;
  int32_t l, d, l_sel, n_sel, w;
  /* optional rotation tables made by nsa_rope_table for positions [rope_t0, rope_t0 + rope_rows): rope_q [rows][H*Dk/2][2] (Q is
   * rotated as one H*Dk-wide vector) and rope_k [rows][Dk/2][2], in the tensors' dtype, (sin, cos) per position and pair, rounded
   * the way the kernel rounds them.  When both are given and cover [t, t + S) (and state is NULL) the kernel loads them instead
   * of evaluating sincosf: bit-identical outputs, the rotation's ~70 instructions per pair become one 16-byte load per 4 pairs. */
  const void *rope_q, *rope_k;
  int32_t rope_t0, rope_rows;
} nsa_decode_produce_t;
int nsa_decode_produce(const nsa_decode_produce_t* a, void* stream);
/* out [rows][pairs][2] in dtype: (sin, cos) of angle (t0 + row) / scale * base^(-2 * pair / rot_dim), computed and rounded exactly as
 * the rotation kernels do (rope.py:14-33 through ATen's reciprocal-multiply forms). */
int nsa_rope_table(int rows, int pairs, int rot_dim, int t0, float base, float scale, int dtype, void* out, void* stream);

/* Emission of a decode step (nsa_attention.py:587-604): with S_raw = state->row_raw + 1 raw tokens present, if S_raw >= l and
 * (S_raw - l) % d == 0 the compressed token phi(raw rows [S_raw - l, S_raw)) (K with RoPE, compress_pool.py:9-38) is written to
 * row state->S_cmp of the K_cmp / V_cmp slabs; otherwise nothing happens.  Runs after nsa_decode_produce of the same step. */
typedef struct nsa_decode_emit {
  const nsa_decode_state_t* state;
  const void *K_raw, *V_raw;   /* [B,G,cap_raw,Dk|Dv] */
  void *K_cmp, *V_cmp;         /* [B,G,cap_cmp,Dk|Dv] */
  int32_t BG, cap_raw, cap_cmp, Dk, Dv, l, d;
  float base, scale;
  int32_t dtype;
  const float *w_k, *w_v;      /* learnable phi taps [Dk][l] / [Dv][l] fp32, or NULL = average pool */
} nsa_decode_emit_t;
int nsa_decode_emit(const nsa_decode_emit_t* a, void* stream);
/* state += one step: t, row_win, row_raw, ctr_idx advance by one, S_cmp by one when this step emitted. */
int nsa_decode_advance(nsa_decode_state_t* state, int l, int d, void* stream);

/* ---- caller-side row kernels of the block around the hot path ---------------------------------------
 * RMSNorm (nsa/model/llama_block_nsa.py:13-22: y = x * rsqrt(mean(x^2) + eps) * w) fused with the residual add feeding it
 * (llama_block_nsa.py:102-106: x = x + attn_out; mlp(norm2(x))) and with the cast of its consumers: x, s_out, dx, ds are
 * [rows, dim] in x_dtype, r in r_dtype, y / dy in y_dtype, w / dw [dim] in w_dtype; fp32 arithmetic, one rounding on the way out.
 *   fwd: s = x (+ r, then s is written to s_out); rstd[row] saved (fp32, may be NULL); y = (s * rstd) * w.
 *   bwd: dx = rstd * (dy*w - xh * mean(dy*w*xh)) (+ ds), xh = s * rstd; dw = sum_rows dy * xh (dw may be NULL; otherwise
 *        dw_partial is a caller-owned [nsa_rmsnorm_partials(rows), dim] fp32 workspace).  dim % 4 == 0. */
int nsa_rmsnorm_fwd(const void* x, const void* r, const void* w, void* s_out, void* y, float* rstd, int rows, int dim, float eps,
                    int x_dtype, int r_dtype, int w_dtype, int y_dtype, void* stream);
int nsa_rmsnorm_bwd(const void* dy, const void* s, const void* w, const float* rstd, const void* ds, void* dx, void* dw,
                    float* dw_partial, int rows, int dim, int x_dtype, int w_dtype, int y_dtype, void* stream);
int nsa_rmsnorm_partials(int rows);

/* ---- observability reductions ---------------------------------------------------------------------
 * One device record holding what _compute_gate_stats (nsa_attention.py:127-165) and _update_sel_stats_from_ranges (:455-507)
 * report: the host reads it with ONE copy and divides by the row counts.  gates [n_gate_rows,3] fp32 (may be NULL), ranges
 * [n_range_rows,K,2] int32 (may be NULL), row_len: workspace of n_range_rows int32.
 *   gate_sum = { sum entropy, sum max_gate, #collapsed (entropy < 0.1 and max_gate > 0.95), sum g_cmp, sum g_sel, sum g_win };
 *   entropy_min_ord / max_gate_max_ord: fp32 bits mapped to a total order (i >= 0 ? i : i ^ 0x7fffffff), the map is its own inverse;
 *   k_sum = sum_rows L, k_max = max_rows L, rows_at_max = #rows with L == k_max, L = sum_k max(end - start, 0). */
typedef struct nsa_stats {
  double gate_sum[6];
  int32_t entropy_min_ord, max_gate_max_ord;
  int64_t k_sum;
  int32_t k_max, pad_;
  int64_t rows_at_max;
} nsa_stats_t;
int nsa_stats(const float* gates, int64_t n_gate_rows, const int32_t* ranges, int64_t n_range_rows, int K, int32_t* row_len,
              nsa_stats_t* out, void* stream);

enum { NSA_WS_SCORE_SELECT = 0, NSA_WS_DECODE = 1, NSA_WS_PREFILL = 2, NSA_WS_SEL_BLOCKMAJOR = 3, NSA_WS_BWD = 4, NSA_WS_PREFILL_FULL = 5 };
int64_t nsa_workspace_bytes(const nsa_dims_t* dm, int which);

#ifdef __cplusplus
}
#endif
#endif /* NSA_B200_H_ */
