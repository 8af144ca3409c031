// C ABI of libnsa_b200.so (see include/nsa_b200.h).  Validates arguments, picks the kernel family
// (tcgen05 kernels when the shape allows, SIMT otherwise -- both hand-written sm_100a CUDA; there is no
// CPU path) and launches on the caller's stream.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "select.cuh"
#include <map>
#include <mutex>
#include <utility>
#include "launchers.h"

namespace nsa {

static thread_local char g_err[512] = "";
static unsigned long long g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return NSA_ERR_CUDA;
  }
  __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED);
  return NSA_OK;
}

static int validate_dims(const nsa_dims_t* dm, const char* who) {
  NSA_REQUIRE(dm != nullptr, "%s: dims is NULL", who);
  NSA_REQUIRE(dm->B >= 0 && dm->S >= 0 && dm->G >= 1 && dm->h >= 1, "%s: bad B/S/G/h = %d/%d/%d/%d", who, dm->B, dm->S,
              dm->G, dm->h);
  NSA_REQUIRE(dm->Dk >= 1 && dm->Dv >= 1, "%s: bad Dk/Dv", who);
  NSA_REQUIRE(dm->l >= 1 && dm->d >= 1 && dm->l_sel >= 1, "%s: block parameters must be positive", who);
  NSA_REQUIRE(dm->l % dm->d == 0 && dm->l_sel % dm->d == 0, "%s: require d|l and d|l_sel (l=%d d=%d l_sel=%d)", who,
              dm->l, dm->d, dm->l_sel);  // nsa_attention.py:210-211
  NSA_REQUIRE(dm->dtype == NSA_F32 || dm->dtype == NSA_BF16 || dm->dtype == NSA_F16, "%s: dtype %d", who, dm->dtype);
  NSA_REQUIRE(dm->S_cmp <= dm->cap_cmp && dm->S_sel_kv <= dm->cap_sel && dm->S_win_kv <= dm->cap_win,
              "%s: cache rows exceed capacity", who);
  return NSA_OK;
}

static bool tc_eligible(const nsa_dims_t& dm) { return dm.impl != NSA_IMPL_SIMT; }

// The block-major selected branch pays for an index build and a merge pass: worth it once the query-major gather is bound by
// L2 bandwidth: from a few thousand query rows on.  NSA_B200_SEL2=0|1 forces the choice (benchmarks / tests).
static bool use_sel2(const nsa_dims_t& dm) {
  static const int env = getenv("NSA_B200_SEL2") ? atoi(getenv("NSA_B200_SEL2")) : -1;
  if (!tc_sel2_supported(dm) || env == 0) return false;
  if (env == 1) return true;
  // measured on B200 (m7c dims): B=1 S=2048 0.204 vs 0.229 ms for the whole prefill_fwd, B=8 S=2048 0.615 vs 1.139 ms
  return (long long)dm.B * dm.S * dm.G >= 4096 && dm.S_sel_kv >= 1024;
}

}  // namespace nsa

using namespace nsa;

extern "C" {

const char* nsa_version(void) { return "nsa_b200 0.1 (sm_100a)"; }
const char* nsa_last_error(void) { return g_err; }
int64_t nsa_kernel_launches(void) { return (int64_t)__atomic_load_n(&g_launches, __ATOMIC_RELAXED); }

int nsa_prefill_range_cols(int S_total, int l_sel, int n_sel) { return prefill_range_cols(S_total, l_sel, n_sel); }
int nsa_prefill_range_cols_ex(int S_total, int l_sel, int n_sel, int force_init, int force_local) {
  return prefill_range_cols_ex(S_total, l_sel, n_sel, force_init, force_local);
}

int nsa_select_ranges_prefill(const float* p_grp, int B, int S, int G, int S_sel, int l_sel, int n_sel, int S_total,
                              int t0, int K, int force_init, int force_local, int32_t* ranges, void* stream) {
  NSA_REQUIRE(p_grp && ranges, "select_prefill: NULL pointer");
  NSA_REQUIRE(l_sel >= 1 && n_sel >= 0 && S_total >= 1, "select_prefill: bad l_sel/n_sel/S_total");
  NSA_REQUIRE(force_local >= 0 && (force_init ? 1 : 0) + force_local <= 3, "select_prefill: force_init + force_local must be <= 3");
  const int Kw = prefill_range_cols_ex(S_total, l_sel, n_sel, force_init, force_local);
  NSA_REQUIRE(K == Kw, "select_prefill: K=%d but the reference emits %d columns", K, Kw);
  return launch_select(p_grp, B * S * G, S, G, S_sel, l_sel, n_sel, 0, forced_code(0, S_total, l_sel, force_init, force_local), K, t0,
                       ranges, (cudaStream_t)stream);
}

int nsa_select_ranges_decode(const float* p_grp, int B, int G, int S_sel, int l_sel, int n_sel, int t, int force_init,
                             int force_local, int32_t* ranges, void* stream) {
  NSA_REQUIRE(p_grp && ranges, "select_decode: NULL pointer");
  NSA_REQUIRE(l_sel >= 1 && n_sel >= 0 && t >= 0, "select_decode: bad l_sel/n_sel/t");
  NSA_REQUIRE(force_local >= 0 && (force_init ? 1 : 0) + force_local <= 3, "select_decode: force_init + force_local must be <= 3");
  return launch_select(p_grp, B * G, 1, G, S_sel, l_sel, n_sel, 1, forced_code(1, t + 1, l_sel, force_init, force_local), n_sel, t,
                       ranges, (cudaStream_t)stream);
}

int nsa_pcmp_all(const nsa_dims_t* dm, const void* Q, const void* K_cmp, float* p_cmp, void* stream) {
  if (int rc = validate_dims(dm, "pcmp_all")) return rc;
  NSA_REQUIRE(Q && p_cmp && (K_cmp || dm->S_cmp == 0), "pcmp_all: NULL pointer");
  return launch_pcmp_all(*dm, Q, K_cmp, p_cmp, (cudaStream_t)stream);
}

int nsa_map_pcmp_to_pslc(const float* p_cmp, int64_t n_rows, int S_cmp, int S_sel, int l, int d, int l_sel, float* p_slc,
                         void* stream) {
  NSA_REQUIRE((p_cmp || n_rows * S_cmp == 0) && (p_slc || n_rows * S_sel == 0), "map_pcmp_to_pslc: NULL pointer");
  NSA_REQUIRE(n_rows >= 0 && S_cmp >= 0 && S_sel >= 0 && l >= 1 && d >= 1 && l_sel >= 1, "map_pcmp_to_pslc: bad sizes");
  return launch_map_pslc(p_cmp, (long long)n_rows, S_cmp, S_sel, l, d, l_sel, p_slc, (cudaStream_t)stream);
}

int nsa_indices_to_ranges(const int32_t* indices, int B, int S, int G, int K, int S_sel, int l_sel, int t0, int32_t* ranges,
                          void* stream) {
  NSA_REQUIRE((indices && ranges) || (long long)B * S * G * K == 0, "indices_to_ranges: NULL pointer");
  NSA_REQUIRE(B >= 0 && S >= 0 && G >= 0 && K >= 0 && l_sel >= 1, "indices_to_ranges: bad sizes");
  return launch_indices_to_ranges(indices, B, S, G, K, S_sel, l_sel, t0, ranges, (cudaStream_t)stream);
}

int nsa_score(const nsa_dims_t* dm, const void* Q, const void* K_cmp, int S_sel, float* p_grp, void* stream) {
  if (int rc = validate_dims(dm, "score")) return rc;
  NSA_REQUIRE(Q && p_grp && (K_cmp || dm->S_cmp == 0), "score: NULL pointer");
  if (tc_eligible(*dm) && tc_score_supported(*dm))
    return launch_score_tc(*dm, Q, K_cmp, S_sel, 0, 0, 0, p_grp, nullptr, nullptr, (cudaStream_t)stream);
  return launch_score_generic(*dm, Q, K_cmp, S_sel, dm->t0 + dm->S, 0, 0, p_grp, nullptr, (cudaStream_t)stream);
}

int nsa_score_select(const nsa_dims_t* dm, const void* Q, const void* K_cmp, int S_sel, int S_total, int mode,
                     int32_t* ranges, void* workspace, void* stream) {
  if (int rc = validate_dims(dm, "score_select")) return rc;
  NSA_REQUIRE(Q && ranges && (K_cmp || dm->S_cmp == 0), "score_select: NULL pointer");
  NSA_REQUIRE(mode == 0 || mode == 1, "score_select: mode %d", mode);
  const int K = mode == 0 ? prefill_range_cols(S_total, dm->l_sel, dm->n_sel) : dm->n_sel;
  NSA_REQUIRE(dm->n_ranges == K, "score_select: dims.n_ranges=%d but this rule emits %d columns", dm->n_ranges, K);
  if (tc_eligible(*dm) && tc_score_supported(*dm))
    return launch_score_tc(*dm, Q, K_cmp, S_sel, S_total, mode, K, nullptr, ranges, workspace, (cudaStream_t)stream);
  return launch_score_generic(*dm, Q, K_cmp, S_sel, S_total, mode, K, nullptr, ranges, (cudaStream_t)stream);
}

static void fill_fwd(FwdArgs& a, const void* Q, const void* K_sel, const void* V_sel, const void* K_win,
                     const void* V_win, const void* K_cmp, const void* V_cmp) {
  memset(&a, 0, sizeof(a));
  a.Q = Q;
  a.K[0] = K_cmp; a.V[0] = V_cmp;
  a.K[1] = K_sel; a.V[1] = V_sel;
  a.K[2] = K_win; a.V[2] = V_win;
}

int nsa_branch_attn_fwd(const nsa_dims_t* dm, int branch, const void* Q, const void* K, const void* V,
                        const int32_t* ranges, void* O_b, float* lse_b, void* stream) {
  if (int rc = validate_dims(dm, "branch_attn_fwd")) return rc;
  NSA_REQUIRE(branch >= 0 && branch <= 2, "branch_attn_fwd: branch %d", branch);
  NSA_REQUIRE(Q && O_b, "branch_attn_fwd: NULL pointer");
  NSA_REQUIRE(branch != 1 || ranges, "branch_attn_fwd: the selected branch needs ranges");
  if (tc_branch_supported(*dm, branch)) return launch_branch_tc(*dm, branch, Q, K, V, ranges, O_b, lse_b, (cudaStream_t)stream);
  NSA_REQUIRE(dm->impl != NSA_IMPL_TC, "branch_attn_fwd: no tcgen05 kernel for this shape/branch (impl=TC was forced)");
  FwdArgs a;
  memset(&a, 0, sizeof(a));
  a.Q = Q;
  a.K[branch] = K;
  a.V[branch] = V;
  a.ranges = ranges;
  a.branch_mask = 1 << branch;
  // write this branch's result into slot `branch` of the [3][...] views by offsetting the base pointers back
  const size_t rows_h = (size_t)dm->B * dm->S * dm->G * dm->h;
  a.O_br = (char*)O_b - (size_t)branch * rows_h * dm->Dv * elt_size(dm->dtype);
  a.lse = lse_b ? lse_b - (size_t)branch * rows_h : nullptr;
  nsa_dims_t d2 = *dm;
  d2.gate_mode = NSA_GATE_UNIFORM;  // gates are irrelevant here (O is not produced)
  return launch_fwd_generic(d2, a, (cudaStream_t)stream);
}

int nsa_sel_attn_fwd_blockmajor(const nsa_dims_t* dm, const void* Q, const void* K_sel, const void* V_sel,
                                const int32_t* ranges, void* O_b, float* lse_b, void* workspace, void* stream) {
  if (int rc = validate_dims(dm, "sel_attn_fwd_blockmajor")) return rc;
  NSA_REQUIRE(Q && K_sel && V_sel && ranges && O_b && workspace, "sel_attn_fwd_blockmajor: NULL pointer");
  if (!tc_sel2_supported(*dm)) {
    set_error("sel_attn_fwd_blockmajor: no block-major kernel for this shape/dtype");
    return NSA_ERR_UNSUPPORTED;
  }
  return launch_sel2_tc(*dm, Q, K_sel, V_sel, ranges, O_b, lse_b, workspace, (cudaStream_t)stream);
}

int nsa_branch_attn_bwd(const nsa_dims_t* dm, int branch, const void* Q, const void* K, const void* V,
                        const int32_t* ranges, const void* O_b, const float* lse_b, const void* dO_b, float* dQ,
                        float* dK, float* dV, void* workspace, void* stream) {
  if (int rc = validate_dims(dm, "branch_attn_bwd")) return rc;
  NSA_REQUIRE(branch >= 0 && branch <= 2, "branch_attn_bwd: branch %d", branch);
  NSA_REQUIRE(Q && O_b && lse_b && dO_b && dQ && dK && dV, "branch_attn_bwd: NULL pointer");
  BwdArgs a;
  memset(&a, 0, sizeof(a));
  const size_t rows_h = (size_t)dm->B * dm->S * dm->G * dm->h;
  a.Q = Q;
  a.K[branch] = K;
  a.V[branch] = V;
  a.ranges = ranges;
  a.O_br = (const char*)O_b - (size_t)branch * rows_h * dm->Dv * elt_size(dm->dtype);
  a.lse = lse_b - (size_t)branch * rows_h;
  a.dO = dO_b;
  a.dQ = dQ;
  a.dK[branch] = dK;
  a.dV[branch] = dV;
  a.branch_mask = 1 << branch;
  return launch_bwd_tc(*dm, a, workspace, (cudaStream_t)stream);
}

int nsa_gate_fwd(const nsa_dims_t* dm, const void* Q, const nsa_gate_params_t* gp, float* gates, void* stream) {
  if (int rc = validate_dims(dm, "gate_fwd")) return rc;
  NSA_REQUIRE(Q && gates && gp, "gate_fwd: NULL pointer");
  NSA_REQUIRE(dm->gate_mode != NSA_GATE_MLP || (gp->fc1_w && gp->fc2_w), "gate_fwd: MLP weights missing");
  return launch_gate_fwd(*dm, Q, *gp, gates, (cudaStream_t)stream);
}

int nsa_gate_bwd(const nsa_dims_t* dm, const void* Q, const nsa_gate_params_t* gp, const float* dgates, float* dQ,
                 float* d_fc1_w, float* d_fc1_b, float* d_fc2_w, float* d_fc2_b, void* stream) {
  if (int rc = validate_dims(dm, "gate_bwd")) return rc;
  NSA_REQUIRE(Q && gp && dgates && d_fc1_w && d_fc1_b && d_fc2_w && d_fc2_b, "gate_bwd: NULL pointer");
  return launch_gate_bwd(*dm, Q, *gp, dgates, dQ, d_fc1_w, d_fc1_b, d_fc2_w, d_fc2_b, (cudaStream_t)stream);
}

}  // extern "C"

// nsa_prefill_fwd; have_mask: branches whose output (and lse) already sit in their staging slot (the fused scorer + compressed
// branch writes slot 0 before the ranges exist)
static int prefill_fwd_impl(const nsa_dims_t* dm, const void* Q, const void* K_sel, const void* V_sel, const void* K_win,
                            const void* V_win, const void* K_cmp, const void* V_cmp, const int32_t* ranges,
                            const nsa_gate_params_t* gp, void* O, float* lse, float* gates, void* O_branches, void* workspace,
                            void* stream, int have_mask, cudaEvent_t join = nullptr, const float* fuse_gates = nullptr) {
  struct JoinGuard {  // the side stream's work (a precomputed branch) joins the caller's stream before anything reads it
    cudaEvent_t ev; cudaStream_t st; bool done;
    void now() { if (ev && !done) { cudaStreamWaitEvent(st, ev, 0); done = true; } }
    ~JoinGuard() { now(); }
  } join_guard{join, (cudaStream_t)stream, false};
  if (int rc = validate_dims(dm, "prefill_fwd")) return rc;
  NSA_REQUIRE(Q && O && ranges && gp, "prefill_fwd: NULL pointer");
  NSA_REQUIRE(dm->gate_mode != NSA_GATE_MLP || (gp->fc1_w && gp->fc2_w), "prefill_fwd: MLP weights missing");
  cudaStream_t st = (cudaStream_t)stream;
  FwdArgs a;
  fill_fwd(a, Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp);
  a.ranges = ranges;
  a.gp = *gp;
  a.lse = lse;
  int tc_mask = 0;
  for (int br = 0; br < 3; ++br)
    if (tc_branch_supported(*dm, br)) tc_mask |= 1 << br;
  NSA_REQUIRE((have_mask & ~tc_mask) == 0, "prefill_fwd: a precomputed branch needs the tensor-core path");
  if (tc_mask == 0) {  // everything in the one fused SIMT kernel: branch outputs never leave the SM
    NSA_REQUIRE(dm->impl != NSA_IMPL_TC, "prefill_fwd: no tcgen05 kernel for this shape (impl=TC was forced)");
    a.O = O;
    a.gates_out = gates;
    a.O_br = O_branches;
    a.branch_mask = 7;
    return launch_fwd_generic(*dm, a, st);
  }
  // tensor-core branches run as their own kernels; branch outputs go through O_branches (or the workspace)
  const size_t rows_h = (size_t)dm->B * dm->S * dm->G * dm->h;
  const size_t per_branch = rows_h * dm->Dv * elt_size(dm->dtype);
  const size_t staging = (3 * per_branch + 255) & ~(size_t)255;
  void* obr = O_branches ? O_branches : workspace;
  NSA_REQUIRE(obr, "prefill_fwd: this shape needs O_branches or a workspace of nsa_workspace_bytes(NSA_WS_PREFILL)");
  const void* Ks[3] = {K_cmp, K_sel, K_win};
  const void* Vs[3] = {V_cmp, V_sel, V_win};
  // No-grad long prefill (nothing saved for backward, all three branches on tensor cores, block-major selected branch) can merge
  // the selected branch's partials, evaluate the gate and combine with the other two branches in ONE pass, so that O_sel and the
  // gates never reach HBM (NSA_B200_FUSE_COMBINE=1).  Measured at 64k it is SLOWER than the two kernels it replaces (0.70 ms
  // against 0.405 + 0.136 ms: two dependent gather round trips per row plus the gate MLP at 16 warps per SM), so it is opt-in.
  static const bool fuse_env = getenv("NSA_B200_FUSE_COMBINE") && atoi(getenv("NSA_B200_FUSE_COMBINE")) == 1;
  // fuse_gates: the gates were evaluated on the side stream (nsa_prefill_full_fwd): the merge of the selected branch's partials
  // blends the three branches itself -- no O_sel round trip, no combine pass
  const bool fuse = (fuse_env || fuse_gates) && tc_mask == 7 && !O_branches && !lse && workspace && use_sel2(*dm) && sel2_fuse_supported(*dm);
  for (int br = 0; br < 3; ++br) {
    if (!(tc_mask & (1 << br)) || (fuse && br == 1) || (have_mask & (1 << br))) continue;
    void* ob = (char*)obr + br * per_branch;
    float* lb = lse ? lse + br * rows_h : nullptr;
    if (br == 1 && workspace && use_sel2(*dm)) {  // long prefill: KV-block-major selected branch
      if (int rc = launch_sel2_tc(*dm, Q, K_sel, V_sel, ranges, ob, lb, (char*)workspace + staging, st)) return rc;
      continue;
    }
    if (int rc = launch_branch_tc(*dm, br, Q, Ks[br], Vs[br], ranges, ob, lb, st)) return rc;
  }
  join_guard.now();
  if (fuse) {
    Sel2Fuse f;
    f.gp = gp;
    f.O_cmp = obr;
    f.O_win = (char*)obr + 2 * per_branch;
    f.O = O;
    f.gates = gates;
    f.gates_in = fuse_gates;
    return launch_sel2_tc(*dm, Q, K_sel, V_sel, ranges, nullptr, nullptr, (char*)workspace + staging, st, &f);
  }
  if (tc_mask != 7) {
    a.O_br = obr;
    a.branch_mask = 7 & ~tc_mask;
    nsa_dims_t d2 = *dm;
    d2.gate_mode = NSA_GATE_UNIFORM;  // gates are applied by the combine kernel
    if (int rc = launch_fwd_generic(d2, a, st)) return rc;
  }
  return launch_combine(*dm, Q, *gp, obr, O, gates, st);
}

extern "C" {

int nsa_prefill_fwd(const nsa_dims_t* dm, const void* Q, const void* K_sel, const void* V_sel, const void* K_win,
                    const void* V_win, const void* K_cmp, const void* V_cmp, const int32_t* ranges,
                    const nsa_gate_params_t* gp, void* O, float* lse, float* gates, void* O_branches, void* workspace,
                    void* stream) {
  return prefill_fwd_impl(dm, Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, ranges, gp, O, lse, gates, O_branches, workspace, stream, 0);
}

int nsa_score_stats(const nsa_dims_t* dm, const void* Q, const void* K_cmp, float* stats, void* stream) {
  if (int rc = validate_dims(dm, "score_stats")) return rc;
  NSA_REQUIRE(Q && K_cmp && stats, "score_stats: NULL pointer");
  if (!(tc_eligible(*dm) && tc_score_supported(*dm) && tc_score_cmp_supported(*dm))) {
    set_error("score_stats: the split scorer serves long 16-bit prefill only (Dk = Dv = 64, l = 2d, l_sel = 4d, h <= 32)");
    return NSA_ERR_UNSUPPORTED;
  }
  return launch_score_stats_tc(*dm, Q, K_cmp, stats, (cudaStream_t)stream);
}

int nsa_score_cmp(const nsa_dims_t* dm, const void* Q, const void* K_cmp, const void* V_cmp, int S_sel, const float* stats,
                  float* p_grp, void* O_cmp, float* lse_cmp, void* stream) {
  if (int rc = validate_dims(dm, "score_cmp")) return rc;
  NSA_REQUIRE(Q && K_cmp && V_cmp && stats && p_grp && O_cmp, "score_cmp: NULL pointer");
  if (!(tc_eligible(*dm) && tc_score_supported(*dm) && tc_score_cmp_supported(*dm))) {
    set_error("score_cmp: the fused scorer + compressed branch serves long 16-bit prefill only");
    return NSA_ERR_UNSUPPORTED;
  }
  return launch_score_cmp_tc(*dm, Q, K_cmp, V_cmp, S_sel, stats, p_grp, O_cmp, lse_cmp, (cudaStream_t)stream);
}

// A side stream per (device, caller stream) for work that may overlap the caller's chain inside one call (fork / join with
// events, which is also how a capturing stream pulls a second stream into its graph).
static cudaStream_t side_stream_for(cudaStream_t main) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, cudaStream_t> pool;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lk(mu);
  auto key = std::make_pair(dev, main);
  auto it = pool.find(key);
  if (it != pool.end()) return it->second;
  cudaStream_t s = nullptr;
  if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  pool[key] = s;
  return s;
}

// workspace slot for the gates of the fused prefill when the caller passes no gates tensor
static int64_t full_gate_bytes(const nsa_dims_t& dm) { return (((int64_t)dm.B * dm.S * dm.G * 3 * 4) + 255) & ~(int64_t)255; }

static bool full_fused(const nsa_dims_t& dm) {
  return tc_eligible(dm) && tc_score_supported(dm) && tc_score_cmp_supported(dm) && tc_branch_supported(dm, 0) &&
         tc_branch_supported(dm, 1) && tc_branch_supported(dm, 2);
}

int nsa_prefill_full_fwd(const nsa_dims_t* dm, const void* Q, const void* K_sel, const void* V_sel, const void* K_win,
                         const void* V_win, const void* K_cmp, const void* V_cmp, const nsa_gate_params_t* gp, int S_sel,
                         int S_total, int sel_mode, int32_t* ranges, void* O, float* lse, float* gates, void* O_branches,
                         void* workspace, void* stream) {
  if (int rc = validate_dims(dm, "prefill_full_fwd")) return rc;
  NSA_REQUIRE(Q && O && ranges && gp, "prefill_full_fwd: NULL pointer");
  NSA_REQUIRE(sel_mode == 0 || sel_mode == 1, "prefill_full_fwd: mode %d", sel_mode);
  const int K = sel_mode == 0 ? prefill_range_cols(S_total, dm->l_sel, dm->n_sel) : dm->n_sel;
  NSA_REQUIRE(dm->n_ranges == K, "prefill_full_fwd: dims.n_ranges=%d but this rule emits %d columns", dm->n_ranges, K);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ws_score = (tc_score_workspace(*dm) + 255) & ~(int64_t)255;
  char* ws = reinterpret_cast<char*>(workspace);
  if (!full_fused(*dm)) {  // the two stand-alone entry points back to back
    if (int rc = nsa_score_select(dm, Q, K_cmp, S_sel, S_total, sel_mode, ranges, ws, stream)) return rc;
    return nsa_prefill_fwd(dm, Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, ranges, gp, O, lse, gates, O_branches,
                           ws ? ws + ws_score : nullptr, stream);
  }
  // Long prefill on tensor cores: pass 1 of the scorer (row statistics), then pass 2 FUSED with the compressed branch (one
  // exponential per (row, compressed key) serves p_grp and O_cmp), selection, selected and sliding branches, gated combine.
  NSA_REQUIRE(ws, "prefill_full_fwd: needs a workspace of nsa_workspace_bytes(NSA_WS_PREFILL_FULL) bytes");
  NSA_REQUIRE((int64_t)dm->B * dm->S * dm->G * S_sel * 4 <= tc_score_workspace(*dm), "prefill_full_fwd: S_sel=%d exceeds the workspace", S_sel);
  float* pg = reinterpret_cast<float*>(ws);
  float* stats = reinterpret_cast<float*>(ws + ws_score);
  char* ws_pre = ws + ws_score + tc_score_cmp_stats_bytes(*dm) + full_gate_bytes(*dm);
  const size_t rows_h = (size_t)dm->B * dm->S * dm->G * dm->h;
  void* o_cmp = O_branches ? O_branches : (void*)ws_pre;  // staging slot 0 of prefill_fwd_impl
  // The sliding branch depends on nothing the scorer or the selection produce: it runs on a side stream, forked here and joined
  // before the combine, so the tails of its kernel and of the chain's first kernels fill each other.  Measured at 64k: step
  // 3.27 -> 3.21 ms, module-level prefill 3.68 -> 3.60 ms (forking after the selection instead: no gain).  NSA_B200_WIN_SIDE=0
  // keeps everything on the caller's stream.
  static const int win_side_env = getenv("NSA_B200_WIN_SIDE") ? atoi(getenv("NSA_B200_WIN_SIDE")) : 1;
  // no-grad long prefill (nothing saved for a backward): the GateMLP runs on the side stream too and the selected branch's merge
  // blends the three branches (sel2_merge_blend_kernel); gates go to the caller's tensor or to a workspace slot
  static const bool gate_side_env = !(getenv("NSA_B200_GATE_SIDE") && atoi(getenv("NSA_B200_GATE_SIDE")) == 0);
  float* gates_buf = gates ? gates : reinterpret_cast<float*>(ws + ws_score + tc_score_cmp_stats_bytes(*dm));
  const bool gate_side = gate_side_env && !O_branches && !lse && gate_fast_supported(*dm, gp, Q) && use_sel2(*dm) && sel2_fuse_supported(*dm);
  const float* fuse_gates = nullptr;
  int have = 1;
  cudaEvent_t ev_join = nullptr;
  struct Joiner {  // whatever path leaves this function, the caller's stream has joined the side stream and the event is gone
    cudaEvent_t& ev; cudaStream_t st;
    ~Joiner() { if (ev) { cudaStreamWaitEvent(st, ev, 0); cudaEventDestroy(ev); ev = nullptr; } }
  } joiner{ev_join, st};
  auto fork_win = [&]() -> int {
    cudaStream_t side = side_stream_for(st);
    cudaEvent_t ev_fork = nullptr;
    if (side && cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) == cudaSuccess &&
        cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) == cudaSuccess) {
      const size_t per_branch = rows_h * dm->Dv * elt_size(dm->dtype);
      void* obr = O_branches ? O_branches : (void*)ws_pre;
      cudaEventRecord(ev_fork, st);
      cudaStreamWaitEvent(side, ev_fork, 0);
      int rc = launch_branch_tc(*dm, 2, Q, K_win, V_win, nullptr, (char*)obr + 2 * per_branch, lse ? lse + 2 * rows_h : nullptr, side);
      if (!rc && gate_side) {
        rc = launch_gate_fast(*dm, Q, *gp, gates_buf, side);
        if (!rc) fuse_gates = gates_buf;
      }
      cudaEventRecord(ev_join, side);
      cudaEventDestroy(ev_fork);
      if (rc) return rc;
      have |= 4;
    } else {
      if (ev_fork) cudaEventDestroy(ev_fork);
      if (ev_join) cudaEventDestroy(ev_join);
      ev_join = nullptr;
    }
    return NSA_OK;
  };
  // long prefill only: at the training shape (8 x 2048 tokens, the step replayed as one CUDA graph) the fork measured +1 % on the step
  if (win_side_env == 1 && dm->S >= 8192)
    if (int rc = fork_win()) return rc;
  if (int rc = launch_score_stats_tc(*dm, Q, K_cmp, stats, st)) return rc;
  if (int rc = launch_score_cmp_tc(*dm, Q, K_cmp, V_cmp, S_sel, stats, pg, o_cmp, lse, st)) return rc;
  (void)rows_h;
  const int nf = forced_code_default(sel_mode, S_total, dm->l_sel);
  if (int rc = launch_select(pg, dm->B * dm->S * dm->G, dm->S, dm->G, S_sel, dm->l_sel, dm->n_sel, sel_mode, nf, K, dm->t0, ranges, st)) return rc;
  return prefill_fwd_impl(dm, Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, ranges, gp, O, lse, gates, O_branches, ws_pre, stream, have,
                          ev_join, fuse_gates);
}

int nsa_prefill_bwd(const nsa_dims_t* dm, const void* Q, const void* K_sel, const void* V_sel, const void* K_win,
                    const void* V_win, const void* K_cmp, const void* V_cmp, const int32_t* ranges,
                    const void* O_branches, const float* lse, const float* gates, const void* dO, float* dQ,
                    float* dK_sel, float* dV_sel, float* dK_win, float* dV_win, float* dK_cmp, float* dV_cmp,
                    float* dgates, void* workspace, void* stream) {
  if (int rc = validate_dims(dm, "prefill_bwd")) return rc;
  NSA_REQUIRE(Q && ranges && O_branches && lse && gates && dO && dQ, "prefill_bwd: NULL pointer");
  NSA_REQUIRE(dK_sel && dV_sel && dK_win && dV_win && dK_cmp && dV_cmp, "prefill_bwd: NULL gradient buffer");
  BwdArgs a;
  memset(&a, 0, sizeof(a));
  a.Q = Q;
  a.K[0] = K_cmp; a.V[0] = V_cmp; a.dK[0] = dK_cmp; a.dV[0] = dV_cmp;
  a.K[1] = K_sel; a.V[1] = V_sel; a.dK[1] = dK_sel; a.dV[1] = dV_sel;
  a.K[2] = K_win; a.V[2] = V_win; a.dK[2] = dK_win; a.dV[2] = dV_win;
  a.ranges = ranges;
  a.O_br = O_branches;
  a.lse = lse;
  a.gates = gates;
  a.dO = dO;
  a.dQ = dQ;
  a.dgates = dgates;
  a.branch_mask = 7;
  return launch_bwd_tc(*dm, a, workspace, (cudaStream_t)stream);
}

int nsa_decode_fwd(const nsa_dims_t* dm, const void* Q, const void* K_sel, const void* V_sel, const void* K_win,
                   const void* V_win, const void* K_cmp, const void* V_cmp, const nsa_gate_params_t* gp, void* O,
                   int32_t* ranges_out, void* workspace, void* stream) {
  if (int rc = validate_dims(dm, "decode_fwd")) return rc;
  NSA_REQUIRE(dm->S == 1, "decode_fwd: decode requires S == 1, got S=%d", dm->S);  // nsa_attention.py:532-535
  NSA_REQUIRE(Q && O && gp, "decode_fwd: NULL pointer");
  NSA_REQUIRE(ranges_out || workspace, "decode_fwd: need ranges_out or a workspace");
  NSA_REQUIRE(dm->n_ranges == dm->n_sel, "decode_fwd: dims.n_ranges must equal n_sel");
  if (tc_eligible(*dm) && tc_decode_supported(*dm))
    return launch_decode_tc(*dm, Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, *gp, O, ranges_out, workspace,
                            (cudaStream_t)stream);
  int32_t* ranges = ranges_out ? ranges_out : (int32_t*)workspace;
  const int t = dm->t0;
  const int cover = t + 1 > dm->l_sel ? t + 1 : dm->l_sel;  // meta covers max(t+1, l_sel) tokens (nsa_attention.py:609-632)
  const int S_sel = ceil_div(cover, dm->l_sel);
  nsa_dims_t d2 = *dm;
  d2.norm_mode = NSA_NORM_FULL_ROW;  // decode sees only emitted blocks: softmax over all of them (:650-651)
  if (int rc = launch_score_generic(d2, Q, K_cmp, S_sel, t + 1, 1, dm->n_sel, nullptr, ranges, (cudaStream_t)stream)) return rc;
  FwdArgs a;
  fill_fwd(a, Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp);
  a.ranges = ranges;
  a.gp = *gp;
  a.O = O;
  a.branch_mask = 7;
  return launch_fwd_generic(*dm, a, (cudaStream_t)stream);
}

int nsa_decode_fwd_stepped(const nsa_dims_t* dm, const void* Q, const void* K_sel, const void* V_sel, const void* K_win,
                           const void* V_win, const void* K_cmp, const void* V_cmp, const nsa_gate_params_t* gp, void* O,
                           int32_t* ranges_out, const nsa_decode_state_t* state, void* stream) {
  NSA_REQUIRE(dm && state, "decode_fwd_stepped: NULL dims / state");
  nsa_dims_t d2 = *dm;  // rows present are read from the device record: validate the static geometry against the capacities
  d2.S_sel_kv = d2.cap_sel; d2.S_win_kv = d2.cap_win; d2.S_cmp = d2.cap_cmp; d2.t0 = 0; d2.win_off = 0;
  if (int rc = validate_dims(&d2, "decode_fwd_stepped")) return rc;
  NSA_REQUIRE(dm->S == 1, "decode_fwd_stepped: decode requires S == 1, got S=%d", dm->S);
  NSA_REQUIRE(Q && O && gp && K_sel && V_sel && K_win && V_win && K_cmp && V_cmp, "decode_fwd_stepped: NULL pointer");
  NSA_REQUIRE(dm->n_ranges == dm->n_sel, "decode_fwd_stepped: dims.n_ranges must equal n_sel");
  if (!tc_decode_stepped_supported(d2)) {
    set_error("decode_fwd_stepped: no fused tcgen05 decode kernel for this shape / capacity");
    return NSA_ERR_UNSUPPORTED;
  }
  return launch_decode_tc_stepped(d2, Q, K_sel, V_sel, K_win, V_win, K_cmp, V_cmp, *gp, O, ranges_out, state, (cudaStream_t)stream);
}

int nsa_decode_stepped_supported(const nsa_dims_t* dm) {
  if (!dm || dm->S != 1 || dm->n_ranges != dm->n_sel) return 0;
  nsa_dims_t d2 = *dm;
  d2.S_sel_kv = d2.cap_sel; d2.S_win_kv = d2.cap_win; d2.S_cmp = d2.cap_cmp; d2.t0 = 0; d2.win_off = 0;
  return tc_eligible(d2) && tc_decode_stepped_supported(d2) ? 1 : 0;
}

int nsa_decode_emit(const nsa_decode_emit_t* a, void* stream) {
  NSA_REQUIRE(a, "decode_emit: NULL argument block");
  return launch_decode_emit(*a, (cudaStream_t)stream);
}

int nsa_decode_advance(nsa_decode_state_t* state, int l, int d, void* stream) {
  return launch_decode_advance(state, l, d, (cudaStream_t)stream);
}

int nsa_rope_shape(const void* x, void* y, int B, int S, int V, int D, int src_layout, int dst_layout, int rot_dim, int t0,
                   float base, float scale, int inverse, int dtype, void* stream) {
  NSA_REQUIRE(dtype == NSA_F32 || dtype == NSA_BF16 || dtype == NSA_F16, "rope_shape: dtype %d", dtype);
  return launch_rope_shape(x, y, B, S, V, D, src_layout, dst_layout, rot_dim, t0, base, scale, inverse, dtype, (cudaStream_t)stream);
}

int nsa_phi_avgpool(const void* x, void* y, int BG, int S, int D, int l, int d, int rope, int t0, float base, float scale,
                    int backward, int dtype, void* stream) {
  NSA_REQUIRE(dtype == NSA_F32 || dtype == NSA_BF16 || dtype == NSA_F16, "phi_avgpool: dtype %d", dtype);
  return launch_phi_avgpool(x, y, BG, S, D, l, d, rope, t0, base, scale, backward, dtype, (cudaStream_t)stream);
}

int nsa_phi_conv(const void* x, const float* w, void* y, const void* dy, int BG, int S, int D, int l, int d, int rope, int t0,
                 float base, float scale, int mode, int dtype, void* stream) {
  NSA_REQUIRE(dtype == NSA_F32 || dtype == NSA_BF16 || dtype == NSA_F16, "phi_conv: dtype %d", dtype);
  NSA_REQUIRE(mode >= 0 && mode <= 2 && (w || mode == 2), "phi_conv: mode %d / NULL weights", mode);
  return launch_phi_avgpool(x, y, BG, S, D, l, d, rope, t0, base, scale, mode, dtype, (cudaStream_t)stream, w, dy);
}

int nsa_rope_table(int rows, int pairs, int rot_dim, int t0, float base, float scale, int dtype, void* out, void* stream) {
  NSA_REQUIRE(dtype == NSA_F32 || dtype == NSA_BF16 || dtype == NSA_F16, "rope_table: dtype %d", dtype);
  return launch_rope_table(rows, pairs, rot_dim, t0, base, scale, dtype, out, (cudaStream_t)stream);
}

int nsa_decode_produce(const nsa_decode_produce_t* a, void* stream) {
  NSA_REQUIRE(a, "decode_produce: NULL argument block");
  NSA_REQUIRE(a->dtype == NSA_F32 || a->dtype == NSA_BF16 || a->dtype == NSA_F16, "decode_produce: dtype %d", a->dtype);
  return launch_decode_produce(*a, (cudaStream_t)stream);
}

int nsa_rmsnorm_fwd(const void* x, const void* r, const void* w, void* s_out, void* y, float* rstd, int rows, int dim, float eps,
                    int x_dtype, int r_dtype, int w_dtype, int y_dtype, void* stream) {
  return launch_rmsnorm_fwd(x, r, w, s_out, y, rstd, rows, dim, eps, x_dtype, r_dtype, w_dtype, y_dtype, (cudaStream_t)stream);
}

int nsa_rmsnorm_bwd(const void* dy, const void* s, const void* w, const float* rstd, const void* ds, void* dx, void* dw,
                    float* dw_partial, int rows, int dim, int x_dtype, int w_dtype, int y_dtype, void* stream) {
  return launch_rmsnorm_bwd(dy, s, w, rstd, ds, dx, dw, dw_partial, rows, dim, x_dtype, w_dtype, y_dtype, (cudaStream_t)stream);
}

int nsa_rmsnorm_partials(int rows) { return rmsnorm_partials(rows); }

int nsa_stats(const float* gates, int64_t n_gate_rows, const int32_t* ranges, int64_t n_range_rows, int K, int32_t* row_len,
              nsa_stats_t* out, void* stream) {
  return launch_stats(gates, (long long)n_gate_rows, ranges, (long long)n_range_rows, K, row_len, out, (cudaStream_t)stream);
}

int nsa_ranges_max_blocks(const int32_t* ranges, int64_t n_rows, int K, int S_kv, int32_t* max_blocks, void* stream) {
  return launch_ranges_max_blocks(ranges, (long long)n_rows, K, S_kv, max_blocks, (cudaStream_t)stream);
}

int64_t nsa_workspace_bytes(const nsa_dims_t* dm, int which) {
  if (!dm) return 0;
  switch (which) {
    case NSA_WS_DECODE:
      return (((int64_t)dm->B * dm->G * dm->n_sel * 2 * sizeof(int32_t) + 15) & ~(int64_t)15) + tc_decode_workspace(*dm);
    case NSA_WS_SCORE_SELECT:
      return tc_score_workspace(*dm);
    case NSA_WS_PREFILL: {
      for (int br = 0; br < 3; ++br)
        if (tc_branch_supported(*dm, br)) {
          const int64_t staging = ((int64_t)3 * dm->B * dm->S * dm->G * dm->h * dm->Dv * (int64_t)elt_size(dm->dtype) + 255) & ~(int64_t)255;
          return staging + (use_sel2(*dm) ? tc_sel2_workspace(*dm) : 0);
        }
      return 0;
    }
    case NSA_WS_PREFILL_FULL:
      return ((tc_score_workspace(*dm) + 255) & ~(int64_t)255) + (full_fused(*dm) ? tc_score_cmp_stats_bytes(*dm) + full_gate_bytes(*dm) : 0) +
             nsa_workspace_bytes(dm, NSA_WS_PREFILL);
    case NSA_WS_SEL_BLOCKMAJOR:
      return tc_sel2_workspace(*dm);
    case NSA_WS_BWD:
      return tc_bwd_workspace(*dm);
    default:
      return 0;
  }
}

}  // extern "C"
