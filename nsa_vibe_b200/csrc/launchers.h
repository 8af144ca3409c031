// Internal launcher declarations shared by api.cu and the kernel translation units.
#pragma once
#include "common.cuh"

namespace nsa {

struct FwdArgs {
  const void *Q, *K[3], *V[3];  // index = branch: 0 cmp, 1 sel, 2 win
  const int32_t* ranges;
  nsa_gate_params_t gp;
  const float* gates_in;  // optional precomputed gates [rows,3]
  void* O;                // combined output (may be NULL)
  float* lse;             // [3][rows][h] (may be NULL)
  float* gates_out;       // [rows,3] (may be NULL)
  void* O_br;             // [3][rows][h][Dv] (may be NULL)
  int branch_mask;
};

struct BwdArgs {
  const void *Q, *K[3], *V[3];
  const int32_t* ranges;
  const void* O_br;     // [3][rows][h][Dv] saved branch outputs
  const float* lse;     // [3][rows][h]
  const float* gates;   // [rows][3]; NULL -> every enabled branch has weight 1 (single-branch API)
  const void* dO;       // [rows][h][Dv]
  float *dQ, *dK[3], *dV[3];
  float* dgates;        // [rows][3] (may be NULL)
  int branch_mask;
};

// select.cu
int launch_select(const float* p_grp, int n_rows, int S_rows, int G, int S_sel, int l_sel, int n_sel, int mode, int nf,
                  int K, int t0, int32_t* ranges, cudaStream_t stream);
// generic.cu (SIMT)
int launch_score_generic(const nsa_dims_t& dm, const void* Q, const void* Kc, int S_sel, int S_total, int sel_mode, int Kr,
                         float* p_grp, int32_t* ranges, cudaStream_t stream);
int launch_fwd_generic(const nsa_dims_t& dm, const FwdArgs& a, cudaStream_t stream);
int launch_bwd_generic(const nsa_dims_t& dm, const BwdArgs& a, cudaStream_t stream);
int launch_pcmp_all(const nsa_dims_t& dm, const void* Q, const void* Kc, float* p_cmp, cudaStream_t stream);
int launch_map_pslc(const float* p_cmp, long long n_rows, int S_cmp, int S_sel, int l, int d, int l_sel, float* p_slc, cudaStream_t stream);
int launch_indices_to_ranges(const int32_t* indices, int B, int S, int G, int K, int S_sel, int l_sel, int t0, int32_t* ranges,
                             cudaStream_t stream);
int launch_gate_fwd(const nsa_dims_t& dm, const void* Q, const nsa_gate_params_t& gp, float* gates, cudaStream_t stream);
int launch_gate_bwd(const nsa_dims_t& dm, const void* Q, const nsa_gate_params_t& gp, const float* dgates, float* dQ,
                    float* d_fc1_w, float* d_fc1_b, float* d_fc2_w, float* d_fc2_b, cudaStream_t stream);
int launch_combine(const nsa_dims_t& dm, const void* Q, const nsa_gate_params_t& gp, const void* O_br, void* O, float* gates,
                   cudaStream_t stream);
// producers.cu (RoPE + re-layout, phi average pool)
int launch_rope_shape(const void* x, void* y, int B, int S, int V, int D, int src_layout, int dst_layout, int rot_dim, int t0,
                      float base, float scale, int inverse, int dtype, cudaStream_t stream);
int launch_phi_avgpool(const void* x, void* y, int BG, int S, int D, int l, int d, int rope, int t0, float base, float scale,
                       int backward, int dtype, cudaStream_t stream, const float* w = nullptr, const void* dy_for_dw = nullptr);
int launch_rope_table(int rows, int pairs, int rot_dim, int t0, float base, float scale, int dtype, void* out, cudaStream_t stream);
int launch_decode_produce(const nsa_decode_produce_t& a, cudaStream_t stream);
int launch_decode_emit(const nsa_decode_emit_t& a, cudaStream_t stream);
int launch_decode_advance(nsa_decode_state_t* state, int l, int d, cudaStream_t stream);
// stats.cu
int launch_stats(const float* gates, long long n_gate_rows, const int32_t* ranges, long long n_range_rows, int K, int32_t* row_len,
                 nsa_stats_t* out, cudaStream_t stream);
int launch_ranges_max_blocks(const int32_t* ranges, long long n_rows, int K, int S_kv, int32_t* out, cudaStream_t stream);
// block_ops.cu
int launch_rmsnorm_fwd(const void* x, const void* r, const void* w, void* s_out, void* y, float* rstd, int rows, int dim, float eps,
                       int x_dtype, int r_dtype, int w_dtype, int y_dtype, cudaStream_t stream);
int launch_rmsnorm_bwd(const void* dy, const void* s, const void* w, const float* rstd, const void* ds, void* dx, void* dw,
                       float* dw_partial, int rows, int dim, int x_dtype, int w_dtype, int y_dtype, cudaStream_t stream);
int rmsnorm_partials(int rows);
// tc_*.cu (tcgen05 / TMA kernels)
bool tc_branch_supported(const nsa_dims_t& dm, int branch);
bool tc_score_supported(const nsa_dims_t& dm);
bool tc_decode_supported(const nsa_dims_t& dm);
bool tc_sel2_supported(const nsa_dims_t& dm);
int64_t tc_sel2_workspace(const nsa_dims_t& dm);
// fuse != NULL: instead of writing the selected branch's O / lse, merge the partials, evaluate the gate and write the gated
// combination with the given compressed / sliding outputs (no-grad prefill: O_sel and the gates stay on chip)
struct Sel2Fuse {
  const nsa_gate_params_t* gp;
  const void *O_cmp, *O_win;
  void* O;
  float* gates;  // [rows,3] (may be NULL)
  const float* gates_in = nullptr;  // gates already evaluated (launch_gate_fast, on a side stream): merge + blend only
};
bool sel2_fuse_supported(const nsa_dims_t& dm);
bool gate_fast_supported(const nsa_dims_t& dm, const nsa_gate_params_t* gp, const void* Q);
int launch_gate_fast(const nsa_dims_t& dm, const void* Q, const nsa_gate_params_t& gp, float* gates, cudaStream_t stream);
int launch_sel2_tc(const nsa_dims_t& dm, const void* Q, const void* K, const void* V, const int32_t* ranges, void* O, float* lse,
                   void* workspace, cudaStream_t stream, const Sel2Fuse* fuse = nullptr);
// tensor-core backward (tc_bwd.cu); falls back to launch_bwd_generic per branch when a shape has no tcgen05 kernel or
// workspace is NULL
bool tc_bwd_supported(const nsa_dims_t& dm, int branch);
int64_t tc_bwd_workspace(const nsa_dims_t& dm);
int launch_bwd_tc(const nsa_dims_t& dm, const BwdArgs& a, void* workspace, cudaStream_t stream);
int64_t tc_score_workspace(const nsa_dims_t& dm);
int64_t tc_decode_workspace(const nsa_dims_t& dm);
int launch_score_tc(const nsa_dims_t& dm, const void* Q, const void* Kc, int S_sel, int S_total, int sel_mode, int Kr,
                    float* p_grp, int32_t* ranges, void* workspace, cudaStream_t stream);
// pass 1 of the scorer alone / pass 2 fused with the compressed branch (tc_score_cmp.cu)
int launch_score_stats_tc(const nsa_dims_t& dm, const void* Q, const void* Kc, float* stats, cudaStream_t stream);
bool tc_score_cmp_supported(const nsa_dims_t& dm);
int64_t tc_score_cmp_stats_bytes(const nsa_dims_t& dm);
int launch_score_cmp_tc(const nsa_dims_t& dm, const void* Q, const void* Kc, const void* Vc, int S_sel, const float* stats,
                        float* p_grp, void* O_cmp, float* lse_cmp, cudaStream_t stream);
int launch_branch_tc(const nsa_dims_t& dm, int branch, const void* Q, const void* K, const void* V, const int32_t* ranges,
                     void* O_b, float* lse_b, cudaStream_t stream);
int launch_decode_tc(const nsa_dims_t& dm, const void* Q, const void* K_sel, const void* V_sel, const void* K_win,
                     const void* V_win, const void* K_cmp, const void* V_cmp, const nsa_gate_params_t& gp, void* O,
                     int32_t* ranges_out, void* workspace, cudaStream_t stream);
bool tc_decode_stepped_supported(const nsa_dims_t& dm);
int launch_decode_tc_stepped(const nsa_dims_t& dm, const void* Q, const void* K_sel, const void* V_sel, const void* K_win,
                             const void* V_win, const void* K_cmp, const void* V_cmp, const nsa_gate_params_t& gp, void* O,
                             int32_t* ranges_out, const nsa_decode_state_t* state, cudaStream_t stream);

}  // namespace nsa
