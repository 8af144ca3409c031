// GateMLP forward evaluated by one warp (shared by the SIMT kernels and the fused decode kernel).
#pragma once
#include "common.cuh"

namespace nsa {

// ------------------------------------------------------------------------------------------------
// Gate MLP (nsa/core/nsa_attention.py:32-82) evaluated by one warp.  qgp: smem float[Dk] (mean over heads),
// xs: smem float[hidden] scratch.  Returns the three gate probabilities in every lane.
// ------------------------------------------------------------------------------------------------
struct Gate3 { float c, s, w; };

__device__ inline Gate3 gate_forward_warp(const float* qgp, float* xs, float* pre, const nsa_gate_params_t& gp, int Dk,
                                          int hidden, float tau, int mode, bool* peaked_out) {
  const int lane = threadIdx.x & 31;
  if (peaked_out) *peaked_out = false;
  if (mode == NSA_GATE_UNIFORM) return {1.0f / 3.0f, 1.0f / 3.0f, 1.0f / 3.0f};
  if (mode == NSA_GATE_CMP) return {1.f, 0.f, 0.f};
  if (mode == NSA_GATE_SEL) return {0.f, 1.f, 0.f};
  if (mode == NSA_GATE_WIN) return {0.f, 0.f, 1.f};
  for (int u = lane; u < hidden; u += 32) {
    float a = gp.fc1_b ? gp.fc1_b[u] : 0.f;
    const float* wrow = gp.fc1_w + (size_t)u * Dk;
    for (int k = 0; k < Dk; ++k) a = fmaf(wrow[k], qgp[k], a);
    if (pre) pre[u] = a;
    xs[u] = a / (1.0f + expf(-a));  // silu
  }
  __syncwarp();
  float g[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float a = 0.f;
    for (int u = lane; u < hidden; u += 32) a = fmaf(gp.fc2_w[c * hidden + u], xs[u], a);
    a = warp_sum(a);
    g[c] = (a + (gp.fc2_b ? gp.fc2_b[c] : 0.f)) / fmaxf(tau, 1e-6f);
  }
  float mx = fmaxf(g[0], fmaxf(g[1], g[2]));
  int am = g[0] >= g[1] ? (g[0] >= g[2] ? 0 : 2) : (g[1] >= g[2] ? 1 : 2);  // first maximum
  float second = am == 0 ? fmaxf(g[1], g[2]) : (am == 1 ? fmaxf(g[0], g[2]) : fmaxf(g[0], g[1]));
  if (mx - second > 50.0f) {  // hard one-hot (nsa_attention.py:74-81)
    if (peaked_out) *peaked_out = true;
    return {am == 0 ? 1.f : 0.f, am == 1 ? 1.f : 0.f, am == 2 ? 1.f : 0.f};
  }
  float e0 = expf(g[0] - mx), e1 = expf(g[1] - mx), e2 = expf(g[2] - mx);
  float inv = 1.0f / (e0 + e1 + e2);
  return {e0 * inv, e1 * inv, e2 * inv};
}

}  // namespace nsa
