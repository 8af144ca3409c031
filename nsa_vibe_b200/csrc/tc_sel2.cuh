// Inverted index of the selected branch (built on the device by tc_sel2.cu): for every 64-key block of every (b, g) slab
// the list of query rows whose ranges contain it, padded to whole M-tiles, plus a run table (one CTA per run).
// Shared by the block-major forward (tc_sel2.cu) and the tensor-core backward (tc_bwd.cu).
#pragma once
#include "common.cuh"

namespace nsa {

constexpr int kS2MaxSlots = 16;   // 64-key blocks per row (n_sel * l_sel <= 1024 keys)
constexpr int kS2Run = 24;        // M-tiles one CTA walks (same block: its K/V tile is loaded once)

struct S2Geom {
  int B, S, G, n_ranges, S_kv, NB, t0, tokp;  // NB = 64-key blocks per slab, tokp = queries per M-tile
};
// run = (list = bg * NB + block, first M-tile, number of M-tiles <= kS2Run)
struct S2Run { int list, tile0, ntiles, pad; };

// pair p = tile * tokp + position: tok[p] = b*S + s of the query (-1 = padding), hi[p] = valid keys of the block for it
struct S2Index {
  S2Geom gm;
  const S2Run* runs;
  const int* n_runs;   // device scalar
  const int* tok;
  const int* hi;
  int max_runs;        // upper bound on *n_runs (grid size)
};

bool sel2_index_supported(const nsa_dims_t& dm);
int64_t sel2_index_workspace(const nsa_dims_t& dm);
// Builds the index into `workspace` (sel2_index_workspace bytes); pointers in *out refer to that workspace.
int sel2_build_index(const nsa_dims_t& dm, const int32_t* ranges, void* workspace, cudaStream_t stream, S2Index* out);

}  // namespace nsa
