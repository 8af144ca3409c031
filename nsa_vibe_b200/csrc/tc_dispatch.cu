// Dispatch into the tcgen05 / TMA kernel family.  Shapes outside their specialisation fall to the SIMT
// kernels in generic.cu (same device, same semantics) -- never to a CPU path.
#include "launchers.h"

namespace nsa {

bool tc_supported(const nsa_dims_t& dm) { (void)dm; return false; }
bool tc_score_supported(const nsa_dims_t& dm) { (void)dm; return false; }
bool tc_decode_supported(const nsa_dims_t& dm) { (void)dm; return false; }
int64_t tc_score_workspace(const nsa_dims_t& dm) { (void)dm; return 0; }
int64_t tc_decode_workspace(const nsa_dims_t& dm) { (void)dm; return 0; }

int launch_score_tc(const nsa_dims_t&, const void*, const void*, int, int, int, int, float*, int32_t*, void*, cudaStream_t) {
  set_error("tcgen05 scorer not built");
  return NSA_ERR_UNSUPPORTED;
}
int launch_branch_tc(const nsa_dims_t&, int, const void*, const void*, const void*, const int32_t*, void*, float*,
                     cudaStream_t) {
  set_error("tcgen05 branch kernel not built");
  return NSA_ERR_UNSUPPORTED;
}
int launch_prefill_tc(const nsa_dims_t&, const void*, const void*, const void*, const void*, const void*, const void*,
                      const void*, const int32_t*, const nsa_gate_params_t&, void*, float*, float*, void*, cudaStream_t) {
  set_error("tcgen05 prefill kernel not built");
  return NSA_ERR_UNSUPPORTED;
}
int launch_decode_tc(const nsa_dims_t&, const void*, const void*, const void*, const void*, const void*, const void*,
                     const void*, const nsa_gate_params_t&, void*, int32_t*, void*, cudaStream_t) {
  set_error("tcgen05 decode kernel not built");
  return NSA_ERR_UNSUPPORTED;
}

}  // namespace nsa
