// Forward-mode tangents through the fused path (JacobianLayer replacement) -- filled in below.
#include "nif_tile.cuh"

int nif_tangent_impl(const Plan& pl, long long B, const float* z, const float* x, const float* packed, int n_dir,
                     const float* zdot, const float* xdot, float* u, float* udot, cudaStream_t st) {
  nif_set_error("nif_forward_tangent: not built yet");
  return NIF_E_UNSUPPORTED;
}
