// icnn_tc.cu -- tcgen05 (TF32 / BF16 / 3xTF32) variants of the fused ICNN decode.  Placeholder until the
// tensor-core kernel lands: every entry reports "unsupported" so callers fail loudly.
#include "common.cuh"
namespace b200vae {
size_t tc_extra_ws_floats(int, int, int) { return 0; }
int tc_prepare(const b200vae_icnn_params*, int, int, int, int, float*, cudaStream_t) { return B200VAE_EUNSUP; }
int tc_fwd(const float*, int, int, int, float, float*, float*, uint32_t*, uint8_t*, int, const float*, cudaStream_t) {
  return B200VAE_EUNSUP;
}
}  // namespace b200vae
