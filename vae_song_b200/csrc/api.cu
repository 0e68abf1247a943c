// api.cu -- extern "C" entry points of libb200vae.so for the ICNN path (validation + dispatch).
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "tc_common.cuh"

namespace b200vae {
int g_last_cuda_error = 0;
long long g_launch_count = 0;

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      n = 148;
  }
  return n;
}

int simt_prepare(const b200vae_icnn_params* p, int d, int H, int mode, float* ws, cudaStream_t st);
int simt_fwd(const float* z, int B, int d, int H, float kappa, float* psi, float* xhat, uint32_t* mask1,
             uint8_t* mask2, const float* ws, cudaStream_t st);
int simt_bwd(const float* z, const float* v, const float* gpsi, const uint32_t* mask1, const uint8_t* mask2,
             int B, int d, int H, const b200vae_icnn_params* p, int mode, float kappa,
             const b200vae_icnn_grads* g, float* dz, float* ws, size_t mid_extra, cudaStream_t st);
// tensor-core (tcgen05) variants, icnn_tc.cu
int tc_prepare(const b200vae_icnn_params* p, int d, int H, int mode, int precision, float* ws, cudaStream_t st);
int tc3_prepare(int d, int H, int precision, float* ws, cudaStream_t st);
int tc3_fwd(const float* z, int B, int d, int H, float kappa, float* psi, float* xhat, uint32_t* mask1, uint8_t* mask2,
            int precision, float* ws, float* accsave, cudaStream_t st);
size_t tc_extra_ws_floats(int B, int d, int H, int precision);
size_t tc_bwd_ws_floats(int B, int d, int H);
int tc_bwd(const float* z, const float* v, const uint32_t* mask1, const uint8_t* mask2, int B, int d, int H,
           const b200vae_icnn_params* p, int mode, float kappa, const b200vae_icnn_grads* g, float* dz, int precision,
           float* ws, const float* accsave, int phase, cudaStream_t st);

// ---- saved GEMM2 accumulators (training, 3xTF32) ------------------------------------------------------------------------
// With a for_backward workspace and both masks requested, the pair forward kernel also stores its GEMM2 accumulators
// (gx1 / s2, [Bp][Hq] fp32) at the end of the workspace; the backward then skips recomputing that GEMM (icnn_tc3.cu, SV
// kernels).  Which workspace holds a valid save -- and for which (z, B, d, H) -- is tracked HERE on the host, in call
// order (= stream order for one stream; a CUDA-graph capture records the same decision it replays): prepare invalidates,
// a saving forward validates, the backward uses the save only on an exact match and otherwise recomputes.
struct SavedAcc { const void* z; int B, d, H; };
static std::mutex g_save_mu;
static std::unordered_map<const void*, SavedAcc> g_saved;
static bool save_enabled() {
  static const bool on = [] { const char* e = getenv("B200VAE_SAVE_GX1"); return !e || atoi(e) != 0; }();
  return on;
}
static void save_forget(const void* ws) {
  std::lock_guard<std::mutex> lk(g_save_mu);
  g_saved.erase(ws);
}
static void save_record(const void* ws, const void* z, int B, int d, int H) {
  std::lock_guard<std::mutex> lk(g_save_mu);
  if (g_saved.size() > 1024) g_saved.clear();
  g_saved[ws] = SavedAcc{z, B, d, H};
}
static bool save_matches(const void* ws, const void* z, int B, int d, int H) {
  std::lock_guard<std::mutex> lk(g_save_mu);
  auto it = g_saved.find(ws);
  return it != g_saved.end() && it->second.z == z && it->second.B == B && it->second.d == d && it->second.H == H;
}
static size_t bwd_floats_without_save(int B, int d, int H, int precision) {
  const size_t extra = (precision != B200VAE_PREC_FP32) ? tc_extra_ws_floats(B, d, H, precision) : 0;
  const WsLayout L = ws_layout(B, d, H, extra);
  return (L.end + (extra ? tc_bwd_ws_floats(B, d, H) : 0) + 63) / 64 * 64;
}
static size_t accsave_floats(int B, int d, int H, int precision) {
  if ((precision != B200VAE_PREC_TF32X3 && precision != B200VAE_PREC_F16X3) || d > 3 || !save_enabled()) return 0;
  const Tc3Layout T3 = tc3_layout(B, d, H);
  return (size_t)T3.Bp * T3.Hq + 64;
}

static bool params_ok(const b200vae_icnn_params* p) {
  return p && p->A0w && p->A0b && p->A1w && p->A1b && p->A2w && p->A2b && p->W0 && p->W1 && aligned4(p->A0w) &&
         aligned4(p->W0);
}
static int shape_ok(int B, int d, int H) {
  if (B <= 0 || d <= 0 || H <= 0) return B200VAE_ESHAPE;
  if (d > 4) return B200VAE_EUNSUP;      // fused small-d path; wider inputs go through the tiled path
  if (H > 4096) return B200VAE_EUNSUP;
  return B200VAE_OK;
}
static bool prec_ok(int precision) {
  return precision == B200VAE_PREC_FP32 || precision == B200VAE_PREC_TF32 || precision == B200VAE_PREC_TF32X3 ||
         precision == B200VAE_PREC_F16X3;   // 2 is reserved
}
}  // namespace b200vae

using namespace b200vae;

extern "C" size_t b200vae_icnn_workspace_bytes(int B, int d, int H, int precision, int for_backward) {
  if (B <= 0 || d <= 0 || H <= 0) return 0;
  const size_t extra = (precision != B200VAE_PREC_FP32) ? tc_extra_ws_floats(B, d, H, precision) : 0;
  const WsLayout L = ws_layout(B, d, H, extra);
  const size_t fl = for_backward ? bwd_floats_without_save(B, d, H, precision) + accsave_floats(B, d, H, precision)
                                 : L.fwd_end + extra + 64;
  return fl * sizeof(float) + 256;
}

static float* ws_base(const void* ws) {
  return reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
}

extern "C" int b200vae_icnn_prepare(const b200vae_icnn_params* p, int d, int H, int weight_mode, int precision,
                                    void* ws, size_t ws_bytes, void* stream) {
  if (!params_ok(p) || !ws) return B200VAE_EALIGN;
  int rc = shape_ok(1, d, H);
  if (rc) return rc;
  if (weight_mode != B200VAE_WEIGHT_EXP && weight_mode != B200VAE_WEIGHT_CLAMP) return B200VAE_EUNSUP;
  if (!prec_ok(precision)) return B200VAE_EUNSUP;
  if (ws_bytes < b200vae_icnn_workspace_bytes(1, d, H, precision, 0)) return B200VAE_EWS;
  save_forget(ws_base(ws));
  rc = simt_prepare(p, d, H, weight_mode, ws_base(ws), (cudaStream_t)stream);
  if (rc || precision == B200VAE_PREC_FP32) return rc;
  rc = tc_prepare(p, d, H, weight_mode, precision, ws_base(ws), (cudaStream_t)stream);
  if (rc) return rc;
  return tc3_prepare(d, H, precision, ws_base(ws), (cudaStream_t)stream);
}

extern "C" int b200vae_icnn_decode_fwd(const float* z, int B, int d, int H, int weight_mode, float kappa, float* psi,
                                       float* xhat, uint32_t* mask1, uint8_t* mask2, int precision, const void* ws,
                                       size_t ws_bytes, void* stream) {
  (void)weight_mode;
  if (!z || !ws || !aligned4(z)) return B200VAE_EALIGN;
  int rc = shape_ok(B, d, H);
  if (rc) return rc;
  if (!prec_ok(precision)) return B200VAE_EUNSUP;
  if (ws_bytes < b200vae_icnn_workspace_bytes(B, d, H, precision, 0)) return B200VAE_EWS;
  if (precision == B200VAE_PREC_FP32)
    return simt_fwd(z, B, d, H, kappa, psi, xhat, mask1, mask2, ws_base(ws), (cudaStream_t)stream);
  // persistent CTA pairs (icnn_tc3.cu); B200VAE_EUNSUP where they do not take the shape (d > 3, H > 1024)
  float* accsave = nullptr;
  if (mask1 && mask2 && xhat && accsave_floats(B, d, H, precision) &&
      ws_bytes >= b200vae_icnn_workspace_bytes(B, d, H, precision, 1))
    accsave = ws_base(ws) + bwd_floats_without_save(B, d, H, precision);
  save_forget(ws_base(ws));
  rc = tc3_fwd(z, B, d, H, kappa, psi, xhat, mask1, mask2, precision, ws_base(ws), accsave, (cudaStream_t)stream);
  if (rc == B200VAE_OK && accsave) save_record(ws_base(ws), z, B, d, H);
  return rc;
}

extern "C" int b200vae_icnn_decode_bwd(const float* z, const float* v, const float* gpsi, const uint32_t* mask1,
                                       const uint8_t* mask2, int B, int d, int H, const b200vae_icnn_params* p,
                                       int weight_mode, float kappa, const b200vae_icnn_grads* grads, float* dz,
                                       int precision, void* ws, size_t ws_bytes, void* stream) {
  if (!z || !mask1 || !mask2 || !ws || !params_ok(p)) return B200VAE_EALIGN;
  if (!v && !gpsi) return B200VAE_ESHAPE;
  int rc = shape_ok(B, d, H);
  if (rc) return rc;
  if (!prec_ok(precision)) return B200VAE_EUNSUP;
  if (ws_bytes < b200vae_icnn_workspace_bytes(B, d, H, precision, 1)) return B200VAE_EWS;
  if (precision != B200VAE_PREC_FP32 && !gpsi) { // tensor-core backward (gpsi path stays on the FP32 kernels)
    const float* accsave = nullptr;
    if (accsave_floats(B, d, H, precision) && save_matches(ws_base(ws), z, B, d, H))
      accsave = ws_base(ws) + bwd_floats_without_save(B, d, H, precision);
    return tc_bwd(z, v, mask1, mask2, B, d, H, p, weight_mode, kappa, grads, dz, precision, ws_base(ws), accsave, 0,
                  (cudaStream_t)stream);
  }
  const size_t extra = (precision != B200VAE_PREC_FP32) ? tc_extra_ws_floats(B, d, H, precision) : 0;
  return simt_bwd(z, v, gpsi, mask1, mask2, B, d, H, p, weight_mode, kappa, grads, dz, ws_base(ws), extra,
                  (cudaStream_t)stream);
}

extern "C" int b200vae_icnn_decode_bwd_params(const float* z, const float* v, const uint32_t* mask1, const uint8_t* mask2,
                                              int B, int d, int H, const b200vae_icnn_params* p, int weight_mode,
                                              const b200vae_icnn_grads* grads, int precision, void* ws, size_t ws_bytes,
                                              void* stream) {
  if (!z || !v || !mask1 || !mask2 || !ws || !grads || !params_ok(p)) return B200VAE_EALIGN;
  int rc = shape_ok(B, d, H);
  if (rc) return rc;
  if (!prec_ok(precision) || precision == B200VAE_PREC_FP32) return B200VAE_EUNSUP;
  if (ws_bytes < b200vae_icnn_workspace_bytes(B, d, H, precision, 1)) return B200VAE_EWS;
  return tc_bwd(z, v, mask1, mask2, B, d, H, p, weight_mode, 0.f, grads, nullptr, precision, ws_base(ws), nullptr, 2,
                (cudaStream_t)stream);
}

extern "C" int b200vae_last_cuda_error(void) { return g_last_cuda_error; }
extern "C" const char* b200vae_version(void) { return "b200vae 0.2 (sm_100a)"; }
extern "C" long long b200vae_launch_count(void) { return g_launch_count; }
