// icnn_tc.cu -- host side of the tensor-core (tcgen05) ICNN path that is shared by the CTA-pair kernels of icnn_tc3.cu:
// the per-unit parameter tables (A0q, A1q, P1q), the workspace layout behind the FP32 arrays, the backward driver
// (rows kernel -> dP0 kernel -> finalize kernels, in one call or as two phases) and the ordered finalize of the row
// partials.  The first-generation single-CTA kernels that lived here (A/B partners of the pair kernels in round 1) are gone:
// the pair kernels cover every shape the tensor-core path accepts (d <= 3, H <= 1024).
#include <cstdlib>

#include "common.cuh"

#include "tc_common.cuh"

namespace b200vae {

size_t tc_extra_ws_floats(int B, int d, int H, int precision) {
  (void)precision;
  return tc_layout(d, H).end + tc3_layout(B, d, H).end;      // parameter tables, then the pair kernels' arrays
}

// one thread per hidden unit: (A0w, A0b), (A1w | P1 in the spare lane, A1b), P1 in natural unit order, zero padded to Hq
__global__ void tc_prepare_kernel(const float* __restrict__ P1, const float* __restrict__ A0p, const float* __restrict__ A1p,
                                  int d, int Hp, int Hq, float4* __restrict__ A0q, float4* __restrict__ A1q,
                                  float* __restrict__ P1q) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Hq) return;
  const bool inr = c < Hp;
  float w[4] = {0.f, 0.f, 0.f, 0.f}, u[4] = {0.f, 0.f, 0.f, 0.f};
  if (inr) {
    for (int j = 0; j < d; ++j) { w[j] = A0p[(size_t)c * (d + 1) + j]; u[j] = A1p[(size_t)c * (d + 1) + j]; }
    w[3] = A0p[(size_t)c * (d + 1) + d]; u[3] = A1p[(size_t)c * (d + 1) + d];
  }
  if (d <= 2) u[2] = inr ? P1[c] : 0.f;     // spare lane of the float4: P1 rides along
  A0q[c] = make_float4(w[0], w[1], w[2], w[3]);
  A1q[c] = make_float4(u[0], u[1], u[2], u[3]);
  P1q[c] = inr ? P1[c] : 0.f;
}


// ordered reduction of the row-kernel partials + chain through the positive reparam for W1.
// grid (Hq/32, NF, 2): one block of 1024 threads sums all `nslots` partial rows of 32 columns (32 slot lanes x 32 columns,
// coalesced 128-byte reads, fixed summation order -> deterministic).
__global__ void __launch_bounds__(1024)
tc_finalize_small_kernel(const float* __restrict__ partA, const float* __restrict__ partB,
                         const float* __restrict__ a2part, int nslots, int nmt, int d, int H, int Hq,
                         const float* __restrict__ P1, const float* __restrict__ W1raw, int mode, b200vae_icnn_grads g) {
  __shared__ float red[32][33];
  const int NF = d + 1;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx, f = blockIdx.y, which = blockIdx.z;
  const float* part = (which ? partB : partA) + (size_t)f * Hq + n;
  // 16 independent partial sums per thread, 32 slot lanes per block: the kernel is latency bound (2048 slots x 128-byte
  // rows per block at B = 65536), so the number of loads in flight is what sets its duration (8 slot lanes: 21 us)
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  int sl = ty;
  const size_t stride = (size_t)NF * Hq;
  for (; sl + 480 < nslots; sl += 512) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] += part[(size_t)(sl + 32 * i) * stride];
  }
  for (; sl < nslots; sl += 32) acc[0] += part[(size_t)sl * stride];
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1)
#pragma unroll
    for (int i = 0; i < w; ++i) acc[i] += acc[i + w];
  red[ty][tx] = acc[0];
  __syncthreads();
  if (ty == 0 && n < H) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 32; ++q) s += red[q][tx];
    if (!which) {
      if (f < d) { if (g.A0w) g.A0w[(size_t)n * d + f] = s; }
      else if (g.A0b) g.A0b[n] = s;
    } else {
      if (f < d) { if (g.A1w) g.A1w[(size_t)n * d + f] = s; }
      else if (g.W1) g.W1[n] = (mode == B200VAE_WEIGHT_EXP) ? s * P1[n] : (W1raw[n] >= kClampMin ? s : 0.f);
      if (f == 0 && g.A1b) g.A1b[n] = 0.f;                   // exact zeros on the <v, xhat> path (Appendix A)
    }
  }
  if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
    if (threadIdx.x < d) {
      float s = 0.f;
      for (int mt = 0; mt < nmt; ++mt) s += a2part[(size_t)mt * d + threadIdx.x];
      if (g.A2w) g.A2w[threadIdx.x] = s;
    }
    if (threadIdx.x == 0 && g.A2b) g.A2b[0] = 0.f;
  }
}

// ------------------------------------------------------------------------------------ host side
int tc_prepare(const b200vae_icnn_params* p, int d, int H, int mode, int precision, float* ws, cudaStream_t st) {
  (void)p; (void)mode;
  if (precision == 2 /* reserved */) return B200VAE_EUNSUP;
  if (d > 3) return B200VAE_EUNSUP;
  const WsLayout L = ws_layout(1, d, H);
  const TcLayout T = tc_layout(d, H);
  float* tb = tc_base(ws, d, H);
  tc_prepare_kernel<<<(T.Hq + 255) / 256, 256, 0, st>>>(ws + L.P1, ws + L.A0p, ws + L.A1p, d, L.Hp, T.Hq,
                                                       reinterpret_cast<float4*>(tb + T.A0q),
                                                       reinterpret_cast<float4*>(tb + T.A1q), tb + T.P1q);
  return check_launch();
}

// ---- backward ----
static int tc_dp0_splits(int B, int Hq) {
  const int tiles = (Hq / kTM) * (Hq / kTN);
  int s = 148 / tiles;                          // one wave of long-running CTAs
  const int maxs = (B + 255) / 256;
  if (s > maxs) s = maxs;
  if (s < 1) s = 1;
  return s;
}

int tc3_bwd_rows(const float* z, const float* v, const uint32_t* mask1, const uint8_t* mask2, int B, int d, int H, float kappa,
                 float* dz, float* partA, float* partB, float* a2part, float* dzpart, int precision, float* ws,
                 const float* accsave, cudaStream_t st);
size_t tc3_bwd_ws_floats(int B, int d, int H);
int tc3_dp0(const float* z, const float* v, const uint32_t* mask1, const uint8_t* mask2, int B, int d, int Hq, int Hw_in,
            const float* A0q, const float* sumV, int precision, int max_splits, float* part, int* splits_out, cudaStream_t st);

size_t tc_bwd_ws_floats(int B, int d, int H) {
  const TcLayout T = tc_layout(d, H);
  const size_t nmt = (size_t)(B + kTM - 1) / kTM;
  return 2 * nmt * 8 * (d + 1) * T.Hq + nmt * 4 + 64 + 16 + (size_t)tc_dp0_splits(B, T.Hq) * T.Hq * T.Hq + 64 +
         tc3_bwd_ws_floats(B, d, H);
}

int finalize_W0_launch(const float* part, int splits, int H, int Hp, int ldp, const float* P0, const float* P1,
                       const float* W0raw, int mode, float scale, float* dW0, cudaStream_t st);

// phase 0: everything; 1: the rows part only (dz + the ordered column partials in the workspace); 2: the parameter
// gradients only (dP0 + the finalize kernels), from the partials a phase-1 call left in the SAME workspace
int tc_bwd(const float* z, const float* v, const uint32_t* mask1, const uint8_t* mask2, int B, int d, int H,
           const b200vae_icnn_params* p, int mode, float kappa, const b200vae_icnn_grads* g, float* dz, int precision,
           float* ws, const float* accsave, int phase, cudaStream_t st) {
  if (precision == 2 /* reserved */ || d > 3 || !v) return B200VAE_EUNSUP;
  const size_t extra = tc_extra_ws_floats(B, d, H, precision);
  const WsLayout L = ws_layout(B, d, H, extra);
  const TcLayout T = tc_layout(d, H);
  const float* tb = tc_base(ws, d, H);
  const size_t nmt = (size_t)(B + kTM - 1) / kTM;
  float* partA = ws + L.end;
  float* partB = partA + nmt * 8 * (d + 1) * T.Hq;
  float* a2part = partB + nmt * 8 * (d + 1) * T.Hq;
  const int Hw_in = L.Hp / 32;
  float* dp0part = a2part + nmt * 4 + 64;
  dp0part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dp0part) + 15) & ~(uintptr_t)15);
  int rc = B200VAE_OK;
  if (phase != 2) {       // rows part: persistent pair kernel (icnn_tc3.cu); EUNSUP where it does not take the shape (H > 1024)
    float* dzpart = dp0part + (size_t)tc_dp0_splits(B, T.Hq) * T.Hq * T.Hq;
    dzpart = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dzpart) + 15) & ~(uintptr_t)15);
    rc = tc3_bwd_rows(z, v, mask1, mask2, B, d, H, kappa, dz, partA, partB, a2part, dzpart, precision, ws, accsave, st);
  }
  if (rc || !g || phase == 1) return rc;
  if (g->W0) {
    float* part = dp0part;
    int splits = tc_dp0_splits(B, T.Hq);          // the pair kernel never needs more slabs than the workspace holds
    rc = tc3_dp0(z, v, mask1, mask2, B, d, T.Hq, Hw_in, tb + T.A0q, tb + T.end + tc3_layout(1, d, H).sumV, precision, splits,
                 part, &splits, st);
    if (rc) return rc;
    rc = finalize_W0_launch(part, splits, H, L.Hp, T.Hq, ws + L.P0, ws + L.P1, p->W0, mode, kSlope, g->W0, st);
    if (rc) return rc;
  }
  dim3 fgrid(T.Hq / 32, d + 1, 2);
  tc_finalize_small_kernel<<<fgrid, 1024, 0, st>>>(partA, partB, a2part, (int)nmt * 8, (int)nmt, d, H, T.Hq, ws + L.P1,
                                                  p->W1, mode, *g);
  return check_launch();
}

}  // namespace b200vae
