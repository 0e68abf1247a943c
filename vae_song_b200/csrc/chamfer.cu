// chamfer.cu -- nearest-neighbour squared distances between two point sets and their backward: the building block of
// model.chamfer_distance (reference model.py:896-912: torch.cdist -> [B,Np,Ng] matrix -> two min reductions).
// Tiled min-reduction: the [B, Np, Ng] distance matrix is never materialised.
#include "common.cuh"

namespace b200vae {

constexpr int kCT = 256;   // points per tile / threads per block
constexpr int kMaxDim = 4;

// minv[b,i] = min_j |A[b,i]-Bp[b,j]|^2 , argm[b,i] = a minimising j.   grid (ceil(Na/256), B)
__global__ void __launch_bounds__(kCT)
nn_min_kernel(const float* __restrict__ A, const float* __restrict__ Bp, int Na, int Nb, int dim,
              float* __restrict__ minv, int* __restrict__ argm) {
  __shared__ float tile[kCT * kMaxDim];
  const int b = blockIdx.y, i = blockIdx.x * kCT + threadIdx.x;
  float a[kMaxDim];
#pragma unroll
  for (int q = 0; q < kMaxDim; ++q) a[q] = (i < Na && q < dim) ? A[((size_t)b * Na + i) * dim + q] : 0.f;
  float best = 3.4e38f;
  int arg = 0;
  for (int j0 = 0; j0 < Nb; j0 += kCT) {
    const int nt = min(kCT, Nb - j0);
    __syncthreads();
    for (int q = threadIdx.x; q < nt * dim; q += kCT) tile[q] = Bp[((size_t)b * Nb + j0) * dim + q];
    __syncthreads();
    for (int j = 0; j < nt; ++j) {
      float d2 = 0.f;
#pragma unroll
      for (int q = 0; q < kMaxDim; ++q)
        if (q < dim) { const float t = a[q] - tile[j * dim + q]; d2 = fmaf(t, t, d2); }
      if (d2 < best) { best = d2; arg = j0 + j; }
    }
  }
  if (i < Na) { minv[(size_t)b * Na + i] = best; argm[(size_t)b * Na + i] = arg; }
}

// dA[b,i] = gA[b,i]*2(A_i - Bp_{argA(i)}) + sum_{j: argB(j)=i} gB[b,j]*2(A_i - Bp_j)     (deterministic: every thread
// scans all j instead of scattering with atomics)
__global__ void __launch_bounds__(kCT)
nn_min_bwd_kernel(const float* __restrict__ A, const float* __restrict__ Bp, const int* __restrict__ argA,
                  const int* __restrict__ argB, const float* __restrict__ gA, const float* __restrict__ gB, int Na, int Nb,
                  int dim, float* __restrict__ dA) {
  __shared__ float tile[kCT * kMaxDim];
  __shared__ int targ[kCT];
  __shared__ float tg[kCT];
  const int b = blockIdx.y, i = blockIdx.x * kCT + threadIdx.x;
  float a[kMaxDim], acc[kMaxDim];
#pragma unroll
  for (int q = 0; q < kMaxDim; ++q) { a[q] = (i < Na && q < dim) ? A[((size_t)b * Na + i) * dim + q] : 0.f; acc[q] = 0.f; }
  if (i < Na && gA) {
    const int j = argA[(size_t)b * Na + i];
    const float g = 2.f * gA[(size_t)b * Na + i];
#pragma unroll
    for (int q = 0; q < kMaxDim; ++q)
      if (q < dim) acc[q] = g * (a[q] - Bp[((size_t)b * Nb + j) * dim + q]);
  }
  if (gB) {
    for (int j0 = 0; j0 < Nb; j0 += kCT) {
      const int nt = min(kCT, Nb - j0);
      __syncthreads();
      for (int q = threadIdx.x; q < nt * dim; q += kCT) tile[q] = Bp[((size_t)b * Nb + j0) * dim + q];
      if (threadIdx.x < nt) { targ[threadIdx.x] = argB[(size_t)b * Nb + j0 + threadIdx.x]; tg[threadIdx.x] = gB[(size_t)b * Nb + j0 + threadIdx.x]; }
      __syncthreads();
      for (int j = 0; j < nt; ++j) {
        if (targ[j] == i) {
          const float g = 2.f * tg[j];
#pragma unroll
          for (int q = 0; q < kMaxDim; ++q)
            if (q < dim) acc[q] = fmaf(g, a[q] - tile[j * dim + q], acc[q]);
        }
      }
    }
  }
  if (i < Na) {
#pragma unroll
    for (int q = 0; q < kMaxDim; ++q)
      if (q < dim) dA[((size_t)b * Na + i) * dim + q] = acc[q];
  }
}

}  // namespace b200vae

using namespace b200vae;

extern "C" int b200vae_nn_sqdist_fwd(const float* A, const float* Bp, int B, int Na, int Nb, int dim, float* minv, int* argm,
                                     void* stream) {
  if (!A || !Bp || !minv || !argm) return B200VAE_EALIGN;
  if (B <= 0 || Na <= 0 || Nb <= 0 || dim < 1 || dim > kMaxDim) return B200VAE_ESHAPE;
  dim3 grid((Na + kCT - 1) / kCT, B);
  nn_min_kernel<<<grid, kCT, 0, (cudaStream_t)stream>>>(A, Bp, Na, Nb, dim, minv, argm);
  return check_launch();
}

extern "C" int b200vae_nn_sqdist_bwd(const float* A, const float* Bp, const int* argA, const int* argB, const float* gA,
                                     const float* gB, int B, int Na, int Nb, int dim, float* dA, void* stream) {
  if (!A || !Bp || !dA || (gA && !argA) || (gB && !argB)) return B200VAE_EALIGN;
  if (B <= 0 || Na <= 0 || Nb <= 0 || dim < 1 || dim > kMaxDim) return B200VAE_ESHAPE;
  dim3 grid((Na + kCT - 1) / kCT, B);
  nn_min_bwd_kernel<<<grid, kCT, 0, (cudaStream_t)stream>>>(A, Bp, argA, argB, gA, gB, Na, Nb, dim, dA);
  return check_launch();
}
