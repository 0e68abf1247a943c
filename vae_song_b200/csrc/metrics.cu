// metrics.cu -- end-of-run latent metrics of utils.py.
//   calc_mi (utils.py:87-107): log q(z_i) = logsumexp_j log N(z_i; mu_j, diag var_j) - log B needs the [B,B,nz] log-density
//   tensor in PyTorch; here it is a tiled all-pairs kernel with an ONLINE logsumexp: one warp per sample row i, lanes
//   stride over j (mu_j, 1/var_j, const_j staged through shared memory in tiles of 64 rows), per-lane (max, sum) merged
//   by shuffles.  Nothing of size B^2 touches memory.
#include "common.cuh"

namespace b200vae {

constexpr int kMiTile = 64;

// logqz[i] = logsumexp_j( -0.5 * sum_k (z[i,k]-mu[j,k])^2 / exp(lv[j,k]) - 0.5 * (nz log 2pi + sum_k lv[j,k]) ) - log B
__global__ void __launch_bounds__(256)
mi_logqz_kernel(const float* __restrict__ z, const float* __restrict__ mu, const float* __restrict__ lv, int B, int nz,
                float* __restrict__ logqz) {
  extern __shared__ float sm[];                 // mu_t[64][nz], iv_t[64][nz], c_t[64]
  float* mu_t = sm; float* iv_t = mu_t + kMiTile * nz; float* c_t = iv_t + kMiTile * nz;
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  const int i = blockIdx.x * 8 + wp;
  const float* zi = z + (size_t)(i < B ? i : 0) * nz;
  float m = -3.4e38f, ssum = 0.f;
  for (int j0 = 0; j0 < B; j0 += kMiTile) {
    __syncthreads();
    for (int e = threadIdx.x; e < kMiTile * nz; e += blockDim.x) {
      const int jj = e / nz, k = e - jj * nz, j = j0 + jj;
      const float l = j < B ? lv[(size_t)j * nz + k] : 0.f;
      mu_t[e] = j < B ? mu[(size_t)j * nz + k] : 0.f;
      iv_t[e] = __expf(-l);
    }
    for (int jj = threadIdx.x; jj < kMiTile; jj += blockDim.x) {
      const int j = j0 + jj;
      float s = 0.f;
      if (j < B) for (int k = 0; k < nz; ++k) s += lv[(size_t)j * nz + k];
      c_t[jj] = -0.5f * ((float)nz * 1.8378770664093453f + s);
    }
    __syncthreads();
    for (int jj = lane; jj < kMiTile && j0 + jj < B; jj += 32) {
      float q = 0.f;
      for (int k = 0; k < nz; ++k) { const float dv = zi[k] - mu_t[jj * nz + k]; q = fmaf(dv * dv, iv_t[jj * nz + k], q); }
      const float ld = fmaf(-0.5f, q, c_t[jj]);
      if (ld > m) { ssum = ssum * __expf(m - ld) + 1.f; m = ld; }
      else ssum += __expf(ld - m);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {      // merge the lanes' (max, sum) pairs
    const float m2 = __shfl_xor_sync(0xffffffffu, m, off), s2 = __shfl_xor_sync(0xffffffffu, ssum, off);
    const float mm = fmaxf(m, m2);
    ssum = ssum * __expf(m - mm) + s2 * __expf(m2 - mm);
    m = mm;
  }
  if (lane == 0 && i < B) logqz[i] = m + logf(ssum) - logf((float)B);
}

// nll_iw (utils.py:109-120): ll_iw = logsumexp over ALL (b, s) of [log p(z) - loss_rec - log q(z|x)] - log S with
// z = mu + eps*exp(lv/2).  The 2 pi terms cancel and (z-mu)^2/var = eps^2, so the summand is
//   w[b,s] = sum_k ( -0.5 z_k^2 + 0.5 eps_k^2 + 0.5 lv_k )  (- loss_rec, a constant shift applied on the host side).
// One pass over eps [B,S,nz] with an online (max, sum) per thread, warp / block merges, ordered per-CTA partials and a
// last-block merge: the [B,S,nz] z tensor and the three [B,S] log-density tensors of the reference are never formed.
constexpr int kIwMaxBlocks = 1024;
__device__ __forceinline__ void lse_merge(float& m, float& s, float m2, float s2) {
  const float mm = fmaxf(m, m2);
  s = (s > 0.f ? s * __expf(m - mm) : 0.f) + (s2 > 0.f ? s2 * __expf(m2 - mm) : 0.f);
  m = mm;
}
__global__ void __launch_bounds__(256)
nll_iw_lse_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps, int B, int S, int nz,
                  float* __restrict__ out, float* __restrict__ scratch) {
  __shared__ float redm[8], reds[8];
  __shared__ int last;
  const long long n = (long long)B * S;
  float m = -3.4e38f, ssum = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(i / S);
    const float* e = eps + (size_t)i * nz;
    float w = 0.f;
    for (int k = 0; k < nz; ++k) {
      const float l = __ldg(lv + (size_t)b * nz + k), ek = e[k];
      const float zk = fmaf(ek, __expf(0.5f * l), __ldg(mu + (size_t)b * nz + k));
      w += 0.5f * (ek * ek - zk * zk + l);
    }
    if (w > m) { ssum = ssum * __expf(m - w) + 1.f; m = w; }
    else ssum += __expf(w - m);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    lse_merge(m, ssum, __shfl_xor_sync(0xffffffffu, m, off), __shfl_xor_sync(0xffffffffu, ssum, off));
  const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
  if (lane == 0) { redm[wp] = m; reds[wp] = ssum; }
  __syncthreads();
  unsigned* ticket = reinterpret_cast<unsigned*>(scratch + 2 * kIwMaxBlocks);
  if (threadIdx.x == 0) {
    float bm = redm[0], bs = reds[0];
    for (int w = 1; w < 8; ++w) lse_merge(bm, bs, redm[w], reds[w]);
    scratch[2 * blockIdx.x] = bm; scratch[2 * blockIdx.x + 1] = bs;
    __threadfence();
    last = (atomicAdd(ticket, 1u) + 1u == gridDim.x) ? 1 : 0;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float bm = -3.4e38f, bs = 0.f;
    for (int q = 0; q < (int)gridDim.x; ++q) lse_merge(bm, bs, __ldcg(scratch + 2 * q), __ldcg(scratch + 2 * q + 1));   // fixed order
    out[0] = bm + logf(bs);
    *ticket = 0u;
  }
}

}  // namespace b200vae

using namespace b200vae;

extern "C" size_t b200vae_nll_iw_scratch_bytes(void) { return (size_t)(2 * kIwMaxBlocks + 4) * sizeof(float); }
extern "C" int b200vae_nll_iw_lse(const float* mu, const float* lv, const float* eps, int B, int S, int nz, float* out,
                                  void* scratch, void* stream) {
  if (!mu || !lv || !eps || !out || !scratch) return B200VAE_EALIGN;
  if (B <= 0 || S <= 0 || nz <= 0) return B200VAE_ESHAPE;
  const long long n = (long long)B * S;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  nll_iw_lse_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(mu, lv, eps, B, S, nz, out, (float*)scratch);
  return check_launch();
}

extern "C" int b200vae_mi_logqz(const float* z, const float* mu, const float* lv, int B, int nz, float* logqz, void* stream) {
  if (!z || !mu || !lv || !logqz) return B200VAE_EALIGN;
  if (B <= 0 || nz <= 0) return B200VAE_ESHAPE;
  const size_t smem = (size_t)(2 * kMiTile * nz + kMiTile) * sizeof(float);
  if (smem > 200 * 1024) return B200VAE_EUNSUP;
  static bool attr_done = false;
  if (!attr_done) { cudaFuncSetAttribute(mi_logqz_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr_done = true; }
  mi_logqz_kernel<<<(B + 7) / 8, 256, smem, (cudaStream_t)stream>>>(z, mu, lv, B, nz, logqz);
  return check_launch();
}
