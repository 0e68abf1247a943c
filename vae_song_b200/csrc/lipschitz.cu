// lipschitz.cu -- pairwise |f(x)-f(y)| / |x-y| estimators.
//   pairs kernel    : reference semantics, utils.py:544-562 (random index pairs given by the caller,
//                     rows flattened, p=2 norms, both norms clamped at eps)
//   all-pairs kernel: north_star kernel 4 -- 64x64 tiles of the upper triangle, warp-shuffle
//                     max / min / sum reduction, optional log2 histogram (for quantiles).
#include "common.cuh"

namespace b200vae {

// one warp per pair; lanes stride over the feature dimension
__global__ void __launch_bounds__(256)
lipschitz_pairs_kernel(const float* __restrict__ X, const float* __restrict__ Y, const int64_t* __restrict__ i1,
                       const int64_t* __restrict__ i2, int P, int N, int dx, int dy, float eps,
                       float* __restrict__ ratio) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int p = warp; p < P; p += nwarps) {
    long long a = i1[p], b = i2[p];
    a = a < 0 ? 0 : (a >= N ? N - 1 : a);
    b = b < 0 ? 0 : (b >= N ? N - 1 : b);
    float sx = 0.f, sy = 0.f;
    for (int j = lane; j < dx; j += 32) { const float t = X[a * dx + j] - X[b * dx + j]; sx = fmaf(t, t, sx); }
    for (int j = lane; j < dy; j += 32) { const float t = Y[a * dy + j] - Y[b * dy + j]; sy = fmaf(t, t, sy); }
    sx = warp_sum(sx); sy = warp_sum(sy);
    if (lane == 0) ratio[p] = fmaxf(sqrtf(sy), eps) / fmaxf(sqrtf(sx), eps);
  }
}

constexpr int kTile = 64;

__device__ __forceinline__ void tile_coords(long long t, int T, int& ti, int& tj) {
  // row-major over the upper triangle (ti <= tj): row ti starts at ti*T - ti*(ti-1)/2
  double tt = (double)t;
  int r = (int)floor(((2.0 * T + 1.0) - sqrt((2.0 * T + 1.0) * (2.0 * T + 1.0) - 8.0 * tt)) * 0.5);
  if (r < 0) r = 0;
  if (r > T - 1) r = T - 1;
  auto start = [T](long long q) { return q * T - q * (q - 1) / 2; };
  while (r > 0 && start(r) > t) --r;
  while (r + 1 < T && start(r + 1) <= t) ++r;
  ti = r;
  tj = r + (int)(t - start(r));
}

// 256 threads: thread (a = tid/16 in 0..15, b = tid%16) handles rows i = a + 16*ii, cols j = b + 16*jj
__global__ void __launch_bounds__(256)
lipschitz_allpairs_kernel(const float* __restrict__ X, const float* __restrict__ Y, int N, int dx, int dy,
                          float eps, long long tile_begin, long long tile_end, double* __restrict__ stats,
                          uint32_t* __restrict__ hist, int nbins, float hist_lo, float hist_hi) {
  extern __shared__ float sm[];   // Xi[64][dx] Xj[64][dx] Yi[64][dy] Yj[64][dy]
  float* Xi = sm; float* Xj = Xi + kTile * dx; float* Yi = Xj + kTile * dx; float* Yj = Yi + kTile * dy;
  __shared__ float red[3][8];
  __shared__ unsigned cnt_red[8];
  const int T = (N + kTile - 1) / kTile;
  const int tid = threadIdx.x, ta = tid >> 4, tb = tid & 15;
  float vmax = 0.f, vmin = 3.4e38f, vsum = 0.f;
  unsigned cnt = 0;
  for (long long t = tile_begin + blockIdx.x; t < tile_end; t += gridDim.x) {
    int ti, tj;
    tile_coords(t, T, ti, tj);
    const int i0 = ti * kTile, j0 = tj * kTile;
    __syncthreads();
    for (int q = tid; q < kTile * dx; q += 256) {
      const int r = q / dx;
      Xi[q] = (i0 + r < N) ? X[(size_t)i0 * dx + q] : 0.f;
      Xj[q] = (j0 + r < N) ? X[(size_t)j0 * dx + q] : 0.f;
    }
    for (int q = tid; q < kTile * dy; q += 256) {
      const int r = q / dy;
      Yi[q] = (i0 + r < N) ? Y[(size_t)i0 * dy + q] : 0.f;
      Yj[q] = (j0 + r < N) ? Y[(size_t)j0 * dy + q] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const int li = ta + 16 * ii, gi = i0 + li;
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int lj = tb + 16 * jj, gj = j0 + lj;
        if (gi < N && gj < N && gi < gj) {
          float sx = 0.f, sy = 0.f;
          for (int q = 0; q < dx; ++q) { const float d = Xi[li * dx + q] - Xj[lj * dx + q]; sx = fmaf(d, d, sx); }
          for (int q = 0; q < dy; ++q) { const float d = Yi[li * dy + q] - Yj[lj * dy + q]; sy = fmaf(d, d, sy); }
          const float r = fmaxf(sqrtf(sy), eps) / fmaxf(sqrtf(sx), eps);
          vmax = fmaxf(vmax, r); vmin = fminf(vmin, r); vsum += r; ++cnt;
          if (hist) {
            const float lg = log2f(r);
            int bin = (int)floorf((lg - hist_lo) / (hist_hi - hist_lo) * (float)nbins);
            bin = bin < 0 ? 0 : (bin >= nbins ? nbins - 1 : bin);
            atomicAdd(hist + bin, 1u);
          }
        }
      }
    }
  }
  vmax = warp_max(vmax); vmin = warp_min(vmin); vsum = warp_sum(vsum);
  cnt = __reduce_add_sync(0xffffffffu, cnt);
  const int w = tid >> 5, l = tid & 31;
  if (l == 0) { red[0][w] = vmax; red[1][w] = vmin; red[2][w] = vsum; cnt_red[w] = cnt; }
  __syncthreads();
  if (tid == 0) {
    float a = red[0][0], b = red[1][0], c = red[2][0];
    unsigned n = cnt_red[0];
    for (int q = 1; q < 8; ++q) { a = fmaxf(a, red[0][q]); b = fminf(b, red[1][q]); c += red[2][q]; n += cnt_red[q]; }
    if (n > 0) {
      // ratios are > 0: the IEEE bit pattern is monotone, so integer atomics give exact float max/min
      atomicMax(reinterpret_cast<unsigned long long*>(stats + 0), (unsigned long long)__double_as_longlong((double)a));
      atomicMin(reinterpret_cast<unsigned long long*>(stats + 1), (unsigned long long)__double_as_longlong((double)b));
      atomicAdd(stats + 2, (double)c);
      atomicAdd(stats + 3, (double)n);
    }
  }
}

__global__ void allpairs_init_kernel(double* stats, uint32_t* hist, int nbins) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) { stats[0] = 0.0; stats[1] = 1.7976931348623157e308; stats[2] = 0.0; stats[3] = 0.0; }
  if (hist && i < nbins) hist[i] = 0u;
}

}  // namespace b200vae

using namespace b200vae;

extern "C" int b200vae_lipschitz_pairs(const float* X, const float* Y, const int64_t* i1, const int64_t* i2, int P,
                                       int N, int dx, int dy, float eps, float* ratio, void* stream) {
  if (!X || !Y || !i1 || !i2 || !ratio) return B200VAE_EALIGN;
  if (P <= 0 || N <= 0 || dx <= 0 || dy <= 0) return B200VAE_ESHAPE;
  int blocks = (P + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  lipschitz_pairs_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(X, Y, i1, i2, P, N, dx, dy, eps, ratio);
  return check_launch();
}

extern "C" long long b200vae_lipschitz_num_tiles(int N) {
  const long long T = (N + kTile - 1) / kTile;
  return T * (T + 1) / 2;
}

extern "C" int b200vae_lipschitz_allpairs(const float* X, const float* Y, int N, int dx, int dy, float eps,
                                          long long tile_begin, long long tile_end, double* stats, uint32_t* hist,
                                          int nbins, float hist_lo, float hist_hi, void* stream) {
  if (!X || !Y || !stats) return B200VAE_EALIGN;
  if (N <= 0 || dx <= 0 || dy <= 0 || tile_begin < 0 || tile_end < tile_begin ||
      tile_end > b200vae_lipschitz_num_tiles(N))
    return B200VAE_ESHAPE;
  if (hist && (nbins <= 0 || !(hist_hi > hist_lo))) return B200VAE_ESHAPE;
  const size_t smem = sizeof(float) * 2 * kTile * ((size_t)dx + dy);
  if (smem > 200 * 1024) return B200VAE_EUNSUP;
  cudaStream_t st = (cudaStream_t)stream;
  allpairs_init_kernel<<<(nbins > 0 && hist ? (nbins + 255) / 256 : 1), 256, 0, st>>>(stats, hist, nbins);
  int rc = check_launch();
  if (rc) return rc;
  const long long nt = tile_end - tile_begin;
  if (nt == 0) return B200VAE_OK;
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(lipschitz_allpairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int blocks = (int)(nt < 148 * 4 ? nt : 148 * 4);
  lipschitz_allpairs_kernel<<<blocks, 256, smem, st>>>(X, Y, N, dx, dy, eps, tile_begin, tile_end, stats, hist, nbins,
                                                       hist_lo, hist_hi);
  return check_launch();
}
