// loss.cu -- fused reparameterisation + Gaussian KL + reconstruction (+ latent reconstruction),
// forward and backward, and the fused Adam step.  HBM-bound elementwise/reduction work: float4
// loads where alignment allows, warp-shuffle + ordered block partials (no float atomics).
//
// Reference lines replaced (see include/b200vae.h): model.py:843, :423-424, utils.py:40-47 (reparam);
// model.py:884/:550/:606, utils.py:140-141 (KL); :870/:542/:589 (MSE); :872-882 (log-MSE);
// :551/:603 (latent recon, mean over dim 0).
#include "common.cuh"

namespace b200vae {

constexpr int kLossThreads = 256;
constexpr int kLossMaxBlocks = 592;   // 4 x 148 SMs
// `out` layout (floats): [0..3] results | [4] ticket (uint) | [8 .. 8+3*kLossMaxBlocks) block partials
constexpr int kPartOff = 8;

__device__ __forceinline__ float block_sum(float v, float* sh) {
  v = warp_sum(v);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) sh[w] = v;
  __syncthreads();
  float s = 0.f;
  if (w == 0) {
    s = (l < kLossThreads / 32) ? sh[l] : 0.f;
    s = warp_sum(s);
  }
  return s;   // valid in warp 0
}

__global__ void __launch_bounds__(kLossThreads)
loss_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                float* __restrict__ z, int L, int B, int D, const float* __restrict__ x,
                const float* __restrict__ xhat, int Dx, int logmse, float* __restrict__ mse_rows,
                const float* __restrict__ z_in, const float* __restrict__ z_rec, int Lz, long long nz,
                float* __restrict__ out) {
  __shared__ float sh[kLossThreads / 32];
  __shared__ bool last;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  const long long BD = (long long)B * D;

  // ---- reparam + KL over [B,D] (z written for every l) ----
  float kl = 0.f;
  if (mu && lv) {
    for (long long i = gtid; i < BD; i += gstride) {
      const float m = mu[i], l = lv[i];
      kl += -0.5f * (1.f + l - m * m - expf(l));
      if (z && eps) {
        const float sd = expf(0.5f * l);
        for (int s = 0; s < L; ++s) z[(long long)s * BD + i] = fmaf(eps[(long long)s * BD + i], sd, m);
      }
    }
  }
  // ---- reconstruction ----
  float rec = 0.f;
  if (x && xhat) {
    if (!logmse) {
      const long long n = (long long)B * Dx;
      if ((n & 3) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(xhat)) & 15) == 0) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        const float4* h4 = reinterpret_cast<const float4*>(xhat);
        for (long long i = gtid; i < n / 4; i += gstride) {
          const float4 a = x4[i], b = h4[i];
          const float d0 = a.x - b.x, d1 = a.y - b.y, d2 = a.z - b.z, d3 = a.w - b.w;
          rec += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
        }
      } else {
        for (long long i = gtid; i < n; i += gstride) { const float dd = x[i] - xhat[i]; rec = fmaf(dd, dd, rec); }
      }
    } else {
      // one warp per row: mse_b then 0.5*Dx*(log(2*pi*mse_b + 1e-5) + 1)
      const int lane = threadIdx.x & 31;
      const long long warp = gtid >> 5, nwarps = gstride >> 5;
      for (long long b = warp; b < B; b += nwarps) {
        float s = 0.f;
        for (int j = lane; j < Dx; j += 32) { const float dd = x[b * Dx + j] - xhat[b * Dx + j]; s = fmaf(dd, dd, s); }
        s = warp_sum(s);
        const float mse = s / (float)Dx;
        if (lane == 0) {
          if (mse_rows) mse_rows[b] = mse;
          rec += 0.5f * (float)Dx * (logf(6.283185307179586f * mse + 1e-5f) + 1.f);
        }
      }
    }
  }
  // ---- latent reconstruction ----
  float lat = 0.f;
  if (z_in && z_rec) {
    for (long long i = gtid; i < nz; i += gstride) { const float dd = z_in[i] - z_rec[i]; lat = fmaf(dd, dd, lat); }
  }

  const float r0 = block_sum(rec, sh), r1 = block_sum(kl, sh), r2 = block_sum(lat, sh);
  float* part = out + kPartOff;
  if (threadIdx.x == 0) {
    part[blockIdx.x * 3 + 0] = r0; part[blockIdx.x * 3 + 1] = r1; part[blockIdx.x * 3 + 2] = r2;
    __threadfence();
    const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(out + 4), 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {   // fixed-order final reduction by the last block to finish
    __threadfence();
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
      a0 += __ldcg(part + i * 3 + 0); a1 += __ldcg(part + i * 3 + 1); a2 += __ldcg(part + i * 3 + 2);
    }
    a0 = block_sum(a0, sh); a1 = block_sum(a1, sh); a2 = block_sum(a2, sh);
    if (threadIdx.x == 0) {
      out[0] = a0 / (float)B;                       // .mean(dim=0).sum()  (log-MSE: .mean() over rows)
      out[1] = a1 / (float)B;
      out[2] = (Lz > 0) ? a2 / (float)Lz : 0.f;     // mean over dim 0 = L (quirk B.2)
      out[3] = 0.f;
      *reinterpret_cast<unsigned*>(out + 4) = 0u;   // leave the ticket clean for the next call
    }
  }
}

__global__ void __launch_bounds__(kLossThreads)
loss_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv, const float* __restrict__ eps,
                const float* __restrict__ gz, int L, int B, int D, const float* __restrict__ x,
                const float* __restrict__ xhat, int Dx, int logmse, const float* __restrict__ mse_rows,
                const float* __restrict__ z_in, const float* __restrict__ z_rec, int Lz, long long nz,
                const float* __restrict__ g_recon, const float* __restrict__ g_kl, const float* __restrict__ g_lat,
                float* __restrict__ d_mu, float* __restrict__ d_lv, float* __restrict__ d_xhat,
                float* __restrict__ d_zrec) {
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  const long long BD = (long long)B * D;
  const float invB = 1.f / (float)B;
  if (d_mu && d_lv && mu && lv) {
    const float gk = g_kl ? *g_kl : 0.f;
    for (long long i = gtid; i < BD; i += gstride) {
      const float m = mu[i], l = lv[i];
      float dm = gk * m * invB;
      float dl = gk * (-0.5f * (1.f - expf(l))) * invB;
      if (gz) {
        const float hsd = 0.5f * expf(0.5f * l);
        for (int s = 0; s < L; ++s) {
          const float g = gz[(long long)s * BD + i];
          dm += g;
          if (eps) dl = fmaf(g * eps[(long long)s * BD + i], hsd, dl);
        }
      }
      d_mu[i] = dm; d_lv[i] = dl;
    }
  }
  if (d_xhat && x && xhat) {
    const float gr = g_recon ? *g_recon : 0.f;
    const long long n = (long long)B * Dx;
    if (!logmse) {
      const float c = gr * (-2.f * invB);
      for (long long i = gtid; i < n; i += gstride) d_xhat[i] = c * (x[i] - xhat[i]);
    } else {
      for (long long i = gtid; i < n; i += gstride) {
        const long long b = i / Dx;
        const float coef = 0.5f * (float)Dx * 6.283185307179586f / (6.283185307179586f * mse_rows[b] + 1e-5f) * invB;
        d_xhat[i] = gr * coef * (-2.f / (float)Dx) * (x[i] - xhat[i]);
      }
    }
  }
  if (d_zrec && z_in && z_rec) {
    const float c = (g_lat ? *g_lat : 0.f) * (-2.f / (float)Lz);
    for (long long i = gtid; i < nz; i += gstride) d_zrec[i] = c * (z_in[i] - z_rec[i]);
  }
}

// torch.optim.Adam (amsgrad=False, maximize=False), single-tensor semantics
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            long long n, float lr, float b1, float b2, float eps, float wd, float bc1, float bc2_sqrt,
            float gscale) {
  const long long gstride = (long long)gridDim.x * blockDim.x;
  const float step_size = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gstride) {
    float gi = g[i] * gscale;
    const float pi = p[i];
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    m[i] = mi; v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = pi - step_size * (mi / denom);
  }
}

// graph-capturable variant: the step counter lives on the device (incremented here by block 0 *after* every
// block has read it is not possible without a grid sync, so a 1-thread tick kernel runs first)
__global__ void adam_tick_kernel(long long* step) { *step += 1; }
// lr of torch.optim.lr_scheduler.CosineAnnealingLR(T_max, eta_min = 0) at 0-based scheduler step k = *step - 1 (main.py:201-203:
// scheduler.step() after every optimizer.step()), closed form: lr0 * (1 + cos(pi k / T)) / 2
__device__ __forceinline__ float sched_lr(float lr0, int kind, long long T, long long step1) {
  if (kind != 1 || T <= 0) return lr0;
  const double k = (double)(step1 - 1);
  return (float)(0.5 * (double)lr0 * (1.0 + cos(3.14159265358979323846 * k / (double)T)));
}
__global__ void __launch_bounds__(256)
adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                long long n, float lr0, float b1, float b2, float eps, float wd, const long long* __restrict__ step,
                float gscale, int sched_kind, long long sched_T) {
  // bias corrections in fp64 (torch semantics) ONCE per block: a double-precision pow per thread was most of this kernel
  __shared__ float s_step_size, s_bc2_sqrt;
  if (threadIdx.x == 0) {
    const double t = (double)*step;
    const float lr = sched_lr(lr0, sched_kind, sched_T, *step);
    s_step_size = lr / (float)(1.0 - pow((double)b1, t));
    s_bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const long long gstride = (long long)gridDim.x * blockDim.x, tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  auto upd = [&](float& pi, float gi, float& mi, float& vi) {
    gi *= gscale;
    if (wd != 0.f) gi = fmaf(wd, pi, gi);
    mi = b1 * mi + (1.f - b1) * gi;
    vi = b2 * vi + (1.f - b2) * gi * gi;
    pi = pi - step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  };
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15u) == 0;
  const long long n4 = vec ? n / 4 : 0;
  for (long long i = tid0; i < n4; i += gstride) {           // 16-byte accesses (the flat buffers are 16-byte aligned)
    float4 pi = reinterpret_cast<float4*>(p)[i], mi = reinterpret_cast<float4*>(m)[i], vi = reinterpret_cast<float4*>(v)[i];
    const float4 gi = reinterpret_cast<const float4*>(g)[i];
    upd(pi.x, gi.x, mi.x, vi.x); upd(pi.y, gi.y, mi.y, vi.y); upd(pi.z, gi.z, mi.z, vi.z); upd(pi.w, gi.w, mi.w, vi.w);
    reinterpret_cast<float4*>(p)[i] = pi; reinterpret_cast<float4*>(m)[i] = mi; reinterpret_cast<float4*>(v)[i] = vi;
  }
  for (long long i = 4 * n4 + tid0; i < n; i += gstride) {
    float pi = p[i], mi = m[i], vi = v[i];
    upd(pi, g[i], mi, vi);
    p[i] = pi; m[i] = mi; v[i] = vi;
  }
}

static int loss_grid(long long work) {
  long long blocks = (work + kLossThreads - 1) / kLossThreads;
  if (blocks < 1) blocks = 1;
  if (blocks > kLossMaxBlocks) blocks = kLossMaxBlocks;
  return (int)blocks;
}

}  // namespace b200vae

using namespace b200vae;

extern "C" int b200vae_loss_fwd(const float* mu, const float* lv, const float* eps, float* z, int L, int B, int D,
                                const float* x, const float* xhat, int Dx, int logmse, float* mse_rows,
                                const float* z_in, const float* z_rec, int Lz, int Bz, int Dz, float* out,
                                void* stream) {
  if (B <= 0 || !out) return out ? B200VAE_ESHAPE : B200VAE_EALIGN;
  if ((mu || lv) && (!mu || !lv || D <= 0)) return B200VAE_ESHAPE;
  if (z && (!eps || L <= 0 || !mu)) return B200VAE_ESHAPE;
  if ((x || xhat) && (!x || !xhat || Dx <= 0)) return B200VAE_ESHAPE;
  if ((z_in || z_rec) && (!z_in || !z_rec || Lz <= 0 || Bz <= 0 || Dz <= 0)) return B200VAE_ESHAPE;
  if (logmse && x && !mse_rows) return B200VAE_EALIGN;
  const long long nz = (z_in ? (long long)Lz * Bz * Dz : 0);
  long long work = (long long)B * (D > 0 ? D : 1);
  const long long wx = x ? (logmse ? (long long)B * 32 : (long long)B * Dx / 4) : 0;
  if (wx > work) work = wx;
  if (nz > work) work = nz;
  loss_fwd_kernel<<<loss_grid(work), kLossThreads, 0, (cudaStream_t)stream>>>(
      mu, lv, eps, z, L, B, D, x, xhat, Dx, logmse, mse_rows, z_in, z_rec, Lz, nz, out);
  return check_launch();
}

extern "C" int b200vae_loss_bwd(const float* mu, const float* lv, const float* eps, const float* gz, int L, int B,
                                int D, const float* x, const float* xhat, int Dx, int logmse, const float* mse_rows,
                                const float* z_in, const float* z_rec, int Lz, int Bz, int Dz, const float* g_recon,
                                const float* g_kl, const float* g_lat, float* d_mu, float* d_lv, float* d_xhat,
                                float* d_zrec, void* stream) {
  if (B <= 0) return B200VAE_ESHAPE;
  if ((d_mu || d_lv) && (!d_mu || !d_lv || !mu || !lv || D <= 0)) return B200VAE_ESHAPE;
  if (d_xhat && (!x || !xhat || Dx <= 0 || (logmse && !mse_rows))) return B200VAE_ESHAPE;
  if (d_zrec && (!z_in || !z_rec || Lz <= 0)) return B200VAE_ESHAPE;
  const long long nz = (d_zrec ? (long long)Lz * Bz * Dz : 0);
  long long work = (long long)B * (D > 0 ? D : 1);
  if (d_xhat && (long long)B * Dx > work) work = (long long)B * Dx;
  if (nz > work) work = nz;
  loss_bwd_kernel<<<loss_grid(work), kLossThreads, 0, (cudaStream_t)stream>>>(
      mu, lv, eps, gz, L, B, D, x, xhat, Dx, logmse, mse_rows, z_in, z_rec, Lz, nz, g_recon, g_kl, g_lat, d_mu,
      d_lv, d_xhat, d_zrec);
  return check_launch();
}

extern "C" int b200vae_adam_step(float* param, const float* grad, float* m, float* v, long long n, float lr,
                                 float beta1, float beta2, float eps, float weight_decay, long long step,
                                 float grad_scale, void* stream) {
  if (!param || !grad || !m || !v) return B200VAE_EALIGN;
  if (n <= 0 || step <= 0) return B200VAE_ESHAPE;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                            (float)bc1, (float)sqrt(bc2), grad_scale);
  return check_launch();
}

extern "C" int b200vae_adam_step_dev(float* param, const float* grad, float* m, float* v, long long n, float lr,
                                     float beta1, float beta2, float eps, float weight_decay, long long* step_dev,
                                     float grad_scale, void* stream) {
  if (!param || !grad || !m || !v || !step_dev) return B200VAE_EALIGN;
  if (n <= 0) return B200VAE_ESHAPE;
  long long blocks = (n / 4 + 255) / 256 + 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
  int rc = check_launch();
  if (rc) return rc;
  adam_dev_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, lr, beta1, beta2, eps, weight_decay,
                                                                step_dev, grad_scale, 0, 0);
  return check_launch();
}

extern "C" int b200vae_adam_step_sched(float* param, const float* grad, float* m, float* v, long long n, float lr0,
                                       float beta1, float beta2, float eps, float weight_decay, long long* step_dev,
                                       float grad_scale, int sched_kind, long long sched_T, void* stream) {
  if (!param || !grad || !m || !v || !step_dev) return B200VAE_EALIGN;
  if (n <= 0 || sched_kind < 0 || sched_kind > 1) return B200VAE_ESHAPE;
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  adam_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_dev);
  int rc = check_launch();
  if (rc) return rc;
  adam_dev_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(param, grad, m, v, n, lr0, beta1, beta2, eps, weight_decay,
                                                                step_dev, grad_scale, sched_kind, sched_T);
  return check_launch();
}
