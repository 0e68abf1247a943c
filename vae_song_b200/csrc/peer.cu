// peer.cu -- host side of the peer-memory exchange (CUDA IPC set-up), the generic small all-gather, and the
// two-shot gradient all-reduce fused with Adam.  See peer.cuh for the exchange protocol and include/b200vae.h for
// the contract.  Replaces, across ranks, lipschitz.py:41-43 (backward -> optimizer.step) of the reference, which is
// single-process.
#include <math.h>
#include <string.h>

#include "peer.cuh"

namespace b200vae {

__global__ void __launch_bounds__(256)
peer_allgather_kernel(PeerComm c, int slot, const float* __restrict__ in, int n, float* __restrict__ out,
                      long long* tick) {
  __shared__ float mine[kPeerPay];
  __shared__ float all[kPeerMaxWorld * kPeerPay];
  for (int i = threadIdx.x; i < n; i += blockDim.x) mine[i] = in[i];
  if (tick && threadIdx.x == 0) *tick += 1;
  __syncthreads();
  peer_exchange(c, slot, mine, n, all);
  for (int i = threadIdx.x; i < c.world * n; i += blockDim.x) out[i] = all[i];
}

struct PtrList { const float4* g[kPeerMaxWorld]; float4* p[kPeerMaxWorld]; };

__device__ __forceinline__ float adam_one(float& p, float g, float& m, float& v, float b1, float b2, float eps, float wd,
                                          float step_size, float bc2_sqrt) {
  if (wd != 0.f) g = fmaf(wd, p, g);
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  p = p - step_size * (m / (sqrtf(v) / bc2_sqrt + eps));
  return p;
}

// chunk [lo4, hi4) in float4 units belongs to this rank
__global__ void __launch_bounds__(256)
peer_reduce_adam_kernel(PtrList L, int world, int rank, long long lo4, long long hi4, float4* __restrict__ m,
                        float4* __restrict__ v, float lr, float b1, float b2, float eps, float wd,
                        const long long* __restrict__ step, float gscale, const unsigned* __restrict__ dead) {
  // an exchange on this rank timed out (sticky flag): the gradients may be incomplete -> poison instead of a stale update
  const float poison = (*reinterpret_cast<const volatile unsigned*>(dead) != 0u) ? __int_as_float(0x7fc00000) : 0.f;
  __shared__ float s_step_size, s_bc2_sqrt;      // fp64 bias corrections once per block, not a double pow per thread
  if (threadIdx.x == 0) {
    const double t = (double)*step;
    s_step_size = lr / (float)(1.0 - pow((double)b1, t));
    s_bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, t));
  }
  __syncthreads();
  const float step_size = s_step_size, bc2_sqrt = s_bc2_sqrt;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += gstride) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = 0; q < world; ++q) {                      // rank order: deterministic, identical everywhere
      const float4 g = __ldcv(L.g[q] + i);                  // peer load over NVLink (never from a stale L1 line)
      s.x += g.x; s.y += g.y; s.z += g.z; s.w += g.w;
    }
    s.x += poison; s.y += poison; s.z += poison; s.w += poison;
    float4 p = L.p[rank][i], mi = m[i], vi = v[i];
    adam_one(p.x, s.x * gscale, mi.x, vi.x, b1, b2, eps, wd, step_size, bc2_sqrt);
    adam_one(p.y, s.y * gscale, mi.y, vi.y, b1, b2, eps, wd, step_size, bc2_sqrt);
    adam_one(p.z, s.z * gscale, mi.z, vi.z, b1, b2, eps, wd, step_size, bc2_sqrt);
    adam_one(p.w, s.w * gscale, mi.w, vi.w, b1, b2, eps, wd, step_size, bc2_sqrt);
    m[i] = mi; v[i] = vi;
    for (int q = 0; q < world; ++q) L.p[q][i] = p;          // updated chunk to every replica (peer stores)
  }
  __threadfence_system();
}

static int launch_allgather(const b200vae_peer_t* comm, int slot, const float* in, int n, float* out, long long* tick,
                            cudaStream_t st) {
  peer_allgather_kernel<<<1, 256, 0, st>>>(make_peer(comm), slot, in, n, out, tick);
  return check_launch();
}

}  // namespace b200vae

using namespace b200vae;

extern "C" size_t b200vae_peer_exchange_bytes(void) { return sizeof(PeerSlot) * (size_t)kPeerSlots; }
extern "C" int b200vae_peer_num_slots(void) { return kPeerSlots; }
extern "C" int b200vae_peer_max_payload(void) { return kPeerPay; }

static int cuda_rc(cudaError_t e) {
  if (e == cudaSuccess) return B200VAE_OK;
  g_last_cuda_error = (int)e;
  cudaGetLastError();
  return B200VAE_ECUDA;
}

extern "C" int b200vae_peer_alloc(size_t bytes, void** buf, unsigned char* handle) {
  if (!buf || !handle) return B200VAE_EALIGN;
  if (bytes == 0) return B200VAE_ESHAPE;
  static_assert(sizeof(cudaIpcMemHandle_t) == B200VAE_PEER_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  int rc = cuda_rc(cudaMalloc(&p, bytes));
  if (rc) return rc;
  rc = cuda_rc(cudaMemset(p, 0, bytes));
  if (!rc) rc = cuda_rc(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  if (!rc) rc = cuda_rc(cudaIpcGetMemHandle(&h, p));
  if (rc) { cudaFree(p); return rc; }
  memcpy(handle, &h, sizeof(h));
  *buf = p;
  return B200VAE_OK;
}
extern "C" int b200vae_peer_open(const unsigned char* handle, void** mapped) {
  if (!handle || !mapped) return B200VAE_EALIGN;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  int rc = cuda_rc(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  if (rc) return rc;
  *mapped = p;
  return B200VAE_OK;
}
extern "C" int b200vae_peer_close(void* mapped) { return mapped ? cuda_rc(cudaIpcCloseMemHandle(mapped)) : B200VAE_EALIGN; }
extern "C" int b200vae_peer_free(void* buf) { return buf ? cuda_rc(cudaFree(buf)) : B200VAE_EALIGN; }

extern "C" int b200vae_peer_timed_out(const b200vae_peer_t* comm, int* out) {
  if (!peer_ok(comm, 0) || !out) return B200VAE_EALIGN;
  unsigned flag = 0;     // sticky flag of this rank: slot 0 (peer.cuh)
  int rc = cuda_rc(cudaMemcpy(&flag, &((PeerSlot*)comm->buf[comm->rank])->timed_out, sizeof(flag), cudaMemcpyDeviceToHost));
  if (rc) return rc;
  *out = flag != 0;
  return B200VAE_OK;
}

extern "C" int b200vae_peer_allgather(const b200vae_peer_t* comm, int slot, const float* in, int n, float* out,
                                      void* stream) {
  if (!peer_ok(comm, slot)) return B200VAE_EALIGN;
  if (n < 0 || n > kPeerPay) return B200VAE_ESHAPE;
  if (n > 0 && (!in || !out)) return B200VAE_EALIGN;
  return launch_allgather(comm, slot, in, n, out, nullptr, (cudaStream_t)stream);
}

extern "C" int b200vae_peer_allreduce_adam(const b200vae_peer_t* comm, int slot, void* const* grads, void* const* params,
                                           float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                                           float weight_decay, long long* step_dev, float grad_scale, void* stream) {
  if (!peer_ok(comm, slot) || !peer_ok(comm, slot + 1) || !grads || !params || !m || !v || !step_dev) return B200VAE_EALIGN;
  if (n <= 0 || (n & 3)) return B200VAE_ESHAPE;
  cudaStream_t st = (cudaStream_t)stream;
  const int W = comm->world, r = comm->rank;
  PtrList L;
  for (int q = 0; q < kPeerMaxWorld; ++q) {
    L.g[q] = q < W ? (const float4*)grads[q] : nullptr;
    L.p[q] = q < W ? (float4*)params[q] : nullptr;
    if (q < W && (!grads[q] || !params[q] || !aligned16(grads[q]) || !aligned16(params[q]))) return B200VAE_EALIGN;
  }
  if (!aligned16(m) || !aligned16(v)) return B200VAE_EALIGN;
  const long long n4 = n / 4, per = (n4 + W - 1) / W;
  const long long lo = per * r < n4 ? per * r : n4, hi = lo + per < n4 ? lo + per : n4;
  // (1) every rank's gradients are complete (and the Adam step counter advances)
  int rc = launch_allgather(comm, slot, nullptr, 0, nullptr, step_dev, st);
  if (rc) return rc;
  // (2) reduce my chunk over the peers, Adam, broadcast the updated chunk
  if (hi > lo) {
    long long blocks = (hi - lo + 255) / 256;
    if (blocks > 148 * 4) blocks = 148 * 4;
    peer_reduce_adam_kernel<<<(int)blocks, 256, 0, st>>>(L, W, r, lo, hi, (float4*)m, (float4*)v, lr, beta1, beta2, eps,
                                                        weight_decay, step_dev, grad_scale,
                                                        &((PeerSlot*)comm->buf[r])->timed_out);
    rc = check_launch();
    if (rc) return rc;
  }
  // (3) every rank's parameter writes have landed (and nobody still reads my gradients)
  return launch_allgather(comm, slot + 1, nullptr, 0, nullptr, nullptr, st);
}
