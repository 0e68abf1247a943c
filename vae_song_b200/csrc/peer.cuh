// peer.cuh -- small-message exchange between the GPUs of one node over NVLink / NVSwitch peer memory.
//
// Every rank owns one "exchange buffer" (cudaMalloc + CUDA IPC, mapped into every other rank's address space by
// b200vae_peer_open).  It is an array of PeerSlot.  An exchange on slot s, executed by ONE CTA per rank:
//   1. read the slot's local epoch counter e-1, use e (all ranks execute the same sequence of exchanges, so the
//      counters agree without communication; they live on the device so that a CUDA-graph replay advances them);
//   2. store this rank's payload into data[e&1][rank] of EVERY rank's slot through the peer mapping -- every float travels
//      as ONE 8-byte word (value, e): the word's arrival is its own flag (the "LL" idea of NCCL), so an exchange costs one
//      one-way NVLink trip instead of payload + system fence (a write-acknowledge round trip) + flag;
//   3. spin (bounded by a timeout -- 20 s, B200VAE_PEER_TIMEOUT_S overrides -- that sets a sticky per-rank flag) on the
//      LOCAL words data[e&1][q][j] until their tag is e, for every rank q.  Results are combined in rank order, so every
//      rank computes bit-identical values.
//   An exchange without payload (n = 0) is a barrier: __threadfence_system, release-store flag[e&1][rank] = e on every
//   rank, acquire-spin on the local flags -- everything written before it (gradients, parameters) is visible after it.
// FAILING LOUDLY: once the sticky flag is set every exchange on this rank returns NaN payloads (BatchNorm statistics and
// hence the loss turn NaN) and the fused all-reduce + Adam kernel writes NaN parameters, so a rank that lagged by more than
// the timeout can never silently train on stale data; the host reads the flag with b200vae_peer_timed_out
// (train.DataParallelTrainer polls it every `check_every` steps and raises).
// The parity double-buffer makes slot reuse safe without a trailing barrier: a rank can only publish epoch e+2 into the
// buffers of epoch e after it passed exchange e+1, i.e. after every rank published e+1, which each does (stream order)
// only after it finished reading epoch e.
#pragma once
#include <cstdlib>

#include "common.cuh"

namespace b200vae {

constexpr int kPeerMaxWorld = B200VAE_PEER_MAX_WORLD;   // 16
constexpr int kPeerPay = 400;                           // floats per rank per exchange (>= 3*128 + 1)
constexpr int kPeerSlots = 64;
constexpr unsigned long long kPeerTimeoutNs = 20000000000ull;  // 20 s default: a missing peer must not hang the GPU

struct PeerSlot {
  unsigned long long data[2][kPeerMaxWorld][kPeerPay];   // (float bits, epoch tag << 32)
  unsigned flag[2][kPeerMaxWorld];
  unsigned epoch;
  unsigned timed_out;
  unsigned pad[30];
};

struct PeerComm {   // kernel-parameter copy of b200vae_peer_t
  int world, rank;
  unsigned long long timeout_ns;
  PeerSlot* buf[kPeerMaxWorld];
};

inline unsigned long long peer_timeout_ns() {
  static const unsigned long long ns = [] {
    const char* e = getenv("B200VAE_PEER_TIMEOUT_S");
    const double s = e ? atof(e) : 0.0;
    return s > 0.0 ? (unsigned long long)(s * 1e9) : kPeerTimeoutNs;
  }();
  return ns;
}

inline PeerComm make_peer(const b200vae_peer_t* c) {
  PeerComm p;
  p.world = c->world; p.rank = c->rank;
  p.timeout_ns = peer_timeout_ns();
  for (int r = 0; r < kPeerMaxWorld; ++r) p.buf[r] = r < c->world ? (PeerSlot*)c->buf[r] : nullptr;
  return p;
}
inline bool peer_ok(const b200vae_peer_t* c, int slot) {
  if (!c || c->world < 1 || c->world > kPeerMaxWorld || c->rank < 0 || c->rank >= c->world) return false;
  if (slot < 0 || slot >= kPeerSlots) return false;
  for (int r = 0; r < c->world; ++r)
    if (!c->buf[r]) return false;
  return true;
}

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}

// Called by all threads of one CTA (blockDim.x >= world).  `mine`: n floats in shared memory; `all`: [world][n] floats
// in shared memory, filled with every rank's payload (rank-major).  n <= kPeerPay.
__device__ __forceinline__ void peer_exchange(const PeerComm& c, int slot, const float* mine, int n, float* all) {
  __shared__ unsigned s_epoch;
  PeerSlot* local = c.buf[c.rank] + slot;
  if (threadIdx.x == 0) s_epoch = local->epoch + 1u;
  __syncthreads();
  const unsigned e = s_epoch;
  const int par = (int)(e & 1u);
  // Once any exchange on this rank has timed out (sticky flag in slot 0) later ones do not wait again: a dead peer
  // costs ONE timeout, not one per exchange, and the host sees the flag through b200vae_peer_timed_out.
  unsigned* dead = &c.buf[c.rank]->timed_out;
  if (n > 0) {
    const unsigned long long tag = (unsigned long long)e << 32;
    for (int i = threadIdx.x; i < c.world * n; i += blockDim.x) {
      const int p = i / n, j = i - p * n;
      st_relaxed_sys_u64(&(c.buf[p] + slot)->data[par][c.rank][j], tag | (unsigned long long)__float_as_uint(mine[j]));
    }
    const unsigned long long t0 = global_timer_ns();
    for (int i = threadIdx.x; i < c.world * n; i += blockDim.x) {
      const int p = i / n, j = i - p * n;
      unsigned long long w;
      while ((unsigned)((w = ld_relaxed_sys_u64(&local->data[par][p][j])) >> 32) != e) {
        if (*reinterpret_cast<volatile unsigned*>(dead) != 0u) break;
        if (global_timer_ns() - t0 > c.timeout_ns) { *reinterpret_cast<volatile unsigned*>(dead) = 1u; break; }
      }
      all[i] = __uint_as_float((unsigned)w);
    }
    __syncthreads();
    if (*reinterpret_cast<volatile unsigned*>(dead) != 0u)                  // NaN: fail loudly
      for (int i = threadIdx.x; i < c.world * n; i += blockDim.x) all[i] = __int_as_float(0x7fc00000);
  } else {
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < c.world) {
      st_release_sys(&(c.buf[threadIdx.x] + slot)->flag[par][c.rank], e);
      const unsigned long long t0 = global_timer_ns();
      while (ld_acquire_sys(&local->flag[par][threadIdx.x]) != e) {
        if (*reinterpret_cast<volatile unsigned*>(dead) != 0u) break;
        if (global_timer_ns() - t0 > c.timeout_ns) { *reinterpret_cast<volatile unsigned*>(dead) = 1u; break; }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) local->epoch = e;
  __syncthreads();
}

}  // namespace b200vae
