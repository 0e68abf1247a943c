// icnn_tc.cu -- tcgen05 / TMEM / TMA version of the fused ICNN potential + Brenier map (forward).
//
// One CTA = 256 samples (two M=128 UMMA row blocks), one accumulator pass = 256 output columns:
// the whole 512-column TMEM of the SM holds D[256 x 256] fp32.  Per pass the K loop streams
//   B (positive weights, [N=256 rows][K] K-major, 64 B swizzled rows)  by TMA from the L2-resident
//     prepared copy (P for h1 = x1.P^T,  P^T for gx1 = g1.P), shared by both row blocks, and
//   A (x1 = leaky(A0 z+b0)^2  or  g1 = s2*P1*s1) which never exists in memory: the 256 "worker"
//     threads own one sample row each and write their row of every K-block straight into the
//     swizzled UMMA layout in shared memory (generic proxy -> fence.proxy.async -> mbarrier).
// The same thread later drains its row of the accumulator with tcgen05.ld, so every reduction of
// the algorithm (h2 = P1.x2, xhat = A0^T g0 + A1^T g1, the LeakyReLU bit masks) is THREAD-LOCAL:
// no shuffles, no atomics, deterministic.
//
// Warp roles (320 threads): warps 0-7 workers (A generation + epilogue; warp%4 = TMEM lane quarter),
// warp 8 TMA producer + TMEM allocator, warp 9 MMA issuer (one elected lane).
// Precisions: TF32 (kind::tf32, operands rounded-to-nearest to tf32) and 3xTF32
// (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, fp32 accumulate in TMEM: fp32-grade accuracy).
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

#include "tc_common.cuh"

namespace b200vae {

size_t tc_extra_ws_floats(int B, int d, int H, int precision) {
  (void)precision;
  return tc_layout(d, H).end + tc3_layout(B, d, H).end;      // single-CTA operand copies, then the pair kernel's arrays
}

__global__ void tc_prepare_kernel(const float* __restrict__ P0, const float* __restrict__ P0T, const float* __restrict__ P1,
                                  const float* __restrict__ A0p, const float* __restrict__ A1p, int d, int Hp, int Hq,
                                  float* __restrict__ B1hi, float* __restrict__ B1lo, float* __restrict__ B2hi,
                                  float* __restrict__ B2lo, float4* __restrict__ A0q, float4* __restrict__ A1q,
                                  float* __restrict__ P1q) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
  if (c < Hq) {
    const bool in = (r < Hp && c < Hp);
    const float a = in ? P0[(size_t)r * Hp + c] : 0.f, b = in ? P0T[(size_t)r * Hp + c] : 0.f;
    const float ah = to_tf32(a), bh = to_tf32(b);
    B1hi[(size_t)r * Hq + c] = ah; B1lo[(size_t)r * Hq + c] = to_tf32(a - ah);
    B2hi[(size_t)r * Hq + c] = bh; B2lo[(size_t)r * Hq + c] = to_tf32(b - bh);
    if (r == 0) {
      const bool inr = c < Hp;
      float w[4] = {0.f, 0.f, 0.f, 0.f}, u[4] = {0.f, 0.f, 0.f, 0.f};
      if (inr) {
        for (int j = 0; j < d; ++j) { w[j] = A0p[(size_t)c * (d + 1) + j]; u[j] = A1p[(size_t)c * (d + 1) + j]; }
        w[3] = A0p[(size_t)c * (d + 1) + d]; u[3] = A1p[(size_t)c * (d + 1) + d];
      }
      if (d <= 2) u[2] = inr ? P1[c] : 0.f;     // spare lane of the float4: P1 rides along (one LDS less per element)
      A0q[c] = make_float4(w[0], w[1], w[2], w[3]);
      A1q[c] = make_float4(u[0], u[1], u[2], u[3]);
      P1q[c] = inr ? P1[c] : 0.f;
    }
  }
}

// ------------------------------------------------------------------------------------ the kernels
struct TcMaps { CUtensorMap b1hi, b1lo, b2hi, b2lo; };

constexpr int kNW = 16;                 // worker warps: 2 threads per sample row (K halves / column halves)
constexpr int kTcThreads = (kNW + 2) * 32;

__device__ __forceinline__ void worker_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kNW * 32) : "memory"); }

template <bool X3>
struct TcCfg {
  // A stage holds KPS K-blocks of 16 (one mbarrier round trip per stage).  At 1xTF32 a 16-wide block is only 512
  // tensor cycles, about what the single MMA-issuing thread needs per loop iteration (wait, fence, issue, commit),
  // so two blocks share a stage; at 3xTF32 a block is 1536 cycles and one per stage suffices.
  static constexpr int KPS = X3 ? 1 : 2;
  static constexpr int S = 2;                                   // pipeline stages
  static constexpr int kSubBytes = (X3 ? 4 : 2) * kTileBytes;   // one K-block: A(hi[,lo]) + B(hi[,lo])
  static constexpr int kStageBytes = KPS * kSubBytes;
  static constexpr int kOffAlo = kTileBytes, kOffB = (X3 ? 2 : 1) * kTileBytes, kOffBlo = 3 * kTileBytes;
};

// shared-memory carve-up common to the forward and backward kernels
struct TcSmem {
  unsigned char* stages;
  float4 *A0s, *A1s;
  float* P1s;
  uint32_t* maskw;     // [Hq/32][256]  word-major so that a thread's own row is bank-conflict free
  float* xch;          // [4][256][4] exchange between the threads that own parts of a row
  uint32_t full0, empty0, accfull, accempty;
  uint32_t* tmem_slot;
};
template <bool X3>
__device__ __forceinline__ TcSmem carve(unsigned char* smem_raw, int Hq) {
  using C = TcCfg<X3>;
  TcSmem m;
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // swizzle needs 1 KB alignment
  m.stages = smem;
  m.A0s = reinterpret_cast<float4*>(smem + C::S * C::kStageBytes);
  m.A1s = m.A0s + Hq;
  m.P1s = reinterpret_cast<float*>(m.A1s + Hq);
  m.maskw = reinterpret_cast<uint32_t*>(m.P1s + Hq);
  m.xch = reinterpret_cast<float*>(m.maskw + (Hq / 32) * 256);
  uint64_t* bars = reinterpret_cast<uint64_t*>(m.xch + 4 * 256 * 4);
  m.full0 = smem_u32(bars); m.empty0 = smem_u32(bars + C::S);
  m.accfull = smem_u32(bars + 2 * C::S); m.accempty = smem_u32(bars + 2 * C::S + 1);
  m.tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::S + 2);
  return m;
}
template <bool X3>
static size_t tc_smem_bytes(int Hq) {
  using C = TcCfg<X3>;
  return (size_t)C::S * C::kStageBytes + (size_t)Hq * (16 + 16 + 4) + (size_t)(Hq / 32) * 256 * 4 + 4 * 256 * 4 * 4 +
         (2 * C::S + 2) * 8 + 16 + 1024;
}

template <bool X3>
__device__ __forceinline__ uint32_t tc_setup(const TcSmem& m, int Hq, const float4* A0q_g, const float4* A1q_g,
                                              const float* P1q_g) {
  using C = TcCfg<X3>;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < C::S; ++s) { mbar_init(m.full0 + 8 * s, (C::KPS == 2 ? kNW : kNW / 2) + 1); mbar_init(m.empty0 + 8 * s, 1); }
    mbar_init(m.accfull, 1);
    mbar_init(m.accempty, kNW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kNW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(m.tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  for (int i = tid; i < Hq; i += kTcThreads) { m.A0s[i] = A0q_g[i]; m.A1s[i] = A1q_g[i]; m.P1s[i] = P1q_g[i]; }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *m.tmem_slot;
}
__device__ __forceinline__ void tc_teardown(uint32_t tmem_base) {
  tc_fence_before();
  __syncthreads();
  if ((threadIdx.x >> 5) == kNW) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// TMA producer: for each GEMM g (maps hi[g]/lo[g]) stream NP x NKB B tiles through the stage ring.
// Called by the WHOLE (converged) warp; one elected lane issues (see elect_one()).
template <bool X3>
__device__ __forceinline__ void tc_tma_role(const TcSmem& m, const CUtensorMap* const* hi, const CUtensorMap* const* lo,
                                            int ngemm, int NP, int NKB) {
  using C = TcCfg<X3>;
  uint32_t it = 0;
  const int NST = NKB / C::KPS;
  for (int g = 0; g < ngemm; ++g)
    for (int p = 0; p < NP; ++p)
      for (int st = 0; st < NST; ++st, ++it) {
        const uint32_t s = it % C::S, ph = (it / C::S) & 1;
        mbar_wait(m.empty0 + 8 * s, ph ^ 1);
        if (elect_one()) {
          const uint32_t bar = m.full0 + 8 * s;
          mbar_arrive_expect_tx(bar, C::KPS * (X3 ? 2 : 1) * kTileBytes);
#pragma unroll
          for (int j = 0; j < C::KPS; ++j) {
            const uint32_t dst = smem_u32(m.stages + s * C::kStageBytes + j * C::kSubBytes);
            const int kb = st * C::KPS + j;
            tma_load_2d(dst + C::kOffB, hi[g], bar, kb * kKB, p * kTN);
            if (X3) tma_load_2d(dst + C::kOffBlo, lo[g], bar, kb * kKB, p * kTN);
          }
        }
        __syncwarp();
      }
}
// MMA issuer: npass accumulator passes of NKB K-blocks; D[256 x 256] = two M=128 blocks sharing B.
// Whole converged warp; one elected lane issues; descriptors advance by constants from one base.
template <bool X3>
__device__ __forceinline__ void tc_mma_role(const TcSmem& m, uint32_t tmem_base, int npass, int NKB) {
  using C = TcCfg<X3>;
  uint32_t it = 0;
  const int NST = NKB / C::KPS;
  const uint64_t desc0 = make_desc_sw64(smem_u32(m.stages));
  for (int pp = 0; pp < npass; ++pp) {
    mbar_wait(m.accempty, (pp & 1) ^ 1);
    tc_fence_after();
    for (int st = 0; st < NST; ++st, ++it) {
      const uint32_t s = it % C::S, ph = (it / C::S) & 1;
      mbar_wait(m.full0 + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int j = 0; j < C::KPS; ++j) {
          const uint64_t sa = desc0 + (uint64_t)((s * C::kStageBytes + j * C::kSubBytes) >> 4);
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t d_t = tmem_base + (uint32_t)(half * kTN);
#pragma unroll
            for (int ks = 0; ks < 2; ++ks) {                         // two K=8 steps per 64-byte row
              const uint64_t a_hi = sa + (uint64_t)((half * (kTileBytes / 2) + ks * 32) >> 4);
              const uint64_t b_hi = sa + (uint64_t)((C::kOffB + ks * 32) >> 4);
              const uint32_t acc = (st | j | ks) ? 1u : 0u;
              if (X3) {
                const uint64_t a_lo = a_hi + (uint64_t)(C::kOffAlo >> 4);
                const uint64_t b_lo = sa + (uint64_t)((C::kOffBlo + ks * 32) >> 4);
                umma_tf32(d_t, a_lo, b_hi, kIdescTf32, acc);
                umma_tf32(d_t, a_hi, b_lo, kIdescTf32, 1u);
                umma_tf32(d_t, a_hi, b_hi, kIdescTf32, 1u);
              } else {
                umma_tf32(d_t, a_hi, b_hi, kIdescTf32, acc);
              }
            }
          }
        }
        umma_commit(m.empty0 + 8 * s);       // frees the smem stage when these MMAs have read it
        if (st == NST - 1) umma_commit(m.accfull);   // accumulator pass complete
      }
      __syncwarp();
    }
  }
}

// per-thread worker context: row = tid & 255; `kh` selects the K half it generates and the column half it drains
struct Worker {
  int row, kh, warp, lane, rsw;
  uint32_t a_row_off, taddr, it, pp;
};
__device__ __forceinline__ Worker make_worker(uint32_t tmem_base) {
  Worker w;
  const int tid = threadIdx.x;
  w.row = tid & 255; w.kh = tid >> 8; w.warp = tid >> 5; w.lane = tid & 31;
  w.rsw = (w.row >> 1) & 3;
  w.a_row_off = (uint32_t)w.row * 64u;
  // TMEM: lane quarter = warp%4, accumulator half (rows 128..255) = (warp>>2)&1, column half = kh
  w.taddr = tmem_base + ((uint32_t)((w.warp & 3) * 32) << 16) + (uint32_t)(((w.warp >> 2) & 1) * kTN + w.kh * (kTN / 2));
  w.it = 0; w.pp = 0;
  return w;
}
// The two threads of a row ALTERNATE K-blocks (thread kh produces the blocks with kb % 2 == kh, all 16 elements of
// the row): every warp then has two MMA stage-times to turn one stage around, which hides the LDS -> FMA -> STS ->
// proxy-fence -> arrive latency chain that otherwise paces the pipeline.  gen(k, e) -> A[row, k], e = k - 16*kb.
template <bool X3, class Gen>
__device__ __forceinline__ void worker_produce(const TcSmem& m, Worker& w, int kb, Gen&& gen) {
  using C = TcCfg<X3>;
  if ((kb & 1) != w.kh) { ++w.it; return; }                    // w.it counts K-blocks (same sequence in every role)
  const uint32_t stg = w.it / C::KPS, s = stg % C::S, ph = (stg / C::S) & 1, sub = w.it % C::KPS;
  float v[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) v[e] = gen(kb * kKB + e, e);
  mbar_wait(m.empty0 + 8 * s, ph ^ 1);
  unsigned char* At = m.stages + s * C::kStageBytes + sub * C::kSubBytes;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const uint32_t off = w.a_row_off + ((uint32_t)(c ^ w.rsw) << 4);
    const float4 hi = make_float4(to_tf32(v[c * 4 + 0]), to_tf32(v[c * 4 + 1]), to_tf32(v[c * 4 + 2]), to_tf32(v[c * 4 + 3]));
    *reinterpret_cast<float4*>(At + off) = hi;
    if (X3)
      *reinterpret_cast<float4*>(At + C::kOffAlo + off) =
          make_float4(to_tf32(v[c * 4 + 0] - hi.x), to_tf32(v[c * 4 + 1] - hi.y), to_tf32(v[c * 4 + 2] - hi.z),
                      to_tf32(v[c * 4 + 3] - hi.w));
  }
  fence_async_smem();
  __syncwarp();
  if (w.lane == 0) mbar_arrive(m.full0 + 8 * s);
  ++w.it;
}
// ---- 4-rows-per-thread generator (forward kernel) ----------------------------------------------------------
// Shared memory bandwidth is the scarce resource (every UMMA reads 12 KB of operands from it; ncu: LSU wavefronts
// 54 % + tensor reads 45 % of the pipe).  With one row per thread every A element costs one broadcast LDS.128 of
// the unit's parameters; here a thread owns ONE 16-byte chunk (4 consecutive k) of FOUR rows (rb, rb+64, rb+128,
// rb+192), so a parameter load is reused by 4 rows: 4x fewer parameter wavefronts, same STS traffic.
struct Gen4 {
  int c, rb;            // chunk (0..3) and base row (0..63)
  uint32_t off;         // byte offset of my chunk in row rb of an A tile (rows rb+64j: + j*4096)
};
__device__ __forceinline__ Gen4 make_gen4() {
  Gen4 g;
  const int t = threadIdx.x & 255;
  g.c = t & 3; g.rb = t >> 2;
  g.off = (uint32_t)g.rb * 64u + ((uint32_t)(g.c ^ ((g.rb >> 1) & 3)) << 4);
  return g;
}
// gen(k, e, out[4]) fills the value of column k for my 4 rows
template <bool X3, class Gen>
__device__ __forceinline__ void worker_produce4(const TcSmem& m, Worker& w, const Gen4& g, int kb, Gen&& gen) {
  using C = TcCfg<X3>;
  if ((kb & 1) != w.kh) { ++w.it; return; }
  const uint32_t stg = w.it / C::KPS, s = stg % C::S, ph = (stg / C::S) & 1, sub = w.it % C::KPS;
  float v[4][4];        // [row j][e]
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    float o[4];
    gen(kb * kKB + g.c * 4 + e, e, o);
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j][e] = o[j];
  }
  mbar_wait(m.empty0 + 8 * s, ph ^ 1);
  unsigned char* At = m.stages + s * C::kStageBytes + sub * C::kSubBytes + g.off;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 hi = make_float4(to_tf32(v[j][0]), to_tf32(v[j][1]), to_tf32(v[j][2]), to_tf32(v[j][3]));
    *reinterpret_cast<float4*>(At + j * 4096) = hi;
    if (X3)
      *reinterpret_cast<float4*>(At + C::kOffAlo + j * 4096) =
          make_float4(to_tf32(v[j][0] - hi.x), to_tf32(v[j][1] - hi.y), to_tf32(v[j][2] - hi.z), to_tf32(v[j][3] - hi.w));
  }
  fence_async_smem();
  __syncwarp();
  if (w.lane == 0) mbar_arrive(m.full0 + 8 * s);
  ++w.it;
}

// drain my row's 128 accumulator columns of the finished pass: chunk(r[32], first_column_in_pass)
template <class Chunk>
__device__ __forceinline__ void worker_drain(const TcSmem& m, Worker& w, Chunk&& chunk) {
  mbar_wait(m.accfull, w.pp & 1);
  tc_fence_after();
#pragma unroll 1
  for (int cc = 0; cc < kTN / 64; ++cc) {
    uint32_t r[32];
    tmem_ld32(w.taddr + cc * 32, r);
    tmem_ld_wait();
    chunk(r, w.kh * (kTN / 2) + cc * 32);
  }
  tc_fence_before();
  __syncwarp();
  if (w.lane == 0) mbar_arrive(m.accempty);
  ++w.pp;
}

// sum over the 32 lanes of a warp of 32 per-lane values: lane l returns sum_lanes e[l] (31 shuffles).
// `a` holds the 16 values left after the caller folded lane bit 4 (columns i / i+16).
__device__ __forceinline__ float fold16(float (&a)[16], int lane) {
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? a[i] : a[i + off];
      const float keep = up ? a[i + off] : a[i];
      a[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return a[0];
}
__device__ __forceinline__ float fold_first(float lo_col, float hi_col, int lane) {   // lane bit 4: columns i vs i+16
  const bool up = (lane & 16) != 0;
  const float send = up ? lo_col : hi_col, keep = up ? hi_col : lo_col;
  return keep + __shfl_xor_sync(0xffffffffu, send, 16);
}

// =========================================== forward =================================================
template <int D, bool X3>
__global__ void __launch_bounds__(kTcThreads, 1)
icnn_tc_fwd_kernel(const __grid_constant__ TcMaps maps, const float* __restrict__ z, int B, int Hq, int Hw_out,
                   float kappa, const float4* __restrict__ A0q_g, const float4* __restrict__ A1q_g,
                   const float* __restrict__ P1q_g, const float* __restrict__ A2p, float* __restrict__ psi,
                   float* __restrict__ xhat, uint32_t* __restrict__ mask1, uint8_t* __restrict__ mask2) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const TcSmem m = carve<X3>(smem_raw, Hq);
  const uint32_t tmem_base = tc_setup<X3>(m, Hq, A0q_g, A1q_g, P1q_g);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);      // provably warp-uniform role index
  const int m0 = blockIdx.x * kTM;
  const int NP = Hq / kTN, NKB = Hq / kKB;
  const int ngemm = (xhat != nullptr) ? 2 : 1;
  (void)lane;

  if (warp_u < kNW) {
    Worker w = make_worker(tmem_base);
    const Gen4 g = make_gen4();
    const bool valid = (m0 + w.row) < B;
    float zr[D];                       // my drain row
    float z4[4][D];                    // my 4 generator rows
#pragma unroll
    for (int j = 0; j < D; ++j) zr[j] = valid ? z[(size_t)(m0 + w.row) * D + j] : 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < D; ++j)
        z4[r][j] = (m0 + g.rb + 64 * r < B) ? z[(size_t)(m0 + g.rb + 64 * r) * D + j] : 0.f;

    // -------- GEMM1: h1 = x1 . P^T  -> masks, h2 --------
    float h2 = 0.f;
    for (int p = 0; p < NP; ++p) {
      for (int kb = 0; kb < NKB; ++kb)
        worker_produce4<X3>(m, w, g, kb, [&](int k, int, float (&o)[4]) {
          const float4 q = m.A0s[k];
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            const float h = lin_of<D>(q, z4[r]);
            const float a0 = fmaxf(h, kSlope * h);          // LeakyReLU(0.2) = max(h, 0.2h)
            o[r] = a0 * a0;
          }
        });
      worker_drain(m, w, [&](uint32_t (&r)[32], int c0) {
        const int nb = p * kTN + c0;
        uint32_t word = 0;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float4 q1 = m.A1s[nb + j];
          const float h1 = __uint_as_float(r[j]) + lin_of<D>(q1, zr);
          const bool pos = h1 > 0.f;
          h2 = fmaf(D <= 2 ? q1.z : m.P1s[nb + j], pos ? h1 : kSlope * h1, h2);
          word |= (pos ? 1u : 0u) << j;
        }
        m.maskw[(nb >> 5) * 256 + w.row] = word;
      });
    }
    m.xch[w.kh * 256 + w.row] = h2;
    worker_bar();                                            // h2 halves + all mask words visible
    auto s2_of = [&](int row, const float (&zz)[D]) {
      float hh = m.xch[row] + m.xch[256 + row];
      float lin = A2p[D];
#pragma unroll
      for (int j = 0; j < D; ++j) lin = fmaf(A2p[j], zz[j], lin);
      return hh + lin;
    };
    h2 = s2_of(w.row, zr);
    const bool pos2 = h2 > 0.f;
    const float s2 = pos2 ? 1.f : kSlope;
    float s24[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) s24[r] = s2_of(g.rb + 64 * r, z4[r]) > 0.f ? 1.f : kSlope;
    if (valid) {
      if (w.kh == 0) {
        if (psi) psi[m0 + w.row] = pos2 ? h2 : kSlope * h2;
        if (mask2) mask2[m0 + w.row] = pos2 ? 1 : 0;
      }
      if (mask1)
        for (int wd = w.kh; wd < Hw_out; wd += 2) mask1[(size_t)(m0 + w.row) * Hw_out + wd] = m.maskw[wd * 256 + w.row];
    }
    if (xhat != nullptr) {
      // -------- GEMM2: gx1 = g1 . P -> g0 -> xhat --------
      float xacc[D], xa4[4][D];
#pragma unroll
      for (int j = 0; j < D; ++j) {
        xacc[j] = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) xa4[r][j] = 0.f;
      }
      for (int p = 0; p < NP; ++p) {
        for (int kb = 0; kb < NKB; ++kb) {
          uint32_t bits4[4];
          if ((kb & 1) == w.kh) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
              bits4[r] = m.maskw[(kb >> 1) * 256 + g.rb + 64 * r] >> ((kb & 1) * 16 + g.c * 4);
          }
          worker_produce4<X3>(m, w, g, kb, [&](int k, int e, float (&o)[4]) {
            const float4 q = m.A1s[k];
            const float p1 = D <= 2 ? q.z : m.P1s[k];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
              const float c1 = s24[r] * p1;
              const float g1 = ((bits4[r] >> e) & 1u) ? c1 : kSlope * c1;
              o[r] = g1;
              if (p == 0) {                                  // xhat += A1^T g1, once
#pragma unroll
                for (int j = 0; j < D; ++j) xa4[r][j] = fmaf(comp(q, j), g1, xa4[r][j]);
              }
            }
          });
        }
        worker_drain(m, w, [&](uint32_t (&r)[32], int c0) {
          const int nb = p * kTN + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float4 q = m.A0s[nb + j];
            const float h = lin_of<D>(q, zr);
            const float s0 = slope_of(h), a0 = h * s0;
            const float g0 = __uint_as_float(r[j]) * (2.f * a0) * s0;
#pragma unroll
            for (int jj = 0; jj < D; ++jj) xacc[jj] = fmaf(comp(q, jj), g0, xacc[jj]);
          }
        });
      }
      // combine: 2 drain threads per row (xacc) + 2 groups x 4 chunk lanes per row (xa4)
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int j = 0; j < D; ++j) {
          float t = xa4[r][j];
          t += __shfl_xor_sync(0xffffffffu, t, 1);
          t += __shfl_xor_sync(0xffffffffu, t, 2);
          xa4[r][j] = t;
        }
      worker_bar();                                          // xch reuse (everybody has read the h2 halves)
#pragma unroll
      for (int j = 0; j < D; ++j) m.xch[(w.kh * 256 + w.row) * 4 + j] = xacc[j];
      if (g.c == 0) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int j = 0; j < D; ++j) m.xch[((2 + w.kh) * 256 + g.rb + 64 * r) * 4 + j] = xa4[r][j];
      }
      worker_bar();
      if (valid && w.kh == 0) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const float sum = (m.xch[w.row * 4 + j] + m.xch[(256 + w.row) * 4 + j]) +
                            (m.xch[(512 + w.row) * 4 + j] + m.xch[(768 + w.row) * 4 + j]);
          xhat[(size_t)(m0 + w.row) * D + j] = fmaf(2.f * kappa, zr[j], fmaf(s2, A2p[j], sum));
        }
      }
    }
  } else if (warp_u == kNW) {
    const CUtensorMap* hi[2] = {&maps.b1hi, &maps.b2hi};
    const CUtensorMap* lo[2] = {&maps.b1lo, &maps.b2lo};
    tc_tma_role<X3>(m, hi, lo, ngemm, NP, NKB);
  } else {
    tc_mma_role<X3>(m, tmem_base, ngemm * NP, NKB);
  }
  tc_teardown(tmem_base);
}

// ====================================== backward, sample rows =========================================
// GEMM-A  gx1 = g1 . P      -> t0, g0: row-local dz, column sums dA0w / dA0b
// GEMM-B  w1  = u1 + q1.P^T -> column sums dP1 / dA1w
// Column sums over the CTA's 256 rows are reduced inside each warp (32 rows) with a transposing shuffle
// reduction and written as ordered partials  part[(mtile*8 + rowgroup)][f][Hq]  (finalize sums them).
template <int D, bool X3>
__global__ void __launch_bounds__(kTcThreads, 1)
icnn_tc_bwd_rows_kernel(const __grid_constant__ TcMaps maps, const float* __restrict__ z, const float* __restrict__ v,
                        const uint32_t* __restrict__ mask1, const uint8_t* __restrict__ mask2, int B, int Hq,
                        int Hw_in, float kappa, const float4* __restrict__ A0q_g, const float4* __restrict__ A1q_g,
                        const float* __restrict__ P1q_g, float* __restrict__ dz, float* __restrict__ partA,
                        float* __restrict__ partB, float* __restrict__ a2part) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const TcSmem m = carve<X3>(smem_raw, Hq);
  const uint32_t tmem_base = tc_setup<X3>(m, Hq, A0q_g, A1q_g, P1q_g);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);      // provably warp-uniform role index
  const int m0 = blockIdx.x * kTM;
  const int NP = Hq / kTN, NKB = Hq / kKB;
  constexpr int NF = D + 1;

  if (warp_u < kNW) {
    Worker w = make_worker(tmem_base);
    const bool valid = (m0 + w.row) < B;
    float zr[D], vr[D];
#pragma unroll
    for (int j = 0; j < D; ++j) {
      zr[j] = valid ? z[(size_t)(m0 + w.row) * D + j] : 0.f;
      vr[j] = valid ? v[(size_t)(m0 + w.row) * D + j] : 0.f;
    }
    const float s2 = valid ? (mask2[m0 + w.row] ? 1.f : kSlope) : 0.f;     // 0 kills every term of padded rows
    for (int wd = w.kh; wd < Hq / 32; wd += 2)
      m.maskw[wd * 256 + w.row] = (valid && wd < Hw_in) ? mask1[(size_t)(m0 + w.row) * Hw_in + wd] : 0u;
    worker_bar();
    float* pA = partA + (size_t)(blockIdx.x * 8 + (w.warp & 7)) * NF * Hq;
    float* pB = partB + (size_t)(blockIdx.x * 8 + (w.warp & 7)) * NF * Hq;
    float dzacc[D];
#pragma unroll
    for (int j = 0; j < D; ++j) dzacc[j] = 0.f;

    // -------- GEMM-A --------
    for (int p = 0; p < NP; ++p) {
      for (int kb = 0; kb < NKB; ++kb) {
        const uint32_t bits = m.maskw[(kb >> 1) * 256 + w.row] >> ((kb & 1) * 16);
        worker_produce<X3>(m, w, kb, [&](int k, int e) {
          const float c1 = s2 * m.P1s[k];
          return ((bits >> e) & 1u) ? c1 : kSlope * c1;
        });
      }
      worker_drain(m, w, [&](uint32_t (&r)[32], int c0) {
        const int nb = p * kTN + c0;
        float acc[NF][16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float e[2][NF];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int j = i + 16 * hh;
            const float4 q = m.A0s[nb + j];
            const float h = lin_of<D>(q, zr), u0 = dot_of<D>(q, vr);
            const float s0 = slope_of(h), a0 = h * s0, gx1 = __uint_as_float(r[j]);
            const float g0 = gx1 * (2.f * a0) * s0;
            const float t0 = u0 * (2.f * gx1) * s0 * s0;
#pragma unroll
            for (int jj = 0; jj < D; ++jj) {
              dzacc[jj] = fmaf(comp(q, jj), t0, dzacc[jj]);
              e[hh][jj] = fmaf(g0, vr[jj], t0 * zr[jj]);
            }
            e[hh][D] = t0;
          }
#pragma unroll
          for (int f = 0; f < NF; ++f) acc[f][i] = fold_first(e[0][f], e[1][f], lane);
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) pA[(size_t)f * Hq + nb + lane] = fold16(acc[f], lane);
      });
    }
    // -------- GEMM-B --------
    for (int p = 0; p < NP; ++p) {
      for (int kb = 0; kb < NKB; ++kb)
        worker_produce<X3>(m, w, kb, [&](int k, int) {
          const float4 q = m.A0s[k];
          const float h = lin_of<D>(q, zr), u0 = dot_of<D>(q, vr);
          const float s0 = slope_of(h), a0 = h * s0;
          return u0 * (2.f * a0) * s0;
        });
      worker_drain(m, w, [&](uint32_t (&r)[32], int c0) {
        const int nb = p * kTN + c0;
        const uint32_t word = m.maskw[(nb >> 5) * 256 + w.row];
        float acc[NF][16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float e[2][NF];
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int j = i + 16 * hh;
            const float4 q = m.A1s[nb + j];
            const float s1 = ((word >> j) & 1u) ? 1.f : kSlope;
            const float w1 = __uint_as_float(r[j]) + dot_of<D>(q, vr);
            const float g1 = (s2 * m.P1s[nb + j]) * s1;
#pragma unroll
            for (int jj = 0; jj < D; ++jj) e[hh][jj] = g1 * vr[jj];
            e[hh][D] = (s2 * s1) * w1;
          }
#pragma unroll
          for (int f = 0; f < NF; ++f) acc[f][i] = fold_first(e[0][f], e[1][f], lane);
        }
#pragma unroll
        for (int f = 0; f < NF; ++f) pB[(size_t)f * Hq + nb + lane] = fold16(acc[f], lane);
      });
    }
    // -------- rows: dz = A0^T t0 + 2 kappa v ; dA2w = sum_m s2 v --------
#pragma unroll
    for (int j = 0; j < D; ++j) m.xch[(w.kh * 256 + w.row) * 4 + j] = dzacc[j];
    worker_bar();
    if (w.kh == 0) {
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const float g = fmaf(2.f * kappa, vr[j], m.xch[w.row * 4 + j] + m.xch[(256 + w.row) * 4 + j]);
        if (valid && dz) dz[(size_t)(m0 + w.row) * D + j] = g;
      }
    }
    worker_bar();
    if (w.kh == 0) {
#pragma unroll
      for (int j = 0; j < D; ++j) {
        const float sred = warp_sum(s2 * vr[j]);
        if (lane == 0) m.xch[w.warp * 4 + j] = sred;
      }
    }
    worker_bar();
    if (threadIdx.x < D) {
      float sred = 0.f;
      for (int q = 0; q < 8; ++q) sred += m.xch[q * 4 + threadIdx.x];
      a2part[(size_t)blockIdx.x * D + threadIdx.x] = sred;
    }
  } else if (warp_u == kNW) {
    const CUtensorMap* hi[2] = {&maps.b2hi, &maps.b1hi};
    const CUtensorMap* lo[2] = {&maps.b2lo, &maps.b1lo};
    tc_tma_role<X3>(m, hi, lo, 2, NP, NKB);
  } else {
    tc_mma_role<X3>(m, tmem_base, 2 * NP, NKB);
  }
  tc_teardown(tmem_base);
}

// ====================================== backward, dP0 (batch-reduced) ===================================
// dP0part[split][o][n] = sum_{m in split} (1 + 4 bit[m,o]) * s2[m] q1[m,n]   (0.2 P1[o] and the exp/clamp chain: finalize)
// Output-stationary 256 x 256 tile in TMEM, K = the CTA's batch slice.  BOTH operands are generated: a sample
// (= K index) is owned by a thread, which emits 8 consecutive o's / n's as 16-byte chunks -> MN-major UMMA
// layout.  For 32-bit MN-major operands the only legal UMMA layout is SWIZZLE_128B_BASE32B (cute
// Layout_MN_SW128_32B_Atom): atoms of 4 k-rows x 128 B (32 MN elements), 32-byte chunks XOR-ed with the k-row;
// LBO = 512 B between 32-wide MN blocks, SBO = 4 KB between groups of 4 k.  No TMA, no B matrix.
constexpr int kDpThreads = (kNW + 1) * 32;
constexpr uint32_t kIdescTf32MN = kIdescTf32 | (1u << 15) | (1u << 16);     // a_major = b_major = MN
__device__ __forceinline__ uint64_t make_desc_mn_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(512 >> 4) << 16) | ((uint64_t)(4096 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)1 << 61);     // layout type 1 = SWIZZLE_128B_BASE32B
}

template <int D, bool X3>
__global__ void __launch_bounds__(kDpThreads, 1)
icnn_tc_dP0_kernel(const float* __restrict__ z, const float* __restrict__ v, const uint32_t* __restrict__ mask1,
                   const uint8_t* __restrict__ mask2, int B, int Hq, int Hw_in, int rows_per_split,
                   const float4* __restrict__ A0q_g, float* __restrict__ dP0part) {
  constexpr int S = X3 ? 2 : 5;
  constexpr int kStage = (X3 ? 4 : 2) * kTileBytes;
  constexpr int kOffAlo = kTileBytes, kOffB = (X3 ? 2 : 1) * kTileBytes, kOffBlo = 3 * kTileBytes;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* stages = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* samp = reinterpret_cast<float*>(stages + S * kStage);            // [2][2D+1][256]  z, v, s2 per sample
  uint32_t* sampw = reinterpret_cast<uint32_t*>(samp + 2 * 256 * (2 * D + 1));   // [2][8][256] mask words
  uint64_t* bars = reinterpret_cast<uint64_t*>(sampw + 2 * 8 * 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * S + 1);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + S), accfull = smem_u32(bars + 2 * S);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int warp_u = __shfl_sync(0xffffffffu, warp, 0);      // provably warp-uniform role index
  const int n0 = blockIdx.x * kTN, o0 = blockIdx.y * kTM, split = blockIdx.z;
  const int b0 = split * rows_per_split;
  const int b1 = min(B, b0 + rows_per_split);
  const int NKB = (max(b1 - b0, 0) + kKB - 1) / kKB;

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full0 + 8 * s, kNW); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(accfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kNW) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp_u < kNW) {
    const int ks = tid & 15, u = tid >> 4, blk = u >> 2, qd = u & 3;     // sample-in-stage, MN block, 8-wide quarter
    const int g4 = ks >> 2, kr = ks & 3;                                 // group of 4 k, k-row inside the atom
    // my 8 n's as 4 pairs (packed f32x2 math): component arrays (w0, w1, w2, bias)
    float2 qx2[4], qy2[4], qz2[4], qw2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 qa = A0q_g[n0 + blk * 32 + qd * 8 + 2 * i], qb = A0q_g[n0 + blk * 32 + qd * 8 + 2 * i + 1];
      qx2[i] = make_float2(qa.x, qb.x); qy2[i] = make_float2(qa.y, qb.y);
      qz2[i] = make_float2(qa.z, qb.z); qw2[i] = make_float2(qa.w, qb.w);
    }
    const uint32_t off0 = (uint32_t)((g4 * 8 + blk) * 512 + kr * 128) + ((uint32_t)(qd ^ kr) << 5), off1 = off0 + 16;
    // per-sample inputs (z, v, s2, the 8 mask words of this o-tile) are staged through shared memory in
    // chunks of 256 samples, fetched one chunk ahead (register staged) so no global latency is exposed
    constexpr int CH = 256, ZV = 2 * D + 1;                  // floats per sample: z, v, s2
    const int nchunk = (NKB * kKB + CH - 1) / CH;
    float pre_f[ZV];
    uint32_t pre_w[4];
    auto fetch = [&](int c) {                                // thread -> sample (tid&255), word half (tid>>8)
      const int mrow = b0 + c * CH + (tid & 255);
      const bool in = mrow < b1;
      if (tid < CH) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
          pre_f[j] = in ? __ldg(z + (size_t)mrow * D + j) : 0.f;
          pre_f[D + j] = in ? __ldg(v + (size_t)mrow * D + j) : 0.f;
        }
        pre_f[2 * D] = in ? (__ldg(mask2 + mrow) ? 1.f : kSlope) : 0.f;
      }
      const int w0 = (o0 >> 5) + (tid >> 8) * 4;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) pre_w[q4] = (in && w0 + q4 < Hw_in) ? __ldg(mask1 + (size_t)mrow * Hw_in + w0 + q4) : 0u;
    };
    auto stash = [&](int buf) {
      float* zf = samp + buf * (CH * ZV);
      uint32_t* mw = sampw + buf * (CH * 8);
      if (tid < CH) {
#pragma unroll
        for (int j = 0; j < ZV; ++j) zf[j * CH + tid] = pre_f[j];
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) mw[((tid >> 8) * 4 + q4) * CH + (tid & 255)] = pre_w[q4];
    };
    if (nchunk > 0) { fetch(0); stash(0); }
    worker_bar();
    for (int kb = 0; kb < NKB; ++kb) {
      const int c = kb >> 4, buf = c & 1, sl = (kb & 15) * kKB + ks;       // sample slot inside the chunk
      if ((kb & 15) == 0 && c + 1 < nchunk) fetch(c + 1);
      const float* zf = samp + buf * (CH * ZV);
      float zr[D], vr[D];
#pragma unroll
      for (int j = 0; j < D; ++j) { zr[j] = zf[j * CH + sl]; vr[j] = zf[(D + j) * CH + sl]; }
      const float s2f = zf[2 * D * CH + sl];
      const uint32_t bits = sampw[buf * (CH * 8) + blk * CH + sl] >> (qd * 8);
      // A = 1 + 4*bit (the LeakyReLU slope / 0.2, exact in tf32; 0.2 and P1 are applied by finalize_W0);
      // B = s2 q1 = (2 s2 A0 v) . max(h0, 0.04 h0), two n's per packed instruction
      float av[8], bv[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) av[e] = ((bits >> e) & 1u) ? 5.f : 1.f;
      float sv[D];
#pragma unroll
      for (int j = 0; j < D; ++j) sv[j] = (2.f * s2f) * vr[j];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 h = __ffma2_rn(qx2[i], make_float2(zr[0], zr[0]), qw2[i]);
        float2 uu = __fmul2_rn(qx2[i], make_float2(sv[0], sv[0]));
        if (D > 1) {
          h = __ffma2_rn(qy2[i], make_float2(zr[D > 1 ? 1 : 0], zr[D > 1 ? 1 : 0]), h);
          uu = __ffma2_rn(qy2[i], make_float2(sv[D > 1 ? 1 : 0], sv[D > 1 ? 1 : 0]), uu);
        }
        if (D > 2) {
          h = __ffma2_rn(qz2[i], make_float2(zr[D > 2 ? 2 : 0], zr[D > 2 ? 2 : 0]), h);
          uu = __ffma2_rn(qz2[i], make_float2(sv[D > 2 ? 2 : 0], sv[D > 2 ? 2 : 0]), uu);
        }
        const float2 l = __fmul2_rn(h, make_float2(kSlope * kSlope, kSlope * kSlope));
        const float2 x = __fmul2_rn(uu, make_float2(fmaxf(h.x, l.x), fmaxf(h.y, l.y)));
        bv[2 * i] = x.x; bv[2 * i + 1] = x.y;
      }
      if ((kb & 15) == 15 && c + 1 < nchunk) {               // next chunk's buffer was last read 16 stages ago
        worker_bar();
        stash(buf ^ 1);
        worker_bar();
      }
      const uint32_t s = kb % S, ph = (kb / S) & 1;
      mbar_wait(empty0 + 8 * s, ph ^ 1);
      unsigned char* st = stages + s * kStage;
      *reinterpret_cast<float4*>(st + off0) = make_float4(av[0], av[1], av[2], av[3]);       // exact: no lo part
      *reinterpret_cast<float4*>(st + off1) = make_float4(av[4], av[5], av[6], av[7]);
      if (X3) {
        const float4 h0 = make_float4(rn_tf32_masked(bv[0]), rn_tf32_masked(bv[1]), rn_tf32_masked(bv[2]), rn_tf32_masked(bv[3]));
        const float4 h1 = make_float4(rn_tf32_masked(bv[4]), rn_tf32_masked(bv[5]), rn_tf32_masked(bv[6]), rn_tf32_masked(bv[7]));
        *reinterpret_cast<float4*>(st + kOffB + off0) = h0;
        *reinterpret_cast<float4*>(st + kOffB + off1) = h1;
        *reinterpret_cast<float4*>(st + kOffBlo + off0) =
            make_float4(rn_tf32_fast(bv[0] - h0.x), rn_tf32_fast(bv[1] - h0.y), rn_tf32_fast(bv[2] - h0.z), rn_tf32_fast(bv[3] - h0.w));
        *reinterpret_cast<float4*>(st + kOffBlo + off1) =
            make_float4(rn_tf32_fast(bv[4] - h1.x), rn_tf32_fast(bv[5] - h1.y), rn_tf32_fast(bv[6] - h1.z), rn_tf32_fast(bv[7] - h1.w));
      } else {
        *reinterpret_cast<float4*>(st + kOffB + off0) = make_float4(rn_tf32_masked(bv[0]), rn_tf32_masked(bv[1]), rn_tf32_masked(bv[2]), rn_tf32_masked(bv[3]));
        *reinterpret_cast<float4*>(st + kOffB + off1) = make_float4(rn_tf32_masked(bv[4]), rn_tf32_masked(bv[5]), rn_tf32_masked(bv[6]), rn_tf32_masked(bv[7]));
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full0 + 8 * s);
    }
    // epilogue: my o-row, 128 of the 256 columns
    const int orow = o0 + ((warp >> 2) & 1) * 128 + (warp & 3) * 32 + lane;
    const int chalf = warp >> 3;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(((warp >> 2) & 1) * kTN + chalf * 128);
    float* out = dP0part + ((size_t)split * Hq + orow) * Hq + n0 + chalf * 128;
    if (NKB > 0) {
      mbar_wait(accfull, 0);
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < 4; ++cc) {
        uint32_t r[32];
        tmem_ld32(taddr + cc * 32, r);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(out + cc * 32 + j) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
      }
    } else {
      for (int j = 0; j < 128; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  } else {
    // MMA issuer: whole converged warp, one elected lane issues; descriptors advance by constants
    const uint64_t desc0 = make_desc_mn_sw128(smem_u32(stages));
    for (int kb = 0; kb < NKB; ++kb) {
      const uint32_t s = kb % S, ph = (kb / S) & 1;
      mbar_wait(full0 + 8 * s, ph);
      tc_fence_after();
      if (elect_one()) {
        const uint64_t sa = desc0 + (uint64_t)((s * kStage) >> 4);
#pragma unroll
        for (int g = 0; g < 2; ++g) {
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const uint32_t d_t = tmem_base + (uint32_t)(half * kTN);
            const uint64_t a_hi = sa + (uint64_t)((g * 8192 + half * 2048) >> 4);     // K=8 = two 4-k groups
            const uint64_t b_hi = sa + (uint64_t)((kOffB + g * 8192) >> 4);
            const uint32_t acc = (kb | g) ? 1u : 0u;
            if (X3) {                                                        // A is exact: a_hi.b_lo + a_hi.b_hi
              const uint64_t b_lo = sa + (uint64_t)((kOffBlo + g * 8192) >> 4);
              umma_tf32(d_t, a_hi, b_lo, kIdescTf32MN, acc);
              umma_tf32(d_t, a_hi, b_hi, kIdescTf32MN, 1u);
            } else {
              umma_tf32(d_t, a_hi, b_hi, kIdescTf32MN, acc);
            }
          }
        }
        umma_commit(empty0 + 8 * s);
        if (kb == NKB - 1) umma_commit(accfull);
      }
      __syncwarp();
    }
  }
  tc_teardown(tmem_base);
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}
static int make_map(CUtensorMap* m, const float* base, int Hq) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return B200VAE_ECUDA;
  const cuuint64_t dims[2] = {(cuuint64_t)Hq, (cuuint64_t)Hq};
  const cuuint64_t strides[1] = {(cuuint64_t)Hq * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)kKB, (cuuint32_t)kTN};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { g_last_cuda_error = 100000 + (int)r; return B200VAE_ECUDA; }
  return B200VAE_OK;
}

int tc_prepare(const b200vae_icnn_params* p, int d, int H, int mode, int precision, float* ws, cudaStream_t st) {
  (void)p; (void)mode;
  if (precision == 2 /* reserved */) return B200VAE_EUNSUP;
  if (d > 3) return B200VAE_EUNSUP;
  const WsLayout L = ws_layout(1, d, H);
  const TcLayout T = tc_layout(d, H);
  float* tb = tc_base(ws, d, H);
  dim3 grid((T.Hq + 255) / 256, T.Hq);
  tc_prepare_kernel<<<grid, 256, 0, st>>>(ws + L.P0, ws + L.P0T, ws + L.P1, ws + L.A0p, ws + L.A1p, d, L.Hp, T.Hq,
                                          tb + T.B1hi, tb + T.B1lo, tb + T.B2hi, tb + T.B2lo,
                                          reinterpret_cast<float4*>(tb + T.A0q), reinterpret_cast<float4*>(tb + T.A1q),
                                          tb + T.P1q);
  return check_launch();
}

template <int D, bool X3>
static int launch_tc(const TcMaps& maps, const float* z, int B, const TcLayout& T, const float* tb, const float* A2p,
                     int Hw_out, float kappa, float* psi, float* xhat, uint32_t* mask1, uint8_t* mask2, cudaStream_t st) {
  const size_t smem = tc_smem_bytes<X3>(T.Hq);
  if (smem > 227 * 1024) return B200VAE_EUNSUP;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(icnn_tc_fwd_kernel<D, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_done = true;
  }
  const int grid = (B + kTM - 1) / kTM;
  icnn_tc_fwd_kernel<D, X3><<<grid, kTcThreads, smem, st>>>(
      maps, z, B, T.Hq, Hw_out, kappa, reinterpret_cast<const float4*>(tb + T.A0q),
      reinterpret_cast<const float4*>(tb + T.A1q), tb + T.P1q, A2p, psi, xhat, mask1, mask2);
  return check_launch();
}

static int get_maps(const float* tb, const TcLayout& T, TcMaps* out) {
  // tensor maps only encode (address, shape): cache them per prepared buffer
  static std::mutex mu;
  static std::unordered_map<uint64_t, TcMaps> cache;
  std::lock_guard<std::mutex> lk(mu);
  const uint64_t key = reinterpret_cast<uint64_t>(tb) ^ ((uint64_t)T.Hq << 48);
  auto itc = cache.find(key);
  if (itc != cache.end()) { *out = itc->second; return B200VAE_OK; }
  TcMaps maps;
  int rc = make_map(&maps.b1hi, tb + T.B1hi, T.Hq);
  if (!rc) rc = make_map(&maps.b1lo, tb + T.B1lo, T.Hq);
  if (!rc) rc = make_map(&maps.b2hi, tb + T.B2hi, T.Hq);
  if (!rc) rc = make_map(&maps.b2lo, tb + T.B2lo, T.Hq);
  if (rc) return rc;
  if (cache.size() > 256) cache.clear();
  cache.emplace(key, maps);
  *out = maps;
  return B200VAE_OK;
}

int tc_fwd(const float* z, int B, int d, int H, float kappa, float* psi, float* xhat, uint32_t* mask1, uint8_t* mask2,
           int precision, const float* ws, cudaStream_t st) {
  if (precision == 2 /* reserved */ || d > 3) return B200VAE_EUNSUP;
  const WsLayout L = ws_layout(1, d, H);
  const TcLayout T = tc_layout(d, H);
  const float* tb = tc_base(const_cast<float*>(ws), d, H);
  TcMaps maps;
  int rc = get_maps(tb, T, &maps);
  if (rc) return rc;
  const bool x3 = (precision == B200VAE_PREC_TF32X3 || precision == B200VAE_PREC_F16X3);   // single-CTA kernels: 3xTF32 for both
  const int Hw_out = L.Hp / 32;
#define B200VAE_TC(DD)                                                                                          \
  return x3 ? launch_tc<DD, true>(maps, z, B, T, tb, ws + L.A2p, Hw_out, kappa, psi, xhat, mask1, mask2, st)   \
            : launch_tc<DD, false>(maps, z, B, T, tb, ws + L.A2p, Hw_out, kappa, psi, xhat, mask1, mask2, st)
  switch (d) {
    case 1: B200VAE_TC(1);
    case 2: B200VAE_TC(2);
    case 3: B200VAE_TC(3);
    default: return B200VAE_EUNSUP;
  }
#undef B200VAE_TC
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

template <int D, bool X3>
static int launch_tc_bwd(const TcMaps& maps, const float* z, const float* v, const uint32_t* mask1, const uint8_t* mask2,
                         int B, const TcLayout& T, const float* tb, int Hw_in, float kappa, float* dz, float* partA,
                         float* partB, float* a2part, cudaStream_t st) {
  const size_t smem = tc_smem_bytes<X3>(T.Hq);
  if (smem > 227 * 1024) return B200VAE_EUNSUP;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(icnn_tc_bwd_rows_kernel<D, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_done = true;
  }
  const int grid = (B + kTM - 1) / kTM;
  icnn_tc_bwd_rows_kernel<D, X3><<<grid, kTcThreads, smem, st>>>(
      maps, z, v, mask1, mask2, B, T.Hq, Hw_in, kappa, reinterpret_cast<const float4*>(tb + T.A0q),
      reinterpret_cast<const float4*>(tb + T.A1q), tb + T.P1q, dz, partA, partB, a2part);
  return check_launch();
}

int finalize_W0_launch(const float* part, int splits, int H, int Hp, int ldp, const float* P0, const float* P1,
                       const float* W0raw, int mode, float scale, float* dW0, cudaStream_t st);

template <int D, bool X3>
static int launch_tc_dp0(const float* z, const float* v, const uint32_t* mask1, const uint8_t* mask2, int B,
                         const TcLayout& T, const float* tb, int Hw_in, int splits, float* part, cudaStream_t st) {
  constexpr int S = X3 ? 2 : 5;
  const size_t smem = (size_t)S * (X3 ? 4 : 2) * kTileBytes + (size_t)2 * 256 * (2 * D + 1) * 4 + 2 * 8 * 256 * 4 +
                      (2 * S + 1) * 8 + 16 + 1024;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(icnn_tc_dP0_kernel<D, X3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    attr_done = true;
  }
  int rows = (B + splits - 1) / splits;
  rows = round_up(rows, kKB);
  dim3 grid(T.Hq / kTN, T.Hq / kTM, splits);
  icnn_tc_dP0_kernel<D, X3><<<grid, kDpThreads, smem, st>>>(z, v, mask1, mask2, B, T.Hq, Hw_in, rows,
                                                           reinterpret_cast<const float4*>(tb + T.A0q), part);
  return check_launch();
}

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
  TcMaps maps;
  int rc = get_maps(tb, T, &maps);
  if (rc) return rc;
  const size_t nmt = (size_t)(B + kTM - 1) / kTM;
  float* partA = ws + L.end;
  float* partB = partA + nmt * 8 * (d + 1) * T.Hq;
  float* a2part = partB + nmt * 8 * (d + 1) * T.Hq;
  const bool x3 = (precision == B200VAE_PREC_TF32X3 || precision == B200VAE_PREC_F16X3);   // single-CTA kernels: 3xTF32 for both
  const int Hw_in = L.Hp / 32;
  // rows part: persistent pair kernel (icnn_tc3.cu) unless B200VAE_BWD=1 or it does not fit (H > 1024)
  static const int variant = [] { const char* e = getenv("B200VAE_BWD"); return e ? atoi(e) : 3; }();
  float* dp0part = a2part + nmt * 4 + 64;
  dp0part = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dp0part) + 15) & ~(uintptr_t)15);
  rc = phase == 2 ? B200VAE_OK : B200VAE_EUNSUP;
  if (variant == 3 && phase != 2) {
    float* dzpart = dp0part + (size_t)tc_dp0_splits(B, T.Hq) * T.Hq * T.Hq;
    dzpart = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dzpart) + 15) & ~(uintptr_t)15);
    rc = tc3_bwd_rows(z, v, mask1, mask2, B, d, H, kappa, dz, partA, partB, a2part, dzpart, precision, ws, accsave, st);
  }
  if (rc == B200VAE_EUNSUP) {
#define B200VAE_TCB(DD)                                                                                                 \
  rc = x3 ? launch_tc_bwd<DD, true>(maps, z, v, mask1, mask2, B, T, tb, Hw_in, kappa, dz, partA, partB, a2part, st)    \
          : launch_tc_bwd<DD, false>(maps, z, v, mask1, mask2, B, T, tb, Hw_in, kappa, dz, partA, partB, a2part, st)
  switch (d) {
    case 1: B200VAE_TCB(1); break;
    case 2: B200VAE_TCB(2); break;
    case 3: B200VAE_TCB(3); break;
    default: return B200VAE_EUNSUP;
  }
#undef B200VAE_TCB
  }
  if (rc || !g || phase == 1) return rc;
  if (g->W0) {
    float* part = dp0part;
    int splits = tc_dp0_splits(B, T.Hq);
    // CTA-pair kernel (icnn_tc3.cu) unless B200VAE_DP0=1; it never needs more slabs than the workspace holds
    static const int dp_variant = [] { const char* e = getenv("B200VAE_DP0"); return e ? atoi(e) : 3; }();
    rc = B200VAE_EUNSUP;
    if (dp_variant == 3)
      rc = tc3_dp0(z, v, mask1, mask2, B, d, T.Hq, Hw_in, tb + T.A0q, tb + T.end + tc3_layout(1, d, H).sumV, precision, splits,
                   part, &splits, st);
    if (rc == B200VAE_EUNSUP) {
#define B200VAE_TCD(DD)                                                                                       \
  rc = x3 ? launch_tc_dp0<DD, true>(z, v, mask1, mask2, B, T, tb, Hw_in, splits, part, st)                    \
          : launch_tc_dp0<DD, false>(z, v, mask1, mask2, B, T, tb, Hw_in, splits, part, st)
    switch (d) {
      case 1: B200VAE_TCD(1); break;
      case 2: B200VAE_TCD(2); break;
      default: B200VAE_TCD(3); break;
    }
#undef B200VAE_TCD
    }
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
