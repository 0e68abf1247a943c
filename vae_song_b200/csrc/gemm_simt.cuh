// gemm_simt.cuh -- FP32 SIMT 128x128x16 CTA tile with GENERATED operands.
//
// The ICNN hot path is two (forward) / three (backward) H x H contractions per sample whose "A"
// operand never exists in memory: it is recomputed from z (d floats), v and the saved LeakyReLU
// bit-masks.  This core therefore takes functors instead of pointers:
//     a.pre(kt, As)        BEFORE the FMA block of step kt-1: issue cp.async into As / global loads
//                          into registers / or compute-and-store the whole tile
//     a.post(kt, As)       AFTER the FMA block: store register-staged data   As[kk][m]
//     b.pre / b.post       same for the 16 x 128 B tile Bs[kk][n]
// 256 threads, 8x8 register micro-tile per thread, double-buffered shared memory, one barrier per
// K-step.  Thread layout (chosen so that COLUMN reductions stay inside a warp, see icnn_simt.cu):
//     warp w = tid/32, lane l:  ty = l>>1 (0..15), tx = 2w + (l&1) (0..15)
//     rows  r(i) = (i<4 ?  ty*4+i : 64+ty*4+i-4),   cols c(j) = (j<4 ? tx*4+j : 64+tx*4+j-4)
#pragma once
#include "common.cuh"

namespace b200vae {

constexpr int kBM = 128, kBN = 128, kBK = 16, kThreads = 256;

struct __align__(16) GemmSmem {
  float As[2][kBK][kBM];
  float Bs[2][kBK][kBN];
};

struct TileCoord {
  int w, lane, ty, tx, txl;
  __device__ __forceinline__ TileCoord() {
    w = threadIdx.x >> 5; lane = threadIdx.x & 31; ty = lane >> 1; txl = lane & 1; tx = 2 * w + txl;
  }
  __device__ __forceinline__ int row(int i) const { return (i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4); }
  __device__ __forceinline__ int col(int j) const { return (j < 4) ? tx * 4 + j : 64 + tx * 4 + (j - 4); }
};

// B tile straight from a row-major matrix M[k][n] (leading dimension ld), via cp.async.
struct BFromMatrix {
  const float* base;   // &M[0][n0]
  int ld;
  __device__ __forceinline__ void post(int, float (*)[kBN]) {}
  __device__ __forceinline__ void pre(int kt, float (*Bs)[kBN]) {
    const int r = threadIdx.x >> 5, c4 = (threadIdx.x & 31) * 4;
    const float* src = base + (size_t)(kt * kBK + r) * ld + c4;
    cp_async16(&Bs[r][c4], src);
    cp_async16(&Bs[r + 8][c4], src + (size_t)8 * ld);
  }
};

template <class AGen, class BGen>
__device__ __forceinline__ void gemm_tile(float (&acc)[8][8], GemmSmem& sm, int KT, AGen& a, BGen& b,
                                          const TileCoord& tc) {
  // accumulators as column pairs: the inner product runs on packed fma.rn.f32x2 (FFMA2: two IEEE fp32 FMAs per issued
  // instruction -- the same arithmetic, in the same order, as 64 scalar FFMAs, at half the issue slots, which is what
  // bounded this loop: the FMA pipe was 69 % busy with the scheduler 79 % busy)
  float2 acc2[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[i][j] = make_float2(0.f, 0.f);

  a.pre(0, sm.As[0]); b.pre(0, sm.Bs[0]);
  a.post(0, sm.As[0]); b.post(0, sm.Bs[0]);
  cp_async_commit();

  for (int kt = 0; kt < KT; ++kt) {
    const int cur = kt & 1;
    cp_async_wait_all();
    __syncthreads();
    const bool more = (kt + 1 < KT);
    if (more) { b.pre(kt + 1, sm.Bs[cur ^ 1]); a.pre(kt + 1, sm.As[cur ^ 1]); cp_async_commit(); }
    float(*As)[kBM] = sm.As[cur];
    float(*Bs)[kBN] = sm.Bs[cur];
#pragma unroll
    for (int kk = 0; kk < kBK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][tc.ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + tc.ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tc.tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tc.tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float2 bp[4] = {make_float2(b0.x, b0.y), make_float2(b0.z, b0.w), make_float2(b1.x, b1.y), make_float2(b1.z, b1.w)};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float2 aa = make_float2(av[i], av[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = __ffma2_rn(aa, bp[j], acc2[i][j]);
      }
    }
    if (more) {
      a.post(kt + 1, sm.As[cur ^ 1]);
      b.post(kt + 1, sm.Bs[cur ^ 1]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][2 * j] = acc2[i][j].x; acc[i][2 * j + 1] = acc2[i][j].y; }
  __syncthreads();   // tiles may be overwritten by the next gemm_tile / epilogue scratch
}

// sum over the 16 ty lanes of a warp (lane bits 1..4): the in-warp COLUMN (over m) reduction
__device__ __forceinline__ float colsum16(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  v += __shfl_xor_sync(0xffffffffu, v, 8);
  v += __shfl_xor_sync(0xffffffffu, v, 16);
  return v;
}

}  // namespace b200vae
