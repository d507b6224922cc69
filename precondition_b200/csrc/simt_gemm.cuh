// CUDA-core fp32 GEMM tile (exact fp32 FFMA semantics).  Used for blocks that are
// too small / oddly shaped for the tcgen05 engine and as the on-device cross
// check of the split-precision tensor-core path.
//
//   acc(i,j) = sum_k A(i,k) * B(j,k)        i in [0,M), j in [0,N)
//
// Operands are supplied through accessor functors so that strided views of
// gradient blocks (two-level K addressing) and the Newton-chain buffers share
// one mainloop.  64x64 output tile, BK = 32, 256 threads, 4x4 micro-tile.
#pragma once
#include "common.cuh"

namespace pc {

constexpr int kSimtBM = 64, kSimtBN = 64, kSimtBK = 32, kSimtThreads = 256;
constexpr int kSimtLd = kSimtBM + 4;  // row stride of the k-major smem tiles

struct SimtSmem {
  float a[kSimtBK][kSimtLd];
  float b[kSimtBK][kSimtLd];
};

// View with two-level K addressing:
//   X(i,k) = base[i*s_i + (k / k_inner)*s_ko + (k % k_inner)*s_ki]
struct OperandView {
  const float* base;
  int64_t s_io, s_i, s_ko, s_ki;
  int i_inner, k_inner;
  int rows, k;  // logical extents (loads outside return 0)
  __device__ __forceinline__ float operator()(int i, int kk) const {
    if (i >= rows || kk >= k) return 0.f;
    const int io = i / i_inner, ii = i - io * i_inner;
    const int ko = kk / k_inner, ki = kk - ko * k_inner;
    return __ldg(base + io * s_io + ii * s_i + ko * s_ko + ki * s_ki);
  }
  __device__ __forceinline__ bool k_fast() const { return s_ki == 1; }
  // Two-level addressing whose outer stride continues the inner one (an unpartitioned
  // contiguous tensor described axis by axis) is one-level addressing: drop the div / mod.
  __device__ __forceinline__ void collapse() {
    if (k_inner < k && s_ko == (int64_t)k_inner * s_ki) k_inner = 0x7fffffff;
    if (i_inner < rows && s_io == (int64_t)i_inner * s_i) i_inner = 0x7fffffff;
  }
};

// Masked square view used by the Newton chain: rows/cols >= limit read as 0.
template <class ALoad, class BLoad, class Epilogue>
__device__ __forceinline__ void simt_gemm_tile(int K, int tile_m, int tile_n,
                                               const ALoad& A, const BLoad& B,
                                               bool a_kfast, bool b_kfast,
                                               SimtSmem& sm, Epilogue epi) {
  const int t = threadIdx.x;
  const int tx = t & 15, ty = t >> 4;
  const int i0 = tile_m * kSimtBM, j0 = tile_n * kSimtBN;
  float acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

  float ra[8], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int ia, ka, jb, kb;
      if (a_kfast) { ia = (t >> 5) + 8 * r; ka = t & 31; }
      else         { ia = t & 63;           ka = (t >> 6) + 4 * r; }
      if (b_kfast) { jb = (t >> 5) + 8 * r; kb = t & 31; }
      else         { jb = t & 63;           kb = (t >> 6) + 4 * r; }
      ra[r] = A(i0 + ia, k0 + ka);
      rb[r] = B(j0 + jb, k0 + kb);
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int ia, ka, jb, kb;
      if (a_kfast) { ia = (t >> 5) + 8 * r; ka = t & 31; }
      else         { ia = t & 63;           ka = (t >> 6) + 4 * r; }
      if (b_kfast) { jb = (t >> 5) + 8 * r; kb = t & 31; }
      else         { jb = t & 63;           kb = (t >> 6) + 4 * r; }
      sm.a[ka][ia] = ra[r];
      sm.b[kb][jb] = rb[r];
    }
  };

  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kSimtBK) {
    __syncthreads();  // previous tile fully consumed
    stash();
    __syncthreads();
    if (k0 + kSimtBK < K) fetch(k0 + kSimtBK);  // prefetch into registers
#pragma unroll
    for (int kk = 0; kk < kSimtBK; ++kk) {
      const float4 av = *reinterpret_cast<const float4*>(&sm.a[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&sm.b[kk][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w};
      const float b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(a4[r], b4[c], acc[r][c]);
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) epi(i0 + ty * 4 + r, j0 + tx * 4, acc[r]);
}

}  // namespace pc
