// Grouped GEMM over strided views (statistics update, preconditioner apply):
//   C = alpha * A B^T-view + beta * C_in   -- see pc_gemm_desc in the header.
// Replaces jnp.tensordot at DS:1468-1470 and DS:1707 (fp32, CUDA cores).
#include <algorithm>

#include "simt_gemm.cuh"

namespace pc {

__global__ void __launch_bounds__(kSimtThreads)
grouped_gemm_simt_kernel(const pc_gemm_desc* __restrict__ descs) {
  __shared__ SimtSmem sm;
  const pc_gemm_desc d = descs[blockIdx.z];
  const int tile_m = blockIdx.y, tile_n = blockIdx.x;
  if (tile_m * kSimtBM >= d.m || tile_n * kSimtBN >= d.n) return;
  const OperandView A{d.a, d.a_sio, d.a_si, d.a_sko, d.a_ski, d.a_iinner > 0 ? d.a_iinner : d.m,
                      d.a_kinner, d.m, d.k};
  const OperandView B{d.b, 0, d.b_sj, d.b_sko, d.b_ski, d.n > 0 ? d.n : 1, d.b_kinner, d.n, d.k};
  const float beta = d.beta_dev ? __ldg(d.beta_dev) : d.beta;
  simt_gemm_tile(d.k, tile_m, tile_n, A, B, A.k_fast(), B.k_fast(), sm,
                 [&](int i, int j0, const float* acc) {
                   if (i >= d.m) return;
                   const int io = i / d.c_iinner, ii = i - io * d.c_iinner;
                   const int64_t row = io * d.c_sio + ii * d.c_sii;
#pragma unroll
                   for (int q = 0; q < 4; ++q) {
                     const int j = j0 + q;
                     if (j >= d.n) continue;
                     float v = d.alpha * acc[q];
                     if (d.c_in) v = fmaf(beta, d.c_in[row + j], v);
                     d.c[row + j] = v;
                   }
                 });
}

// Split-K variant for descriptors with a small output and a very long contraction (the
// 9 x 9 statistic of a 3 x 3 x 512 x 512 kernel contracts 262144 elements in ONE tile):
// blockIdx.z = descriptor * splits + split; every split writes alpha * (its partial product)
// into part[z][split][m x n], then splitk_reduce_kernel adds the partials in a fixed order
// (deterministic) and applies beta * C_in.
__global__ void __launch_bounds__(kSimtThreads)
grouped_gemm_splitk_kernel(const pc_gemm_desc* __restrict__ descs, int splits, int max_m,
                           int max_n, float* __restrict__ part) {
  __shared__ SimtSmem sm;
  const int z = blockIdx.z / splits, sp = blockIdx.z - z * splits;
  const pc_gemm_desc d = descs[z];
  const int tile_m = blockIdx.y, tile_n = blockIdx.x;
  if (tile_m * kSimtBM >= d.m || tile_n * kSimtBN >= d.n) return;
  // K range of this split, in whole 32-wide slabs
  const int slabs = (d.k + kSimtBK - 1) / kSimtBK;
  const int per = (slabs + splits - 1) / splits;
  const int k_lo = sp * per * kSimtBK, k_hi = min(d.k, (sp + 1) * per * kSimtBK);
  float* out = part + ((size_t)z * splits + sp) * max_m * max_n;
  const OperandView A{d.a, d.a_sio, d.a_si, d.a_sko, d.a_ski, d.a_iinner > 0 ? d.a_iinner : d.m,
                      d.a_kinner, d.m, k_hi};
  const OperandView B{d.b, 0, d.b_sj, d.b_sko, d.b_ski, d.n > 0 ? d.n : 1, d.b_kinner, d.n, k_hi};
  struct Shift {  // views shifted to the split's first column
    const OperandView& v; int k0;
    __device__ __forceinline__ float operator()(int i, int kk) const { return v(i, kk + k0); }
  };
  const Shift As{A, k_lo}, Bs{B, k_lo};
  simt_gemm_tile(k_hi > k_lo ? k_hi - k_lo : 0, tile_m, tile_n, As, Bs, A.k_fast(), B.k_fast(), sm,
                 [&](int i, int j0, const float* acc) {
                   if (i >= d.m) return;
#pragma unroll
                   for (int q = 0; q < 4; ++q)
                     if (j0 + q < d.n) out[(size_t)i * max_n + j0 + q] = d.alpha * acc[q];
                 });
}

// Loads of the thin kernels.  kSimple (one-level addressing, the common case) is a template
// parameter on purpose: a run-time switch between the two address computations puts every load
// into its own basic block, and the loads of an unrolled loop then complete one after the other
// instead of all being in flight (measured: 4.5 us per loop iteration of 20 loads).
template <bool kSimple>
__device__ __forceinline__ float thin_load(const OperandView& v, int i, int kk) {
  if (!kSimple) return v(i, kk);
  const bool ok = i < v.rows && kk < v.k;
  const float* p = v.base + (int64_t)i * v.s_i + (int64_t)kk * v.s_ki;
  return ok ? __ldg(p) : 0.f;
}

// Thin form of the split-K kernel for outputs of at most 16 x 16 (the 9 x 9 statistic of a
// 3 x 3 convolution kernel): a 64 x 64 tile would compute 50 products to keep one.  Block =
// 64 k-lanes x 4 row groups; a thread keeps rows {ty, ty+4, ty+8, ty+12} x 16 columns of the
// partial product over its k's, the k-lanes are summed by shuffles and a fixed-order pass over
// the 8 warps (deterministic), and the split's partial goes to part[z][split] like the tile
// kernel's, for the same splitk_reduce_kernel.
template <bool kSimple>
__device__ __forceinline__ void thin_splitk_accumulate(const OperandView& A, const OperandView& B,
                                                       int k_lo, int k_hi, int tx, int ty,
                                                       float (&acc)[4][16]) {
  for (int kk = k_lo + tx; kk < k_hi; kk += 64) {
    float a[4], b[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) b[j] = thin_load<kSimple>(B, j, kk);
#pragma unroll
    for (int r = 0; r < 4; ++r) a[r] = thin_load<kSimple>(A, ty + 4 * r, kk);
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[r][j] = fmaf(a[r], b[j], acc[r][j]);
  }
}

__global__ void __launch_bounds__(256)
grouped_gemm_thin_splitk_kernel(const pc_gemm_desc* __restrict__ descs, int splits, int max_m,
                                int max_n, float* __restrict__ part) {
  __shared__ float red[8][64];
  const int z = blockIdx.y, sp = blockIdx.x;
  const pc_gemm_desc d = descs[z];
  const int slabs = (d.k + kSimtBK - 1) / kSimtBK;
  const int per = (slabs + splits - 1) / splits;
  const int k_lo = sp * per * kSimtBK, k_hi = min(d.k, (sp + 1) * per * kSimtBK);
  OperandView A{d.a, d.a_sio, d.a_si, d.a_sko, d.a_ski, d.a_iinner > 0 ? d.a_iinner : d.m,
                d.a_kinner > 0 ? d.a_kinner : d.k, d.m, d.k};
  OperandView B{d.b, 0, d.b_sj, d.b_sko, d.b_ski, d.n > 0 ? d.n : 1,
                d.b_kinner > 0 ? d.b_kinner : d.k, d.n, d.k};
  A.collapse();
  B.collapse();
  A.k = B.k = k_hi;  // loads past this split's range return 0
  const bool simple = A.i_inner >= d.m && A.k_inner >= d.k && B.k_inner >= d.k;
  const int tx = threadIdx.x, ty = threadIdx.y;
  float acc[4][16];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[r][j] = 0.f;
  if (simple) thin_splitk_accumulate<true>(A, B, k_lo, k_hi, tx, ty, acc);
  else thin_splitk_accumulate<false>(A, B, k_lo, k_hi, tx, ty, acc);
  const int lane = tx & 31, warp = ty * 2 + (tx >> 5);
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float v = acc[r][j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[warp][r * 16 + j] = v;
    }
  __syncthreads();
  float* out = part + ((size_t)z * splits + sp) * max_m * max_n;
  const int t = ty * 64 + tx;  // output (i, j) = (rg + 4 r, j), t = rg * 64 + r * 16 + j
  const int rg = t >> 6, r = (t >> 4) & 3, j = t & 15;
  const int i = rg + 4 * r;
  if (i < d.m && j < d.n)
    out[(size_t)i * max_n + j] = d.alpha * (red[rg * 2][r * 16 + j] + red[rg * 2 + 1][r * 16 + j]);
}

// Products with at most 4 rows (a rank-1 parameter times its preconditioner, DS:1707 with a
// [1, n] block): out(i, j) = alpha sum_k A(i, k) B(j, k) + beta C_in.  One CTA per 128 columns;
// the 8 warps split k and are summed in a fixed order.  With B stored j-fast (B(j, k) = P[k][j],
// the optimizer's layout) a lane owns 4 consecutive columns, so a warp reads 512 contiguous
// bytes per k (16-byte loads; unaligned or odd-sized matrices take scalar loads); with B k-fast
// the lanes run along k.  HBM-bound: the matrix is read once.
// kVec: B(j, k) = base[k * s_ki + j] with 16-byte aligned rows.
template <bool kSimple, bool kVec, int kRows>
__device__ __forceinline__ void gemv_jfast(const pc_gemm_desc& d, const OperandView& A,
                                           const OperandView& B, int i0, int j, int warp,
                                           float (&acc)[4][4]) {
  for (int kb = warp; kb < d.k; kb += 64) {
    float4 bv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int kk = kb + 8 * u;
      if (kVec) {
        const bool ok = kk < d.k && j < d.n;
        const float4* p = reinterpret_cast<const float4*>(d.b + (int64_t)kk * d.b_ski + j);
        bv[u] = ok ? __ldg(p) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        bv[u] = make_float4(thin_load<kSimple>(B, j, kk), thin_load<kSimple>(B, j + 1, kk),
                            thin_load<kSimple>(B, j + 2, kk), thin_load<kSimple>(B, j + 3, kk));
      }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float av = thin_load<kSimple>(A, i0 + r, kb + 8 * u);
        acc[r][0] = fmaf(av, bv[u].x, acc[r][0]);
        acc[r][1] = fmaf(av, bv[u].y, acc[r][1]);
        acc[r][2] = fmaf(av, bv[u].z, acc[r][2]);
        acc[r][3] = fmaf(av, bv[u].w, acc[r][3]);
      }
    }
  }
}

template <bool kSimple>
__device__ __forceinline__ void gemv_kfast(const pc_gemm_desc& d, const OperandView& A,
                                           const OperandView& B, int i0, int j, int lane,
                                           float (&s4)[4]) {
  for (int kk = lane; kk < d.k; kk += 32) {
    const float bv = thin_load<kSimple>(B, j, kk);
#pragma unroll
    for (int r = 0; r < 4; ++r) s4[r] = fmaf(thin_load<kSimple>(A, i0 + r, kk), bv, s4[r]);
  }
}

__global__ void __launch_bounds__(256, 2)
grouped_gemv_kernel(const pc_gemm_desc* __restrict__ descs) {
  __shared__ float red[8][4][128];
  const pc_gemm_desc d = descs[blockIdx.y];
  const int j0 = blockIdx.x * 128;
  if (j0 >= d.n) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  OperandView A{d.a, d.a_sio, d.a_si, d.a_sko, d.a_ski, d.a_iinner > 0 ? d.a_iinner : d.m,
                d.a_kinner > 0 ? d.a_kinner : d.k, d.m, d.k};
  OperandView B{d.b, 0, d.b_sj, d.b_sko, d.b_ski, d.n > 0 ? d.n : 1,
                d.b_kinner > 0 ? d.b_kinner : d.k, d.n, d.k};
  A.collapse();
  B.collapse();
  const bool simple = A.i_inner >= d.m && A.k_inner >= d.k && B.k_inner >= d.k;
  const float beta = d.beta_dev ? __ldg(d.beta_dev) : d.beta;
  const int c_iinner = d.c_iinner > 0 ? d.c_iinner : d.m;
  const bool vec = simple && d.b_sj == 1 && (d.b_ski & 3) == 0 && (d.n & 3) == 0 &&
                   (reinterpret_cast<uintptr_t>(d.b) & 15) == 0;
  for (int i0 = 0; i0 < d.m; i0 += 4) {
    const int mrows = min(4, d.m - i0);
    if (d.b_ski != 1 || d.k == 1) {
      // lanes = columns j0 + 4 lane .. + 3, warp w takes k = w, w + 8, ... (8 rows in flight)
      const int j = j0 + 4 * lane;
      float acc[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
      if (vec && mrows == 1) gemv_jfast<true, true, 1>(d, A, B, i0, j, warp, acc);
      else if (vec) gemv_jfast<true, true, 4>(d, A, B, i0, j, warp, acc);
      else if (simple) gemv_jfast<true, false, 4>(d, A, B, i0, j, warp, acc);
      else gemv_jfast<false, false, 4>(d, A, B, i0, j, warp, acc);
#pragma unroll
      for (int r = 0; r < 4; ++r)
        *reinterpret_cast<float4*>(&red[warp][r][4 * lane]) =
            make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
    } else {
      // B k-fast: warp w takes columns j0 + w, j0 + w + 8, ...; lanes run along k
      for (int q = 0; q < 16; ++q) {
        const int jl = warp + 8 * q, j = j0 + jl;
        float s4[4] = {0.f, 0.f, 0.f, 0.f};
        if (j < d.n) {
          if (simple) gemv_kfast<true>(d, A, B, i0, j, lane, s4);
          else gemv_kfast<false>(d, A, B, i0, j, lane, s4);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) s4[r] += __shfl_xor_sync(0xffffffffu, s4[r], o);
          // slot (0, r, column); the other warps' slots hold zeros for the final pass below
          if (lane == 0) red[0][r][jl] = s4[r];
        }
      }
      if (warp > 0)
#pragma unroll
        for (int r = 0; r < 4; ++r)
          *reinterpret_cast<float4*>(&red[warp][r][4 * lane]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 4 * 128; e += 256) {
      const int r = e >> 7, jl = e & 127, i = i0 + r, j = j0 + jl;
      if (r < mrows && j < d.n) {
        float v = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) v += red[w][r][jl];
        v *= d.alpha;
        const int io = i / c_iinner, ii = i - io * c_iinner;
        const int64_t row = io * d.c_sio + ii * d.c_sii;
        if (d.c_in) v = fmaf(beta, d.c_in[row + j], v);
        d.c[row + j] = v;
      }
    }
    __syncthreads();
  }
}

// Products with a contraction of at most 4 (the statistic of a rank-1 parameter, DS:1468-1470 on
// a [n] block: S <- w1 S + w2 g g^T): out(i, j) = alpha sum_k A(i, k) B(j, k) + beta C_in is a
// streaming pass over the output.  One CTA per 16 rows; a thread owns 4 consecutive columns
// (16-byte C_in loads / C stores when rows are aligned), B(j, :) stays in registers over the rows.
// With A == B and a bitwise symmetric C_in the result is bitwise symmetric (a * b == b * a).
template <bool kVec>
__device__ __forceinline__ void outer_rows(const pc_gemm_desc& d, const OperandView& A,
                                           const OperandView& B, int r0, float beta,
                                           float (&as)[16][4]) {
  const int c_iinner = d.c_iinner > 0 ? d.c_iinner : d.m;
  // (a contraction longer than 4 is walked in chunks that accumulate into C: correct, not what
  // the kernel is for)
  for (int kc = 0; kc < d.k; kc += 4) {
    __syncthreads();
    if (threadIdx.x < 64)  // A(r0 .. r0+15, kc .. kc+3) of this CTA's rows (0 outside the view)
      as[threadIdx.x >> 2][threadIdx.x & 3] = A(r0 + (threadIdx.x >> 2), kc + (threadIdx.x & 3));
    __syncthreads();
    for (int j = 4 * threadIdx.x; j < d.n; j += 4 * blockDim.x) {
      float b[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) b[q][kk] = B(j + q, kc + kk);  // (0 outside the view)
      const float* cin = kc == 0 ? d.c_in : d.c;
      const float bw = kc == 0 ? beta : 1.0f;
      const int rend = min(r0 + 16, d.m);
      for (int i0 = r0; i0 < rend; i0 += 4) {  // four rows at a time: their C_in loads in flight
        float4 o[4];
        int64_t row[4];
        bool ok[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const int i = i0 + r;
          ok[r] = i < rend;
          const int io = i / c_iinner, ii = i - io * c_iinner;
          row[r] = io * d.c_sio + ii * d.c_sii;
          o[r] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (kVec) {
            const float4* p = reinterpret_cast<const float4*>(cin + row[r] + j);
            if (cin && ok[r]) o[r] = *p;
          } else if (cin && ok[r]) {
            o[r].x = cin[row[r] + j];
            if (j + 1 < d.n) o[r].y = cin[row[r] + j + 1];
            if (j + 2 < d.n) o[r].z = cin[row[r] + j + 2];
            if (j + 3 < d.n) o[r].w = cin[row[r] + j + 3];
          }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          float a[4];
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) a[kk] = as[(i0 + r - r0) & 15][kk];
          float v[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float acc = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) acc = fmaf(a[kk], b[q][kk], acc);
            v[q] = d.alpha * acc;
          }
          if (cin) {
            v[0] = fmaf(bw, o[r].x, v[0]); v[1] = fmaf(bw, o[r].y, v[1]);
            v[2] = fmaf(bw, o[r].z, v[2]); v[3] = fmaf(bw, o[r].w, v[3]);
          }
          if (!ok[r]) continue;
          if (kVec) {
            *reinterpret_cast<float4*>(d.c + row[r] + j) = make_float4(v[0], v[1], v[2], v[3]);
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (j + q < d.n) d.c[row[r] + j + q] = v[q];
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256, 3)
grouped_outer_kernel(const pc_gemm_desc* __restrict__ descs) {
  __shared__ float as[16][4];
  const pc_gemm_desc d = descs[blockIdx.y];
  const int r0 = blockIdx.x * 16;
  if (r0 >= d.m) return;
  OperandView A{d.a, d.a_sio, d.a_si, d.a_sko, d.a_ski, d.a_iinner > 0 ? d.a_iinner : d.m,
                d.a_kinner > 0 ? d.a_kinner : d.k, d.m, d.k};
  OperandView B{d.b, 0, d.b_sj, d.b_sko, d.b_ski, d.n > 0 ? d.n : 1,
                d.b_kinner > 0 ? d.b_kinner : d.k, d.n, d.k};
  A.collapse();
  B.collapse();
  const float beta = d.beta_dev ? __ldg(d.beta_dev) : d.beta;
  const bool vec = (d.n & 3) == 0 && (d.c_sii & 3) == 0 && (d.c_sio & 3) == 0 &&
                   (reinterpret_cast<uintptr_t>(d.c) & 15) == 0 &&
                   (!d.c_in || (reinterpret_cast<uintptr_t>(d.c_in) & 15) == 0);
  if (vec) outer_rows<true>(d, A, B, r0, beta, as);
  else outer_rows<false>(d, A, B, r0, beta, as);
}

// Products with a tiny contraction AND a tiny output width (n, k <= 16) over very many rows:
// the mode product of a [9, 512, 512] convolution kernel with its 9 x 9 preconditioner is a
// [262144, 9] x [9, 9] product -- a streaming pass, one thread per row.  Larger n / k are
// walked in chunks of 16 (correct, not what the kernel is for).
template <bool kSimple>
__device__ __forceinline__ void rowmap_accumulate(const OperandView& A, int i, int kc,
                                                  const float (&bs)[16][17], float (&acc)[16]) {
  float a[16];
#pragma unroll
  for (int q = 0; q < 16; ++q) a[q] = thin_load<kSimple>(A, i, kc + q);
#pragma unroll
  for (int j = 0; j < 16; ++j)
#pragma unroll
    for (int q = 0; q < 16; ++q) acc[j] = fmaf(a[q], bs[j][q], acc[j]);
}

__global__ void __launch_bounds__(256)
grouped_gemm_rowmap_kernel(const pc_gemm_desc* __restrict__ descs) {
  __shared__ float bs[16][17];
  __shared__ float outs[256 * 16];
  const pc_gemm_desc d = descs[blockIdx.y];
  OperandView A{d.a, d.a_sio, d.a_si, d.a_sko, d.a_ski, d.a_iinner > 0 ? d.a_iinner : d.m,
                d.a_kinner > 0 ? d.a_kinner : d.k, d.m, d.k};
  const OperandView B{d.b, 0, d.b_sj, d.b_sko, d.b_ski, d.n > 0 ? d.n : 1,
                      d.b_kinner > 0 ? d.b_kinner : d.k, d.n, d.k};
  A.collapse();
  const bool a_simple = A.i_inner >= d.m && A.k_inner >= d.k;
  const float beta = d.beta_dev ? __ldg(d.beta_dev) : d.beta;
  const int c_iinner = d.c_iinner > 0 ? d.c_iinner : d.m;
  // rows of the result back to back in memory (the mode product appends its axis last): the
  // CTA's 256 x n values leave through shared memory as one contiguous, coalesced run
  const bool packed_rows = d.n <= 16 && d.c_sii == d.n &&
                           (c_iinner >= d.m || d.c_sio == (int64_t)c_iinner * d.c_sii);
  for (int base = blockIdx.x * 256; base < d.m; base += gridDim.x * 256) {
    const int i = base + threadIdx.x;
    const int io = i / c_iinner, ii = i - io * c_iinner;
    const int64_t row = io * d.c_sio + ii * d.c_sii;
    for (int jc = 0; jc < d.n; jc += 16) {
      float acc[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] = 0.f;
      for (int kc = 0; kc < d.k; kc += 16) {
        __syncthreads();
        bs[threadIdx.x >> 4][threadIdx.x & 15] = B(jc + (threadIdx.x >> 4), kc + (threadIdx.x & 15));
        __syncthreads();
        if (a_simple) rowmap_accumulate<true>(A, i, kc, bs, acc);
        else rowmap_accumulate<false>(A, i, kc, bs, acc);
      }
      if (packed_rows) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j < d.n) outs[threadIdx.x * d.n + j] = d.alpha * acc[j];
        __syncthreads();
        const int64_t first = (int64_t)base * d.n;  // offset of row `base` (rows are back to back)
        const int count = min(256, d.m - base) * d.n;
        for (int e = threadIdx.x; e < count; e += 256) {
          float v = outs[e];
          if (d.c_in) v = fmaf(beta, d.c_in[first + e], v);
          d.c[first + e] = v;
        }
      } else if (i < d.m) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (jc + j >= d.n) break;
          float v = d.alpha * acc[j];
          if (d.c_in) v = fmaf(beta, d.c_in[row + jc + j], v);
          d.c[row + jc + j] = v;
        }
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const pc_gemm_desc* __restrict__ descs, int splits, int max_m,
                                     int max_n, const float* __restrict__ part) {
  const pc_gemm_desc d = descs[blockIdx.y];
  const float* p0 = part + (size_t)blockIdx.y * splits * max_m * max_n;
  const float beta = d.beta_dev ? __ldg(d.beta_dev) : d.beta;
  // one warp per output element: lane l adds splits l, l + 32, ... in order, then a fixed
  // shuffle tree (deterministic; the partial loads of a warp are all in flight at once)
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; e < d.m * d.n; e += warps) {
    const int i = e / d.n, j = e - i * d.n;
    float v = 0.f;
    for (int sp = lane; sp < splits; sp += 32) v += p0[((size_t)sp * max_m + i) * max_n + j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) {
      const int io = i / d.c_iinner, ii = i - io * d.c_iinner;
      const int64_t row = io * d.c_sio + ii * d.c_sii;
      if (d.c_in) v = fmaf(beta, d.c_in[row + j], v);
      d.c[row + j] = v;
    }
  }
}

}  // namespace pc

namespace pc {
__global__ void select_copy_kernel(const float* __restrict__ src, const float* __restrict__ metrics,
                                   float threshold, float* __restrict__ dst, int src_cols,
                                   size_t src_elems, int rows, int cols, int batch) {
  for (int b = blockIdx.y; b < batch; b += gridDim.y) {
    const float err = metrics[(size_t)b * PC_NUM_METRICS + PC_METRIC_ERROR];
    if (isnan(err) || err >= threshold) continue;  // keep the old preconditioner, DS:2936-2943
    const size_t total = (size_t)rows * cols;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (size_t)gridDim.x * blockDim.x) {
      const int r = (int)(e / cols), c = (int)(e - (size_t)r * cols);
      dst[(size_t)b * total + e] = src[(size_t)b * src_elems + (size_t)r * src_cols + c];
    }
  }
}
}  // namespace pc

extern "C" int pc_select_preconditioners(const float* src, const float* metrics, float threshold,
                                         float* dst, int batch, int src_rows, int src_cols,
                                         int rows, int cols, void* stream) {
  PC_REQUIRE(batch >= 0 && rows >= 0 && cols >= 0 && rows <= src_rows && cols <= src_cols,
             "bad select sizes");
  if (batch == 0 || rows == 0 || cols == 0) return PC_OK;
  PC_REQUIRE(src && metrics && dst, "null pointer argument");
  const size_t total = (size_t)rows * cols;
  dim3 grid((unsigned)((total + 255) / 256 < 256 ? (total + 255) / 256 : 256),
            (unsigned)(batch < 65535 ? batch : 65535));
  pc::select_copy_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      src, metrics, threshold, dst, src_cols, (size_t)src_rows * src_cols, rows, cols, batch);
  pc::count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

namespace pc {
// Row j of a gathered buffer -> row dst_idx[j] of the state, unless its metrics row reports a
// failed root (DS:2936-2950).  Rows are raw bytes so that fp32 roots and the int16 / int8 +
// diagonal + bucket triples of quantised preconditioners share the kernel.
__global__ void select_scatter_kernel(const uint8_t* __restrict__ src,
                                      const int64_t* __restrict__ src_off,
                                      const float* __restrict__ metrics_base,
                                      const int64_t* __restrict__ met_off,
                                      const int32_t* __restrict__ dst_idx, float threshold,
                                      uint8_t* __restrict__ dst, int64_t row_bytes,
                                      float* __restrict__ metrics_dst, int count) {
  for (int j = blockIdx.y; j < count; j += gridDim.y) {
    const int di = dst_idx[j];
    if (di < 0) continue;  // filler row (DS:2844-2850)
    const float* m = metrics_base + met_off[j];
    if (metrics_dst && blockIdx.x == 0 && threadIdx.x < PC_NUM_METRICS)
      metrics_dst[(size_t)di * PC_NUM_METRICS + threadIdx.x] = m[threadIdx.x];
    const float err = m[PC_METRIC_ERROR];
    if (isnan(err) || err >= threshold) continue;  // keep the old preconditioner
    const uint8_t* s = src + src_off[j];
    uint8_t* d = dst + (size_t)di * row_bytes;
    if ((row_bytes & 15) == 0 && ((reinterpret_cast<uintptr_t>(s) | reinterpret_cast<uintptr_t>(d)) & 15) == 0) {
      const int64_t n16 = row_bytes >> 4;
      for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n16;
           e += (int64_t)gridDim.x * blockDim.x)
        reinterpret_cast<uint4*>(d)[e] = reinterpret_cast<const uint4*>(s)[e];
    } else {
      for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < row_bytes;
           e += (int64_t)gridDim.x * blockDim.x)
        d[e] = s[e];
    }
  }
}
}  // namespace pc

extern "C" int pc_select_scatter(const void* src, const int64_t* src_offset_bytes,
                                 const float* metrics_base, const int64_t* metrics_offset,
                                 const int32_t* dst_index, float threshold, void* dst,
                                 int64_t row_bytes, float* metrics_dst, int count, void* stream) {
  PC_REQUIRE(count >= 0 && row_bytes >= 0, "bad select sizes");
  if (count == 0) return PC_OK;
  PC_REQUIRE(src && src_offset_bytes && metrics_base && metrics_offset && dst_index && dst,
             "null pointer argument");
  const int64_t units = (row_bytes + 15) / 16;
  const unsigned gx = (unsigned)std::max<int64_t>(1, std::min<int64_t>((units + 255) / 256, 64));
  dim3 grid(gx, (unsigned)std::min(count, 65535));
  pc::select_scatter_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const uint8_t*)src, src_offset_bytes, metrics_base, metrics_offset, dst_index, threshold,
      (uint8_t*)dst, row_bytes, metrics_dst, count);
  pc::count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

extern "C" size_t pc_grouped_gemm_splitk_workspace_bytes(int count, int max_m, int max_n,
                                                        int splits) {
  if (count <= 0 || max_m <= 0 || max_n <= 0 || splits <= 0) return 0;
  return (size_t)count * splits * max_m * max_n * sizeof(float) + 256;
}

extern "C" int pc_grouped_gemm_splitk(const pc_gemm_desc* descs, int count, int max_m, int max_n,
                                      int splits, void* workspace, size_t workspace_bytes,
                                      void* stream) {
  PC_REQUIRE(count >= 0 && max_m >= 0 && max_n >= 0 && splits >= 1, "bad split-K sizes");
  if (count == 0 || max_m == 0 || max_n == 0) return PC_OK;
  PC_REQUIRE(descs && workspace, "null pointer argument");
  PC_REQUIRE((long long)count * splits <= 65535, "count * splits must be <= 65535");
  PC_REQUIRE(workspace_bytes >= pc_grouped_gemm_splitk_workspace_bytes(count, max_m, max_n, splits),
             "split-K workspace too small");
  float* part = reinterpret_cast<float*>(pc::align_up((size_t)workspace, 256));
  const int tm = (max_m + pc::kSimtBM - 1) / pc::kSimtBM, tn = (max_n + pc::kSimtBN - 1) / pc::kSimtBN;
  if (max_m <= 16 && max_n <= 16) {  // thin outputs: no 64 x 64 tile around a 9 x 9 result
    pc::grouped_gemm_thin_splitk_kernel<<<dim3(splits, count), dim3(64, 4), 0,
                                          (cudaStream_t)stream>>>(descs, splits, max_m, max_n,
                                                                  part);
  } else {
    dim3 grid(tn, tm, count * splits);
    pc::grouped_gemm_splitk_kernel<<<grid, pc::kSimtThreads, 0, (cudaStream_t)stream>>>(
        descs, splits, max_m, max_n, part);
  }
  const int rb = (max_m * max_n + 7) / 8;  // 8 warps per block, one output element per warp
  pc::splitk_reduce_kernel<<<dim3(rb < 512 ? rb : 512, count), 256, 0, (cudaStream_t)stream>>>(
      descs, splits, max_m, max_n, part);
  pc::count_launch(2);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_grouped_gemm_thin(const pc_gemm_desc* descs, int count, int max_m, int max_n,
                                    int kind, void* stream) {
  PC_REQUIRE(count >= 0 && max_m >= 0 && max_n >= 0, "bad grouped gemm sizes");
  PC_REQUIRE(kind == PC_THIN_GEMV || kind == PC_THIN_ROWMAP || kind == PC_THIN_OUTER,
             "unknown thin product kind %d", kind);
  if (count == 0 || max_m == 0 || max_n == 0) return PC_OK;
  PC_REQUIRE(descs != nullptr, "null descriptor array");
  for (int z0 = 0; z0 < count; z0 += 65535) {
    const int nz = count - z0 < 65535 ? count - z0 : 65535;
    if (kind == PC_THIN_GEMV) {
      pc::grouped_gemv_kernel<<<dim3((max_n + 127) / 128, nz), 256, 0, (cudaStream_t)stream>>>(
          descs + z0);
    } else if (kind == PC_THIN_OUTER) {
      pc::grouped_outer_kernel<<<dim3((max_m + 15) / 16, nz), 256, 0, (cudaStream_t)stream>>>(
          descs + z0);
    } else {
      const int bx = std::max(1, std::min((max_m + 255) / 256, 2048));
      pc::grouped_gemm_rowmap_kernel<<<dim3(bx, nz), 256, 0, (cudaStream_t)stream>>>(descs + z0);
    }
  }
  pc::count_launch((count + 65534) / 65535);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

extern "C" int pc_grouped_gemm(const pc_gemm_desc* descs, int count, int max_m, int max_n,
                               void* stream) {
  PC_REQUIRE(count >= 0 && max_m >= 0 && max_n >= 0, "bad grouped gemm sizes");
  if (count == 0 || max_m == 0 || max_n == 0) return PC_OK;
  PC_REQUIRE(descs != nullptr, "null descriptor array");
  const int tm = (max_m + pc::kSimtBM - 1) / pc::kSimtBM;
  const int tn = (max_n + pc::kSimtBN - 1) / pc::kSimtBN;
  for (int z0 = 0; z0 < count; z0 += 65535) {
    const int nz = count - z0 < 65535 ? count - z0 : 65535;
    dim3 grid(tn, tm, nz);
    pc::grouped_gemm_simt_kernel<<<grid, pc::kSimtThreads, 0, (cudaStream_t)stream>>>(descs + z0);
  }
  pc::count_launch((count + 65534) / 65535);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}
