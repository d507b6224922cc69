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

__global__ void splitk_reduce_kernel(const pc_gemm_desc* __restrict__ descs, int splits, int max_m,
                                     int max_n, const float* __restrict__ part) {
  const pc_gemm_desc d = descs[blockIdx.y];
  const float* p0 = part + (size_t)blockIdx.y * splits * max_m * max_n;
  const float beta = d.beta_dev ? __ldg(d.beta_dev) : d.beta;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < d.m * d.n; e += gridDim.x * blockDim.x) {
    const int i = e / d.n, j = e - i * d.n;
    float v = 0.f;
    for (int sp = 0; sp < splits; ++sp) v += p0[((size_t)sp * max_m + i) * max_n + j];
    const int io = i / d.c_iinner, ii = i - io * d.c_iinner;
    const int64_t row = io * d.c_sio + ii * d.c_sii;
    if (d.c_in) v = fmaf(beta, d.c_in[row + j], v);
    d.c[row + j] = v;
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
  dim3 grid(tn, tm, count * splits);
  pc::grouped_gemm_splitk_kernel<<<grid, pc::kSimtThreads, 0, (cudaStream_t)stream>>>(
      descs, splits, max_m, max_n, part);
  const int rb = (max_m * max_n + 255) / 256;
  pc::splitk_reduce_kernel<<<dim3(rb < 64 ? rb : 64, count), 256, 0, (cudaStream_t)stream>>>(
      descs, splits, max_m, max_n, part);
  pc::count_launch(2);
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
