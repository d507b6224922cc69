// QuantizedValue (QU:49-113) on device: per-column symmetric bucket quantisation
// to int16 / int8 with optional diagonal extraction, and bf16 casts.
#include <cuda_bf16.h>

#include <algorithm>
#include "common.cuh"

namespace pc {

constexpr int kMaxGridY = 65535;

template <typename Q>
__device__ __forceinline__ Q to_q(float r);
template <>
__device__ __forceinline__ int16_t to_q<int16_t>(float r) { return (int16_t)r; }
template <>
__device__ __forceinline__ int8_t to_q<int8_t>(float r) { return (int8_t)r; }

// grid (ceil(cols/32), batch); block (32, 8): each warp-row strides over rows so
// that global reads are coalesced along the column index.
template <typename Q>
__global__ void __launch_bounds__(256)
quantize_kernel(const float* __restrict__ x, int rows, int cols, float num_buckets,
                int extract_diagonal, Q* __restrict__ q, float* __restrict__ diag,
                float* __restrict__ bucket) {
  __shared__ uint32_t smax[8][32];
  const int b = blockIdx.y;
  const int col = blockIdx.x * 32 + threadIdx.x;
  const float* xb = x + (size_t)b * rows * cols;
  Q* qb = q + (size_t)b * rows * cols;
  uint32_t m = 0;
  if (col < cols) {
    for (int r = threadIdx.y; r < rows; r += 8) {
      float v = xb[(size_t)r * cols + col];
      if (extract_diagonal && r == col) {  // QU:72-76
        diag[(size_t)b * rows + r] = v;
        v = v - v;                        // fvalue - diag(fvalue): NaN/inf stay non-finite
      }
      const uint32_t ab = absbits(v);
      m = ab > m ? ab : m;
    }
  }
  smax[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int k = 1; k < 8; ++k) m = smax[k][threadIdx.x] > m ? smax[k][threadIdx.x] : m;
    smax[0][threadIdx.x] = m;
  }
  __syncthreads();
  if (col >= cols) return;
  const float max_abs = __uint_as_float(smax[0][threadIdx.x]);  // QU:86
  const float bs = max_abs / num_buckets;                       // QU:87
  const float bs_nz = bs > 0.f ? bs : 1.f;                      // QU:90-91
  if (threadIdx.y == 0) bucket[(size_t)b * cols + col] = bs;
  for (int r = threadIdx.y; r < rows; r += 8) {
    float v = xb[(size_t)r * cols + col];
    if (extract_diagonal && r == col) v = v - v;
    qb[(size_t)r * cols + col] = to_q<Q>(rintf(v / bs_nz));     // QU:92-95
  }
}

// Requantisation when the column maxima are already known (they were reduced in the
// epilogue of the kernel that produced x): one pass, same arithmetic as quantize_kernel.
template <typename Q>
__global__ void quantize_from_colmax_kernel(const float* __restrict__ x,
                                            const uint32_t* __restrict__ colmax, int n,
                                            float num_buckets, Q* __restrict__ q,
                                            float* __restrict__ diag, float* __restrict__ bucket) {
  const int b = blockIdx.y;
  const size_t total = (size_t)n * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / n), c = (int)(e - (size_t)r * n);
    float v = x[(size_t)b * total + e];
    const float bs = __uint_as_float(colmax[(size_t)b * n + c]) / num_buckets;  // QU:86-87
    const float bs_nz = bs > 0.f ? bs : 1.f;                                    // QU:90-91
    if (r == c) {
      diag[(size_t)b * n + r] = v;  // QU:72-76
      v = v - v;
      bucket[(size_t)b * n + c] = bs;
    }
    q[(size_t)b * total + e] = to_q<Q>(rintf(v / bs_nz));                       // QU:92-95
  }
}

template <typename Q>
__global__ void dequantize_kernel(const Q* __restrict__ q, const float* __restrict__ diag,
                                  const float* __restrict__ bucket, int rows, int cols,
                                  int extract_diagonal, float* __restrict__ x) {
  const int b = blockIdx.y;
  const size_t total = (size_t)rows * cols;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / cols), c = (int)(e - (size_t)r * cols);
    float v = (float)q[(size_t)b * total + e] * bucket[(size_t)b * cols + c];  // QU:110
    if (extract_diagonal && r == c) v += diag[(size_t)b * rows + r];           // QU:111-112
    x[(size_t)b * total + e] = v;
  }
}

// Two-pass form for tall matrices (momenta of shape [d0, rest], large statistics): the column
// kernel above walks a whole column per thread, i.e. cols / 32 CTAs of 1024+ dependent loads.
// Pass 1 reduces max |x| per column over row chunks (atomicMax on the float bits, all values >= 0),
// pass 2 quantises element-wise over the whole grid.  Same arithmetic, bit-identical output.
constexpr int kQRowChunk = 128;
__global__ void __launch_bounds__(256)
quant_colmax_kernel(const float* __restrict__ x, int rows, int cols, int extract_diagonal,
                    uint32_t* __restrict__ colmax) {
  __shared__ uint32_t smax[8][32];
  const int b = blockIdx.z;
  const int col = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * kQRowChunk, r1 = min(rows, r0 + kQRowChunk);
  const float* xb = x + (size_t)b * rows * cols;
  uint32_t m = 0;
  if (col < cols) {
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      float v = xb[(size_t)r * cols + col];
      if (extract_diagonal && r == col) v = v - v;  // NaN / inf stay non-finite, QU:72-76
      const uint32_t ab = absbits(v);
      m = ab > m ? ab : m;
    }
  }
  smax[threadIdx.y][threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.y == 0 && col < cols) {
    for (int k = 1; k < 8; ++k) m = smax[k][threadIdx.x] > m ? smax[k][threadIdx.x] : m;
    if (m) atomicMax(colmax + (size_t)b * cols + col, m);
  }
}
template <typename Q>
__global__ void __launch_bounds__(256)
quant_apply_kernel(const float* __restrict__ x, const uint32_t* __restrict__ colmax, int rows,
                   int cols, float num_buckets, int extract_diagonal, Q* __restrict__ q,
                   float* __restrict__ diag, float* __restrict__ bucket) {
  const int b = blockIdx.y;
  const size_t total = (size_t)rows * cols;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(e / cols), c = (int)(e - (size_t)r * cols);
    float v = x[(size_t)b * total + e];
    const float bs = __uint_as_float(colmax[(size_t)b * cols + c]) / num_buckets;  // QU:86-87
    const float bs_nz = bs > 0.f ? bs : 1.f;                                       // QU:90-91
    if (r == 0) bucket[(size_t)b * cols + c] = bs;
    if (extract_diagonal && r == c) {
      diag[(size_t)b * rows + r] = v;  // QU:72-76
      v = v - v;
    }
    q[(size_t)b * total + e] = to_q<Q>(rintf(v / bs_nz));                          // QU:92-95
  }
}

__global__ void to_bf16_kernel(const float* __restrict__ x, size_t total,
                               __nv_bfloat16* __restrict__ q) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x)
    q[e] = __float2bfloat16_rn(x[e]);
}
__global__ void from_bf16_kernel(const __nv_bfloat16* __restrict__ q, size_t total,
                                 float* __restrict__ x) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x)
    x[e] = __bfloat162float(q[e]);
}

// ---------------------------------------------------------------------------
// Grouped (de)quantisation of the int8 momenta of a whole model (DS:3582-3586, DS:3620-3621):
// the reference maps to_float / from_float over the parameter tree; here every momentum is a
// segment {int8 data, bucket sizes [cols], fp32 view in a flat buffer, rows, cols} and three
// launches serve all of them.  Work items = chunks of kQgChunk consecutive elements.
// ---------------------------------------------------------------------------
constexpr int kQgThreads = 256;
constexpr int kQgChunk = kQgThreads * 32;

__global__ void __launch_bounds__(kQgThreads)
qgroup_dequantize_kernel(const pc_quant_segment* __restrict__ segs,
                         const int32_t* __restrict__ chunk_seg, int total_chunks) {
  for (int c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    const pc_quant_segment sg = segs[chunk_seg[c]];
    const int64_t numel = (int64_t)sg.rows * sg.cols;
    const int64_t begin = (int64_t)(c - sg.first_chunk) * kQgChunk;
    const int64_t end = begin + kQgChunk < numel ? begin + kQgChunk : numel;
    const int8_t* q = reinterpret_cast<const int8_t*>(sg.q);
    for (int64_t e = begin + threadIdx.x; e < end; e += kQgThreads)
      sg.x[e] = (float)q[e] * sg.bucket[e % sg.cols];  // QU:107-108
  }
}

// per-column max |x| (bit patterns, atomicMax) into sg.colmax (zero on entry)
__global__ void __launch_bounds__(kQgThreads)
qgroup_colmax_kernel(const pc_quant_segment* __restrict__ segs,
                     const int32_t* __restrict__ chunk_seg, int total_chunks) {
  __shared__ uint32_t red[32];
  for (int c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    const pc_quant_segment sg = segs[chunk_seg[c]];
    const int64_t numel = (int64_t)sg.rows * sg.cols;
    const int64_t begin = (int64_t)(c - sg.first_chunk) * kQgChunk;
    const int64_t end = begin + kQgChunk < numel ? begin + kQgChunk : numel;
    if (sg.cols == 1) {  // vectors: one bucket for everything -> block reduction, one atomic
      uint32_t mx = 0;
      for (int64_t e = begin + threadIdx.x; e < end; e += kQgThreads) {
        const uint32_t ab = absbits(sg.x[e]);
        mx = ab > mx ? ab : mx;
      }
      mx = block_max_u32(mx, red);
      if (threadIdx.x == 0 && mx) atomicMax(sg.colmax, mx);
      __syncthreads();
    } else if (sg.cols <= kQgThreads && kQgThreads % sg.cols == 0) {
      // a thread always meets the same column: running maximum in a register
      uint32_t mx = 0;
      for (int64_t e = begin + threadIdx.x; e < end; e += kQgThreads) {
        const uint32_t ab = absbits(sg.x[e]);
        mx = ab > mx ? ab : mx;
      }
      if (mx) atomicMax(sg.colmax + (begin + threadIdx.x) % sg.cols, mx);
    } else {
      for (int64_t e = begin + threadIdx.x; e < end; e += kQgThreads) {
        const uint32_t ab = absbits(sg.x[e]);
        if (ab) atomicMax(sg.colmax + e % sg.cols, ab);
      }
    }
  }
}

__global__ void __launch_bounds__(kQgThreads)
qgroup_quantize_kernel(const pc_quant_segment* __restrict__ segs,
                       const int32_t* __restrict__ chunk_seg, int total_chunks) {
  for (int c = blockIdx.x; c < total_chunks; c += gridDim.x) {
    const pc_quant_segment sg = segs[chunk_seg[c]];
    const int64_t numel = (int64_t)sg.rows * sg.cols;
    const int64_t begin = (int64_t)(c - sg.first_chunk) * kQgChunk;
    const int64_t end = begin + kQgChunk < numel ? begin + kQgChunk : numel;
    int8_t* q = reinterpret_cast<int8_t*>(sg.q);
    for (int64_t e = begin + threadIdx.x; e < end; e += kQgThreads) {
      const int col = (int)(e % sg.cols);
      const float bs = __uint_as_float(sg.colmax[col]) / 127.0f;  // QU:86-87
      const float bs_nz = bs > 0.f ? bs : 1.f;                     // QU:90-91
      if (e < sg.cols) sg.bucket[col] = bs;                        // row 0 publishes the buckets
      q[e] = to_q<int8_t>(rintf(sg.x[e] / bs_nz));                 // QU:92-95
    }
  }
}

}  // namespace pc

extern "C" {

int64_t pc_quant_group_chunk_elems(void) { return pc::kQgChunk; }

int pc_dequantize_grouped(const pc_quant_segment* segments, const int32_t* chunk_segment,
                          int num_segments, int64_t total_chunks, void* stream) {
  PC_REQUIRE(num_segments >= 0 && total_chunks >= 0 && total_chunks < (1ll << 31), "bad counts");
  if (num_segments == 0 || total_chunks == 0) return PC_OK;
  PC_REQUIRE(segments && chunk_segment, "null pointer argument");
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)std::min<int64_t>(total_chunks, (int64_t)sms * 8);
  pc::qgroup_dequantize_kernel<<<grid, pc::kQgThreads, 0, (cudaStream_t)stream>>>(
      segments, chunk_segment, (int)total_chunks);
  pc::count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

int pc_quantize_grouped(const pc_quant_segment* segments, const int32_t* chunk_segment,
                        int num_segments, int64_t total_chunks, uint32_t* colmax_all,
                        size_t colmax_bytes, void* stream) {
  PC_REQUIRE(num_segments >= 0 && total_chunks >= 0 && total_chunks < (1ll << 31), "bad counts");
  if (num_segments == 0 || total_chunks == 0) return PC_OK;
  PC_REQUIRE(segments && chunk_segment && colmax_all, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  PC_CUDA_CHECK(cudaMemsetAsync(colmax_all, 0, colmax_bytes, st));
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)std::min<int64_t>(total_chunks, (int64_t)sms * 8);
  pc::qgroup_colmax_kernel<<<grid, pc::kQgThreads, 0, st>>>(segments, chunk_segment,
                                                           (int)total_chunks);
  pc::qgroup_quantize_kernel<<<grid, pc::kQgThreads, 0, st>>>(segments, chunk_segment,
                                                             (int)total_chunks);
  pc::count_launch(2);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

int pc_quantize_batched(const float* x, int batch, int rows, int cols, int qdtype,
                        int extract_diagonal, void* q, float* diag, float* bucket,
                        void* stream) {
  PC_REQUIRE(batch >= 0 && rows >= 0 && cols >= 0, "bad sizes");
  if (batch == 0 || rows == 0 || cols == 0) return PC_OK;
  PC_REQUIRE(x && q, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)batch * rows * cols;
  if (qdtype == PC_QDTYPE_BF16) {
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pc::to_bf16_kernel<<<blocks, 256, 0, st>>>(x, total, (__nv_bfloat16*)q);
    PC_CUDA_CHECK(cudaGetLastError());
    return PC_OK;
  }
  PC_REQUIRE(qdtype == PC_QDTYPE_INT16 || qdtype == PC_QDTYPE_INT8,
             "Quantized dtype %d not supported.", qdtype);
  PC_REQUIRE(bucket != nullptr, "bucket output required");
  PC_REQUIRE(!extract_diagonal || (diag != nullptr && rows == cols),
             "extract_diagonal needs a square matrix and a diagonal output");
  const size_t per = (size_t)rows * cols;
  if (rows >= 256 && batch <= 65535) {
    // tall matrices: two fully parallel passes (column maxima, then element-wise)
    uint32_t* colmax = nullptr;
    PC_CUDA_CHECK(cudaMallocAsync(&colmax, sizeof(uint32_t) * (size_t)batch * cols, st));
    PC_CUDA_CHECK(cudaMemsetAsync(colmax, 0, sizeof(uint32_t) * (size_t)batch * cols, st));
    dim3 g1((cols + 31) / 32, (rows + pc::kQRowChunk - 1) / pc::kQRowChunk, batch), b1(32, 8);
    pc::quant_colmax_kernel<<<g1, b1, 0, st>>>(x, rows, cols, extract_diagonal, colmax);
    dim3 g2((unsigned)((per + 255) / 256 < 2048 ? (per + 255) / 256 : 2048), batch);
    if (qdtype == PC_QDTYPE_INT16)
      pc::quant_apply_kernel<int16_t><<<g2, 256, 0, st>>>(x, colmax, rows, cols, 32767.f,
                                                         extract_diagonal, (int16_t*)q, diag, bucket);
    else
      pc::quant_apply_kernel<int8_t><<<g2, 256, 0, st>>>(x, colmax, rows, cols, 127.f,
                                                        extract_diagonal, (int8_t*)q, diag, bucket);
    PC_CUDA_CHECK(cudaFreeAsync(colmax, st));
    PC_CUDA_CHECK(cudaGetLastError());
    return PC_OK;
  }
  for (int b0 = 0; b0 < batch; b0 += pc::kMaxGridY) {  // grid.y carries the batch
    const int nb = batch - b0 < pc::kMaxGridY ? batch - b0 : pc::kMaxGridY;
    dim3 grid((cols + 31) / 32, nb), block(32, 8);
    float* dg = diag ? diag + (size_t)b0 * rows : nullptr;
    if (qdtype == PC_QDTYPE_INT16)
      pc::quantize_kernel<int16_t><<<grid, block, 0, st>>>(
          x + b0 * per, rows, cols, 32767.f, extract_diagonal, (int16_t*)q + b0 * per, dg,
          bucket + (size_t)b0 * cols);
    else
      pc::quantize_kernel<int8_t><<<grid, block, 0, st>>>(
          x + b0 * per, rows, cols, 127.f, extract_diagonal, (int8_t*)q + b0 * per, dg,
          bucket + (size_t)b0 * cols);
  }
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

int pc_quantize_from_colmax_batched(const float* x, const uint32_t* colmax, int batch, int n,
                                    int qdtype, void* q, float* diag, float* bucket,
                                    void* stream) {
  PC_REQUIRE(batch >= 0 && n >= 0, "bad sizes");
  if (batch == 0 || n == 0) return PC_OK;
  PC_REQUIRE(x && colmax && q && diag && bucket, "null pointer argument");
  PC_REQUIRE(qdtype == PC_QDTYPE_INT16 || qdtype == PC_QDTYPE_INT8,
             "Quantized dtype %d not supported.", qdtype);
  const size_t per = (size_t)n * n;
  for (int b0 = 0; b0 < batch; b0 += pc::kMaxGridY) {
    const int nb = batch - b0 < pc::kMaxGridY ? batch - b0 : pc::kMaxGridY;
    dim3 grid((unsigned)((per + 255) / 256 < 2048 ? (per + 255) / 256 : 2048), nb);
    const size_t v = (size_t)b0 * n;
    if (qdtype == PC_QDTYPE_INT16)
      pc::quantize_from_colmax_kernel<int16_t><<<grid, 256, 0, (cudaStream_t)stream>>>(
          x + b0 * per, colmax + v, n, 32767.f, (int16_t*)q + b0 * per, diag + v, bucket + v);
    else
      pc::quantize_from_colmax_kernel<int8_t><<<grid, 256, 0, (cudaStream_t)stream>>>(
          x + b0 * per, colmax + v, n, 127.f, (int8_t*)q + b0 * per, diag + v, bucket + v);
  }
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

int pc_dequantize_batched(const void* q, const float* diag, const float* bucket, int batch,
                          int rows, int cols, int qdtype, int extract_diagonal, float* x,
                          void* stream) {
  PC_REQUIRE(batch >= 0 && rows >= 0 && cols >= 0, "bad sizes");
  if (batch == 0 || rows == 0 || cols == 0) return PC_OK;
  PC_REQUIRE(x && q, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)batch * rows * cols;
  if (qdtype == PC_QDTYPE_BF16) {
    const int blocks = (int)((total + 255) / 256 < 4096 ? (total + 255) / 256 : 4096);
    pc::from_bf16_kernel<<<blocks, 256, 0, st>>>((const __nv_bfloat16*)q, total, x);
    PC_CUDA_CHECK(cudaGetLastError());
    return PC_OK;
  }
  PC_REQUIRE(qdtype == PC_QDTYPE_INT16 || qdtype == PC_QDTYPE_INT8,
             "Quantized dtype %d not supported.", qdtype);
  PC_REQUIRE(bucket != nullptr, "bucket required");
  PC_REQUIRE(!extract_diagonal || (diag != nullptr && rows == cols),
             "extract_diagonal needs a square matrix and a diagonal");
  const size_t per = (size_t)rows * cols;
  for (int b0 = 0; b0 < batch; b0 += pc::kMaxGridY) {
    const int nb = batch - b0 < pc::kMaxGridY ? batch - b0 : pc::kMaxGridY;
    dim3 grid((unsigned)((per + 255) / 256 < 1024 ? (per + 255) / 256 : 1024), nb);
    const float* dg = diag ? diag + (size_t)b0 * rows : nullptr;
    if (qdtype == PC_QDTYPE_INT16)
      pc::dequantize_kernel<int16_t><<<grid, 256, 0, st>>>(
          (const int16_t*)q + b0 * per, dg, bucket + (size_t)b0 * cols, rows, cols,
          extract_diagonal, x + b0 * per);
    else
      pc::dequantize_kernel<int8_t><<<grid, 256, 0, st>>>(
          (const int8_t*)q + b0 * per, dg, bucket + (size_t)b0 * cols, rows, cols,
          extract_diagonal, x + b0 * per);
  }
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

}  // extern "C"
