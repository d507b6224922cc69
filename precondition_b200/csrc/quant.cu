// QuantizedValue (QU:49-113) on device: per-column symmetric bucket quantisation
// to int16 / int8 with optional diagonal extraction, and bf16 casts.
#include <cuda_bf16.h>

#include <algorithm>
#include <stdint.h>
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

// ---------------------------------------------------------------------------
// Vectorised forms (cols % 4 == 0, 16-byte aligned fp32, 4-element aligned integers): one
// thread = four consecutive columns of one row, rows of all matrices of the batch strided over
// grid.y -- no 64-bit index division per element, 16-byte fp32 and 8 / 4-byte integer accesses.
// Same arithmetic as the scalar kernels (bit-identical output).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void q_load4(const int16_t* p, float (&v)[4]) {
  const short4 s = *reinterpret_cast<const short4*>(p);
  v[0] = (float)s.x; v[1] = (float)s.y; v[2] = (float)s.z; v[3] = (float)s.w;
}
__device__ __forceinline__ void q_load4(const int8_t* p, float (&v)[4]) {
  const char4 s = *reinterpret_cast<const char4*>(p);
  v[0] = (float)s.x; v[1] = (float)s.y; v[2] = (float)s.z; v[3] = (float)s.w;
}
__device__ __forceinline__ void q_store4(int16_t* p, const float (&r)[4]) {
  *reinterpret_cast<short4*>(p) = make_short4((short)r[0], (short)r[1], (short)r[2], (short)r[3]);
}
__device__ __forceinline__ void q_store4(int8_t* p, const float (&r)[4]) {
  *reinterpret_cast<char4*>(p) = make_char4((signed char)r[0], (signed char)r[1],
                                            (signed char)r[2], (signed char)r[3]);
}

template <typename Q>
__global__ void __launch_bounds__(256)
dequantize_vec_kernel(const Q* __restrict__ q, const float* __restrict__ diag,
                      const float* __restrict__ bucket, int batch, int rows, int cols,
                      int extract_diagonal, float* __restrict__ x) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= cols) return;
  const int total_rows = batch * rows;
  for (int R = blockIdx.y; R < total_rows; R += gridDim.y) {
    const int b = R / rows, r = R - b * rows;
    const size_t off = (size_t)R * cols + c;
    float v[4];
    q_load4(q + off, v);
    const float4 bs = *reinterpret_cast<const float4*>(bucket + (size_t)b * cols + c);
    float4 o = make_float4(v[0] * bs.x, v[1] * bs.y, v[2] * bs.z, v[3] * bs.w);  // QU:110
    if (extract_diagonal && r >= c && r < c + 4) {                               // QU:111-112
      const float dg = diag[(size_t)b * rows + r];
      if (r == c) o.x += dg; else if (r == c + 1) o.y += dg;
      else if (r == c + 2) o.z += dg; else o.w += dg;
    }
    *reinterpret_cast<float4*>(x + off) = o;
  }
}

// kDiagBucket: the diagonal element's thread publishes the bucket (quantize_from_colmax, square
// matrices with an extracted diagonal); otherwise row 0 does (quant_apply)
template <typename Q, bool kDiagBucket>
__global__ void __launch_bounds__(256)
quantize_vec_kernel(const float* __restrict__ x, const uint32_t* __restrict__ colmax, int batch,
                    int rows, int cols, float num_buckets, int extract_diagonal,
                    Q* __restrict__ q, float* __restrict__ diag, float* __restrict__ bucket) {
  const int c = (blockIdx.x * 256 + threadIdx.x) * 4;
  if (c >= cols) return;
  const int total_rows = batch * rows;
  for (int R = blockIdx.y; R < total_rows; R += gridDim.y) {
    const int b = R / rows, r = R - b * rows;
    const size_t off = (size_t)R * cols + c;
    const float4 xv = *reinterpret_cast<const float4*>(x + off);
    const uint4 cm = *reinterpret_cast<const uint4*>(colmax + (size_t)b * cols + c);
    float v[4] = {xv.x, xv.y, xv.z, xv.w};
    const float bs[4] = {__uint_as_float(cm.x) / num_buckets, __uint_as_float(cm.y) / num_buckets,
                         __uint_as_float(cm.z) / num_buckets, __uint_as_float(cm.w) / num_buckets};
    float out[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const bool on_diag = extract_diagonal && r == c + k;
      if (on_diag) {
        diag[(size_t)b * rows + r] = v[k];  // QU:72-76
        v[k] = v[k] - v[k];
      }
      if (kDiagBucket ? (r == c + k) : (r == 0)) bucket[(size_t)b * cols + c + k] = bs[k];
      const float bs_nz = bs[k] > 0.f ? bs[k] : 1.f;  // QU:90-91
      out[k] = rintf(v[k] / bs_nz);                   // QU:92-95
    }
    q_store4(q + off, out);
  }
}

static bool quant_vec_ok(const void* x, const void* q, const void* aux, int cols, int elem_q,
                         long long total_rows) {
  return cols % 4 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)q & (4 * elem_q - 1)) == 0 &&
         ((uintptr_t)aux & 15) == 0 && total_rows < (1ll << 31);
}
static dim3 quant_vec_grid(int cols, long long total_rows) {
  return dim3((unsigned)((cols / 4 + 255) / 256),
              (unsigned)std::min<long long>(total_rows, 8192));
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
    const bool vec = sg.cols % 4 == 0 && (((uintptr_t)sg.x | (uintptr_t)sg.bucket) & 15) == 0 &&
                     ((uintptr_t)sg.q & 3) == 0;
    if (vec) {  // four consecutive elements share a row: 32-bit column arithmetic per group
      const int col0 = (int)(begin % sg.cols), n4 = (int)((end - begin) >> 2);
      for (int g = threadIdx.x; g < n4; g += kQgThreads) {
        const int col = (col0 + 4 * g) % sg.cols;
        const char4 qv = *reinterpret_cast<const char4*>(q + begin + 4 * g);
        const float4 bs = *reinterpret_cast<const float4*>(sg.bucket + col);
        *reinterpret_cast<float4*>(sg.x + begin + 4 * g) =
            make_float4((float)qv.x * bs.x, (float)qv.y * bs.y, (float)qv.z * bs.z,
                        (float)qv.w * bs.w);  // QU:107-108
      }
    } else {
      for (int64_t e = begin + threadIdx.x; e < end; e += kQgThreads)
        sg.x[e] = (float)q[e] * sg.bucket[e % sg.cols];
    }
  }
}

// per-column max |x| (bit patterns, atomicMax) into sg.colmax (zero on entry).  Work items are
// tiles of kQgTileRows rows x 128 columns: a lane owns four consecutive columns, a warp strides
// over the rows, the eight warps meet in shared memory, one atomic per column and tile.
constexpr int kQgTileRows = 128;
__global__ void __launch_bounds__(kQgThreads)
qgroup_colmax_kernel(const pc_quant_segment* __restrict__ segs,
                     const int32_t* __restrict__ tile_seg, int total_tiles) {
  __shared__ uint32_t smax[8][128];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const pc_quant_segment sg = segs[tile_seg[t]];
    const int local = t - sg.first_tile;
    const int tr = local / sg.col_tiles, tc = local - tr * sg.col_tiles;
    const int r0 = tr * kQgTileRows, r1 = min(sg.rows, r0 + kQgTileRows);
    const int c = tc * 128 + lane * 4;
    uint32_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
    const bool vec = sg.cols % 4 == 0 && ((uintptr_t)sg.x & 15) == 0;
    if (c < sg.cols) {
      for (int r = r0 + warp; r < r1; r += 8) {
        const float* row = sg.x + (size_t)r * sg.cols + c;
        if (vec) {
          const float4 v = *reinterpret_cast<const float4*>(row);
          m0 = max(m0, absbits(v.x)); m1 = max(m1, absbits(v.y));
          m2 = max(m2, absbits(v.z)); m3 = max(m3, absbits(v.w));
        } else {
          m0 = max(m0, absbits(row[0]));
          if (c + 1 < sg.cols) m1 = max(m1, absbits(row[1]));
          if (c + 2 < sg.cols) m2 = max(m2, absbits(row[2]));
          if (c + 3 < sg.cols) m3 = max(m3, absbits(row[3]));
        }
      }
    }
    __syncthreads();
    smax[warp][lane * 4] = m0; smax[warp][lane * 4 + 1] = m1;
    smax[warp][lane * 4 + 2] = m2; smax[warp][lane * 4 + 3] = m3;
    __syncthreads();
    if (threadIdx.x < 128) {
      uint32_t m = 0;
#pragma unroll
      for (int w = 0; w < 8; ++w) m = max(m, smax[w][threadIdx.x]);
      const int col = tc * 128 + threadIdx.x;
      if (col < sg.cols && m) atomicMax(sg.colmax + col, m);
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
    const bool vec = sg.cols % 4 == 0 && (((uintptr_t)sg.x | (uintptr_t)sg.colmax) & 15) == 0 &&
                     ((uintptr_t)sg.q & 3) == 0;
    if (vec) {
      const int col0 = (int)(begin % sg.cols), n4 = (int)((end - begin) >> 2);
      for (int g = threadIdx.x; g < n4; g += kQgThreads) {
        const int col = (col0 + 4 * g) % sg.cols;
        const float4 xv = *reinterpret_cast<const float4*>(sg.x + begin + 4 * g);
        const uint4 cm = *reinterpret_cast<const uint4*>(sg.colmax + col);
        const float b0 = __uint_as_float(cm.x) / 127.0f, b1 = __uint_as_float(cm.y) / 127.0f,
                    b2 = __uint_as_float(cm.z) / 127.0f, b3 = __uint_as_float(cm.w) / 127.0f;
        if (begin + 4 * g < sg.cols) {  // row 0 publishes the buckets
          sg.bucket[col] = b0; sg.bucket[col + 1] = b1;
          sg.bucket[col + 2] = b2; sg.bucket[col + 3] = b3;
        }
        *reinterpret_cast<char4*>(q + begin + 4 * g) = make_char4(
            (signed char)rintf(xv.x / (b0 > 0.f ? b0 : 1.f)),
            (signed char)rintf(xv.y / (b1 > 0.f ? b1 : 1.f)),
            (signed char)rintf(xv.z / (b2 > 0.f ? b2 : 1.f)),
            (signed char)rintf(xv.w / (b3 > 0.f ? b3 : 1.f)));  // QU:86-95
      }
    } else {
      for (int64_t e = begin + threadIdx.x; e < end; e += kQgThreads) {
        const int col = (int)(e % sg.cols);
        const float bs = __uint_as_float(sg.colmax[col]) / 127.0f;  // QU:86-87
        const float bs_nz = bs > 0.f ? bs : 1.f;                     // QU:90-91
        if (e < sg.cols) sg.bucket[col] = bs;                        // row 0 publishes the buckets
        q[e] = to_q<int8_t>(rintf(sg.x[e] / bs_nz));                 // QU:92-95
      }
    }
  }
}

}  // namespace pc

extern "C" {

int64_t pc_quant_group_chunk_elems(void) { return pc::kQgChunk; }
int pc_quant_group_tile_rows(void) { return pc::kQgTileRows; }

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
                        int num_segments, int64_t total_chunks, const int32_t* tile_segment,
                        int64_t total_tiles, uint32_t* colmax_all, size_t colmax_bytes,
                        void* stream) {
  PC_REQUIRE(num_segments >= 0 && total_chunks >= 0 && total_chunks < (1ll << 31) &&
             total_tiles >= 0 && total_tiles < (1ll << 31), "bad counts");
  if (num_segments == 0 || total_chunks == 0) return PC_OK;
  PC_REQUIRE(segments && chunk_segment && tile_segment && colmax_all, "null pointer argument");
  cudaStream_t st = (cudaStream_t)stream;
  PC_CUDA_CHECK(cudaMemsetAsync(colmax_all, 0, colmax_bytes, st));
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)std::min<int64_t>(total_chunks, (int64_t)sms * 8);
  const unsigned tgrid = (unsigned)std::min<int64_t>(total_tiles, (int64_t)sms * 8);
  pc::qgroup_colmax_kernel<<<tgrid, pc::kQgThreads, 0, st>>>(segments, tile_segment,
                                                            (int)total_tiles);
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
    const long long trows = (long long)batch * rows;
    const bool vec = pc::quant_vec_ok(x, q, bucket, cols, qdtype == PC_QDTYPE_INT16 ? 2 : 1, trows);
    if (vec && qdtype == PC_QDTYPE_INT16)
      pc::quantize_vec_kernel<int16_t, false><<<pc::quant_vec_grid(cols, trows), 256, 0, st>>>(
          x, colmax, batch, rows, cols, 32767.f, extract_diagonal, (int16_t*)q, diag, bucket);
    else if (vec)
      pc::quantize_vec_kernel<int8_t, false><<<pc::quant_vec_grid(cols, trows), 256, 0, st>>>(
          x, colmax, batch, rows, cols, 127.f, extract_diagonal, (int8_t*)q, diag, bucket);
    else if (qdtype == PC_QDTYPE_INT16)
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
  {
    const long long trows = (long long)batch * n;
    if (pc::quant_vec_ok(x, q, colmax, n, qdtype == PC_QDTYPE_INT16 ? 2 : 1, trows) &&
        ((uintptr_t)bucket & 3) == 0) {
      cudaStream_t st = (cudaStream_t)stream;
      if (qdtype == PC_QDTYPE_INT16)
        pc::quantize_vec_kernel<int16_t, true><<<pc::quant_vec_grid(n, trows), 256, 0, st>>>(
            x, colmax, batch, n, n, 32767.f, 1, (int16_t*)q, diag, bucket);
      else
        pc::quantize_vec_kernel<int8_t, true><<<pc::quant_vec_grid(n, trows), 256, 0, st>>>(
            x, colmax, batch, n, n, 127.f, 1, (int8_t*)q, diag, bucket);
      PC_CUDA_CHECK(cudaGetLastError());
      return PC_OK;
    }
  }
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
  {
    const long long trows = (long long)batch * rows;
    if (pc::quant_vec_ok(x, q, bucket, cols, qdtype == PC_QDTYPE_INT16 ? 2 : 1, trows)) {
      if (qdtype == PC_QDTYPE_INT16)
        pc::dequantize_vec_kernel<int16_t><<<pc::quant_vec_grid(cols, trows), 256, 0, st>>>(
            (const int16_t*)q, diag, bucket, batch, rows, cols, extract_diagonal, x);
      else
        pc::dequantize_vec_kernel<int8_t><<<pc::quant_vec_grid(cols, trows), 256, 0, st>>>(
            (const int8_t*)q, diag, bucket, batch, rows, cols, extract_diagonal, x);
      PC_CUDA_CHECK(cudaGetLastError());
      return PC_OK;
    }
  }
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
