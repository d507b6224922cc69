// Kernel templates shared by the fp32 and the tcgen05 engines of the Newton
// solver: (re)initialisation of a try and the final gather of roots / metrics.
// `Bufs` abstracts how a logical matrix element is stored (fp32, or three bf16
// planes whose sum is the fp32 value).
#pragma once
#include "root_common.cuh"

namespace pc {

// ---------------------------------------------------------------------------
// element store/load policies: fp32 buffers (SIMT engine) or 3 bf16 planes (TC)
// ---------------------------------------------------------------------------
struct F32Bufs {
  float* base[kNumBufs];
  size_t mat_elems;  // n*n
  __device__ __forceinline__ float* mat(int phys, int b, int batch) const {
    return base[phys] + (size_t)b * mat_elems;
  }
};

// ---------------------------------------------------------------------------
// (re)initialise a try: DS:866-875
// ---------------------------------------------------------------------------
template <class Bufs>
__global__ void __launch_bounds__(1024)
root_init_kernel(const float* __restrict__ xs, RootCtl* ctl, Bufs bufs, int batch, int n,
                 RootParams prm, float* __restrict__ roots) {
  __shared__ float scratch[32];
  __shared__ uint32_t uscratch[32];
  const int b = blockIdx.x;
  RootCtl c = ctl[b];
  if (!c.need_init) return;
  const float* A = xs + (size_t)b * n * n;
  const int pad = c.pad, p = c.p;
  if (c.tries == 0) {
    const float ev = prm.relative_eps ? c.max_ev : 1.0f;
    c.max_ev = ev;
    c.ridge = prm.ridge_epsilon * fmaxf(ev, 1e-25f);  // DS:830
  }
  const float alpha = -1.0f / (float)p;  // DS:774
  if (n == 1) {  // DS:850-855
    if (threadIdx.x == 0) {
      const float a = pad > 0 ? A[0] : 0.f;
      // power iteration on a 1x1 matrix returns the entry itself (v = +-1)
      const float ev = prm.relative_eps ? a : 1.0f;
      c.max_ev = ev;
      c.ridge = prm.ridge_epsilon * fmaxf(ev, 1e-25f);
      roots[(size_t)b] = powf(a + c.ridge, alpha);
      c.need_init = 0; c.done = 1; c.active = 0;
      c.m_err = 0.f; c.m_iters = 0.f; c.m_ratio = 0.f; c.m_retries = 0.f;
      c.result_h = -1;  // already written
      ctl[b] = c;
    }
    return;
  }
  float tenpow = 1.f;
  for (int t = 0; t < c.tries; ++t) tenpow *= 10.f;
  const float eps = c.ridge * tenpow;  // DS:869
  // pass 1: Frobenius norm of the damped, masked matrix (DS:870)
  float ss = 0.f;
  const size_t total = (size_t)pad * pad;
  for (size_t e = threadIdx.x; e < total; e += blockDim.x) {
    const int i = (int)(e / pad), j = (int)(e - (size_t)i * pad);
    // the lower triangle is authoritative (statistics are symmetric; this makes
    // the iterates bitwise symmetric even if the caller's matrix is not)
    float a = __ldg(A + (size_t)max(i, j) * n + min(i, j));
    if (i == j) a += eps;
    ss = fmaf(a, a, ss);
  }
  const float norm = sqrtf(block_sum(ss, scratch));
  const float z = (float)(1 + p) / (2.0f * norm);
  const float h0 = powf(z, (float)(1.0 / (double)p));  // DS:873
  float hmul = 1.0f;
  const float hdiag = bufs.h_init_scale(h0, powf(fmaxf(z * eps, 1e-37f), alpha), &hmul);
  // pass 2: M0 = z A_d, M_i0 = (1-alpha) I_m + alpha M0, H0 = z^(1/p) I_m,
  //         err0 = max|M0 - I_m|
  const float one_minus_alpha = 1.0f - alpha;
  uint32_t emax = 0;
  const size_t nn = (size_t)n * n;
  for (size_t e = threadIdx.x; e < nn; e += blockDim.x) {
    const int i = (int)(e / n), j = (int)(e - (size_t)i * n);
    float m0 = 0.f, mi = 0.f, h = 0.f;
    if (i < pad && j < pad) {
      float a = __ldg(A + (size_t)max(i, j) * n + min(i, j));
      if (i == j) a += eps;
      m0 = a * z;                                   // DS:871
      const float e0 = m0 - (i == j ? 1.f : 0.f);   // M0 - I_m
      const uint32_t ab = absbits(e0);
      emax = ab > emax ? ab : emax;
      mi = mi_from_m(m0, i == j, alpha, one_minus_alpha);
      h = (i == j) ? hdiag : 0.f;
    }
    bufs.store(0, b, i, j, n, m0);  // M[0]
    bufs.store(2, b, i, j, n, mi);  // M_i[0]
    bufs.store(4, b, i, j, n, h);   // H[0]
  }
  emax = block_max_u32(emax, uscratch);
  if (threadIdx.x == 0) {
    c.need_init = 0;
    c.iter = 0;
    c.cur = 0;
    c.err = __uint_as_float(emax);  // DS:872
    c.ratio = 1.0f;
    c.hmul = hmul;
    root_after_error_update(c, prm);
    ctl[b] = c;
  }
}


struct F32Store : F32Bufs {
  __device__ __forceinline__ float h_init_scale(float h0, float, float* hmul) const {
    *hmul = 1.0f;
    return h0;
  }
  __device__ __forceinline__ void store(int phys, int b, int i, int j, int n, float v) const {
    base[phys][(size_t)b * mat_elems + (size_t)i * n + j] = v;
  }
  __device__ __forceinline__ float load(int phys, int b, int i, int j, int n) const {
    return base[phys][(size_t)b * mat_elems + (size_t)i * n + j];
  }
};


// ---------------------------------------------------------------------------
// First initialisation of a call, spread over the whole GPU (n >= 256): the one-CTA-per-
// matrix kernel above streams 20 MB per 1024^2 matrix through a single SM.  Same maths,
// two launches over a grid of (32-row strips, batch):
//   root_norm_strip_kernel   partial Frobenius sums of the damped, masked matrix from its
//                            lower triangle (2 a_ij^2 below the diagonal, a_ii^2 on it)
//   root_init_strip_kernel   every CTA adds the partials in the same order (identical z),
//                            then fills its strip of M0 / M_i0 / H0; strictly-upper 32 x 32
//                            tiles are read as the transposed lower tile through shared
//                            memory, so all global reads are coalesced and the lower triangle
//                            stays authoritative.  The last CTA of a matrix to finish
//                            publishes err0 and runs the loop-predicate bookkeeping.
// scratch: [batch * strips] partial sums, then [batch] err bits, [batch] done counters.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float root_try_eps(const RootCtl& c, const RootParams& prm) {
  float ridge = c.ridge;
  if (c.tries == 0) {
    const float ev = prm.relative_eps ? c.max_ev : 1.0f;
    ridge = prm.ridge_epsilon * fmaxf(ev, 1e-25f);  // DS:830
  }
  float tenpow = 1.f;
  for (int t = 0; t < c.tries; ++t) tenpow *= 10.f;
  return ridge * tenpow;  // DS:869
}

static __global__ void __launch_bounds__(256)
root_norm_strip_kernel(const float* __restrict__ xs, const RootCtl* __restrict__ ctl, int n,
                       int strips, RootParams prm, float* __restrict__ scratch, int batch) {
  __shared__ float red[32];
  const int b = blockIdx.y, strip = blockIdx.x;
  const RootCtl c = ctl[b];
  if (!c.need_init) return;
  if (strip == 0 && threadIdx.x == 0) {  // reset this matrix's err / done slots
    reinterpret_cast<uint32_t*>(scratch + (size_t)batch * strips)[b] = 0u;
    reinterpret_cast<uint32_t*>(scratch + (size_t)batch * strips)[batch + b] = 0u;
  }
  const float eps = root_try_eps(c, prm);
  const float* A = xs + (size_t)b * n * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float ss = 0.f;
  for (int r = warp; r < 32; r += 8) {
    const int i = strip * 32 + r;
    if (i >= c.pad) continue;
    for (int j = lane; j <= i; j += 32) {
      float a = __ldg(A + (size_t)i * n + j);
      if (i == j) { a += eps; ss = fmaf(a, a, ss); }
      else ss = fmaf(2.0f * a, a, ss);
    }
  }
  ss = block_sum(ss, red);
  if (threadIdx.x == 0) scratch[(size_t)b * strips + strip] = ss;
}

template <class Bufs>
__global__ void __launch_bounds__(1024)
root_init_strip_kernel(const float* __restrict__ xs, RootCtl* ctl, Bufs bufs, int batch, int n,
                       int strips, RootParams prm, float* __restrict__ scratch) {
  __shared__ float tile[32][33];
  __shared__ uint32_t ured[32];
  __shared__ int is_last;
  const int b = blockIdx.y, strip = blockIdx.x;
  RootCtl c = ctl[b];
  if (!c.need_init) return;
  const int pad = c.pad, p = c.p;
  const float eps = root_try_eps(c, prm);
  float ssum = 0.f;
  for (int r = 0; r < strips; ++r) ssum += scratch[(size_t)b * strips + r];  // same order everywhere
  const float norm = sqrtf(ssum);
  const float alpha = -1.0f / (float)p, one_minus_alpha = 1.0f - alpha;
  const float z = (float)(1 + p) / (2.0f * norm);
  const float h0 = powf(z, (float)(1.0 / (double)p));  // DS:873
  float hmul = 1.0f;
  const float hdiag = bufs.h_init_scale(h0, powf(fmaxf(z * eps, 1e-37f), alpha), &hmul);
  const float* A = xs + (size_t)b * n * n;
  const int r = threadIdx.x >> 5, cidx = threadIdx.x & 31;
  const int i = strip * 32 + r;
  uint32_t emax = 0;
  for (int tj = 0; tj * 32 < n; ++tj) {
    const int j = tj * 32 + cidx;
    float a = 0.f;
    if (tj < strip) {  // below the diagonal: direct, coalesced rows
      if (i < pad && j < pad) a = __ldg(A + (size_t)i * n + j);
    } else {           // diagonal or above: load the mirrored tile, read it transposed
      __syncthreads();
      const int mi = tj * 32 + r, mj = strip * 32 + cidx;  // element (mi, mj) of the lower part
      tile[r][cidx] = (mi < pad && mj < pad && mi < n && mj < n) ? __ldg(A + (size_t)mi * n + mj) : 0.f;
      __syncthreads();
      if (tj == strip) a = (cidx <= r) ? tile[r][cidx] : tile[cidx][r];  // A[max][min]
      else a = tile[cidx][r];
      if (!(i < pad && j < pad)) a = 0.f;
    }
    if (i >= n || j >= n) continue;
    float m0 = 0.f, mi0 = 0.f, h = 0.f;
    if (i < pad && j < pad) {
      if (i == j) a += eps;
      m0 = a * z;                                   // DS:871
      const uint32_t ab = absbits(m0 - (i == j ? 1.f : 0.f));
      emax = ab > emax ? ab : emax;
      mi0 = mi_from_m(m0, i == j, alpha, one_minus_alpha);
      h = (i == j) ? hdiag : 0.f;
    }
    bufs.store(0, b, i, j, n, m0);
    bufs.store(2, b, i, j, n, mi0);
    bufs.store(4, b, i, j, n, h);
  }
  emax = block_max_u32(emax, ured);
  uint32_t* slots = reinterpret_cast<uint32_t*>(scratch + (size_t)batch * strips);
  if (threadIdx.x == 0) {
    atomicMax(slots + b, emax);
    __threadfence();
    is_last = atomicAdd(slots + batch + b, 1u) == (uint32_t)(strips - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    const uint32_t e = atomicMax(slots + b, 0u);  // atomic read of the final maximum
    if (c.tries == 0) {
      const float ev = prm.relative_eps ? c.max_ev : 1.0f;
      c.max_ev = ev;
      c.ridge = prm.ridge_epsilon * fmaxf(ev, 1e-25f);
    }
    c.need_init = 0;
    c.iter = 0;
    c.cur = 0;
    c.err = __uint_as_float(e);  // DS:872
    c.ratio = 1.0f;
    c.hmul = hmul;
    root_after_error_update(c, prm);
    ctl[b] = c;
  }
}


template <class Bufs>
__global__ void root_final_kernel(const RootCtl* __restrict__ ctl, Bufs bufs, int n,
                                  float* __restrict__ roots, float* __restrict__ metrics) {
  const int b = blockIdx.y;
  const RootCtl& c = ctl[b];
  const size_t nn = (size_t)n * n;
  float* out = roots + (size_t)b * nn;
  if (c.result_h != -1) {  // -1: already written by the n == 1 closed form
    const bool zero = (c.result_h == -2) || (c.pad == 0) || !c.done;  // DS:930-937
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn;
         e += (size_t)gridDim.x * blockDim.x) {
      const int i = (int)(e / n), j = (int)(e - (size_t)i * n);
      out[e] = zero ? 0.f : bufs.load(4 + c.result_h, b, i, j, n) * c.hmul;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float* m = metrics + (size_t)b * PC_NUM_METRICS;
    m[PC_METRIC_ERROR] = c.pad == 0 ? 0.f : c.m_err;
    m[PC_METRIC_ITERS] = c.m_iters;
    m[PC_METRIC_ERROR_RATIO] = c.m_ratio;
    m[PC_METRIC_MAX_EV] = c.max_ev;
    m[PC_METRIC_RETRIES] = c.m_retries;
  }
}


}  // namespace pc
