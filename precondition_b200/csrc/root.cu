// Batched matrix_inverse_pth_root (DS:702-940): power iteration, ridge damping,
// coupled Newton iteration with on-device convergence / retry state machine.
// GEMM chain engines: CUDA-core fp32 (this file) and tcgen05 split-bf16
// (tc_gemm.cu), both behind the same per-matrix step programs.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <vector>

#include "root_common.cuh"
#include "simt_gemm.cuh"
#include "root_kernels.cuh"
#include "tc_engine.cuh"

namespace pc {

// persistent per-matrix solver for n <= 128 (small_root.cu)
size_t small_root_workspace_bytes(int batch, int n);
bool small_root_supported_exponents(const int32_t* ps_host, int batch);
int run_small_root(const float* xs, const int32_t* ps, const int32_t* pads, int batch, int n,
                   const pc_root_options* opt, const float* v0_device, float* roots,
                   float* metrics, void* workspace, size_t workspace_bytes, cudaStream_t stream);

// ---------------------------------------------------------------------------
// step programs
// ---------------------------------------------------------------------------
bool build_program(int p, Program* out) {
  memset(out, 0, sizeof(*out));
  if (p < 1 || p > kMaxP) return false;
  bool used[4] = {false, false, false, false};
  auto alloc = [&]() -> int {
    for (int q = 0; q < 4; ++q)
      if (!used[q]) { used[q] = true; return LB_Q0 + q; }
    return -100;
  };
  int mat = LB_MI, power = LB_NONE, n = 0;
  auto release_if_unreferenced = [&](int buf) {
    if (buf >= LB_Q0 && buf != mat && buf != power) used[buf - LB_Q0] = false;
  };
  auto push = [&](Step s) -> bool {
    if (n >= kMaxSteps) return false;
    out->steps[n++] = s;
    return true;
  };
  int i = p;
  while (i > 0) {
    if (i & 1) {
      if (power == LB_NONE) {
        power = mat;  // mat @ I (DS:661, DS:670-672) is exact: alias instead
      } else {
        int dst = alloc();
        if (dst < 0) return false;
        if (!push(Step{(int8_t)dst, (int8_t)mat, (int8_t)power, 0})) return false;  // DS:671
        int old = power;
        power = dst;
        release_if_unreferenced(old);
      }
    }
    i >>= 1;
    if (i > 0) {  // the reference squares once more at i == 0 and discards it
      int dst = alloc();
      if (dst < 0) return false;
      if (!push(Step{(int8_t)dst, (int8_t)mat, (int8_t)mat, 0})) return false;  // DS:674
      int old = mat;
      mat = dst;
      release_if_unreferenced(old);
    }
  }
  // M' = M_i^p M (DS:845); epilogue also emits M_i' (DS:844) and err (DS:847)
  if (!push(Step{LB_MN, (int8_t)power, LB_M, 1})) return false;
  out->nsteps = n;
  return true;
}

__constant__ Program c_programs[kMaxP + 1];
// H' = H M_i  (DS:846), co-scheduled with step 0 of every program
__constant__ Step c_hstep = {LB_HN, LB_H, LB_MI, 0};

static Program h_programs[kMaxP + 1];
static bool programs_uploaded[64] = {false};

static int ensure_programs() {
  int dev = 0;
  PC_CUDA_CHECK(cudaGetDevice(&dev));
  if (dev < 64 && programs_uploaded[dev]) return PC_OK;
  for (int p = 1; p <= kMaxP; ++p) {
    if (!build_program(p, &h_programs[p])) {
      set_error("internal: cannot build Newton program for p=%d", p);
      return PC_ERR_INVALID;
    }
  }
  PC_CUDA_CHECK(cudaMemcpyToSymbol(c_programs, h_programs, sizeof(h_programs)));
  if (dev < 64) programs_uploaded[dev] = true;
  return PC_OK;
}

// ---------------------------------------------------------------------------
// setup
// ---------------------------------------------------------------------------
__global__ void root_setup_kernel(RootCtl* ctl, const int32_t* ps, const int32_t* pads,
                                  int batch, int n, uint32_t* errbits) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  RootCtl c;
  memset(&c, 0, sizeof(c));
  c.p = ps[b];
  int pad = pads ? pads[b] : n;
  c.pad = pad < 0 ? 0 : (pad > n ? n : pad);
  c.max_ev = 1.0f;
  c.ratio = 1.0f;
  c.err = 1000.0f;   // DS:860 initial outer state
  c.m_err = 1000.0f;
  c.m_iters = 100.f;
  c.m_ratio = 1.0f;
  if (c.pad == 0 || c.p < 1 || c.p > kMaxP) {
    c.done = 1;  // DS:930-937 (all padding) / unsupported exponent -> NaN error
    c.result_h = -2;  // root = zeros
    if (c.pad != 0) c.m_err = __int_as_float(0x7fc00000);
  } else {
    c.need_init = 1;
  }
  ctl[b] = c;
  errbits[b] = 0u;
}

// ---------------------------------------------------------------------------
// power iteration (DS:595-652): <= 100 strictly dependent mat-vecs per matrix.
// A cluster of `csize` CTAs owns one matrix: CTA r computes rows [r*n/csize, ...) of
// y = A v/|v|, publishes its slice in a double-buffered global vector and a
// cluster barrier makes it visible; every CTA then forms |y|, v.y redundantly in
// the same order, so all CTAs take identical control decisions.  The matrix
// (4 MiB at n = 1024) streams from L2/HBM once per iteration across csize SMs
// instead of one.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pi_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void pi_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(1024)
power_iteration_kernel(const float* __restrict__ xs, const int32_t* __restrict__ pads,
                       const float* __restrict__ v0, int n, int num_iters, float tol,
                       float* __restrict__ lambdas, int32_t* __restrict__ iters,
                       RootCtl* ctl, float* __restrict__ ybuf, int csize) {
  extern __shared__ float smem[];
  float* v = smem;       // current iterate (full vector)
  float* nv = smem + n;  // normalised iterate
  __shared__ float scratch[32];
  const int b = blockIdx.x / csize;
  const int crank = csize > 1 ? (int)pi_cluster_rank() : 0;
  // `pad` masks the MATRIX (solver path: DS:777-783 masks it before DS:820);
  // `vpad` masks only the start vector (stand-alone power_iteration, DS:645-646).
  int pad = n, vpad = n;
  bool skip = false;
  if (ctl) {
    skip = ctl[b].done != 0;
    pad = vpad = ctl[b].pad;
  } else if (pads) {
    vpad = min(max(pads[b], 0), n);
  }
  if (skip) return;  // uniform across the cluster
  const float* A = xs + (size_t)b * n * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int rows_per = (n + csize - 1) / csize;
  const int r_lo = crank * rows_per, r_hi = min(n, r_lo + rows_per);
  for (int i = threadIdx.x; i < n; i += blockDim.x) v[i] = i < vpad ? v0[i] : 0.f;
  __syncthreads();
  float s = 0.f;
  int it = 0;
  bool run = true;
  const bool vec4 = (n % 4 == 0);
  while (it < num_iters && run) {
    float ss = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) ss += v[i] * v[i];
    const float norm = sqrtf(block_sum(ss, scratch));
    for (int i = threadIdx.x; i < n; i += blockDim.x) nv[i] = v[i] / norm;  // DS:634
    __syncthreads();
    // one CTA per matrix: the product stays in shared memory (no L2 round trip per step)
    float* yout = csize > 1 ? ybuf + ((size_t)b * 2 + (it & 1)) * n : smem + 2 * n;
    for (int row = r_lo + warp; row < r_hi; row += nwarp) {
      float acc = 0.f;
      if (row < pad) {
        const float* ar = A + (size_t)row * n;
        if (vec4) {
          for (int c = lane * 4; c < pad; c += 128) {
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(ar + c));
            const float m1 = c + 1 < pad ? a4.y : 0.f, m2 = c + 2 < pad ? a4.z : 0.f,
                        m3 = c + 3 < pad ? a4.w : 0.f;
            acc = fmaf(a4.x, nv[c], acc);
            acc = fmaf(m1, nv[c + 1], acc);
            acc = fmaf(m2, nv[c + 2], acc);
            acc = fmaf(m3, nv[c + 3], acc);
          }
        } else {
          for (int c = lane; c < pad; c += 32) acc = fmaf(__ldg(ar + c), nv[c], acc);
        }
        acc = warp_sum(acc);
      }
      if (lane == 0) yout[row] = acc;  // DS:636
    }
    if (csize > 1) {
      pi_cluster_sync();
    } else {
      __threadfence_block();
      __syncthreads();
    }
    float dot = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float yi = csize > 1 ? __ldcg(yout + i) : yout[i];
      v[i] = yi;
      dot += nv[i] * yi;
    }
    const float s_new = block_sum(dot, scratch);  // DS:637
    run = fabsf(s_new - s) > tol;                 // DS:639 (NaN -> stop)
    s = s_new;
    ++it;
    __syncthreads();
  }
  if (threadIdx.x == 0 && crank == 0) {
    if (lambdas) lambdas[b] = s;
    if (iters) iters[b] = it;
    if (ctl) ctl[b].max_ev = s;
  }
}

// ---------------------------------------------------------------------------
// Power iteration over the LOWER TRIANGLE only (solver path: the statistics are symmetric and
// everything after this reads them that way).  Each a(r,c), c < r, is loaded once per step and
// used twice -- y_r += a v_c (warp reduction per row) and y_c += a v_r (per-lane column
// accumulators, reduced across warps through shared memory in a fixed order, so the result
// is reproducible) -- which halves the bytes of the HBM-bound sweep.  Rows are paired
// (t, pad-1-t) so every warp task has pad+1 elements; the sweep direction alternates between
// steps so the tail of one sweep is the head of the next and is still in L2.
// n <= 128 * NJ, n % 4 == 0.  ybuf: [batch, 2, csize, n].
// ---------------------------------------------------------------------------
__device__ __forceinline__ void pi_prefetch_l2(const float* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

__device__ __forceinline__ uint32_t pi_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void pi_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "PI_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra PI_DONE;\n\t"
      "bra PI_WAIT;\n\t"
      "PI_DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// 1-D bulk copy global -> shared, completion counted in bytes on `bar`
__device__ __forceinline__ void pi_bulk_load(uint32_t dst, const float* src, uint32_t bytes,
                                             uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst), "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// row r of the lower triangle staged in shared memory at `a`; NC column chunks of 128.
// Chunks left of the one holding the diagonal need no masks (a warp-uniform test).
template <int NC, int NJ>
__device__ __forceinline__ float pi_sym_fma(const float* a, int r, int lane, const float* nv,
                                            float (&cacc)[NJ * 4]) {
  const float vr = nv[r];
  const int jd = r >> 7;  // chunk of the diagonal element
  const float4* a4 = reinterpret_cast<const float4*>(a) + lane;
  const float4* x4 = reinterpret_cast<const float4*>(nv) + lane;
  float racc = 0.f;
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    if (j < jd) {
      // every element feeds its row and its column
      const float4 e = a4[32 * j];
      const float4 x = x4[32 * j];
      racc = fmaf(e.x, x.x, racc);
      racc = fmaf(e.y, x.y, racc);
      racc = fmaf(e.z, x.z, racc);
      racc = fmaf(e.w, x.w, racc);
      cacc[j * 4 + 0] = fmaf(e.x, vr, cacc[j * 4 + 0]);
      cacc[j * 4 + 1] = fmaf(e.y, vr, cacc[j * 4 + 1]);
      cacc[j * 4 + 2] = fmaf(e.z, vr, cacc[j * 4 + 2]);
      cacc[j * 4 + 3] = fmaf(e.w, vr, cacc[j * 4 + 3]);
    } else if (j == jd) {
      // elements right of the diagonal belong to the mirrored half; the diagonal counts once
      const int c = lane * 4 + 128 * j;
      if (c <= r) {
        const float4 e = a4[32 * j];
        const float4 x = x4[32 * j];
        const float a0 = e.x, a1 = c + 1 <= r ? e.y : 0.f, a2 = c + 2 <= r ? e.z : 0.f,
                    a3 = c + 3 <= r ? e.w : 0.f;
        racc = fmaf(a0, x.x, racc);
        racc = fmaf(a1, x.y, racc);
        racc = fmaf(a2, x.z, racc);
        racc = fmaf(a3, x.w, racc);
        cacc[j * 4 + 0] = fmaf(c + 0 < r ? a0 : 0.f, vr, cacc[j * 4 + 0]);
        cacc[j * 4 + 1] = fmaf(c + 1 < r ? a1 : 0.f, vr, cacc[j * 4 + 1]);
        cacc[j * 4 + 2] = fmaf(c + 2 < r ? a2 : 0.f, vr, cacc[j * 4 + 2]);
        cacc[j * 4 + 3] = fmaf(c + 3 < r ? a3 : 0.f, vr, cacc[j * 4 + 3]);
      }
    }
  }
  return racc;
}

// two block-wide sums at once, every thread gets both (fixed order).  `scratch` >= 64 floats;
// the caller keeps a __syncthreads between two uses.
__device__ __forceinline__ float2 pi_block_sum2(float a, float b, float* scratch, int nwarp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (lane == 0) {
    scratch[warp] = a;
    scratch[32 + warp] = b;
  }
  __syncthreads();
  a = lane < nwarp ? scratch[lane] : 0.f;
  b = lane < nwarp ? scratch[32 + lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  return make_float2(a, b);
}

// Every warp streams its own tasks through a private S-stage shared-memory ring filled by
// 1-D bulk copies (one elected lane issues them, an mbarrier per stage counts the bytes), so
// the bytes in flight per SM are set by the ring depth, not by registers.  The task stream
// runs on across the steps (the matrix does not change), so the copies also cover the
// reductions and the cluster exchange between two sweeps.
template <int NJ, int T, int S>
__global__ void __launch_bounds__(T, (T <= 256 && NJ <= 8) ? 2 : 1)
power_iteration_sym_kernel(const float* __restrict__ xs, const float* __restrict__ v0, int n,
                           int num_iters, float tol, RootCtl* ctl, float* __restrict__ ybuf,
                           int csize) {
  extern __shared__ __align__(16) float smem[];
  constexpr int kWarps = T / 32;
  float* v = smem;               // current iterate
  float* nv = smem + n;          // normalised iterate
  float* rowy = smem + 2 * n;    // row parts of this CTA's tasks
  float* colred = smem + 3 * n;  // [kWarps][n] column parts per warp
  const int stage_floats = n + 8;
  float* ring = colred + (size_t)kWarps * n;  // [kWarps][S][n + 8]
  __shared__ float scratch[64];
  __shared__ __align__(8) uint64_t bars[kWarps * S];
  const int b = blockIdx.x / csize;
  const int crank = csize > 1 ? (int)pi_cluster_rank() : 0;
  if (ctl[b].done != 0) return;  // uniform across the cluster
  const int pad = ctl[b].pad;
  const float* A = xs + (size_t)b * n * n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < S; ++q)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(pi_smem_u32(&bars[warp * S + q])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = threadIdx.x; i < n; i += T) v[i] = i < pad ? v0[i] : 0.f;
  __syncthreads();
  const int ntask = (pad + 1) / 2;
  // tasks of this warp: t = crank + csize * (warp + kWarps * k), k = 0 .. nk-1; rows (t, pad-1-t)
  const int first = crank + csize * warp, stride = csize * kWarps;
  const int nk = first < ntask ? (ntask - 1 - first) / stride + 1 : 0;
  float* my_ring = ring + (size_t)warp * S * stage_floats;
  const uint32_t my_bars = pi_smem_u32(&bars[warp * S]);
  // producer cursor (lane 0): sweep `p_it`, position `p_kk` in it, ring slot `p_slot`
  int p_it = 0, p_kk = 0, p_slot = 0;
  auto issue = [&]() {
    if (nk == 0 || p_it >= num_iters) return;
    const int k = (p_it & 1) ? nk - 1 - p_kk : p_kk;
    const int t = first + stride * k, r2 = pad - 1 - t;
    const uint32_t len2 = (uint32_t)(r2 + 4) & ~3u, len1 = r2 != t ? (uint32_t)(t + 4) & ~3u : 0u;
    const uint32_t bar = my_bars + 8u * p_slot;
    const uint32_t dst = pi_smem_u32(my_ring + (size_t)p_slot * stage_floats);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                 "r"((len1 + len2) * 4u)
                 : "memory");
    pi_bulk_load(dst, A + (size_t)r2 * n, len2 * 4u, bar);
    if (len1) pi_bulk_load(dst + len2 * 4u, A + (size_t)t * n, len1 * 4u, bar);
    if (++p_kk == nk) { p_kk = 0; ++p_it; }
    p_slot = p_slot + 1 == S ? 0 : p_slot + 1;
  };
  if (lane == 0)
    for (int q = 0; q < S; ++q) issue();
  int c_slot = 0;
  uint32_t c_phase = 0;
  float s = 0.f;
  int it = 0;
  bool run = true;
  float ss = 0.f;
  for (int i = threadIdx.x; i < n; i += T) ss += v[i] * v[i];
  ss = pi_block_sum2(ss, 0.f, scratch, kWarps).x;
  while (it < num_iters && run) {
    const float norm = sqrtf(ss);
    for (int i = threadIdx.x; i < n; i += T) {
      nv[i] = v[i] / norm;  // DS:634
      rowy[i] = 0.f;
    }
    __syncthreads();
    float cacc[NJ * 4];
#pragma unroll
    for (int q = 0; q < NJ * 4; ++q) cacc[q] = 0.f;
    for (int kk = 0; kk < nk; ++kk) {
      const int k = (it & 1) ? nk - 1 - kk : kk;
      const int t = first + stride * k;
      const int r2 = pad - 1 - t;
      const float* st = my_ring + (size_t)c_slot * stage_floats;
      pi_mbar_wait(my_bars + 8u * c_slot, c_phase);
      // the short row t <= (pad-1)/2 spans at most half of the column chunks
      float q2 = pi_sym_fma<NJ, NJ>(st, r2, lane, nv, cacc);
      float q1 = r2 != t ? pi_sym_fma<(NJ + 1) / 2, NJ>(st + ((r2 + 4) & ~3), t, lane, nv, cacc)
                         : 0.f;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        q1 += __shfl_xor_sync(0xffffffffu, q1, o);
        q2 += __shfl_xor_sync(0xffffffffu, q2, o);
      }
      // the shuffles also order every lane's reads of the slot before it is refilled
      if (lane == 0) {
        rowy[r2] = q2;
        if (r2 != t) rowy[t] = q1;
        issue();
      }
      if (++c_slot == S) { c_slot = 0; c_phase ^= 1u; }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = lane * 4 + 128 * j;
      if (c < n)
        *reinterpret_cast<float4*>(colred + (size_t)warp * n + c) =
            make_float4(cacc[j * 4], cacc[j * 4 + 1], cacc[j * 4 + 2], cacc[j * 4 + 3]);
    }
    __syncthreads();
    float* slot = ybuf + (((size_t)b * 2 + (it & 1)) * csize) * n;
    for (int i = threadIdx.x; i < n; i += T) {
      float y = rowy[i];
#pragma unroll
      for (int w = 0; w < kWarps; ++w) y += colred[(size_t)w * n + i];
      if (csize > 1) slot[(size_t)crank * n + i] = y;
      else v[i] = y;
    }
    if (csize > 1) pi_cluster_sync();
    else __syncthreads();
    float dot = 0.f, yy = 0.f;
    for (int i = threadIdx.x; i < n; i += T) {
      float yi;
      if (csize > 1) {
        yi = 0.f;
        for (int r = 0; r < csize; ++r) yi += __ldcg(slot + (size_t)r * n + i);
        v[i] = yi;  // DS:636
      } else {
        yi = v[i];
      }
      dot += nv[i] * yi;
      yy += yi * yi;
    }
    // v.y for this step and |y|^2 for the next normalisation in one reduction
    const float2 red = pi_block_sum2(dot, yy, scratch, kWarps);
    const float s_new = red.x;     // DS:637
    ss = red.y;
    run = fabsf(s_new - s) > tol;  // DS:639 (NaN -> stop)
    s = s_new;
    ++it;
  }
  // copies issued for sweeps that will not run must land before the CTA gives up its memory
  {
    const int issued = __shfl_sync(0xffffffffu, p_it * nk + p_kk, 0);
    for (int g = it * nk; g < issued; ++g) {
      pi_mbar_wait(my_bars + 8u * c_slot, c_phase);
      if (++c_slot == S) { c_slot = 0; c_phase ^= 1u; }
    }
  }
  if (threadIdx.x == 0 && crank == 0) ctl[b].max_ev = s;
}

// ---------------------------------------------------------------------------
// SIMT engine: one launch executes step `s` of every active matrix's program
// (+ the H update alongside step 0).
// ---------------------------------------------------------------------------
struct SquareView {
  const float* base;
  int n, lim;
  __device__ __forceinline__ float operator()(int i, int k) const {
    return (i < lim && k < lim) ? base[(size_t)i * n + k] : 0.f;
  }
};

__global__ void __launch_bounds__(kSimtThreads)
root_phase_simt_kernel(F32Store bufs, const RootCtl* __restrict__ ctl,
                       uint32_t* __restrict__ errbits, int n, int s) {
  __shared__ SimtSmem sm;
  const int b = blockIdx.z >> 1, lane_op = blockIdx.z & 1;
  const RootCtl& c = ctl[b];
  if (!c.active) return;
  Step st;
  if (lane_op == 1) {
    if (s != 0) return;
    st = c_hstep;
  } else {
    const Program& pr = c_programs[c.p];
    if (s >= pr.nsteps) return;
    st = pr.steps[s];
  }
  const int lim = c.pad;  // everything is zero outside [0,pad)^2
  // Only lower-triangular tiles (tile_m >= tile_n) are computed; the result is
  // mirrored so that every iterate is EXACTLY symmetric.  That makes reading B by
  // rows exact (B^T == B bitwise) -- with merely "nearly symmetric" iterates the
  // antisymmetric rounding noise doubles per Newton step instead of being squared
  // away -- and removes (T-1)/(2T) of the GEMM work.
  int tile_m, tile_n;
  tri_decode(blockIdx.x, tile_m, tile_n);
  const int cur = c.cur;
  float* out = bufs.mat(physical_buf(st.dst, cur), b, 0);
  float* out_mi = st.emit_mi ? bufs.mat(physical_buf(LB_MIN, cur), b, 0) : nullptr;
  if (tile_m * kSimtBM >= lim) {
    // tile fully in the padding: result (and its mirror) is exactly zero
    for (int e = threadIdx.x; e < kSimtBM * kSimtBN; e += blockDim.x) {
      int i = tile_m * kSimtBM + e / kSimtBN, j = tile_n * kSimtBN + e % kSimtBN;
      if (i < n && j < n) {
        out[(size_t)i * n + j] = 0.f;
        out[(size_t)j * n + i] = 0.f;
        if (out_mi) { out_mi[(size_t)i * n + j] = 0.f; out_mi[(size_t)j * n + i] = 0.f; }
      }
    }
    return;
  }
  const SquareView A{bufs.mat(physical_buf(st.a, cur), b, 0), n, lim};
  const SquareView B{bufs.mat(physical_buf(st.b, cur), b, 0), n, lim};
  const float alpha = -1.0f / (float)c.p, one_minus_alpha = 1.0f - alpha;
  const bool diag_tile = tile_m == tile_n;
  uint32_t emax = 0;
  simt_gemm_tile(lim, tile_m, tile_n, A, B, true, true, sm,
                 [&](int i, int j0, const float* acc) {
                   if (i >= n) return;
#pragma unroll
                   for (int q = 0; q < 4; ++q) {
                     const int j = j0 + q;
                     if (j >= n || (diag_tile && j > i)) continue;
                     const bool in = (i < lim && j < lim);
                     const float v = in ? acc[q] : 0.f;
                     out[(size_t)i * n + j] = v;
                     if (i != j) out[(size_t)j * n + i] = v;
                     if (out_mi) {
                       const float mi = in ? mi_from_m(v, i == j, alpha, one_minus_alpha) : 0.f;
                       out_mi[(size_t)i * n + j] = mi;
                       if (i != j) out_mi[(size_t)j * n + i] = mi;
                       const uint32_t ab = absbits(v - ((in && i == j) ? 1.f : 0.f));
                       emax = ab > emax ? ab : emax;
                     }
                   }
                 });
  if (st.emit_mi) {
    emax = warp_max_u32(emax);
    if ((threadIdx.x & 31) == 0 && emax) atomicMax(errbits + b, emax);
  }
}

// ---------------------------------------------------------------------------
// per-iteration control (DS:836-848 bookkeeping, DS:876-885 retry logic)
// ---------------------------------------------------------------------------
// One block walks the whole batch and publishes {unfinished, waiting for (re)initialisation}
// straight into pinned host memory (UVA): no memset / copy launches around it.
__global__ void __launch_bounds__(1024)
root_control_kernel(RootCtl* ctl, uint32_t* errbits, int batch, RootParams prm,
                    int* host_slot) {
  __shared__ int counts[2];
  if (threadIdx.x < 2) counts[threadIdx.x] = 0;
  __syncthreads();
  int unfinished = 0, waiting = 0;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    RootCtl c = ctl[b];
    if (c.active) {
      const float new_err = __uint_as_float(errbits[b]);  // max|M' - I_m|, DS:847
      errbits[b] = 0u;
      c.ratio = new_err / c.err;                // DS:848
      c.err = new_err;
      c.iter += 1;
      c.cur ^= 1;
      root_after_error_update(c, prm);
      ctl[b] = c;
    }
    unfinished += c.done ? 0 : 1;
    waiting += c.need_init ? 1 : 0;
  }
  if (unfinished) atomicAdd(&counts[0], unfinished);
  if (waiting) atomicAdd(&counts[1], waiting);
  __syncthreads();
  if (threadIdx.x == 0) {
    host_slot[0] = counts[0];
    host_slot[1] = counts[1];
    __threadfence_system();
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
// numpy.random.RandomState(1729).uniform(-1, 1, n) -- MT19937, 53-bit doubles
// (DS:642-643).  Generated on the host so the start vector is bit-identical.
static void mt19937_uniform(uint32_t seed, int n, float* out) {
  uint32_t mt[624];
  mt[0] = seed;
  for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i;
  int idx = 624;
  auto next = [&]() -> uint32_t {
    if (idx >= 624) {
      for (int k = 0; k < 624; ++k) {
        uint32_t yv = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (yv >> 1) ^ ((yv & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t yv = mt[idx++];
    yv ^= yv >> 11;
    yv ^= (yv << 7) & 0x9d2c5680u;
    yv ^= (yv << 15) & 0xefc60000u;
    yv ^= yv >> 18;
    return yv;
  };
  for (int i = 0; i < n; ++i) {
    const uint32_t a = next() >> 5, bq = next() >> 6;
    const double r = (a * 67108864.0 + bq) / 9007199254740992.0;  // [0,1)
    out[i] = (float)(-1.0 + 2.0 * r);  // uniform(low, high) = low + (high-low)*r
  }
}

struct RootWorkspace {
  RootCtl* ctl;
  uint32_t* errbits;
  int* unfinished;  // [1]
  float* v0;        // [n]
  float* ybuf;      // [batch, 2, n] power-iteration exchange vectors
  char* engine_mem;
  size_t engine_bytes;
};

static size_t engine_bytes_simt(int batch, int n) {
  return (size_t)kNumBufs * batch * n * n * sizeof(float);
}

static int pick_cluster_size(int batch) {
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();  // workspace queries also run without a device
    sms = 148;
  }
  int c = 8;
  while (c > 1 && (long long)batch * c > sms) c >>= 1;
  return c;
}

// power-iteration exchange vectors: [batch, 2, cluster size, n] (the strip kernels of the
// first initialisation reuse the space)
static size_t pi_exchange_bytes(int batch, int n) {
  // (the symmetric kernel picks any cluster size up to 8 with batch * csize <= 2 * SMs)
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) {
    cudaGetLastError();
    sms = 148;
  }
  const int cmax = std::max(pick_cluster_size(batch), std::max(1, std::min(8, 2 * sms / std::max(batch, 1))));
  return sizeof(float) * 2 * (size_t)batch * cmax * (n < 4 ? 4 : n);
}

static size_t header_bytes(int batch, int n) {
  size_t s = 0;
  s += align_up(sizeof(RootCtl) * batch, 256);
  s += align_up(sizeof(uint32_t) * batch, 256);
  s += 256;
  s += align_up(sizeof(float) * n, 256);
  s += align_up(pi_exchange_bytes(batch, n), 256);
  return s;
}

static int resolve_engine(int engine, int n) {
  if (engine == PC_ENGINE_AUTO) {
    // n = 128 stays on fp32 FFMA: with so short a K the fp32 reference loses little to
    // accumulation and the 22-bit operands show (residual 2.06x the reference's, measured)
    if (tc_engine_available() && n >= 256 && n % 128 == 0) return PC_ENGINE_TC_FP16X3;
    return PC_ENGINE_SIMT_FP32;
  }
  return engine;
}

size_t root_workspace_bytes(int batch, int n, int engine) {
  if (engine == PC_ENGINE_TC_SMALL) return small_root_workspace_bytes(batch, n) + 1024;
  const bool maybe_small = engine == PC_ENGINE_AUTO && n <= 128 && tc_engine_available();
  engine = resolve_engine(engine, n);
  size_t e = engine == PC_ENGINE_SIMT_FP32 ? engine_bytes_simt(batch, n)
                                           : tc_engine_bytes(batch, n, engine == PC_ENGINE_TC_FP16X3 ? 2 : 3);
  size_t total = header_bytes(batch, n) + align_up(e, 256) + 1024;
  // PC_ENGINE_AUTO may pick the persistent small-block solver (per-CTA H slots)
  if (maybe_small) total = std::max(total, small_root_workspace_bytes(batch, n) + 1024);
  return total;
}


// numpy's start vector per (device, size), uploaded ONCE into device memory that is never
// freed, and the kernel attributes: everything that must not happen inside a stream capture.
// (A per-call upload, however small, queues on the host-to-device copy engine behind whatever
// the application is uploading on other streams -- a 310 MB input copy of the next step held
// every solve back by its full 5.6 ms.)
static std::mutex v0_mu;
static std::map<std::pair<int, int>, float*> v0_cache;
// v, v/|v|, row parts, column parts per warp, and the per-warp rings of (n + 8)-float slots
static size_t pi_sym_smem_bytes(int n, int warps, int stages) {
  return sizeof(float) * ((size_t)n * (3 + warps) + (size_t)warps * stages * (n + 8));
}

int prepare_power_iteration(int n) {
  std::lock_guard<std::mutex> lock(v0_mu);
  int dev = 0;
  PC_CUDA_CHECK(cudaGetDevice(&dev));
  if (v0_cache.find({dev, n}) != v0_cache.end()) return PC_OK;
  std::vector<float> host((size_t)n);
  mt19937_uniform(1729u, n, host.data());
  float* v0 = nullptr;
  PC_CUDA_CHECK(cudaMalloc(&v0, sizeof(float) * n));
  PC_CUDA_CHECK(cudaMemcpy(v0, host.data(), sizeof(float) * n, cudaMemcpyHostToDevice));
  v0_cache[{dev, n}] = v0;
  const size_t smem = sizeof(float) * 3 * (size_t)n;
  if (smem > 48 * 1024)
    PC_CUDA_CHECK(cudaFuncSetAttribute(power_iteration_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  static bool sym_configured = false;
  if (!sym_configured) {
    const cudaFuncAttribute kMax = cudaFuncAttributeMaxDynamicSharedMemorySize;
    PC_CUDA_CHECK(cudaFuncSetAttribute(power_iteration_sym_kernel<2, 512, 4>, kMax,
                                       (int)pi_sym_smem_bytes(256, 16, 4)));
    PC_CUDA_CHECK(cudaFuncSetAttribute(power_iteration_sym_kernel<4, 512, 4>, kMax,
                                       (int)pi_sym_smem_bytes(512, 16, 4)));
    PC_CUDA_CHECK(cudaFuncSetAttribute(power_iteration_sym_kernel<8, 512, 2>, kMax,
                                       (int)pi_sym_smem_bytes(1024, 16, 2)));
    PC_CUDA_CHECK(cudaFuncSetAttribute(power_iteration_sym_kernel<16, 256, 2>, kMax,
                                       (int)pi_sym_smem_bytes(2048, 8, 2)));
    PC_CUDA_CHECK(cudaFuncSetAttribute(power_iteration_sym_kernel<2, 256, 2>, kMax,
                                       (int)pi_sym_smem_bytes(256, 8, 2)));
    PC_CUDA_CHECK(cudaFuncSetAttribute(power_iteration_sym_kernel<4, 256, 2>, kMax,
                                       (int)pi_sym_smem_bytes(512, 8, 2)));
    PC_CUDA_CHECK(cudaFuncSetAttribute(power_iteration_sym_kernel<8, 256, 2>, kMax,
                                       (int)pi_sym_smem_bytes(1024, 8, 2)));
    sym_configured = true;
  }
  return PC_OK;
}

// device pointer of the start vector on the current device (after prepare_power_iteration)
static const float* power_iteration_v0_device(int n) {
  std::lock_guard<std::mutex> lock(v0_mu);
  int dev = 0;
  cudaGetDevice(&dev);
  auto it = v0_cache.find({dev, n});
  return it == v0_cache.end() ? nullptr : it->second;
}

int run_power_iteration(const float* xs, const int32_t* pads, int batch, int n,
                        int num_iters, float tol, float* lambdas, int32_t* iters,
                        RootCtl* ctl, float* v0_dev, float* ybuf, cudaStream_t stream) {
  int rc = prepare_power_iteration(n);
  if (rc != PC_OK) return rc;
  (void)v0_dev;  // the start vector is resident on the device
  const float* v0 = power_iteration_v0_device(n);
  const size_t smem = sizeof(float) * 3 * (size_t)n;  // v, v / |v|, and A v when csize == 1
  int csize = n >= 256 ? pick_cluster_size(batch) : 1;
  const int threads = n >= 512 ? 1024 : (n >= 128 ? 512 : 128);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(batch * csize));
  cfg.blockDim = dim3((unsigned)threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, power_iteration_kernel, xs, pads, v0,
                                   n, num_iters, tol, lambdas, iters, ctl, ybuf, csize));
  return PC_OK;
}

// symmetric-half sweep (solver path); false when the shape is outside its range
static bool pi_sym_supported(int n) {
  static const bool off = [] {
    const char* e = getenv("PC_PI_SYM");
    return e && e[0] == '0';
  }();
  return !off && n >= 256 && n <= 2048 && n % 4 == 0;
}

template <int NJ, int T, int S>
static int launch_pi_sym(const float* xs, int batch, int n, RootCtl* ctl, const float* v0_dev,
                         float* ybuf, int csize, cudaStream_t stream) {
  const size_t smem = pi_sym_smem_bytes(n, T / 32, S);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(batch * csize));
  cfg.blockDim = dim3((unsigned)T);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, power_iteration_sym_kernel<NJ, T, S>, xs, v0_dev, n,
                                   100, 1e-6f, ctl, ybuf, csize));  // DS:820-825
  return PC_OK;
}

static int run_power_iteration_sym(const float* xs, int batch, int n, RootCtl* ctl, float* v0_dev,
                                   float* ybuf, cudaStream_t stream) {
  int rc = prepare_power_iteration(n);
  if (rc != PC_OK) return rc;
  (void)v0_dev;
  const float* v0 = power_iteration_v0_device(n);
  // Two shapes of the same kernel: (A) 16 warps, one CTA per SM; (B) 8 warps and half the
  // ring, two CTAs per SM.  Clusters of any size up to 8 share a matrix; the shape whose grid
  // puts more warps on the machine wins (A on a tie: fewer partial vectors to exchange).  E.g.
  // 74 statistics: A with clusters of 2 (148 CTAs); 79: B with clusters of 3 (237 CTAs on 296
  // slots) instead of 79 CTAs on 148 SMs; 40: B with clusters of 7.
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  static const bool only_a = [] {
    const char* e = getenv("PC_PI_SHAPE");
    return e && e[0] == 'A';
  }();
  const int ca = std::max(1, std::min(8, sms / std::max(batch, 1)));
  const int cb = std::max(1, std::min(8, 2 * sms / std::max(batch, 1)));
  const bool shape_b = !only_a && n <= 1024 && (long long)batch * cb * 8 > (long long)batch * ca * 16;
  if (shape_b) {
    if (n <= 256) return launch_pi_sym<2, 256, 2>(xs, batch, n, ctl, v0, ybuf, cb, stream);
    if (n <= 512) return launch_pi_sym<4, 256, 2>(xs, batch, n, ctl, v0, ybuf, cb, stream);
    return launch_pi_sym<8, 256, 2>(xs, batch, n, ctl, v0, ybuf, cb, stream);
  }
  const int csize = only_a ? pick_cluster_size(batch) : ca;
  if (n <= 256) return launch_pi_sym<2, 512, 4>(xs, batch, n, ctl, v0, ybuf, csize, stream);
  if (n <= 512) return launch_pi_sym<4, 512, 4>(xs, batch, n, ctl, v0, ybuf, csize, stream);
  if (n <= 1024) return launch_pi_sym<8, 512, 2>(xs, batch, n, ctl, v0, ybuf, csize, stream);
  return launch_pi_sym<16, 256, 2>(xs, batch, n, ctl, v0, ybuf, csize, stream);
}

// ---------------------------------------------------------------------------
// One solver call = pre (setup, power iteration, first initialisation) -> Newton
// iterations until every matrix is done -> final gather.  Two drivers share the pieces:
//   * graph mode (default): the call is ONE CUDA graph whose middle is a conditional WHILE
//     node; the control kernel sets the loop condition on the device
//     (cudaGraphSetConditional), so the host enqueues and returns -- no event polling, no
//     host round trip per iteration, and independent calls on different streams overlap.
//     Executable graphs are cached per argument set (a training loop re-launches them).
//   * host-polled mode (PC_ROOT_MODE=poll, GEMM timing runs, PC_TC_SYNC=1): the host launches
//     iteration by iteration and reads the "unfinished" counter two iterations behind.
// ---------------------------------------------------------------------------
struct RootCall {
  const float* xs; const int32_t* ps; const int32_t* pads;
  int batch, n, engine, max_steps;
  float* roots; float* metrics;
  RootWorkspace ws;
  RootParams prm;
  F32Store f32;
  TcEngine tc;
};

static int max_program_steps(const int32_t* ps_host, int batch) {
  int max_steps = 1;
  if (!ps_host) {
    for (int p = 1; p <= kMaxP; ++p) max_steps = std::max(max_steps, h_programs[p].nsteps);
    return max_steps;
  }
  for (int b = 0; b < batch; ++b)
    if (ps_host[b] >= 1 && ps_host[b] <= kMaxP)
      max_steps = std::max(max_steps, h_programs[ps_host[b]].nsteps);
  return max_steps;
}

static int root_enqueue_pre(RootCall& c, cudaStream_t stream) {
  root_setup_kernel<<<(c.batch + 127) / 128, 128, 0, stream>>>(c.ws.ctl, c.ps, c.pads, c.batch,
                                                              c.n, c.ws.errbits);
  PC_CUDA_CHECK(cudaGetLastError());
  int launches = 1;
  if (c.prm.relative_eps && c.n > 1) {
    int rc = pi_sym_supported(c.n)
                 ? run_power_iteration_sym(c.xs, c.batch, c.n, c.ws.ctl, c.ws.v0, c.ws.ybuf, stream)
                 : run_power_iteration(c.xs, nullptr, c.batch, c.n, 100, 1e-6f, nullptr, nullptr,
                                       c.ws.ctl, c.ws.v0, c.ws.ybuf, stream);  // DS:820-825
    if (rc != PC_OK) return rc;
    ++launches;
  }
  if (c.engine == PC_ENGINE_SIMT_FP32) {
    for (int k = 0; k < kNumBufs; ++k)
      c.f32.base[k] = reinterpret_cast<float*>(c.ws.engine_mem) + (size_t)k * c.batch * c.n * c.n;
    c.f32.mat_elems = (size_t)c.n * c.n;
    if (c.n >= 256) {  // first initialisation: whole-GPU strip kernels
      const int strips = (c.n + 31) / 32;
      root_norm_strip_kernel<<<dim3(strips, c.batch), 256, 0, stream>>>(
          c.xs, c.ws.ctl, c.n, strips, c.prm, c.ws.ybuf, c.batch);
      root_init_strip_kernel<F32Store><<<dim3(strips, c.batch), 1024, 0, stream>>>(
          c.xs, c.ws.ctl, c.f32, c.batch, c.n, strips, c.prm, c.ws.ybuf);
      launches += 2;
    }
  } else {
    int rc = tc_engine_init(&c.tc, c.ws.engine_mem, c.batch, c.n,
                            c.engine == PC_ENGINE_TC_BF16X6 ? 6 : 3,
                            c.engine == PC_ENGINE_TC_FP16X3 ? 1 : 0, stream);
    if (rc != PC_OK) return rc;
    // max_steps = 0: only the first (strip) initialisation
    rc = tc_engine_iteration(&c.tc, c.xs, c.ws.ctl, c.ws.errbits, c.prm, c.roots, 0, c.ws.ybuf,
                             false, stream);
    if (rc != PC_OK) return rc;
  }
  count_launch(launches);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

// one Newton iteration of every active matrix (+ the (re)initialisation of whoever waits)
static int root_enqueue_iteration(RootCall& c, bool do_init, cudaStream_t stream) {
  if (c.engine == PC_ENGINE_SIMT_FP32) {
    if (do_init) {
      root_init_kernel<F32Store><<<c.batch, c.n >= 512 ? 1024 : 256, 0, stream>>>(
          c.xs, c.ws.ctl, c.f32, c.batch, c.n, c.prm, c.roots);
      count_launch(1);
    }
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    if (gemm_timing_enabled()) {
      cudaEventCreate(&ev0); cudaEventCreate(&ev1);
      cudaEventRecord(ev0, stream);
    }
    const int tiles = (c.n + kSimtBM - 1) / kSimtBM;
    for (int s = 0; s < c.max_steps; ++s) {
      dim3 grid(tiles * (tiles + 1) / 2, 1, c.batch * 2);
      root_phase_simt_kernel<<<grid, kSimtThreads, 0, stream>>>(c.f32, c.ws.ctl, c.ws.errbits,
                                                               c.n, s);
    }
    if (ev0) { cudaEventRecord(ev1, stream); gemm_timing_record(ev0, ev1); }
    count_launch(c.max_steps);
    gemm_count(c.max_steps);
    PC_CUDA_CHECK(cudaGetLastError());
    return PC_OK;
  }
  return tc_engine_iteration(&c.tc, c.xs, c.ws.ctl, c.ws.errbits, c.prm, c.roots, c.max_steps,
                             nullptr, do_init, stream);
}

static int root_enqueue_final(RootCall& c, cudaStream_t stream) {
  if (c.engine == PC_ENGINE_SIMT_FP32) {
    dim3 fgrid((unsigned)std::min<size_t>(((size_t)c.n * c.n + 255) / 256, 64), c.batch);
    root_final_kernel<F32Store><<<fgrid, 256, 0, stream>>>(c.ws.ctl, c.f32, c.n, c.roots,
                                                          c.metrics);
  } else {
    int rc = tc_engine_final(&c.tc, c.ws.ctl, c.roots, c.metrics, stream);
    if (rc != PC_OK) return rc;
  }
  count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

// ---- graph mode ------------------------------------------------------------
struct RootGraphKey {
  const void* xs; const void* ps; const void* pads; const void* roots; const void* metrics;
  const void* workspace;
  int batch, n, engine, max_steps, num_iters, relative_eps, device, variant;
  float ridge, tol;
  bool operator<(const RootGraphKey& o) const { return memcmp(this, &o, sizeof(*this)) < 0; }
};
struct RootGraphEntry {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  int nodes = 0;
  uint64_t last_use = 0;
};
static std::mutex g_graph_mu;
static std::map<RootGraphKey, RootGraphEntry> g_graph_cache;
static uint64_t g_graph_clock = 0;
constexpr size_t kGraphCacheMax = 64;

static bool root_graph_mode_allowed() {
  const char* m = getenv("PC_ROOT_MODE");
  if (m && (m[0] == 'p' || m[0] == 'P')) return false;  // "poll"
  const char* gs = getenv("PC_TC_SYNC");  // the group rendezvous alternates counters per launch
  if (gs && gs[0] == '1') return false;
  return !gemm_timing_enabled();
}

// Device-side convergence check of a captured iteration: same bookkeeping as
// root_control_kernel, the loop condition goes to the graph's WHILE node.
__global__ void __launch_bounds__(1024)
root_control_graph_kernel(RootCtl* ctl, uint32_t* errbits, int batch, RootParams prm,
                          cudaGraphConditionalHandle handle) {
  __shared__ int unfinished_total;
  if (threadIdx.x == 0) unfinished_total = 0;
  __syncthreads();
  int unfinished = 0;
  for (int b = threadIdx.x; b < batch; b += blockDim.x) {
    RootCtl c = ctl[b];
    if (c.active) {
      const float new_err = __uint_as_float(errbits[b]);  // max|M' - I_m|, DS:847
      errbits[b] = 0u;
      c.ratio = new_err / c.err;                // DS:848
      c.err = new_err;
      c.iter += 1;
      c.cur ^= 1;
      root_after_error_update(c, prm);
      ctl[b] = c;
    }
    unfinished += c.done ? 0 : 1;
  }
  if (unfinished) atomicAdd(&unfinished_total, unfinished);
  __syncthreads();
  if (threadIdx.x == 0) cudaGraphSetConditional(handle, unfinished_total > 0 ? 1u : 0u);
}

static int root_build_graph(RootCall& c, RootGraphEntry* out) {
  // private capture stream of this host thread (per device)
  static thread_local cudaStream_t capture_streams[64] = {nullptr};
  int dev = 0;
  PC_CUDA_CHECK(cudaGetDevice(&dev));
  cudaStream_t& cs = capture_streams[dev & 63];
  if (!cs) PC_CUDA_CHECK(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
  cudaGraph_t g = nullptr;
  PC_CUDA_CHECK(cudaGraphCreate(&g, 0));
  auto fail = [&](int rc) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(cs, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) {
      cudaGraph_t dummy = nullptr;
      cudaStreamEndCapture(cs, &dummy);
    }
    cudaGraphDestroy(g);
    cudaGetLastError();
    return rc;
  };
#define PC_G(expr)                                                                         \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return fail(PC_ERR_CUDA);                                                            \
    }                                                                                      \
  } while (0)
  // (1) pre
  PC_G(cudaStreamBeginCaptureToGraph(cs, g, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  int rc = root_enqueue_pre(c, cs);
  if (rc != PC_OK) return fail(rc);
  cudaStreamCaptureStatus status;
  const cudaGraphNode_t* leaves = nullptr;
  size_t nleaves = 0;
  PC_G(cudaStreamGetCaptureInfo(cs, &status, nullptr, nullptr, &leaves, &nleaves));
  std::vector<cudaGraphNode_t> deps(leaves, leaves + nleaves);
  cudaGraph_t same = nullptr;
  PC_G(cudaStreamEndCapture(cs, &same));
  // (2) WHILE node: iterate while any matrix is unfinished
  cudaGraphConditionalHandle handle;
  PC_G(cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault));
  cudaGraphNodeParams cp = {};
  cp.type = cudaGraphNodeTypeConditional;
  cp.conditional.handle = handle;
  cp.conditional.type = cudaGraphCondTypeWhile;
  cp.conditional.size = 1;
  cudaGraphNode_t while_node = nullptr;
  PC_G(cudaGraphAddNode(&while_node, g, deps.data(), deps.size(), &cp));
  cudaGraph_t body = cp.conditional.phGraph_out[0];
  PC_G(cudaStreamBeginCaptureToGraph(cs, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
  rc = root_enqueue_iteration(c, true, cs);
  if (rc != PC_OK) return fail(rc);
  root_control_graph_kernel<<<1, 1024, 0, cs>>>(c.ws.ctl, c.ws.errbits, c.batch, c.prm, handle);
  PC_G(cudaGetLastError());
  PC_G(cudaStreamEndCapture(cs, &same));
  // (3) final gather
  PC_G(cudaStreamBeginCaptureToGraph(cs, g, &while_node, nullptr, 1, cudaStreamCaptureModeRelaxed));
  rc = root_enqueue_final(c, cs);
  if (rc != PC_OK) return fail(rc);
  PC_G(cudaStreamEndCapture(cs, &same));
  cudaGraphExec_t exec = nullptr;
  PC_G(cudaGraphInstantiate(&exec, g, 0));
#undef PC_G
  size_t nn = 0, nb = 0;
  cudaGraphGetNodes(g, nullptr, &nn);
  cudaGraphGetNodes(body, nullptr, &nb);
  out->graph = g;
  out->exec = exec;
  out->nodes = (int)(nn + nb);
  return PC_OK;
}

static int run_root_graph(RootCall& c, const RootGraphKey& key, cudaStream_t stream) {
  std::lock_guard<std::mutex> lock(g_graph_mu);  // graph construction / launch order per device
  auto it = g_graph_cache.find(key);
  if (it == g_graph_cache.end()) {
    RootGraphEntry e;
    int rc = root_build_graph(c, &e);
    if (rc != PC_OK) return rc;
    if (g_graph_cache.size() >= kGraphCacheMax) {  // evict the least recently used entry
      auto victim = g_graph_cache.begin();
      for (auto jt = g_graph_cache.begin(); jt != g_graph_cache.end(); ++jt)
        if (jt->second.last_use < victim->second.last_use) victim = jt;
      cudaGraphExecDestroy(victim->second.exec);
      cudaGraphDestroy(victim->second.graph);
      g_graph_cache.erase(victim);
    }
    it = g_graph_cache.emplace(key, e).first;
  } else {
    count_launch(it->second.nodes);  // (building counted its launches while capturing)
  }
  it->second.last_use = ++g_graph_clock;
  PC_CUDA_CHECK(cudaGraphLaunch(it->second.exec, stream));
  return PC_OK;
}

int run_root(const float* xs, const int32_t* ps, const int32_t* ps_host, const int32_t* pads,
             int batch, int n, const pc_root_options* opt, float* roots, float* metrics,
             void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  // Small statistics (n <= 128, exponents 2^s known on the host): the whole solve of a matrix
  // runs inside one persistent CTA on the tensor cores -- one launch, nothing to poll.
  {
    const char* sm = getenv("PC_SMALL_ROOT");
    const bool allowed = !(sm && sm[0] == '0');
    const bool eligible = n <= 128 && tc_engine_available() &&
                          small_root_supported_exponents(ps_host, batch);
    if (opt->engine == PC_ENGINE_TC_SMALL) {
      PC_REQUIRE(eligible, "PC_ENGINE_TC_SMALL needs sm_100, n <= 128 and host exponents in "
                           "{1, 2, 4, 8, 16} (pc_inverse_pth_root_enqueue with ps_host)");
    }
    if (opt->engine == PC_ENGINE_TC_SMALL || (opt->engine == PC_ENGINE_AUTO && eligible && allowed)) {
      int rc0 = prepare_power_iteration(n);
      if (rc0 != PC_OK) return rc0;
      if (workspace_bytes < small_root_workspace_bytes(batch, n)) {
        set_error("workspace too small: %zu < %zu", workspace_bytes,
                  small_root_workspace_bytes(batch, n));
        return PC_ERR_WORKSPACE;
      }
      return run_small_root(xs, ps, pads, batch, n, opt, power_iteration_v0_device(n), roots,
                            metrics, workspace, workspace_bytes, stream);
    }
  }
  const int engine = resolve_engine(opt->engine, n);
  PC_REQUIRE(engine == PC_ENGINE_SIMT_FP32 || engine == PC_ENGINE_TC_BF16X6 ||
                 engine == PC_ENGINE_TC_BF16X3 || engine == PC_ENGINE_TC_FP16X3,
             "unknown engine %d", opt->engine);
  if (engine != PC_ENGINE_SIMT_FP32) {
    PC_REQUIRE(n % 128 == 0 && n >= 128, "tcgen05 engine needs n %% 128 == 0 (n=%d)", n);
    if (!tc_engine_available()) {
      set_error("tcgen05 engine requested but device is not sm_100");
      return PC_ERR_UNSUPPORTED;
    }
  }
  if (workspace_bytes < root_workspace_bytes(batch, n, opt->engine)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes,
              root_workspace_bytes(batch, n, opt->engine));
    return PC_ERR_WORKSPACE;
  }
  int rc = ensure_programs();
  if (rc != PC_OK) return rc;
  if (ps_host)
    for (int b = 0; b < batch; ++b)
      PC_REQUIRE(ps_host[b] >= 1 && ps_host[b] <= kMaxP,
                 "exponent p = %d of matrix %d is outside [1, %d]", ps_host[b], b, kMaxP);

  RootCall c;
  c.xs = xs; c.ps = ps; c.pads = pads; c.batch = batch; c.n = n; c.engine = engine;
  c.roots = roots; c.metrics = metrics;
  c.prm = RootParams{opt->ridge_epsilon, opt->error_tolerance, opt->num_iters,
                     opt->relative_matrix_epsilon};
  // carve the workspace
  char* w = reinterpret_cast<char*>(align_up((size_t)workspace, 256));
  c.ws.ctl = reinterpret_cast<RootCtl*>(w); w += align_up(sizeof(RootCtl) * batch, 256);
  c.ws.errbits = reinterpret_cast<uint32_t*>(w); w += align_up(sizeof(uint32_t) * batch, 256);
  c.ws.unfinished = reinterpret_cast<int*>(w); w += 256;
  c.ws.v0 = reinterpret_cast<float*>(w); w += align_up(sizeof(float) * n, 256);
  c.ws.ybuf = reinterpret_cast<float*>(w); w += align_up(pi_exchange_bytes(batch, n), 256);
  c.ws.engine_mem = w;

  // everything that may not happen inside a stream capture: one-time allocations, symbol
  // uploads, kernel attributes
  rc = prepare_power_iteration(n);
  if (rc != PC_OK) return rc;
  if (engine != PC_ENGINE_SIMT_FP32) {
    rc = tc_engine_prepare();
    if (rc != PC_OK) return rc;
  }

  if (root_graph_mode_allowed()) {
    c.max_steps = max_program_steps(ps_host, batch);
    RootGraphKey key;
    memset(&key, 0, sizeof(key));
    key.xs = xs; key.ps = ps; key.pads = pads; key.roots = roots; key.metrics = metrics;
    key.workspace = workspace;
    key.batch = batch; key.n = n; key.engine = engine; key.max_steps = c.max_steps;
    key.num_iters = opt->num_iters; key.relative_eps = opt->relative_matrix_epsilon;
    key.ridge = opt->ridge_epsilon; key.tol = opt->error_tolerance;
    cudaGetDevice(&key.device);
    key.variant = engine == PC_ENGINE_SIMT_FP32 ? 0 : tc_engine_variant_mask();
    return run_root_graph(c, key, stream);
  }

  // ---- host-polled mode ----
  // Per-host-thread pinned scratch: exponents (decide how many GEMM launches one Newton
  // iteration needs) and a ring of poll slots for the device-side "unfinished" counter.
  constexpr int kPollRing = 4, kPollLag = 2;
  struct HostScratch {
    int32_t* ps = nullptr;
    int ps_cap = 0;
    int* poll = nullptr;
    cudaEvent_t ev[kPollRing] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ps_ev = nullptr;
  };
  static thread_local HostScratch hsx;
  if (!hsx.poll) {
    PC_CUDA_CHECK(cudaMallocHost(&hsx.poll, sizeof(int) * 2 * kPollRing));
    for (int i = 0; i < kPollRing; ++i)
      PC_CUDA_CHECK(cudaEventCreateWithFlags(&hsx.ev[i], cudaEventDisableTiming));
    PC_CUDA_CHECK(cudaEventCreateWithFlags(&hsx.ps_ev, cudaEventDisableTiming));
  }
  if (hsx.ps_cap < batch) {
    if (hsx.ps) cudaFreeHost(hsx.ps);
    PC_CUDA_CHECK(cudaMallocHost(&hsx.ps, sizeof(int32_t) * batch));
    hsx.ps_cap = batch;
  }
  int32_t* hps = hsx.ps;
  // the copy is queued BEFORE the power iteration and waited for after it has been
  // launched, so the host round trip hides behind that kernel
  if (ps_host) {
    memcpy(hps, ps_host, sizeof(int32_t) * batch);
  } else {
    PC_CUDA_CHECK(cudaMemcpyAsync(hps, ps, sizeof(int32_t) * batch, cudaMemcpyDeviceToHost, stream));
    PC_CUDA_CHECK(cudaEventRecord(hsx.ps_ev, stream));
  }
  rc = root_enqueue_pre(c, stream);
  if (rc != PC_OK) return rc;
  if (!ps_host) PC_CUDA_CHECK(cudaEventSynchronize(hsx.ps_ev));
  c.max_steps = max_program_steps(hps, batch);

  // Convergence is polled kPollLag iterations behind the launches: the host never waits
  // for the iteration it has just enqueued, so the device queue stays non-empty.  The
  // (at most kPollLag) surplus iterations find no active matrix and return at once.
  // 6 tries x num_iters iterations, plus the iterations a retry may idle before the lagged
  // poll notices it and launches its (re)initialisation
  const int max_total = opt->num_iters * 6 + 8 + 6 * (kPollLag + 1);
  // (Re)initialisation is launched on the first iteration and afterwards only when a lagged
  // poll reports matrices waiting for a retry (they idle until then; retries are rare).
  bool init_pending = true;
  for (int it = 0; it < max_total + kPollLag; ++it) {
    if (it >= kPollLag) {
      const int slot = (it - kPollLag) % kPollRing;
      cudaError_t e = cudaEventSynchronize(hsx.ev[slot]);
      if (e != cudaSuccess) {
        set_error("root iteration failed: %s", cudaGetErrorString(e));
        return PC_ERR_CUDA;
      }
      if (hsx.poll[2 * slot] == 0) break;
      if (hsx.poll[2 * slot + 1] > 0) init_pending = true;
    }
    if (it >= max_total) continue;
    const bool do_init = init_pending;
    init_pending = false;
    rc = root_enqueue_iteration(c, do_init, stream);
    if (rc != PC_OK) return rc;
    count_launch(1);
    const int slot = it % kPollRing;
    root_control_kernel<<<1, 1024, 0, stream>>>(c.ws.ctl, c.ws.errbits, batch, c.prm,
                                                hsx.poll + 2 * slot);
    cudaEventRecord(hsx.ev[slot], stream);
  }
  rc = root_enqueue_final(c, stream);
  if (rc != PC_OK) return rc;
  // algorithmic GEMM flops actually needed: iterations x G(p) x 2 n^3
  if (gemm_timing_enabled()) {
    std::vector<float> hm((size_t)batch * PC_NUM_METRICS);
    cudaMemcpyAsync(hm.data(), metrics, hm.size() * sizeof(float), cudaMemcpyDeviceToHost,
                    stream);
    cudaStreamSynchronize(stream);
    double fl = 0.0;
    for (int b = 0; b < batch; ++b) {
      const int p = hps[b];
      if (p < 1 || p > kMaxP) continue;
      const double gp = h_programs[p].nsteps + 1;  // chain steps + the H update
      fl += (double)hm[(size_t)b * PC_NUM_METRICS + PC_METRIC_ITERS] * gp * 2.0 * n * (double)n * n;
    }
    gemm_add_flops(fl);
  }
  return PC_OK;
}

}  // namespace pc

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

void pc_root_options_default(pc_root_options* opt) {
  opt->ridge_epsilon = 1e-6f;
  opt->error_tolerance = 1e-6f;
  opt->num_iters = 100;
  opt->relative_matrix_epsilon = 1;
  opt->engine = PC_ENGINE_AUTO;
  opt->reserved = 0;
}

int pc_resolve_engine(int n, int engine) { return pc::resolve_engine(engine, n); }

size_t pc_inverse_pth_root_workspace_bytes(int batch, int n, int engine) {
  if (batch <= 0 || n <= 0) return 0;
  return pc::root_workspace_bytes(batch < 32767 ? batch : 32767, n, engine);
}

// grid.y / grid.z carry the batch in several kernels: larger batches run as consecutive chunks
static int run_root_chunked(const float* xs, const int32_t* ps, const int32_t* ps_host,
                            const int32_t* padding_starts, int batch, int n,
                            const pc_root_options* opt, float* roots, float* metrics,
                            void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  constexpr int kMaxChunk = 32767;  // the SIMT engine launches 2 * batch blocks in grid.z
  const size_t nn = (size_t)n * n;
  for (int b0 = 0; b0 < batch; b0 += kMaxChunk) {
    const int nb = batch - b0 < kMaxChunk ? batch - b0 : kMaxChunk;
    int rc = pc::run_root(xs + (size_t)b0 * nn, ps + b0, ps_host ? ps_host + b0 : nullptr,
                          padding_starts ? padding_starts + b0 : nullptr, nb, n, opt,
                          roots + (size_t)b0 * nn, metrics + (size_t)b0 * PC_NUM_METRICS, workspace,
                          workspace_bytes, stream);
    if (rc != PC_OK) return rc;
  }
  return PC_OK;
}

int pc_inverse_pth_root_batched(const float* xs, const int32_t* ps,
                                const int32_t* padding_starts, int batch, int n,
                                const pc_root_options* opt, float* roots, float* metrics,
                                void* workspace, size_t workspace_bytes, void* stream) {
  return pc_inverse_pth_root_enqueue(xs, ps, nullptr, padding_starts, batch, n, opt, roots,
                                     metrics, workspace, workspace_bytes, stream);
}

int pc_inverse_pth_root_enqueue(const float* xs, const int32_t* ps, const int32_t* ps_host,
                                const int32_t* padding_starts, int batch, int n,
                                const pc_root_options* opt, float* roots, float* metrics,
                                void* workspace, size_t workspace_bytes, void* stream) {
  PC_REQUIRE(batch >= 0 && n >= 1, "bad batch/n (%d, %d)", batch, n);
  if (batch == 0) return PC_OK;
  PC_REQUIRE(xs && ps && roots && metrics && workspace && opt, "null pointer argument");
  PC_REQUIRE(opt->num_iters >= 0, "num_iters < 0");
  return run_root_chunked(xs, ps, ps_host, padding_starts, batch, n, opt, roots, metrics,
                          workspace, workspace_bytes, (cudaStream_t)stream);
}

int pc_root_mode(void) { return pc::root_graph_mode_allowed() ? 1 : 0; }

int pc_power_iteration_batched(const float* xs, const int32_t* padding_starts, int batch,
                               int n, int num_iters, float error_tolerance, float* lambdas,
                               int32_t* iters, void* stream) {
  PC_REQUIRE(batch >= 0 && n >= 1, "bad batch/n");
  if (batch == 0) return PC_OK;
  PC_REQUIRE(xs && lambdas, "null pointer argument");
  float* v0 = nullptr;
  PC_CUDA_CHECK(cudaMallocAsync(&v0, sizeof(float) * ((size_t)n + 2 * (size_t)batch * n),
                                (cudaStream_t)stream));
  int rc = pc::run_power_iteration(xs, padding_starts, batch, n, num_iters, error_tolerance,
                                   lambdas, iters, nullptr, v0, v0 + n, (cudaStream_t)stream);
  cudaFreeAsync(v0, (cudaStream_t)stream);
  return rc;
}

}  // extern "C"
