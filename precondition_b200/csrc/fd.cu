// Sketchy / frequent-directions sketch update on the GPU  (reference: _fd_update_root,
// DS:1123-1290; pack / unpack DS:555-592; caller new_mi_pth_root DS:2706-2738).
//
// The reference stacks  [sqrt(beta2) U sqrt(lambda + eps) | G]  (d x (r + d)) and takes a
// LAPACK SVD.  Only the left singular vectors u_i and s_i^2 are used, and those are the
// eigenpairs of the d x d covariance
//        C = beta2 * U diag(lambda + eps) U^T + G G^T,
// so the B200 path is GEMM-shaped:
//   1. fd_prepare_kernel   unpack the previous sketch, damping, masks      (DS:1150-1174)
//   2. batched GEMMs       C = Bs Bs^T + F F^T   (or + the Gram itself)
//   3. top-(r+1) eigenpairs of C
//        d <= full_eigh_max_dim : cyclic one-sided Jacobi on all of C (exact "small eigh")
//        otherwise              : block subspace iteration warm-started from the previous
//                                 sketch, Rayleigh-Ritz on a k x k matrix (k = r+1+oversample)
//                                 solved by the same Jacobi kernel
//   4. fd_finalize_kernel  deflation by s_r^2, tail accumulation, inversion, safety
//                          masks and packing                             (DS:1195-1262)
// Vectors are kept as ROWS ("transposed" layout) so Jacobi rotations and GEMM operands
// are contiguous.  All GEMMs run on the fp32 CUDA-core tile of simt_gemm.cuh.
#include <math.h>

#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "simt_gemm.cuh"

namespace pc {

// tcgen05 grouped GEMM (tc_gemm.cu): used for the covariance build when d % 128 == 0
bool tc_engine_available();
size_t tc_grouped_gemm_workspace_bytes(const pc_gemm_desc* descs, int count);
int tc_grouped_gemm(const pc_gemm_desc* descs, const pc_gemm_quant* quant, int count,
                    void* workspace, size_t workspace_bytes, int reuse_plan, cudaStream_t stream);

// ---------------------------------------------------------------------------
// batched strided GEMM, descriptor by value:
//   C[z](i,j) = rs[z][i] * alpha * sum_k A[z](i,k) B[z](j,k) + beta * Cin[z](i,j)
// ---------------------------------------------------------------------------
struct FdGemm {
  const float* a; const float* b; const float* cin; float* c;
  const float* row_scale;  // optional [batch, m]
  int64_t a_bs, b_bs, c_bs, rs_bs;
  int64_t a_si, a_sk, b_sj, b_sk, c_si;
  int m, n, k;
  float alpha, beta;
};

__global__ void __launch_bounds__(kSimtThreads) fd_gemm_kernel(const FdGemm g) {
  __shared__ SimtSmem sm;
  const int z = blockIdx.z;
  const int tile_m = blockIdx.y, tile_n = blockIdx.x;
  const OperandView A{g.a + z * g.a_bs, 0, g.a_si, 0, g.a_sk, g.m > 0 ? g.m : 1,
                      g.k > 0 ? g.k : 1, g.m, g.k};
  const OperandView B{g.b + z * g.b_bs, 0, g.b_sj, 0, g.b_sk, g.n > 0 ? g.n : 1,
                      g.k > 0 ? g.k : 1, g.n, g.k};
  const float* cin = g.cin ? g.cin + z * g.c_bs : nullptr;
  float* c = g.c + z * g.c_bs;
  const float* rs = g.row_scale ? g.row_scale + z * g.rs_bs : nullptr;
  simt_gemm_tile(g.k, tile_m, tile_n, A, B, A.k_fast(), B.k_fast(), sm,
                 [&](int i, int j0, const float* acc) {
                   if (i >= g.m) return;
                   const float sc = rs ? rs[i] * g.alpha : g.alpha;
#pragma unroll
                   for (int q = 0; q < 4; ++q) {
                     const int j = j0 + q;
                     if (j >= g.n) continue;
                     float v = sc * acc[q];
                     if (cin) v = fmaf(g.beta, cin[(int64_t)i * g.c_si + j], v);
                     c[(int64_t)i * g.c_si + j] = v;
                   }
                 });
}

static void fd_gemm(const FdGemm& g, int batch, cudaStream_t stream) {
  if (g.m <= 0 || g.n <= 0 || batch <= 0) return;
  dim3 grid((g.n + kSimtBN - 1) / kSimtBN, (g.m + kSimtBM - 1) / kSimtBM, batch);
  fd_gemm_kernel<<<grid, kSimtThreads, 0, stream>>>(g);
  count_launch(1);
}

// ---------------------------------------------------------------------------
// per-matrix scalars handed from prepare to finalize
// ---------------------------------------------------------------------------
struct FdScalars {
  float tail_decayed;  // tail * decay, DS:1201
  float ridge;
  int pad;
  int p;
};

// tearfree's Sketchy (TF/sketchy.py:380-470) keeps SINGULAR values where distributed_shampoo
// keeps eigenvalues, decays its tail by sqrt(decay), adds no ridge to the sketch and instead
// regularises at the inversion: (undeflated + eps)^(-1/p), eps = epsilon (* max undeflated).
struct FdTearfree {
  int on;
  float epsilon;
  int relative;
};

__device__ __forceinline__ float fd_hash_uniform(uint32_t b, uint32_t j, uint32_t i) {
  uint32_t x = b * 0x9E3779B1u ^ (j + 0x7F4A7C15u) * 0x85EBCA77u ^ (i + 0x165667B1u) * 0xC2B2AE3Du;
  x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
  return (float)(x >> 8) * (1.0f / 8388608.0f) - 1.0f;
}

// prev [d, r+2] -> Bs [d, r] = sqrt(decay) * (U masked) * sqrt((lambda + ridge) masked)
// (DS:1155-1171, DS:1180-1192); optionally the start basis of the subspace iteration
// Yt [k, d]: rows < r = previous eigenvectors (random if that slot was empty), rows >= r
// pseudo-random, all masked to the unpadded rows.
__global__ void __launch_bounds__(256)
fd_prepare_kernel(const float* __restrict__ prev, const int32_t* __restrict__ ps,
                  const int32_t* __restrict__ pads, int d, int r, float ridge_epsilon,
                  float error_tolerance, int relative_eps, float decay, float* __restrict__ bs,
                  FdScalars* __restrict__ scal, float* __restrict__ yt, int k, int k_ld,
                  FdTearfree tf) {
  __shared__ float scratch[32];
  const int b = blockIdx.x;
  const int pd = r + 2;
  const float* P = prev + (size_t)b * d * pd;
  const int pad = pads ? min(max(pads[b], 0), d) : d;
  const float max_ev = relative_eps ? P[(size_t)(d - r) * pd + r + 1] : 1.0f;   // DS:1155-1158
  const float ridge = tf.on ? 0.f : ridge_epsilon * fmaxf(max_ev, error_tolerance);  // DS:1159
  if (threadIdx.x == 0 && blockIdx.y == 0) {
    FdScalars s;
    // DS:1201; TF/sketchy.py:390,425: the tail decays by sqrt(second_moment_decay)
    s.tail_decayed = P[(size_t)1 * pd + r + 1] * (tf.on ? sqrtf(decay) : decay);
    s.ridge = ridge;
    s.pad = pad;
    s.p = ps[b];
    scal[b] = s;
  }
  const float sdecay = sqrtf(decay);
  float* B = bs + (size_t)b * d * r;
  for (size_t e = (size_t)blockIdx.y * blockDim.x + threadIdx.x; e < (size_t)d * r;
       e += (size_t)gridDim.y * blockDim.x) {
    const int i = (int)(e / r), j = (int)(e - (size_t)i * r);
    const bool on = i < pad && j < pad;                                  // DS:1167-1168
    const float stored = P[(size_t)(d - r + j) * pd + r + 1];
    // sqrt(eigenvalue + ridge) -- or the stored singular value itself (TF/sketchy.py:387)
    const float sv = tf.on ? stored : sqrtf(stored + ridge);
    const float w = (on ? P[(size_t)i * pd + j] : 0.f) * sv * (j < pad ? 1.f : 0.f);  // DS:1171
    B[e] = sdecay * w;
  }
  if (!yt) return;
  float* Y = yt + (size_t)b * k_ld * d;  // k rows used of k_ld allocated
  for (int j = blockIdx.y; j < k; j += gridDim.y) {  // whole rows per CTA (block-uniform test)
    bool use_prev = false;
    if (j < r) {  // block-uniform decision: is the previous eigenvector slot populated?
      float ss = 0.f;
      for (int i = threadIdx.x; i < pad; i += blockDim.x) {
        const float v = P[(size_t)i * pd + j];
        ss = fmaf(v, v, ss);
      }
      use_prev = block_sum(ss, scratch) > 0.25f;
    }
    for (int i = threadIdx.x; i < d; i += blockDim.x) {
      float v = 0.f;
      if (i < pad) v = use_prev ? P[(size_t)i * pd + j] : fd_hash_uniform(b, j, i);
      Y[(size_t)j * d + i] = v;
    }
  }
}

// masked copy of the new-gradient input (DS:1172-1174): factor [d, m] -> Fm (columns are
// masked as well when the factor is square, as in the reference), or Gram [d, d] -> C.
__global__ void fd_mask_kernel(const float* __restrict__ src, const FdScalars* __restrict__ scal,
                               int rows, int cols, int mask_cols, float* __restrict__ dst) {
  const int b = blockIdx.y;
  const int pad = scal[b].pad;
  const size_t total = (size_t)rows * cols;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / cols), j = (int)(e - (size_t)i * cols);
    const bool on = i < pad && (!mask_cols || j < pad);
    dst[(size_t)b * total + e] = on ? src[(size_t)b * total + e] : 0.f;
  }
}

// each row of X [rows, len] scaled to unit 2-norm (zero rows stay zero)
__global__ void __launch_bounds__(256)
fd_row_normalize_kernel(float* __restrict__ x, int rows_ld, int len) {
  __shared__ float scratch[32];
  float* row = x + ((size_t)blockIdx.y * rows_ld + blockIdx.x) * len;
  float ss = 0.f;
  for (int i = threadIdx.x; i < len; i += blockDim.x) ss = fmaf(row[i], row[i], ss);
  const float nrm = sqrtf(block_sum(ss, scratch));
  const float inv = nrm > 0.f ? 1.0f / nrm : 0.f;
  for (int i = threadIdx.x; i < len; i += blockDim.x) row[i] *= inv;
}

// ---------------------------------------------------------------------------
// Symmetric eigensolver: cyclic one-sided Jacobi (Hestenes) on the ROWS of A [n, n]
// (symmetric, destroyed) with the rotations accumulated in Vt [n, n] (rows = eigenvectors).
// Round-robin ordering: n/2 disjoint pairs per round, one warp per pair; a thread-block
// cluster shares one matrix (rows live in L2, ld/st.cg) and a cluster barrier separates
// rounds.  theta_i = <A_i, Vt_i> (Rayleigh quotient; A_i = theta_i v_i at convergence).
// ---------------------------------------------------------------------------
constexpr int kJacMaxN = 512;    // with accumulated eigenvectors (FD paths)
constexpr int kEighMaxN = 2048;  // factor form (eigh-based roots)
constexpr int kJacMaxSweeps = 15;
constexpr int kJacCnt = kJacMaxSweeps + 1;  // per-matrix counters: rotations per sweep + max row norm^2

__device__ __forceinline__ void jac_cluster_sync(int csize) {
  if (csize > 1) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  } else {
    __threadfence_block();
    __syncthreads();
  }
}

// Q = row elements per lane (n <= 32 Q), kJacThreads = CTA size: shorter rows leave registers for
// more warps per CTA, so that one round needs a single pass over the pairs.
template <int Q, int kJacThreads, bool kWithV>
__global__ void __launch_bounds__(kJacThreads)
fd_jacobi_kernel(float* __restrict__ a_all, float* __restrict__ vt_all, int n, float tol,
                 unsigned* __restrict__ rot_count, float* __restrict__ theta_all, int csize,
                 int max_sweeps) {
  const int b = blockIdx.x / csize;
  int crank = 0;
  if (csize > 1) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  float* A = a_all + (size_t)b * n * n;
  float* V = vt_all ? vt_all + (size_t)b * n * n : nullptr;
  unsigned* cnt = rot_count + (size_t)b * kJacCnt;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarp = (kJacThreads / 32) * csize, gw = crank * (kJacThreads / 32) + warp;
  constexpr bool with_v = kWithV;  // without V: rows of A are a FACTOR (A = G, G^T G = T)
  if (with_v) {                            // Vt <- I
    for (size_t e = (size_t)crank * kJacThreads + threadIdx.x; e < (size_t)n * n;
         e += (size_t)csize * kJacThreads)
      __stcg(V + e, (e / n == e % n) ? 1.f : 0.f);
  }
  jac_cluster_sync(csize);
  const int m2 = (n + 1) & ~1;  // players (one phantom if n is odd)
  const int rounds = m2 - 1, npairs = m2 / 2;
  // Rows whose norm is below 1e-5 of the largest one are rounding noise (they belong to
  // eigenvalues that are numerically zero): two such rows are never rotated against each
  // other, otherwise their random mutual angles keep every sweep busy.  The largest squared
  // row norm of the previous sweep is kept next to the rotation counters.
  unsigned* amax_bits = cnt + kJacMaxSweeps;
  for (int sweep = 0; sweep < max_sweeps; ++sweep) {
    unsigned rotated = 0;
    const float noise2 = 1e-10f * __uint_as_float(__ldcg(amax_bits));
    float seen_max = 0.f;
    for (int t = 0; t < rounds; ++t) {
      for (int p = gw; p < npairs; p += nwarp) {
        int i, j;  // circle method: player m2-1 is fixed, the others rotate
        if (p == 0) { i = t; j = m2 - 1; }
        else { i = (t + p) % rounds; j = (t - p + rounds) % rounds; }
        if (i > j) { const int tmp = i; i = j; j = tmp; }
        if (j >= n) continue;  // phantom
        float ai[Q], aj[Q], vi[kWithV ? Q : 1], vj[kWithV ? Q : 1];
        float al = 0.f, be = 0.f, ga = 0.f;
        // the V rows are fetched together with the A rows: one L2 round trip per pair instead
        // of two on the critical path of the round (most pairs rotate in the early sweeps)
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int c = lane + 32 * q;
          ai[q] = c < n ? __ldcg(A + (size_t)i * n + c) : 0.f;
          aj[q] = c < n ? __ldcg(A + (size_t)j * n + c) : 0.f;
          if (with_v) {
            vi[q] = c < n ? __ldcg(V + (size_t)i * n + c) : 0.f;
            vj[q] = c < n ? __ldcg(V + (size_t)j * n + c) : 0.f;
          }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          al = fmaf(ai[q], ai[q], al);
          be = fmaf(aj[q], aj[q], be);
          ga = fmaf(ai[q], aj[q], ga);
        }
        al = warp_sum(al); be = warp_sum(be); ga = warp_sum(ga);
        seen_max = fmaxf(seen_max, fmaxf(al, be));
        const float lim = tol * sqrtf(al) * sqrtf(be);
        if (!(fabsf(ga) > lim) || lim == 0.f) continue;  // warp-uniform
        if (fmaxf(al, be) < noise2) continue;            // both rows are noise
        ++rotated;
        const float zeta = (be - al) / (2.0f * ga);
        const float tt = copysignf(1.0f, zeta) / (fabsf(zeta) + sqrtf(1.0f + zeta * zeta));
        const float cs = rsqrtf(1.0f + tt * tt), sn = cs * tt;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
          const int c = lane + 32 * q;
          if (c < n) {
            __stcg(A + (size_t)i * n + c, cs * ai[q] - sn * aj[q]);
            __stcg(A + (size_t)j * n + c, sn * ai[q] + cs * aj[q]);
            if (with_v) {
              __stcg(V + (size_t)i * n + c, cs * vi[q] - sn * vj[q]);
              __stcg(V + (size_t)j * n + c, sn * vi[q] + cs * vj[q]);
            }
          }
        }
      }
      jac_cluster_sync(csize);
    }
    if (lane == 0 && rotated) atomicAdd(cnt + sweep, rotated);
    if (lane == 0 && seen_max > 0.f) atomicMax(amax_bits, __float_as_uint(seen_max));
    __threadfence();
    jac_cluster_sync(csize);
    if (__ldcg(cnt + sweep) == 0u) break;  // uniform across the cluster
  }
  float* theta = theta_all + (size_t)b * n;
  for (int i = gw; i < n; i += nwarp) {
    float dot = 0.f;
    for (int c = lane; c < n; c += 32) {
      const float a = __ldcg(A + (size_t)i * n + c);
      dot = fmaf(a, with_v ? __ldcg(V + (size_t)i * n + c) : a, dot);  // <A_i, V_i> or |A_i|^2
    }
    dot = warp_sum(dot);
    if (lane == 0) theta[i] = dot;
  }
}

// order[rank] = index of the rank-th largest theta (ties by index); sorted[rank] = theta
__global__ void __launch_bounds__(512)
fd_sort_kernel(const float* __restrict__ theta_all, int n, int* __restrict__ order_all,
               float* __restrict__ sorted_all) {
  const int b = blockIdx.x;
  const float* th = theta_all + (size_t)b * n;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float raw = th[i];
    const float x = raw != raw ? INFINITY : raw;  // NaN sorts first so that it is seen
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float yr = th[j];
      const float y = yr != yr ? INFINITY : yr;
      rank += (y > x || (y == x && j < i)) ? 1 : 0;
    }
    order_all[(size_t)b * n + rank] = i;
    sorted_all[(size_t)b * n + rank] = raw;
  }
}

// 1/sqrt(theta) row scales for the Gram-based orthonormalisation (directions whose Gram
// eigenvalue is numerically zero are dropped)
__global__ void fd_orth_scale_kernel(const float* __restrict__ theta_all, int n,
                                     float* __restrict__ scale_all) {
  __shared__ float smax;
  const int b = blockIdx.x;
  const float* th = theta_all + (size_t)b * n;
  if (threadIdx.x == 0) {
    float m = 0.f;
    for (int i = 0; i < n; ++i) m = fmaxf(m, th[i]);
    smax = m;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = th[i];
    scale_all[(size_t)b * n + i] = (x > 1e-5f * smax && x > 0.f) ? rsqrtf(x) : 0.f;
  }
}

// dst[b][t][:] = src[b][order[b][t]][:] for t < count
__global__ void fd_gather_rows_kernel(const float* __restrict__ src, const int* __restrict__ order,
                                      int n_rows, int len, int count, float* __restrict__ dst) {
  const int b = blockIdx.y, t = blockIdx.x;
  const float* s = src + ((size_t)b * n_rows + order[(size_t)b * n_rows + t]) * len;
  float* o = dst + ((size_t)b * count + t) * len;
  for (int i = threadIdx.x; i < len; i += blockDim.x) o[i] = s[i];
}

// the same with `rows_ld` rows allocated per matrix in dst (rows >= gridDim.x stay as they are)
__global__ void fd_gather_rows_pad_kernel(const float* __restrict__ src,
                                          const int* __restrict__ order, int n_rows, int len,
                                          int rows_ld, float* __restrict__ dst) {
  const int b = blockIdx.y, t = blockIdx.x;
  const float* s = src + ((size_t)b * n_rows + order[(size_t)b * n_rows + t]) * len;
  float* o = dst + ((size_t)b * rows_ld + t) * len;
  for (int i = threadIdx.x; i < len; i += blockDim.x) o[i] = s[i];
}

// ---------------------------------------------------------------------------
// deflation, tail, inversion, safety masks and packing: DS:1195-1262, DS:572-592
//   vt      [batch, nv, d]  candidate eigenvectors as rows, sorted by eigenvalue (>= r rows)
//   theta   [batch, ld_theta] eigenvalues s_i^2, descending (>= r + 1 entries)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fd_finalize_kernel(const float* __restrict__ vt_all, int nv, const float* __restrict__ theta_all,
                   int ld_theta, const FdScalars* __restrict__ scal, int d, int r,
                   float* __restrict__ out_all, float* __restrict__ metrics, FdTearfree tf) {
  extern __shared__ float sh[];
  float* deflated = sh;        // [r]
  float* inverted = sh + r;    // [r]
  float* keep = sh + 2 * r;    // [r] 1 = direction kept
  float* scale = sh + 3 * r;   // [r] 1 / norm
  __shared__ int has_zero_flag;
  const int b = blockIdx.x;
  const FdScalars s = scal[b];
  const float* vt = vt_all + (size_t)b * nv * d;
  const float* theta = theta_all + (size_t)b * ld_theta;
  float* out = out_all + (size_t)b * d * (r + 2);
  const int pd = r + 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  if (threadIdx.x == 0) has_zero_flag = 0;
  __syncthreads();
  const float alpha = -1.0f / (float)s.p;
  // Eigenvalues of the fp32 covariance carry absolute noise ~ eps * sqrt(d) * theta_0, where
  // LAPACK's SVD of the stacked factor returns (near-)exact zeros for a rank-deficient
  // update.  Values under that floor are treated as the zeros they stand for, so the
  // has_zeros / skip logic (DS:1209-1251) fires like it does in the reference.
  const float floor_ev = 4.0f * 1.1920929e-7f * sqrtf((float)d) * fmaxf(theta[0], 0.f);
  auto ev = [&](int j) { const float t = theta[j]; return t > floor_ev ? t : 0.f; };
  const float cutoff = sqrtf(ev(r));  // s[rank], DS:1195
  const float rho = cutoff * cutoff;
  float new_tail = s.tail_decayed + rho;             // DS:1202
  // tearfree: eps = epsilon * max(undeflated) with undeflated_0 = theta_0 + decayed tail the
  // largest (TF/sketchy.py:444-449); distributed_shampoo regularised the sketch instead
  float inv_eps = 0.f;
  if (tf.on) inv_eps = tf.relative && tf.epsilon > 0.f ? (ev(0) + s.tail_decayed) * tf.epsilon
                                                        : tf.epsilon;
  const float new_const = new_tail <= 0.f ? 0.f : powf(new_tail + inv_eps, alpha);  // DS:1205
  new_tail = new_tail <= 0.f ? 0.f : new_tail;
  // one warp per direction: deflation, norm / padding safety (DS:1199-1246)
  for (int j = warp; j < r; j += nwarp) {
    const float top = sqrtf(ev(j));
    float defl = (top - cutoff) * (top + cutoff);    // DS:1199
    defl = defl <= 0.f ? 0.f : defl;                 // DS:1209
    float ss = 0.f, padmass = 0.f;
    const float on = defl > 0.f ? 1.f : 0.f;         // DS:1210
    for (int i = lane; i < d; i += 32) {
      const float v = vt[(size_t)j * d + i] * on;
      ss = fmaf(v, v, ss);
    }
    const float nrm = sqrtf(warp_sum(ss));
    const bool safe = 0.99f <= nrm && nrm <= 1.01f;  // DS:1214-1216
    const float inv = safe ? 1.0f / nrm : 1.0f;
    for (int i = lane; i < d; i += 32)
      if (i >= s.pad)  // DS:1224-1226: L1 mass of the (normalised) vector on padding rows
        padmass += fabsf(vt[(size_t)j * d + i] * on * (safe ? inv : 0.f));
    padmass = warp_sum(padmass);
    const bool haspad = padmass > 0.01f;
    const float kp = (safe && !haspad) ? on : 0.f;
    defl = defl * (safe ? 1.f : 0.f) * (haspad ? 0.f : 1.f);
    float up = (top * top + s.tail_decayed) * (defl > 0.f ? 1.f : 0.f);  // DS:1247-1248
    up = up <= 0.f ? 0.f : up;
    const float invd = up <= 0.f ? 0.f : powf(up + inv_eps, alpha);
    if (lane == 0) {
      deflated[j] = tf.on ? sqrtf(defl) : defl;  // TF/sketchy.py:404-406: singular values
      inverted[j] = invd;
      keep[j] = kp;
      scale[j] = inv;
      if (defl <= 0.f) has_zero_flag = 1;  // DS:1251
    }
  }
  __syncthreads();
  // tearfree has no skip rule: dropped directions and an empty tail simply contribute nothing
  const bool has_zeros = !tf.on && (has_zero_flag != 0 || new_tail <= 0.f);
  const bool zero_all = s.pad == 0;  // DS:1265-1268
  for (size_t e = threadIdx.x; e < (size_t)d * pd; e += blockDim.x) {
    const int i = (int)(e / pd), j = (int)(e - (size_t)i * pd);
    float v = 0.f;
    if (!zero_all) {
      if (j < r) {
        v = vt[(size_t)j * d + i] * keep[j] * scale[j];
        if (v == 0.f) v = 0.f;  // -0 -> +0 like the reference's masked product
      } else if (j == r) {      // column -2: inverted eigenvalues, has_zeros in the last row
        if (i < r) v = inverted[i];
        if (i == d - 1) v = has_zeros ? 1.f : 0.f;
      } else {                  // column -1: const, tail, deflated eigenvalues (DS:585-591)
        if (i == 0) v = new_const;
        if (i == 1) v = new_tail;
        if (i >= d - r) v = deflated[i - (d - r)];
      }
    }
    out[e] = v;
  }
  if (threadIdx.x == 0 && metrics) {
    float* m = metrics + (size_t)b * PC_NUM_METRICS;  // DS:1263-1264: error 0, rest default
    m[0] = 0.f; m[1] = 0.f; m[2] = 0.f; m[3] = 0.f; m[4] = 0.f;
  }
}

// ---------------------------------------------------------------------------
// Cheap orthonormalisation of the intermediate subspace blocks (shifted Cholesky QR):
//   G = Yt Yt^T + shift I = L L^T,   Qt = L^-1 Yt.
// Only the SPAN of the block matters between iterations (the Rayleigh-Ritz rotation inside the
// span is redundant there), so the k x k eigen-solves are kept for the last iteration only.
// ---------------------------------------------------------------------------
// in place on a [k, k] Gram matrix (lower triangle <- L); one CTA (32 x 32 threads) per matrix.
// kSmem: the packed lower triangle (k (k + 1) / 2 floats) lives in shared memory for the whole
// factorisation.  The scaled pivot column is staged in its own vector so that the trailing
// update reads it without bank conflicts; threads tile the trailing block in 2-D (no div / mod).
template <bool kSmem>
__global__ void __launch_bounds__(1024)
fd_cholesky_shift_kernel(float* __restrict__ a_all, int k, float shift) {
  extern __shared__ float chol_smem[];
  float* col = chol_smem;            // [k] scaled pivot column
  float* packed = chol_smem + ((k + 31) & ~31);
  __shared__ float piv_s;
  float* A = a_all + (size_t)blockIdx.x * k * k;
  const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
  auto at = [&](int i, int j) -> float& {  // j <= i
    return kSmem ? packed[(size_t)i * (i + 1) / 2 + j] : A[(size_t)i * k + j];
  };
  if (kSmem) {
    for (int i = ty; i < k; i += 32)
      for (int j = tx; j <= i; j += 32)
        packed[(size_t)i * (i + 1) / 2 + j] = A[(size_t)i * k + j] + (i == j ? shift : 0.f);
  } else {
    for (int i = tid; i < k; i += blockDim.x) A[(size_t)i * k + i] += shift;
  }
  __syncthreads();
  for (int c = 0; c < k; ++c) {
    if (tid == 0) {
      const float p = at(c, c);
      piv_s = p > 0.f ? sqrtf(p) : __int_as_float(0x7fc00000);
      at(c, c) = piv_s;
    }
    __syncthreads();
    const float inv = 1.0f / piv_s;
    for (int i = c + 1 + tid; i < k; i += blockDim.x) {
      const float v = at(i, c) * inv;
      at(i, c) = v;
      col[i] = v;
    }
    __syncthreads();
    for (int i = c + 1 + ty; i < k; i += 32) {
      const float li = col[i];
      for (int j = c + 1 + tx; j <= i; j += 32) at(i, j) -= li * col[j];
    }
    __syncthreads();
  }
  if (kSmem) {
    for (int i = ty; i < k; i += 32)
      for (int j = tx; j <= i; j += 32) A[(size_t)i * k + j] = packed[(size_t)i * (i + 1) / 2 + j];
  }
}
// dst [k, k] <- top-left corner of src [ld, ld]
__global__ void fd_compact_kernel(const float* __restrict__ src, int ld, float* __restrict__ dst,
                                  int k) {
  const int b = blockIdx.y;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < k * k; e += gridDim.x * blockDim.x) {
    const int i = e / k, j = e - i * k;
    dst[(size_t)b * k * k + e] = src[((size_t)b * ld + i) * ld + j];
  }
}
constexpr size_t kCholSmemMax = 200 * 1024;
static int fd_cholesky_shift(float* a, int k, int batch, float shift, cudaStream_t stream) {
  const size_t colb = (size_t)((k + 31) & ~31) * sizeof(float);
  const size_t bytes = colb + (size_t)k * (k + 1) / 2 * sizeof(float);
  if (bytes <= kCholSmemMax) {
    static bool configured = false;
    if (!configured) {
      PC_CUDA_CHECK(cudaFuncSetAttribute(fd_cholesky_shift_kernel<true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kCholSmemMax));
      configured = true;
    }
    fd_cholesky_shift_kernel<true><<<batch, 1024, bytes, stream>>>(a, k, shift);
  } else {
    fd_cholesky_shift_kernel<false><<<batch, 1024, colb, stream>>>(a, k, shift);
  }
  count_launch(1);
  return PC_OK;
}

// L^-1 of the [k, k] Cholesky factor into the top-left corner of a zero [k_ld, k_ld] matrix: with
// it, Qt = L^-1 Yt is ONE tensor-core GEMM instead of k dependent substitution steps over d
// columns (21 of 76 ms of the Sketchy step at 16 x 4096^2).  One CTA per matrix, the packed
// triangle in shared memory, one warp per column of the inverse (x_jj = 1 / l_jj,
// x_ij = -(sum_{j <= m < i} l_im x_mj) / l_ii, the sum split over the lanes in a fixed order).
__global__ void __launch_bounds__(1024)
fd_tri_inverse_kernel(const float* __restrict__ l_all, float* __restrict__ inv_all, int k,
                      int k_ld) {
  extern __shared__ float tri_smem[];
  float* packed = tri_smem;                               // k (k + 1) / 2
  float* xcol = tri_smem + (size_t)k * (k + 1) / 2;       // [32 warps][k]
  const float* L = l_all + (size_t)blockIdx.x * k * k;
  float* X = inv_all + (size_t)blockIdx.x * k_ld * k_ld;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = warp; i < k; i += 32)
    for (int j = lane; j <= i; j += 32) packed[(size_t)i * (i + 1) / 2 + j] = L[(size_t)i * k + j];
  __syncthreads();
  float* x = xcol + (size_t)warp * k;
  // gridDim.y CTAs share one matrix: the columns are dealt round-robin over all their warps (the
  // early columns are the long ones), so with k <= 32 gridDim.y every warp solves one column
  for (int j = blockIdx.y * 32 + warp; j < k; j += gridDim.y * 32) {
    const float xjj = 1.0f / packed[(size_t)j * (j + 1) / 2 + j];
    if (lane == 0) {
      x[j] = xjj;
      X[(size_t)j * k_ld + j] = xjj;
    }
    __syncwarp();
    for (int i = j + 1; i < k; ++i) {
      const float* li = packed + (size_t)i * (i + 1) / 2;
      float acc = 0.f;
      for (int m = j + lane; m < i; m += 32) acc = fmaf(li[m], x[m], acc);
      acc = warp_sum(acc);
      const float xij = -acc / li[i];
      if (lane == 0) {
        x[i] = xij;
        X[(size_t)i * k_ld + j] = xij;
      }
      __syncwarp();
    }
  }
}
static bool fd_tri_inverse_fits(int k) {
  return ((size_t)k * (k + 1) / 2 + 32 * (size_t)k) * sizeof(float) <= kCholSmemMax;
}
static int fd_tri_inverse(const float* l, float* inv, int k, int k_ld, int batch,
                          cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    PC_CUDA_CHECK(cudaFuncSetAttribute(fd_tri_inverse_kernel,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)kCholSmemMax));
    configured = true;
  }
  const size_t bytes = ((size_t)k * (k + 1) / 2 + 32 * (size_t)k) * sizeof(float);
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int share = std::max(1, std::min((k + 31) / 32, sms / std::max(batch, 1)));
  fd_tri_inverse_kernel<<<dim3(batch, share), 1024, bytes, stream>>>(l, inv, k, k_ld);
  count_launch(1);
  return PC_OK;
}

// Qt = L^-1 Yt by forward substitution, in place allowed; one thread per column of the [k, d]
// block (the k^2 / 2 steps of a column are sequential, the d columns are independent).  The
// finished part of a thread's column stays in shared memory ([k][128] floats) and the rows of L
// are staged through shared memory eight at a time, so the inner loop is two LDS and one FMA.
constexpr int kTrsmRows = 8;
__global__ void __launch_bounds__(128)
fd_trsm_rows_kernel(const float* __restrict__ l_all, const float* yt_all, float* qt_all, int k,
                    int d, int k_ld) {
  extern __shared__ float trsm_smem[];
  const int kp = (k + 3) & ~3;
  float* lrows = trsm_smem;                    // [kTrsmRows][kp]
  float* cols = trsm_smem + kTrsmRows * kp;    // [k][128]
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool live = c < d;
  const float* L = l_all + (size_t)b * k * k;
  const float* Y = yt_all + (size_t)b * k_ld * d;
  float* Q = qt_all + (size_t)b * k_ld * d;
  for (int i0 = 0; i0 < k; i0 += kTrsmRows) {
    const int nr = min(kTrsmRows, k - i0);
    __syncthreads();
    for (int e = threadIdx.x; e < nr * kp; e += blockDim.x) {
      const int r = e / kp, j = e - r * kp;
      lrows[e] = (j < k && j <= i0 + r) ? __ldg(L + (size_t)(i0 + r) * k + j) : 0.f;
    }
    __syncthreads();
    if (!live) continue;
    for (int r = 0; r < nr; ++r) {
      const int i = i0 + r;
      const float* li = lrows + r * kp;
      float a0 = Y[(size_t)i * d + c], a1 = 0.f, a2 = 0.f, a3 = 0.f;
      float a4 = 0.f, a5 = 0.f, a6 = 0.f, a7 = 0.f;
      int j = 0;
      for (; j + 8 <= i; j += 8) {
        const float4 l0 = *reinterpret_cast<const float4*>(li + j);
        const float4 l1 = *reinterpret_cast<const float4*>(li + j + 4);
        a0 = fmaf(-l0.x, cols[j * 128 + threadIdx.x], a0);
        a1 = fmaf(-l0.y, cols[(j + 1) * 128 + threadIdx.x], a1);
        a2 = fmaf(-l0.z, cols[(j + 2) * 128 + threadIdx.x], a2);
        a3 = fmaf(-l0.w, cols[(j + 3) * 128 + threadIdx.x], a3);
        a4 = fmaf(-l1.x, cols[(j + 4) * 128 + threadIdx.x], a4);
        a5 = fmaf(-l1.y, cols[(j + 5) * 128 + threadIdx.x], a5);
        a6 = fmaf(-l1.z, cols[(j + 6) * 128 + threadIdx.x], a6);
        a7 = fmaf(-l1.w, cols[(j + 7) * 128 + threadIdx.x], a7);
      }
      for (; j < i; ++j) a0 = fmaf(-li[j], cols[j * 128 + threadIdx.x], a0);
      const float q = (((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7))) / li[i];
      cols[i * 128 + threadIdx.x] = q;
      Q[(size_t)i * d + c] = q;
    }
  }
}
// fallback for k too large for the shared-memory column block
__global__ void __launch_bounds__(128)
fd_trsm_rows_global_kernel(const float* __restrict__ l_all, const float* yt_all, float* qt_all,
                           int k, int d, int k_ld) {
  const int b = blockIdx.y;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= d) return;
  const float* L = l_all + (size_t)b * k * k;
  const float* Y = yt_all + (size_t)b * k_ld * d;
  float* Q = qt_all + (size_t)b * k_ld * d;
  for (int i = 0; i < k; ++i) {
    const float* li = L + (size_t)i * k;
    float a0 = Y[(size_t)i * d + c], a1 = 0.f, a2 = 0.f, a3 = 0.f;
    int j = 0;
    for (; j + 4 <= i; j += 4) {
      a0 = fmaf(-__ldg(li + j), Q[(size_t)j * d + c], a0);
      a1 = fmaf(-__ldg(li + j + 1), Q[(size_t)(j + 1) * d + c], a1);
      a2 = fmaf(-__ldg(li + j + 2), Q[(size_t)(j + 2) * d + c], a2);
      a3 = fmaf(-__ldg(li + j + 3), Q[(size_t)(j + 3) * d + c], a3);
    }
    for (; j < i; ++j) a0 = fmaf(-__ldg(li + j), Q[(size_t)j * d + c], a0);
    Q[(size_t)i * d + c] = ((a0 + a1) + (a2 + a3)) / __ldg(li + i);
  }
}
static int fd_trsm_rows(const float* l, const float* yt, float* qt, int k, int k_ld, int d,
                        int batch, cudaStream_t stream) {
  const size_t bytes = ((size_t)k * 128 + (size_t)kTrsmRows * ((k + 3) & ~3)) * sizeof(float);
  dim3 grid((d + 127) / 128, batch);
  if (bytes <= kCholSmemMax) {
    static bool configured = false;
    if (!configured) {
      PC_CUDA_CHECK(cudaFuncSetAttribute(fd_trsm_rows_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kCholSmemMax));
      configured = true;
    }
    fd_trsm_rows_kernel<<<grid, 128, bytes, stream>>>(l, yt, qt, k, d, k_ld);
  } else {
    fd_trsm_rows_global_kernel<<<grid, 128, 0, stream>>>(l, yt, qt, k, d, k_ld);
  }
  count_launch(1);
  return PC_OK;
}

// ---------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------
struct FdPlan {
  bool subspace;
  int k;  // Jacobi size: d (full) or r + 1 + oversample
};

static FdPlan fd_plan(int d, int rank, const pc_fd_options* opt) {
  FdPlan p;
  p.subspace = d > opt->full_eigh_max_dim;
  p.k = d;
  if (p.subspace) {
    int k = rank + 1 + std::max(opt->oversample, 0);
    k = std::min(k, std::min(d, kJacMaxN));
    p.k = k;
  }
  return p;
}

struct FdWorkspace {
  FdScalars* scal; unsigned* rot; float* bs; float* fm; float* cmat; float* vt; float* theta;
  float* sorted; int* order; float* rscale; float* yt; float* qt; float* pt; float* small;
  float* zsel; float* vtop;
  char* tcws; size_t tcws_bytes;  // tcgen05 grouped-GEMM workspace of the covariance build
  // subspace path on the tensor cores: the [k, d] blocks are allocated with k_ld = k rounded up
  // to 128 rows (zero rows) so that the long-contraction products run on the grouped GEMM
  int k_ld;
  float* gram_pad;                 // [batch, k_ld, k_ld]
  char* tcws_gram; size_t tcws_gram_bytes;  // Gram of Yt (plan reused across iterations)
  char* tcws_c; size_t tcws_c_bytes;        // Qt C (plan reused across iterations)
  char* tcws_misc; size_t tcws_misc_bytes;  // one-off plans (Gram of Qt, Ritz matrix)
  float* linv;                     // [batch, k_ld, k_ld] inverse Cholesky factor (zero padded)
  char* tcws_x; size_t tcws_x_bytes;        // L^-1 Yt (plan reused across iterations)
  char* tcws_x2; size_t tcws_x2_bytes;      // one-off [k_ld, d] products (second pass, final rotation)
};

// descriptors of the subspace products on the tcgen05 grouped GEMM
static void fd_sub_descs(const float* x, const float* y, float* dst, int batch, int k_ld, int d,
                         std::vector<pc_gemm_desc>* out) {  // dst [k_ld, k_ld] = X Y^T over d
  for (int b = 0; b < batch; ++b) {
    pc_gemm_desc g{};
    g.a = x + (size_t)b * k_ld * d; g.b = y + (size_t)b * k_ld * d;
    g.c = dst + (size_t)b * k_ld * k_ld; g.c_in = nullptr;
    g.a_iinner = k_ld; g.a_sio = 0; g.a_si = d; g.a_kinner = g.b_kinner = d;
    g.a_sko = g.b_sko = 0; g.a_ski = g.b_ski = 1; g.b_sj = d;
    g.c_iinner = k_ld; g.c_sio = 0; g.c_sii = k_ld;
    g.m = g.n = k_ld; g.k = d; g.alpha = 1.f; g.beta = 0.f;
    out->push_back(g);
  }
}
static void fd_timesc_descs(const float* x, const float* cmat, float* dst, int batch, int k_ld,
                            int d, std::vector<pc_gemm_desc>* out) {  // dst [k_ld, d] = X C
  for (int b = 0; b < batch; ++b) {
    pc_gemm_desc g{};
    g.a = x + (size_t)b * k_ld * d; g.b = cmat + (size_t)b * d * d;
    g.c = dst + (size_t)b * k_ld * d; g.c_in = nullptr;
    g.a_iinner = k_ld; g.a_sio = 0; g.a_si = d; g.a_kinner = g.b_kinner = d;
    g.a_sko = g.b_sko = 0; g.a_ski = g.b_ski = 1; g.b_sj = d;  // C is symmetric: rows of C
    g.c_iinner = k_ld; g.c_sio = 0; g.c_sii = d;
    g.m = k_ld; g.n = d; g.k = d; g.alpha = 1.f; g.beta = 0.f;
    out->push_back(g);
  }
}

// dst [m_ld, d] = coef [m_ld, kk] (row stride coef_ld) times X [kk rows of a k_ld-row block, d]
static void fd_left_descs(const float* coef, int coef_ld, int64_t coef_bs, const float* x,
                          float* dst, int batch, int m_ld, int kk, int k_ld, int d,
                          std::vector<pc_gemm_desc>* out) {
  for (int b = 0; b < batch; ++b) {
    pc_gemm_desc g{};
    g.a = coef + (size_t)b * coef_bs; g.b = x + (size_t)b * k_ld * d;
    g.c = dst + (size_t)b * m_ld * d; g.c_in = nullptr;
    g.a_iinner = m_ld; g.a_sio = 0; g.a_si = coef_ld; g.a_kinner = kk; g.a_sko = 0; g.a_ski = 1;
    g.b_sj = 1; g.b_kinner = kk; g.b_sko = 0; g.b_ski = d;  // B(j, k) = X[k, j]
    g.c_iinner = m_ld; g.c_sio = 0; g.c_sii = d;
    g.m = m_ld; g.n = d; g.k = kk; g.alpha = 1.f; g.beta = 0.f;
    out->push_back(g);
  }
}

// descriptors of the covariance build on the tcgen05 grouped GEMM (host side, per matrix)
static void fd_cov_descs(const float* fm, const float* bs, float* cmat, int batch, int d, int m,
                         int rank, bool gram, std::vector<pc_gemm_desc>* first,
                         std::vector<pc_gemm_desc>* second) {
  for (int b = 0; b < batch; ++b) {
    pc_gemm_desc g{};
    g.c = cmat + (size_t)b * d * d;
    g.c_iinner = d; g.c_sio = 0; g.c_sii = d;
    g.a_iinner = d; g.a_sio = 0; g.a_sko = g.b_sko = 0; g.a_ski = g.b_ski = 1;
    g.m = g.n = d;
    if (!gram) {  // C = F F^T
      pc_gemm_desc f = g;
      f.a = f.b = fm + (size_t)b * d * m;
      f.a_si = f.b_sj = m; f.a_kinner = f.b_kinner = m; f.k = m;
      f.c_in = nullptr; f.alpha = 1.f; f.beta = 0.f;
      first->push_back(f);
    }
    pc_gemm_desc s2 = g;  // C += Bs Bs^T
    s2.a = s2.b = bs + (size_t)b * d * rank;
    s2.a_si = s2.b_sj = rank; s2.a_kinner = s2.b_kinner = rank; s2.k = rank;
    s2.c_in = g.c; s2.alpha = 1.f; s2.beta = 1.f;
    second->push_back(s2);
  }
}

static bool fd_use_tc(int d) { return d % 128 == 0 && d >= 256 && tc_engine_available(); }

static size_t fd_carve(FdWorkspace* w, char* base, int batch, int d, int m, int rank,
                       const FdPlan& pl, bool gram) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off += align_up(bytes, 256);
    return p;
  };
  const size_t B = (size_t)batch;
  const int k = pl.k;
  const int rot_slots = 64;  // Jacobi calls per update (each uses kJacMaxSweeps counters)
  w->scal = (FdScalars*)take(B * sizeof(FdScalars));
  w->rot = (unsigned*)take(B * kJacCnt * rot_slots * sizeof(unsigned));
  w->bs = (float*)take(B * d * rank * 4);
  w->fm = gram ? nullptr : (float*)take(B * d * m * 4);
  w->cmat = (float*)take(B * d * d * 4);
  w->vt = (float*)take(B * k * k * 4);
  w->theta = (float*)take(B * k * 4);
  w->sorted = (float*)take(B * k * 4);
  w->order = (int*)take(B * k * 4);
  w->rscale = (float*)take(B * k * 4);
  w->k_ld = k;
  w->gram_pad = nullptr;
  w->tcws_gram = w->tcws_c = w->tcws_misc = nullptr;
  w->tcws_gram_bytes = w->tcws_c_bytes = w->tcws_misc_bytes = 0;
  if (pl.subspace) {
    if (fd_use_tc(d)) w->k_ld = (k + 127) / 128 * 128;
    const size_t kl = (size_t)w->k_ld;
    w->yt = (float*)take(B * kl * d * 4);
    w->qt = (float*)take(B * kl * d * 4);
    w->pt = (float*)take(B * kl * d * 4);
    w->small = (float*)take(B * k * k * 4);
    // tensor-core path: the selected Ritz vectors and the rotated block are padded to k_ld rows
    const size_t zrows = (w->k_ld != k) ? kl : (size_t)(rank + 1);
    w->zsel = (float*)take(B * zrows * k * 4);
    w->vtop = (float*)take(B * zrows * d * 4);
    w->linv = nullptr;
    w->tcws_x = w->tcws_x2 = nullptr;
    w->tcws_x_bytes = w->tcws_x2_bytes = 0;
    if (w->k_ld != k) {
      w->linv = (float*)take(B * kl * kl * 4);
      {
        std::vector<pc_gemm_desc> gx;
        fd_left_descs(nullptr, w->k_ld, 0, nullptr, nullptr, batch, w->k_ld, w->k_ld, w->k_ld, d,
                      &gx);
        for (auto& gd : gx) gd.b = reinterpret_cast<const float*>(1);
        w->tcws_x_bytes = tc_grouped_gemm_workspace_bytes(gx.data(), (int)gx.size()) + 1024;
        w->tcws_x2_bytes = w->tcws_x_bytes;
        w->tcws_x = take(w->tcws_x_bytes);
        w->tcws_x2 = take(w->tcws_x2_bytes);
      }
      w->gram_pad = (float*)take(B * kl * kl * 4);
      std::vector<pc_gemm_desc> g1, g2;
      fd_sub_descs(nullptr, nullptr, nullptr, batch, w->k_ld, d, &g1);
      fd_timesc_descs(nullptr, nullptr, nullptr, batch, w->k_ld, d, &g2);
      // (a symmetric plan packs one operand, a general one two: size for the general form)
      for (auto& gd : g1) gd.b = reinterpret_cast<const float*>(1);
      w->tcws_gram_bytes = tc_grouped_gemm_workspace_bytes(g1.data(), (int)g1.size()) + 1024;
      w->tcws_misc_bytes = w->tcws_gram_bytes;
      w->tcws_c_bytes = tc_grouped_gemm_workspace_bytes(g2.data(), (int)g2.size()) + 1024;
      w->tcws_gram = take(w->tcws_gram_bytes);
      w->tcws_misc = take(w->tcws_misc_bytes);
      w->tcws_c = take(w->tcws_c_bytes);
    }
  } else {
    w->yt = w->qt = w->pt = w->small = w->zsel = nullptr;
    w->linv = nullptr;
    w->tcws_x = w->tcws_x2 = nullptr;
    w->tcws_x_bytes = w->tcws_x2_bytes = 0;
    w->vtop = (float*)take(B * (rank + 1) * d * 4);
  }
  w->tcws = nullptr;
  w->tcws_bytes = 0;
  if (fd_use_tc(d)) {
    std::vector<pc_gemm_desc> first, second;
    fd_cov_descs(nullptr, nullptr, nullptr, batch, d, m, rank, gram, &first, &second);
    // sized from dummy descriptors: only the extents matter for the plan
    size_t need = tc_grouped_gemm_workspace_bytes(second.data(), (int)second.size());
    if (!first.empty())
      need = std::max(need, tc_grouped_gemm_workspace_bytes(first.data(), (int)first.size()));
    w->tcws_bytes = need + 1024;
    w->tcws = take(w->tcws_bytes);
  }
  return off + 256;
}

template <int Q, int kThreads, bool kWithV>
static int fd_jacobi_launch_v(float* a, float* vt, int n, int batch, unsigned* rot, float* theta,
                              cudaStream_t stream, float tol, int max_sweeps) {
  int csize = 1;
  while (csize < 8 && (kThreads / 32) * csize < (n + 1) / 2) csize <<= 1;
  // Large batches (tearfree's blocked Shampoo: hundreds of 256 x 256 statistics): a cluster per
  // matrix only pays while there are SMs to spare; beyond that fewer CTAs per matrix (more pairs
  // per warp and round, block barriers instead of cluster barriers) give the same rotations --
  // bitwise, the pairs of a round are disjoint -- at a multiple of the throughput.
  {
    static int occ = 0;  // resident CTAs per SM of this instantiation
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (occ == 0) {
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fd_jacobi_kernel<Q, kThreads, kWithV>,
                                                    kThreads, 0);
      if (occ < 1) occ = 1;
    }
    while (csize > 1 && (long long)batch * csize > (long long)sms * occ) csize >>= 1;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(batch * csize));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PC_CUDA_CHECK(cudaLaunchKernelEx(&cfg, fd_jacobi_kernel<Q, kThreads, kWithV>, a, vt, n, tol, rot,
                                   theta, csize, max_sweeps));
  count_launch(1);
  return PC_OK;
}
template <int Q, int kThreads>
static int fd_jacobi_launch(float* a, float* vt, int n, int batch, unsigned* rot, float* theta,
                            cudaStream_t stream, float tol, int max_sweeps) {
  return vt ? fd_jacobi_launch_v<Q, kThreads, true>(a, vt, n, batch, rot, theta, stream, tol,
                                                    max_sweeps)
            : fd_jacobi_launch_v<Q, kThreads, false>(a, vt, n, batch, rot, theta, stream, tol,
                                                     max_sweeps);
}

static int fd_jacobi(float* a, float* vt, int n, int batch, unsigned* rot, float* theta,
                     cudaStream_t stream, float tol = 3e-6f, int max_sweeps = kJacMaxSweeps) {
  max_sweeps = std::min(max_sweeps, kJacMaxSweeps);
  if (n <= 128) return fd_jacobi_launch<4, 512>(a, vt, n, batch, rot, theta, stream, tol, max_sweeps);
  if (n <= 256) return fd_jacobi_launch<8, 640>(a, vt, n, batch, rot, theta, stream, tol, max_sweeps);
  if (n <= 384) return fd_jacobi_launch<12, 640>(a, vt, n, batch, rot, theta, stream, tol, max_sweeps);
  if (n <= 512) return fd_jacobi_launch<16, 512>(a, vt, n, batch, rot, theta, stream, tol, max_sweeps);
  // Larger factors (the eigh-based roots up to 2048 x 2048): rows of 32 / 64 elements per lane,
  // factor form only (no accumulated V: the rotated rows ARE sigma_i u_i^T).  Correct but slow --
  // one cluster of 8 CTAs per matrix, every rotation streams two rows through L2.
  if (vt != nullptr || n > kEighMaxN) {
    set_error("Jacobi eigen-solve with accumulated vectors supports n <= %d (n = %d)", kJacMaxN, n);
    return PC_ERR_UNSUPPORTED;
  }
  if (n <= 1024)
    return fd_jacobi_launch_v<32, 512, false>(a, nullptr, n, batch, rot, theta, stream, tol, max_sweeps);
  return fd_jacobi_launch_v<64, 256, false>(a, nullptr, n, batch, rot, theta, stream, tol, max_sweeps);
}

int run_fd_update(const float* new_grad, const float* prev, const int32_t* ps,
                  const int32_t* pads, int batch, int d, int m, int rank,
                  const pc_fd_options* opt, float* out, float* metrics, void* workspace,
                  size_t workspace_bytes, cudaStream_t stream) {
  const bool gram = opt->input_is_gram != 0;
  const FdTearfree tf{opt->tearfree, opt->tearfree_epsilon, opt->tearfree_relative_epsilon};
  const FdPlan pl = fd_plan(d, rank, opt);
  FdWorkspace w;
  const size_t need = fd_carve(&w, nullptr, batch, d, m, rank, pl, gram);
  if (workspace_bytes < need) {
    set_error("fd workspace too small: %zu < %zu", workspace_bytes, need);
    return PC_ERR_WORKSPACE;
  }
  fd_carve(&w, reinterpret_cast<char*>(align_up((size_t)workspace, 256)), batch, d, m, rank, pl,
           gram);
  const int k = pl.k;
  PC_CUDA_CHECK(cudaMemsetAsync(w.rot, 0, (size_t)batch * kJacCnt * 64 * sizeof(unsigned),
                                stream));
  int jac_calls = 0;
  auto jacobi = [&](float* a, float* vt, int n, bool orth_only = false) -> int {
    unsigned* rot = w.rot + (size_t)(jac_calls++ % 64) * batch * kJacCnt;
    // orthonormalisation solves act on a Gram matrix of unit rows (nearly the identity once
    // the basis has settled) and are followed by a Rayleigh-Ritz solve: a loose tolerance
    // and a few sweeps suffice there
    return orth_only ? fd_jacobi(a, vt, n, batch, rot, w.theta, stream, 3e-5f, 8)
                     : fd_jacobi(a, vt, n, batch, rot, w.theta, stream);
  };

  // grid.y CTAs share one matrix (the kernel used to run one CTA per matrix: 5 ms at d = 4096)
  fd_prepare_kernel<<<dim3(batch, 64), 256, 0, stream>>>(prev, ps, pads, d, rank, opt->ridge_epsilon,
                                              opt->error_tolerance, opt->relative_matrix_epsilon,
                                              opt->decay, w.bs, w.scal, pl.subspace ? w.yt : nullptr,
                                              k, w.k_ld, tf);
  count_launch(1);
  const unsigned mgrid = (unsigned)std::min<size_t>(((size_t)d * std::max(d, m) + 255) / 256, 1024);
  // ---- covariance C = Bs Bs^T + (masked) F F^T ----
  if (gram) {
    fd_mask_kernel<<<dim3(mgrid, batch), 256, 0, stream>>>(new_grad, w.scal, d, d, 1, w.cmat);
  } else {
    fd_mask_kernel<<<dim3(mgrid, batch), 256, 0, stream>>>(new_grad, w.scal, d, m, m == d ? 1 : 0,
                                                          w.fm);
  }
  count_launch(1);
  if (fd_use_tc(d)) {
    // symmetric rank-k updates on the tensor cores (scaled-fp16 three-pass, 22-bit operands)
    std::vector<pc_gemm_desc> first, second;
    fd_cov_descs(w.fm, w.bs, w.cmat, batch, d, m, rank, gram, &first, &second);
    if (!first.empty()) {
      int rc = tc_grouped_gemm(first.data(), nullptr, (int)first.size(), w.tcws, w.tcws_bytes, 0,
                               stream);
      if (rc != PC_OK) return rc;
    }
    int rc = tc_grouped_gemm(second.data(), nullptr, (int)second.size(), w.tcws, w.tcws_bytes, 0,
                             stream);
    if (rc != PC_OK) return rc;
  } else {
    FdGemm g{};
    g.alpha = 1.f;
    if (!gram) {
      g.a = g.b = w.fm; g.cin = nullptr; g.c = w.cmat;
      g.a_bs = g.b_bs = (int64_t)d * m; g.c_bs = (int64_t)d * d;
      g.a_si = g.b_sj = m; g.a_sk = g.b_sk = 1; g.c_si = d;
      g.m = g.n = d; g.k = m; g.beta = 0.f;
      fd_gemm(g, batch, stream);
    }
    g = FdGemm{};
    g.alpha = 1.f; g.beta = 1.f;
    g.a = g.b = w.bs; g.cin = w.cmat; g.c = w.cmat;
    g.a_bs = g.b_bs = (int64_t)d * rank; g.c_bs = (int64_t)d * d;
    g.a_si = g.b_sj = rank; g.a_sk = g.b_sk = 1; g.c_si = d;
    g.m = g.n = d; g.k = rank;
    fd_gemm(g, batch, stream);
  }

  const float* vt_final = nullptr;
  int nv = 0;
  if (!pl.subspace) {
    // ---- exact: all eigenpairs of C ----
    int rc = jacobi(w.cmat, w.vt, d);
    if (rc != PC_OK) return rc;
    fd_sort_kernel<<<batch, 512, 0, stream>>>(w.theta, d, w.order, w.sorted);
    fd_gather_rows_kernel<<<dim3(rank + 1, batch), 256, 0, stream>>>(w.vt, w.order, d, d, rank + 1,
                                                                    w.vtop);
    count_launch(2);
    vt_final = w.vtop;
    nv = rank + 1;
  } else {
    // ---- block subspace iteration with Rayleigh-Ritz ----
    const int iters = std::max(opt->subspace_iters, 1);
    const int kl = w.k_ld;
    const bool tc = kl != k;  // long-contraction products on the tcgen05 grouped GEMM
    if (tc) {  // rows k .. k_ld - 1 of the blocks are zero and stay zero
      const size_t blk = (size_t)batch * kl * d * sizeof(float);
      PC_CUDA_CHECK(cudaMemsetAsync(w.qt, 0, blk, stream));
      PC_CUDA_CHECK(cudaMemsetAsync(w.pt, 0, blk, stream));
      // (yt: fd_prepare_kernel wrote rows < k; clear the padding rows)
      for (int b = 0; b < batch; ++b)
        PC_CUDA_CHECK(cudaMemsetAsync(w.yt + ((size_t)b * kl + k) * d, 0,
                                      (size_t)(kl - k) * d * sizeof(float), stream));
    }
    std::vector<pc_gemm_desc> d_gram_y, d_gram_q, d_ritz, d_c, d_x, d_x2, d_rot;
    bool plan_gram = false, plan_c = false, plan_x = false;
    // L^-1 as a matrix + one GEMM instead of the forward substitution (PC_FD_TRSM=1: old path)
    const char* trsm_env = getenv("PC_FD_TRSM");
    const bool inv_gemm = tc && fd_tri_inverse_fits(k) && !(trsm_env && trsm_env[0] == '1');
    if (inv_gemm) {
      PC_CUDA_CHECK(cudaMemsetAsync(w.linv, 0, (size_t)batch * kl * kl * sizeof(float), stream));
      PC_CUDA_CHECK(cudaMemsetAsync(w.zsel, 0, (size_t)batch * kl * k * sizeof(float), stream));
      fd_left_descs(w.linv, kl, (int64_t)kl * kl, w.yt, w.qt, batch, kl, kl, kl, d, &d_x);
      fd_left_descs(w.linv, kl, (int64_t)kl * kl, w.qt, w.pt, batch, kl, kl, kl, d, &d_x2);
      fd_left_descs(w.zsel, k, (int64_t)kl * k, w.qt, w.vtop, batch, kl, k, kl, d, &d_rot);
    }
    // Qt <- L^-1 X for X = Yt (plan reused) or X = Qt (through Pt, which is free at that point)
    auto solve_rows = [&](const float* x) -> int {
      if (!inv_gemm) return fd_trsm_rows(w.small, x, w.qt, k, kl, d, batch, stream);
      int rc = fd_tri_inverse(w.small, w.linv, k, kl, batch, stream);
      if (rc != PC_OK) return rc;
      if (x == w.yt) {
        rc = tc_grouped_gemm(d_x.data(), nullptr, batch, w.tcws_x, w.tcws_x_bytes, plan_x ? 1 : 0,
                             stream);
        plan_x = true;
        return rc;
      }
      rc = tc_grouped_gemm(d_x2.data(), nullptr, batch, w.tcws_x2, w.tcws_x2_bytes, 0, stream);
      if (rc != PC_OK) return rc;
      PC_CUDA_CHECK(cudaMemcpyAsync(w.qt, w.pt, (size_t)batch * kl * d * sizeof(float),
                                    cudaMemcpyDeviceToDevice, stream));
      return PC_OK;
    };
    if (tc) {
      fd_sub_descs(w.yt, w.yt, w.gram_pad, batch, kl, d, &d_gram_y);
      fd_sub_descs(w.qt, w.qt, w.gram_pad, batch, kl, d, &d_gram_q);
      fd_sub_descs(w.qt, w.pt, w.gram_pad, batch, kl, d, &d_ritz);
      fd_timesc_descs(w.qt, w.cmat, w.pt, batch, kl, d, &d_c);
    }
    auto compact = [&]() {
      fd_compact_kernel<<<dim3(64, batch), 256, 0, stream>>>(w.gram_pad, kl, w.small, k);
      count_launch(1);
    };
    auto gemm_small_from_rows = [&](const float* x, const float* y, float* dst) -> int {
      if (tc) {  // dst [k,k] = X Y^T over the long dimension d (22-bit split products)
        int rc;
        if (x == w.yt) {
          rc = tc_grouped_gemm(d_gram_y.data(), nullptr, batch, w.tcws_gram, w.tcws_gram_bytes,
                               plan_gram ? 1 : 0, stream);
          plan_gram = true;
        } else {
          std::vector<pc_gemm_desc>& dd = (y == w.pt) ? d_ritz : d_gram_q;
          rc = tc_grouped_gemm(dd.data(), nullptr, batch, w.tcws_misc, w.tcws_misc_bytes, 0, stream);
        }
        if (rc != PC_OK) return rc;
        compact();
        return PC_OK;
      }
      FdGemm q{};
      q.alpha = 1.f; q.a = x; q.b = y; q.c = dst;
      q.a_bs = q.b_bs = (int64_t)kl * d; q.c_bs = (int64_t)k * k;
      q.a_si = q.b_sj = d; q.a_sk = q.b_sk = 1; q.c_si = k;
      q.m = q.n = k; q.k = d;
      fd_gemm(q, batch, stream);
      return PC_OK;
    };
    auto rotate_rows = [&](const float* coef, int rows, int64_t coef_bs, const float* rs,
                           const float* x, float* dst, int64_t dst_bs) {
      FdGemm q{};  // dst [rows, d] = diag(rs) coef [rows, k] X [k, d]
      q.alpha = 1.f; q.a = coef; q.b = x; q.c = dst; q.row_scale = rs;
      q.a_bs = coef_bs; q.b_bs = (int64_t)kl * d; q.c_bs = dst_bs; q.rs_bs = k;
      q.a_si = k; q.a_sk = 1; q.b_sj = 1; q.b_sk = d; q.c_si = d;
      q.m = rows; q.n = d; q.k = k;
      fd_gemm(q, batch, stream);
    };
    // Pt = Qt C (rows of Pt = C q_i); always qt -> pt so that the tensor-core plan is reused
    auto times_c = [&]() -> int {
      if (tc) {
        // C does not change during the call: after the first product its packed planes stay
        int rc = tc_grouped_gemm(d_c.data(), nullptr, batch, w.tcws_c, w.tcws_c_bytes,
                                 plan_c ? 3 : 0, stream);
        plan_c = true;
        return rc;
      }
      FdGemm q{};
      q.alpha = 1.f; q.a = w.qt; q.b = w.cmat; q.c = w.pt;
      q.a_bs = (int64_t)kl * d; q.b_bs = (int64_t)d * d; q.c_bs = (int64_t)kl * d;
      q.a_si = d; q.a_sk = 1; q.b_sj = d; q.b_sk = 1; q.c_si = d;
      q.m = k; q.n = d; q.k = d;
      fd_gemm(q, batch, stream);
      return PC_OK;
    };
    auto next_block_from_pt = [&]() -> int {  // Yt <- Pt
      PC_CUDA_CHECK(cudaMemcpyAsync(w.yt, w.pt, (size_t)batch * kl * d * sizeof(float),
                                    cudaMemcpyDeviceToDevice, stream));
      return PC_OK;
    };
    // PC_FD_RR_EVERY=1 restores a Rayleigh-Ritz eigen-solve in every iteration (round 1)
    const char* rr_env = getenv("PC_FD_RR_EVERY");
    const bool rr_every = rr_env && rr_env[0] == '1';
    for (int it = 0; it < iters; ++it) {
      const bool last = it + 1 == iters;
      int rc;
      fd_row_normalize_kernel<<<dim3(k, batch), 256, 0, stream>>>(w.yt, kl, d);
      count_launch(1);
      if ((rc = gemm_small_from_rows(w.yt, w.yt, w.small)) != PC_OK) return rc;
      if (!rr_every) {
        // Shifted Cholesky QR (unit rows: the shift is relative to a unit diagonal).  Between
        // iterations only the SPAN of the block matters, one pass is enough; the last block is
        // orthonormalised twice (Cholesky QR 2: orthogonality ~1e-6) before the Rayleigh-Ritz
        // solve, which is the only k x k eigen-solve left.
        rc = fd_cholesky_shift(w.small, k, batch, 1e-5f, stream);
        if (rc == PC_OK) rc = solve_rows(w.yt);
        if (rc != PC_OK) return rc;
        if (!last) {
          if ((rc = times_c()) != PC_OK) return rc;
          if ((rc = next_block_from_pt()) != PC_OK) return rc;
          continue;
        }
        if ((rc = gemm_small_from_rows(w.qt, w.qt, w.small)) != PC_OK) return rc;
        rc = fd_cholesky_shift(w.small, k, batch, 0.f, stream);
        if (rc == PC_OK) rc = solve_rows(w.qt);
        if (rc != PC_OK) return rc;
      } else {
        // round-1 path: orthonormalise through an eigen-solve of the Gram matrix
        rc = jacobi(w.small, w.vt, k, true);
        if (rc != PC_OK) return rc;
        fd_orth_scale_kernel<<<batch, 256, 0, stream>>>(w.theta, k, w.rscale);
        count_launch(1);
        rotate_rows(w.vt, k, (int64_t)k * k, w.rscale, w.yt, w.qt, (int64_t)kl * d);
      }
      if ((rc = times_c()) != PC_OK) return rc;
      // Rayleigh-Ritz: T = Qt Pt^T, eigh -> Zt (rows = Ritz coefficient vectors)
      if ((rc = gemm_small_from_rows(w.qt, w.pt, w.small)) != PC_OK) return rc;
      rc = jacobi(w.small, w.vt, k);
      if (rc != PC_OK) return rc;
      if (!last)  // next block: Yt = Zt Pt = (C Q Z)^T
        rotate_rows(w.vt, k, (int64_t)k * k, nullptr, w.pt, w.yt, (int64_t)kl * d);
    }
    fd_sort_kernel<<<batch, 512, 0, stream>>>(w.theta, k, w.order, w.sorted);
    if (inv_gemm) {
      // selected Ritz vectors into the first rank + 1 rows of the zero [k_ld, k] block, then
      // Vtop = Zsel Qt on the tensor cores (rows >= rank + 1 of the result are zero)
      fd_gather_rows_pad_kernel<<<dim3(rank + 1, batch), 256, 0, stream>>>(w.vt, w.order, k, k, kl,
                                                                          w.zsel);
      count_launch(2);
      int rc = tc_grouped_gemm(d_rot.data(), nullptr, batch, w.tcws_x2, w.tcws_x2_bytes, 0, stream);
      if (rc != PC_OK) return rc;
      vt_final = w.vtop;
      nv = kl;  // only the batch stride of the rows is taken from nv
    } else {
      fd_gather_rows_kernel<<<dim3(rank + 1, batch), 256, 0, stream>>>(w.vt, w.order, k, k, rank + 1,
                                                                      w.zsel);
      count_launch(2);
      rotate_rows(w.zsel, rank + 1, (int64_t)(rank + 1) * k, nullptr, w.qt, w.vtop,
                  (int64_t)(rank + 1) * d);
      vt_final = w.vtop;
      nv = rank + 1;
    }
  }
  fd_finalize_kernel<<<batch, 256, sizeof(float) * 4 * rank, stream>>>(
      vt_final, nv, w.sorted, k, w.scal, d, rank, out, metrics, tf);
  count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

// ---------------------------------------------------------------------------
// packed low-rank preconditioner -> the dense operator it applies (DS:1690-1705):
//   g -> c (g - g V V^T) + (g V lambda^-) V^T  ==  g (c I + V diag(lambda^- - c) V^T),
// identity when has_zeros is set (the reference bypasses the block then).
// ---------------------------------------------------------------------------
__global__ void fd_lowrank_scale_kernel(const float* __restrict__ packed, int d, int r,
                                        float* __restrict__ ws, float* __restrict__ c_out) {
  const int b = blockIdx.y, pd = r + 2;
  const float* P = packed + (size_t)b * d * pd;
  const float c = P[r + 1];
  const bool skip = P[(size_t)(d - 1) * pd + r] != 0.f;
  if (c_out && blockIdx.x == 0 && threadIdx.x == 0) c_out[b] = skip ? 1.0f : c;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)d * r;
       e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / r), j = (int)(e - (size_t)i * r);
    ws[(size_t)b * d * r + e] = skip ? 0.f : P[(size_t)i * pd + j] * (P[(size_t)j * pd + r] - c);
  }
}
__global__ void fd_add_diag_kernel(const float* __restrict__ packed, int d, int r,
                                   float* __restrict__ dense) {
  const int b = blockIdx.y, pd = r + 2;
  const float* P = packed + (size_t)b * d * pd;
  const bool skip = P[(size_t)(d - 1) * pd + r] != 0.f;
  const float c = skip ? 1.0f : P[r + 1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < d; i += gridDim.x * blockDim.x)
    dense[((size_t)b * d + i) * d + i] += c;
}

// dense = (V diag(lambda^- - c)) V^T: descriptors of the tcgen05 grouped GEMM (d % 128 == 0)
static void fd_dense_descs(const float* scaled, const float* packed, float* dense, int batch,
                           int d, int rank, std::vector<pc_gemm_desc>* out) {
  for (int b = 0; b < batch; ++b) {
    pc_gemm_desc g{};
    g.a = scaled + (size_t)b * d * rank; g.b = packed + (size_t)b * d * (rank + 2);
    g.c = dense + (size_t)b * d * d; g.c_in = nullptr;
    g.a_iinner = d; g.a_sio = 0; g.a_si = rank; g.a_kinner = g.b_kinner = rank;
    g.a_sko = g.b_sko = 0; g.a_ski = g.b_ski = 1; g.b_sj = rank + 2;
    g.c_iinner = d; g.c_sio = 0; g.c_sii = d;
    g.m = g.n = d; g.k = rank; g.alpha = 1.f; g.beta = 0.f;
    out->push_back(g);
  }
}

size_t low_rank_to_dense_bytes(int batch, int d, int rank) {
  size_t need = align_up((size_t)batch * d * rank * sizeof(float), 256) + 512;
  if (fd_use_tc(d)) {
    std::vector<pc_gemm_desc> g;
    fd_dense_descs(nullptr, reinterpret_cast<const float*>(1), nullptr, batch, d, rank, &g);
    need += tc_grouped_gemm_workspace_bytes(g.data(), (int)g.size()) + 1024;
  }
  return need;
}

int run_low_rank_to_dense(const float* packed, int batch, int d, int rank, float* dense,
                          void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  const size_t need = low_rank_to_dense_bytes(batch, d, rank);
  if (workspace_bytes < need) {
    set_error("low-rank workspace too small: %zu < %zu", workspace_bytes, need);
    return PC_ERR_WORKSPACE;
  }
  float* ws = reinterpret_cast<float*>(align_up((size_t)workspace, 256));
  const unsigned grid = (unsigned)std::min<size_t>(((size_t)d * rank + 255) / 256, 512);
  fd_lowrank_scale_kernel<<<dim3(grid, batch), 256, 0, stream>>>(packed, d, rank, ws, nullptr);
  if (fd_use_tc(d)) {
    std::vector<pc_gemm_desc> descs;
    fd_dense_descs(ws, packed, dense, batch, d, rank, &descs);
    char* tcws = reinterpret_cast<char*>(ws) + align_up((size_t)batch * d * rank * sizeof(float), 256);
    const size_t tcb = workspace_bytes - (size_t)(tcws - reinterpret_cast<char*>(workspace));
    int rc = tc_grouped_gemm(descs.data(), nullptr, batch, tcws, tcb, 0, stream);
    if (rc != PC_OK) return rc;
    fd_add_diag_kernel<<<dim3((d + 255) / 256, batch), 256, 0, stream>>>(packed, d, rank, dense);
    count_launch(2);
    PC_CUDA_CHECK(cudaGetLastError());
    return PC_OK;
  }
  FdGemm g{};
  g.alpha = 1.f;
  g.a = ws; g.b = packed; g.c = dense;
  g.a_bs = (int64_t)d * rank; g.b_bs = (int64_t)d * (rank + 2); g.c_bs = (int64_t)d * d;
  g.a_si = rank; g.a_sk = 1; g.b_sj = rank + 2; g.b_sk = 1; g.c_si = d;
  g.m = g.n = d; g.k = rank;
  fd_gemm(g, batch, stream);
  fd_add_diag_kernel<<<dim3((d + 255) / 256, batch), 256, 0, stream>>>(packed, d, rank, dense);
  count_launch(2);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}


// ===========================================================================
// eigh-based low-rank root (compression_rank != 0 without frequent_directions):
// _low_rank_root, DS:1033-1120.  reg = masked A + ridge I_m; all eigenpairs by the cluster
// Jacobi kernel (d <= 512); keep the |rank| largest (rank > 0) or smallest (rank < 0)
// eigenvalues inverted to the power -1/p, average the rest into `const`, pack (DS:548-552).
// The reported error is max |U^T reg U - diag(e)| like DS:1076-1081.
// ===========================================================================
int run_power_iteration(const float* xs, const int32_t* pads, int batch, int n, int num_iters,
                        float tol, float* lambdas, int32_t* iters, struct RootCtl* ctl,
                        float* v0_dev, float* ybuf, cudaStream_t stream);  // root.cu

// reg (two copies) <- masked A; lambda-independent part of DS:1052-1060
__global__ void lr_mask_kernel(const float* __restrict__ xs, const int32_t* __restrict__ pads,
                               int d, float* __restrict__ a_masked) {
  const int b = blockIdx.y;
  const int pad = pads ? min(max(pads[b], 0), d) : d;
  const size_t nn = (size_t)d * d;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn;
       e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / d), j = (int)(e - (size_t)i * d);
    // lower triangle authoritative, like the Newton solver
    const float v = xs[(size_t)b * nn + (size_t)max(i, j) * d + min(i, j)];
    a_masked[(size_t)b * nn + e] = (i < pad && j < pad) ? v : 0.f;
  }
}
// reg = a + ridge I_m (in place), copy kept for the error check; ridge per matrix (DS:1069)
__global__ void lr_damp_kernel(float* __restrict__ a, float* __restrict__ copy,
                               const float* __restrict__ lambdas, const int32_t* __restrict__ pads,
                               int d, float ridge_epsilon, float error_tolerance, int relative,
                               float* __restrict__ ridge_out) {
  const int b = blockIdx.y;
  const int pad = pads ? min(max(pads[b], 0), d) : d;
  const float max_ev = relative ? lambdas[b] : 1.0f;
  const float ridge = ridge_epsilon * fmaxf(max_ev, error_tolerance);
  if (blockIdx.x == 0 && threadIdx.x == 0) ridge_out[b] = ridge;
  const size_t nn = (size_t)d * d;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn;
       e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / d), j = (int)(e - (size_t)i * d);
    float v = a[(size_t)b * nn + e];
    if (i == j && i < pad) v += ridge;
    a[(size_t)b * nn + e] = v;
    copy[(size_t)b * nn + e] = v;
  }
}
// Inverse roots are dominated by the SMALL eigenvalues.  One-sided Jacobi on the matrix itself
// leaves every row with absolute noise ~ eps * sqrt(#rotations) * theta_max, i.e. the rows of
// small eigenvalues (norm theta_i) lose their direction (measured: 1 % error in the root at
// cond 1e4).  Run it on the Cholesky factor instead: reg = L L^T, the rows of G = L^T become
// sigma_i u_i^T with sigma_i = sqrt(theta_i), which squares the conditioning away.
// One CTA per matrix, right-looking, in place on the leading pad x pad block; g <- L^T (upper
// triangular, zero elsewhere).  A non-positive pivot (matrix not positive definite) poisons
// the matrix with NaN, which surfaces as a NaN error metric -> failure fallback.
__global__ void __launch_bounds__(1024)
lr_cholesky_kernel(float* __restrict__ reg_all, const int32_t* __restrict__ pads, int d,
                   float* __restrict__ g_all) {
  __shared__ float piv_s;
  const int b = blockIdx.x;
  const int pad = pads ? min(max(pads[b], 0), d) : d;
  float* A = reg_all + (size_t)b * d * d;
  float* G = g_all + (size_t)b * d * d;
  for (int k = 0; k < pad; ++k) {
    if (threadIdx.x == 0) {
      const float p = A[(size_t)k * d + k];
      piv_s = p > 0.f ? sqrtf(p) : __int_as_float(0x7fc00000);
      A[(size_t)k * d + k] = piv_s;
    }
    __syncthreads();
    const float inv = 1.0f / piv_s;
    for (int i = k + 1 + threadIdx.x; i < pad; i += blockDim.x) A[(size_t)i * d + k] *= inv;
    __syncthreads();
    const int m = pad - k - 1;  // trailing block: rows / cols k+1 .. pad-1, lower part
    for (int e = threadIdx.x; e < m * m; e += blockDim.x) {
      const int i = k + 1 + e / m, j = k + 1 + e % m;
      if (j <= i) A[(size_t)i * d + j] -= A[(size_t)i * d + k] * A[(size_t)j * d + k];
    }
    __syncthreads();
  }
  for (size_t e = threadIdx.x; e < (size_t)d * d; e += blockDim.x) {
    const int i = (int)(e / d), j = (int)(e - (size_t)i * d);
    G[e] = (i < pad && j < pad && j >= i) ? A[(size_t)j * d + i] : 0.f;  // G = L^T
  }
}
// G = L^T from the lower triangle of an in-place factorisation (no padding)
__global__ void lr_upper_from_lower_kernel(const float* __restrict__ l_all, int d,
                                           float* __restrict__ g_all) {
  const int b = blockIdx.y;
  const size_t nn = (size_t)d * d;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn;
       e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / d), j = (int)(e - (size_t)i * d);
    g_all[(size_t)b * nn + e] = j >= i ? l_all[(size_t)b * nn + (size_t)j * d + i] : 0.f;
  }
}
// vs[b][t][:] = rows[b][order[t]][:] / |row|  (u_i = row_i / sigma_i; zero rows stay zero)
__global__ void __launch_bounds__(256)
lr_gather_normalize_kernel(const float* __restrict__ rows, const int* __restrict__ order, int d,
                           float* __restrict__ vs) {
  __shared__ float red[32];
  const int b = blockIdx.y, t = blockIdx.x;
  const float* src = rows + ((size_t)b * d + order[(size_t)b * d + t]) * d;
  float ss = 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x) ss = fmaf(src[i], src[i], ss);
  const float nrm = sqrtf(block_sum(ss, red));
  const float inv = nrm > 0.f ? 1.0f / nrm : 0.f;
  for (int i = threadIdx.x; i < d; i += blockDim.x)
    vs[((size_t)b * d + t) * d + i] = src[i] * inv;
}
// err[b] = max over rows i and real eigen-columns j (sorted rank < pad) of
// |recovered[i][j] - delta_ij e_j|   (DS:1076-1081)
__global__ void __launch_bounds__(256)
lr_error_kernel(const float* __restrict__ recovered, const float* __restrict__ sorted,
                const int32_t* __restrict__ pads, int d, uint32_t* __restrict__ errbits) {
  const int b = blockIdx.y;
  const int pad = pads ? min(max(pads[b], 0), d) : d;
  uint32_t mx = 0;
  const size_t nn = (size_t)d * d;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn;
       e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / d), j = (int)(e - (size_t)i * d);
    if (j >= pad) continue;
    const float v = recovered[(size_t)b * nn + e] - (i == j ? sorted[(size_t)b * d + j] : 0.f);
    const uint32_t ab = absbits(v);
    mx = ab > mx ? ab : mx;
  }
  mx = warp_max_u32(mx);
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(errbits + b, mx);
}
// selection, inversion, averaging and packing: DS:1083-1112, DS:548-552
__global__ void __launch_bounds__(256)
lr_pack_kernel(const float* __restrict__ vs, const float* __restrict__ sorted,
               const float* __restrict__ ridge_all, const int32_t* __restrict__ ps,
               const int32_t* __restrict__ pads, const uint32_t* __restrict__ errbits, int d,
               int rank_signed, float* __restrict__ out_all, float* __restrict__ metrics) {
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int k = rank_signed < 0 ? -rank_signed : rank_signed, pd = k + 2;
  const int pad = pads ? min(max(pads[b], 0), d) : d;
  const float ridge = ridge_all[b];
  const float alpha = -1.0f / (float)ps[b];
  const float* e = sorted + (size_t)b * d;            // descending
  const float* V = vs + (size_t)b * d * d;             // rows = eigenvectors, descending order
  float* out = out_all + (size_t)b * d * pd;
  auto inv_e = [&](int idx) {  // idx in descending order; padded eigenvalues (idx >= pad) are 0
    return idx < pad ? powf(fmaxf(e[idx], ridge), alpha) : 0.f;
  };
  // kept slot t <-> descending index: rank > 0: t; rank < 0: pad - 1 - t (ascending real ones)
  auto kept_idx = [&](int t) { return rank_signed > 0 ? t : pad - 1 - t; };
  float part = 0.f;  // sum of inv_e over everything that is not kept (DS:1101-1106)
  for (int idx = threadIdx.x; idx < pad; idx += blockDim.x) {
    const bool kept = rank_signed > 0 ? idx < k : idx >= pad - k;
    if (!kept) part += inv_e(idx);
  }
  const float total = block_sum(part, red);
  const int n_avg = pad - k;
  const float cst = total / (n_avg > 0 ? (float)n_avg : 1.0f);
  const bool zero_all = pad == 0;  // DS:1115-1118
  for (size_t x = threadIdx.x; x < (size_t)d * pd; x += blockDim.x) {
    const int i = (int)(x / pd), j = (int)(x - (size_t)i * pd);
    float v = 0.f;
    if (!zero_all) {
      if (j < k) {
        const int idx = kept_idx(j);
        v = (idx >= 0 && idx < d) ? V[(size_t)idx * d + i] : 0.f;
      } else if (j == k) {
        if (i < k) { const int idx = kept_idx(i); v = (idx >= 0 && idx < d) ? inv_e(idx) : 0.f; }
      } else if (i == 0) {
        v = cst;
      }
    }
    out[x] = v;
  }
  if (threadIdx.x == 0 && metrics) {
    float* m = metrics + (size_t)b * PC_NUM_METRICS;
    m[0] = zero_all ? 0.f : __uint_as_float(errbits[b]);
    m[1] = 0.f; m[2] = 0.f; m[3] = 0.f; m[4] = 0.f;
  }
}

size_t low_rank_root_bytes(int batch, int d) {
  const size_t B = (size_t)batch, nn = (size_t)d * d * 4;
  return 6 * align_up(B * nn, 256) + 4 * align_up(B * d * 4, 256) + align_up(B * kJacCnt * 4, 256) +
         align_up((size_t)d * 4, 256) + align_up(2 * B * d * 4, 256) + 2048;
}

// `full_root`: instead of packing, form U diag(inv_e) U^T (matrix_inverse_pth_root_eigh,
// DS:943-1030) into out [batch, d, d].
__global__ void lr_scale_rows_kernel(const float* __restrict__ vs, const float* __restrict__ sorted,
                                     const float* __restrict__ ridge_all,
                                     const int32_t* __restrict__ ps,
                                     const int32_t* __restrict__ pads, int d,
                                     float pinv_cutoff, float* __restrict__ rt) {
  const int b = blockIdx.y, idx = blockIdx.x;  // row idx = eigenvector of descending rank idx
  const int pad = pads ? min(max(pads[b], 0), d) : d;
  const float alpha = -1.0f / (float)ps[b];
  float inv_e;
  if (pinv_cutoff > 0.f) {
    // pseudo-inverse root (TF/shampoo.py:440-448): the shift that made the factorisation
    // possible comes off again; eigenvalues <= cutoff * max are dropped, not clamped
    const float ridge = ridge_all[b];
    const float w = sorted[(size_t)b * d + idx] - ridge, wmax = sorted[(size_t)b * d] - ridge;
    inv_e = (w <= pinv_cutoff * wmax) ? 0.f : powf(w, alpha);
  } else {
    inv_e = idx < pad ? powf(fmaxf(sorted[(size_t)b * d + idx], ridge_all[b]), alpha) : 0.f;
  }
  const float sc = sqrtf(inv_e);  // root = u * sqrt(inv_e), DS:1018
  for (int i = threadIdx.x; i < d; i += blockDim.x)
    rt[((size_t)b * d + idx) * d + i] = vs[((size_t)b * d + idx) * d + i] * sc;
}
__global__ void lr_root_metrics_kernel(const uint32_t* __restrict__ errbits,
                                       const int32_t* __restrict__ pads, int d, int batch,
                                       float* __restrict__ out, float* __restrict__ metrics) {
  const int b = blockIdx.x;
  const int pad = pads ? min(max(pads[b], 0), d) : d;
  if (pad == 0)  // DS:1026-1030
    for (size_t e = threadIdx.x; e < (size_t)d * d; e += blockDim.x) out[(size_t)b * d * d + e] = 0.f;
  if (threadIdx.x == 0 && metrics) {
    float* m = metrics + (size_t)b * PC_NUM_METRICS;
    m[0] = pad == 0 ? 0.f : __uint_as_float(errbits[b]);
    m[1] = 0.f; m[2] = 0.f; m[3] = 0.f; m[4] = 0.f;
  }
}

int run_low_rank_root(const float* xs, const int32_t* ps, const int32_t* pads, int batch, int d,
                      int rank_signed, bool full_root, float ridge_epsilon, float error_tolerance,
                      int relative, float* out, float* metrics, void* workspace,
                      size_t workspace_bytes, cudaStream_t stream, float pinv_cutoff = 0.f,
                      float* warm = nullptr, int warm_valid = 0) {
  // `warm` [batch, d, d] (rows = eigenvectors of the previous call's matrix, in / out): the
  // solve then runs on M = W reg W^T, which is nearly diagonal when the matrix moved little
  // since -- its Cholesky factor has nearly orthogonal rows and the Jacobi sweeps drop from
  // ~8 to 2-3 (quadratic convergence); eigenvectors of reg = (eigenvectors of M) W.
  const bool use_warm = warm != nullptr && warm_valid != 0 && pads == nullptr;
  if (workspace_bytes < low_rank_root_bytes(batch, d)) {
    set_error("eigh root workspace too small: %zu < %zu", workspace_bytes,
              low_rank_root_bytes(batch, d));
    return PC_ERR_WORKSPACE;
  }
  const size_t B = (size_t)batch, nn = (size_t)d * d;
  char* w = reinterpret_cast<char*>(align_up((size_t)workspace, 256));
  auto take = [&](size_t bytes) { char* p = w; w += align_up(bytes, 256); return p; };
  float* reg = (float*)take(B * nn * 4);       // destroyed by the factorisation
  float* reg_copy = (float*)take(B * nn * 4);
  float* vt = (float*)take(B * nn * 4);        // G = L^T, then sigma_i u_i^T as rows
  float* vs = (float*)take(B * nn * 4);        // eigenvectors as rows, sorted descending
  float* t1 = (float*)take(B * nn * 4);
  float* rec = (float*)take(B * nn * 4);
  float* theta = (float*)take(B * d * 4);
  float* sorted = (float*)take(B * d * 4);
  int* order = (int*)take(B * d * 4);
  float* lambdas = (float*)take(B * d * 4);    // [batch] lambdas, [batch] ridge, [batch] err bits
  unsigned* rot = (unsigned*)take(B * kJacCnt * 4);
  float* v0 = (float*)take((size_t)d * 4);
  float* ybuf = (float*)take(2 * B * d * 4);
  float* ridge = lambdas + batch;
  uint32_t* errbits = reinterpret_cast<uint32_t*>(lambdas + 2 * (size_t)batch);
  const unsigned g = (unsigned)std::min<size_t>((nn + 255) / 256, 512);
  PC_CUDA_CHECK(cudaMemsetAsync(rot, 0, B * kJacCnt * 4, stream));
  PC_CUDA_CHECK(cudaMemsetAsync(lambdas, 0, B * d * 4, stream));
  lr_mask_kernel<<<dim3(g, batch), 256, 0, stream>>>(xs, pads, d, reg);
  if (relative) {
    int rc = run_power_iteration(reg, pads, batch, d, 100, error_tolerance, lambdas, nullptr,
                                 nullptr, v0, ybuf, stream);  // DS:1061-1067, DS:998-1004
    if (rc != PC_OK) return rc;
  }
  lr_damp_kernel<<<dim3(g, batch), 256, 0, stream>>>(reg, reg_copy, lambdas, pads, d,
                                                    ridge_epsilon, error_tolerance, relative, ridge);
  if (use_warm) {  // reg <- W reg W^T (reg_copy keeps the matrix itself)
    FdGemm q{};
    q.alpha = 1.f; q.a = warm; q.b = reg_copy; q.c = t1;  // t1 = W reg (reg symmetric)
    q.a_bs = q.b_bs = q.c_bs = (int64_t)nn;
    q.a_si = d; q.a_sk = 1; q.b_sj = d; q.b_sk = 1; q.c_si = d;
    q.m = q.n = q.k = d;
    fd_gemm(q, batch, stream);
    q.a = t1; q.b = warm; q.c = reg;                       // reg = t1 W^T
    fd_gemm(q, batch, stream);
  }
  // factor, then one-sided Jacobi on the rows of G = L^T: rows -> sigma_i u_i^T, theta = sigma^2
  const size_t chol_smem = ((size_t)((d + 31) & ~31) + (size_t)d * (d + 1) / 2) * sizeof(float);
  if (pads == nullptr && chol_smem <= kCholSmemMax) {
    // the packed triangle fits in shared memory: the tiled factorisation of the sketch path
    int rcc = fd_cholesky_shift(reg, d, batch, 0.f, stream);
    if (rcc != PC_OK) return rcc;
    lr_upper_from_lower_kernel<<<dim3(g, batch), 256, 0, stream>>>(reg, d, vt);
  } else {
    lr_cholesky_kernel<<<batch, 1024, 0, stream>>>(reg, pads, d, vt);
  }
  int rc = fd_jacobi(vt, nullptr, d, batch, rot, theta, stream);
  if (rc != PC_OK) return rc;
  fd_sort_kernel<<<batch, 512, 0, stream>>>(theta, d, order, sorted);
  lr_gather_normalize_kernel<<<dim3(d, batch), 256, 0, stream>>>(vt, order, d, vs);
  if (use_warm) {  // back to the original basis: Vs <- Vs W
    FdGemm q{};
    q.alpha = 1.f; q.a = vs; q.b = warm; q.c = t1;
    q.a_bs = q.b_bs = q.c_bs = (int64_t)nn;
    q.a_si = d; q.a_sk = 1; q.b_sj = 1; q.b_sk = d; q.c_si = d;
    q.m = q.n = q.k = d;
    fd_gemm(q, batch, stream);
    PC_CUDA_CHECK(cudaMemcpyAsync(vs, t1, B * nn * 4, cudaMemcpyDeviceToDevice, stream));
  }
  if (warm != nullptr && pads == nullptr)
    PC_CUDA_CHECK(cudaMemcpyAsync(warm, vs, B * nn * 4, cudaMemcpyDeviceToDevice, stream));
  if (pinv_cutoff <= 0.f) {
    // recovered = Vs reg Vs^T
    FdGemm q{};
    q.alpha = 1.f; q.a = vs; q.b = reg_copy; q.c = t1;
    q.a_bs = q.b_bs = q.c_bs = (int64_t)nn;
    q.a_si = d; q.a_sk = 1; q.b_sj = d; q.b_sk = 1; q.c_si = d;
    q.m = q.n = q.k = d;
    fd_gemm(q, batch, stream);
    q.a = t1; q.b = vs; q.c = rec;
    fd_gemm(q, batch, stream);
    lr_error_kernel<<<dim3(g, batch), 256, 0, stream>>>(rec, sorted, pads, d, errbits);
  }
  if (full_root) {
    // val = root root^T with root = U sqrt(inv_e): out(i,j) = sum_k Rt(k,i) Rt(k,j)
    lr_scale_rows_kernel<<<dim3(d, batch), 256, 0, stream>>>(vs, sorted, ridge, ps, pads, d,
                                                            pinv_cutoff, t1);
    FdGemm r{};
    r.alpha = 1.f; r.a = t1; r.b = t1; r.c = out;
    r.a_bs = r.b_bs = r.c_bs = (int64_t)nn;
    r.a_si = 1; r.a_sk = d; r.b_sj = 1; r.b_sk = d; r.c_si = d;
    r.m = r.n = r.k = d;
    fd_gemm(r, batch, stream);
    lr_root_metrics_kernel<<<batch, 256, 0, stream>>>(errbits, pads, d, batch, out, metrics);
  } else {
    lr_pack_kernel<<<batch, 256, 0, stream>>>(vs, sorted, ridge, ps, pads, errbits, d, rank_signed,
                                             out, metrics);
  }
  count_launch(8);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

}  // namespace pc

extern "C" {

size_t pc_low_rank_to_dense_workspace_bytes(int batch, int d, int rank) {
  if (batch <= 0 || d <= 0 || rank <= 0) return 0;
  return pc::low_rank_to_dense_bytes(batch, d, rank) + 512;
}

int pc_low_rank_factors(const float* packed, int batch, int d, int rank, float* w, float* c,
                        void* stream) {
  PC_REQUIRE(batch >= 0 && d > 0 && rank > 0 && rank + 2 < d, "bad low-rank sizes");
  if (batch == 0) return PC_OK;
  PC_REQUIRE(packed && w && c, "null pointer argument");
  const unsigned grid = (unsigned)std::min<size_t>(((size_t)d * rank + 255) / 256, 512);
  pc::fd_lowrank_scale_kernel<<<dim3(grid, batch), 256, 0, (cudaStream_t)stream>>>(packed, d, rank,
                                                                                  w, c);
  pc::count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

int pc_low_rank_to_dense(const float* packed, int batch, int d, int rank, float* dense,
                         void* workspace, size_t workspace_bytes, void* stream) {
  PC_REQUIRE(batch >= 0 && d > 0 && rank > 0 && rank + 2 < d, "bad low-rank sizes");
  if (batch == 0) return PC_OK;
  PC_REQUIRE(packed && dense && workspace, "null pointer argument");
  return pc::run_low_rank_to_dense(packed, batch, d, rank, dense, workspace, workspace_bytes,
                                   (cudaStream_t)stream);
}

size_t pc_low_rank_root_workspace_bytes(int batch, int d) {
  if (batch <= 0 || d <= 0) return 0;
  return pc::low_rank_root_bytes(batch, d) + 512;
}

int pc_low_rank_root_batched(const float* xs, const int32_t* ps, const int32_t* padding_starts,
                             int batch, int d, int compression_rank, float ridge_epsilon,
                             float error_tolerance, int relative_matrix_epsilon, float* out,
                             float* metrics, void* workspace, size_t workspace_bytes,
                             void* stream) {
  const int k = compression_rank < 0 ? -compression_rank : compression_rank;
  PC_REQUIRE(batch >= 0 && d > 0 && k > 0, "bad low-rank root sizes");
  if (batch == 0) return PC_OK;
  PC_REQUIRE(xs && ps && out && workspace, "null pointer argument");
  PC_REQUIRE(k + 2 < d, "low-rank root needs |rank| + 2 < d (DS:535-537), got rank=%d d=%d",
             compression_rank, d);
  PC_REQUIRE(d <= pc::kEighMaxN, "low-rank root supports d <= %d (one Jacobi solve), got %d",
             pc::kEighMaxN, d);
  return pc::run_low_rank_root(xs, ps, padding_starts, batch, d, compression_rank, false,
                               ridge_epsilon, error_tolerance, relative_matrix_epsilon, out,
                               metrics, workspace, workspace_bytes, (cudaStream_t)stream);
}

int pc_inverse_pth_root_eigh_batched(const float* xs, const int32_t* ps,
                                     const int32_t* padding_starts, int batch, int d,
                                     float ridge_epsilon, float error_tolerance,
                                     int relative_matrix_epsilon, float* roots, float* metrics,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  PC_REQUIRE(batch >= 0 && d > 0, "bad eigh root sizes");
  if (batch == 0) return PC_OK;
  PC_REQUIRE(xs && ps && roots && workspace, "null pointer argument");
  PC_REQUIRE(d <= pc::kEighMaxN, "eigh root supports d <= %d (one Jacobi solve), got %d",
             pc::kEighMaxN, d);
  return pc::run_low_rank_root(xs, ps, padding_starts, batch, d, 0, true, ridge_epsilon,
                               error_tolerance, relative_matrix_epsilon, roots, metrics, workspace,
                               workspace_bytes, (cudaStream_t)stream);
}

int pc_pinv_pth_root_eigh_batched(const float* xs, const int32_t* ps, int batch, int d,
                                  float rel_cutoff, float* roots, void* workspace,
                                  size_t workspace_bytes, void* stream) {
  return pc_pinv_pth_root_eigh_warm_batched(xs, ps, batch, d, rel_cutoff, roots, nullptr, 0,
                                            workspace, workspace_bytes, stream);
}

int pc_pinv_pth_root_eigh_warm_batched(const float* xs, const int32_t* ps, int batch, int d,
                                       float rel_cutoff, float* roots, float* eigvecs,
                                       int eigvecs_valid, void* workspace,
                                       size_t workspace_bytes, void* stream) {
  PC_REQUIRE(batch >= 0 && d > 0 && rel_cutoff > 0.f, "bad pseudo-inverse root arguments");
  if (batch == 0) return PC_OK;
  PC_REQUIRE(xs && ps && roots && workspace, "null pointer argument");
  PC_REQUIRE(d <= pc::kEighMaxN, "eigh root supports d <= %d (one Jacobi solve), got %d",
             pc::kEighMaxN, d);
  // the eigensolver factors the matrix first, which a rank-deficient statistic (the usual
  // state of the first steps) does not survive in fp32: shift by a few d * eps * lambda_max,
  // take the shift off the eigenvalues again before the cutoff
  const float shift = fmaxf(1e-6f, 4.0f * (float)d * 5.96e-8f);
  return pc::run_low_rank_root(xs, ps, nullptr, batch, d, 0, true, shift, 1e-30f, 1, roots,
                               nullptr, workspace, workspace_bytes, (cudaStream_t)stream,
                               rel_cutoff, eigvecs, eigvecs_valid);
}

void pc_fd_options_default(pc_fd_options* opt) {
  opt->ridge_epsilon = 1e-6f;
  opt->error_tolerance = 1e-6f;
  opt->relative_matrix_epsilon = 1;
  opt->decay = 1.0f;
  opt->input_is_gram = 0;
  opt->subspace_iters = 6;
  opt->oversample = 32;
  opt->full_eigh_max_dim = 512;
  opt->tearfree = 0;
  opt->tearfree_epsilon = 0.f;
  opt->tearfree_relative_epsilon = 0;
}

size_t pc_fd_update_workspace_bytes(int batch, int d, int m, int rank, const pc_fd_options* opt) {
  if (batch <= 0 || d <= 0 || rank <= 0 || !opt) return 0;
  pc::FdWorkspace w;
  return pc::fd_carve(&w, nullptr, batch, d, opt->input_is_gram ? d : m, rank,
                      pc::fd_plan(d, rank, opt), opt->input_is_gram != 0) + 512;
}

int pc_fd_update_batched(const float* new_grad, const float* prev, const int32_t* ps,
                         const int32_t* padding_starts, int batch, int d, int m, int rank,
                         const pc_fd_options* opt, float* out, float* metrics, void* workspace,
                         size_t workspace_bytes, void* stream) {
  PC_REQUIRE(batch >= 0 && d > 0 && rank > 0, "bad fd sizes");
  if (batch == 0) return PC_OK;
  PC_REQUIRE(new_grad && prev && ps && opt && out && workspace, "null pointer argument");
  PC_REQUIRE(rank + 2 < d, "frequent directions needs rank + 2 < d (DS:535-537), got rank=%d d=%d",
             rank, d);
  if (opt->input_is_gram) m = d;
  PC_REQUIRE(m > 0, "factor needs at least one column");
  PC_REQUIRE(opt->subspace_iters <= 30, "subspace_iters must be <= 30");
  PC_REQUIRE(opt->full_eigh_max_dim <= pc::kJacMaxN, "full_eigh_max_dim must be <= %d",
             pc::kJacMaxN);
  if (d > opt->full_eigh_max_dim)
    PC_REQUIRE(rank + 1 <= pc::kJacMaxN, "rank + 1 must be <= %d for the subspace path",
               pc::kJacMaxN);
  return pc::run_fd_update(new_grad, prev, ps, padding_starts, batch, d, m, rank, opt, out, metrics,
                           workspace, workspace_bytes, (cudaStream_t)stream);
}

}  // extern "C"
