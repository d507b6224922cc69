// Top-k deflation around the inverse p-th root (the `lobpcg_topk_precondition` branch of
// matrix_inverse_pth_root, DS:789-812 and DS:889-928) and the diagnostics types of DS:109-195.
// The top-k eigenpairs themselves come from the library's block subspace iteration
// (pc_fd_update_batched on the matrix as a Gram with an empty previous sketch -- the reference
// calls jax.experimental.sparse.linalg.lobpcg_standard, any accurate top-k solver serves);
// these kernels do the O(n k) / O(n^2) glue on the device so that the whole branch stays
// enqueue-only:
//   deflate_prep     eigenvalues from the packed sketch, max / min, the absolute ridge
//                    ridge_epsilon * max(max_ev, 1e-25) (DS:814-830), S1 = V sqrt((l - l_min) / m)
//                    and the matrix scaled by 1 / m (m = max eigenvalue: the solver then runs
//                    with an absolute epsilon on A / m, which is the same problem)
//   redeflate_prep   root <- root * m^(-1/p), S2 = V sqrt(pth_root_difference) (DS:681-699, DS:896-900)
//   root_diag        max / mean |diag(M) - 1| and max / mean |offdiag(M)| of M = B^p A (DS:127-141)
//   lobpcg_diag      consistency / orthogonality errors of the eigenpairs (DS:172-195)
#include <algorithm>

#include "common.cuh"

namespace pc {

// packed sketch of pc_fd_update_batched: [n, k + 2]; vectors in [:, :k], deflated eigenvalues in
// [n - k :, k + 1], tail (= the (k+1)-th eigenvalue here) in [1, k + 1]
__global__ void __launch_bounds__(256)
lob_deflate_prep_kernel(const float* __restrict__ packed, const float* __restrict__ a, int n, int k,
                        float ridge_epsilon, int relative, float* __restrict__ scal,
                        float* __restrict__ eigvals, float* __restrict__ s1,
                        float* __restrict__ a_scaled) {
  __shared__ float lam[512];
  __shared__ float mm[2];
  const int b = blockIdx.y, pd = k + 2;
  const float* P = packed + (size_t)b * n * pd;
  const float tail = P[(size_t)1 * pd + k + 1];
  for (int j = threadIdx.x; j < k; j += blockDim.x) lam[j] = P[(size_t)(n - k + j) * pd + k + 1] + tail;
  __syncthreads();
  if (threadIdx.x == 0) {
    float mx = lam[0], mn = lam[0];
    for (int j = 1; j < k; ++j) { mx = fmaxf(mx, lam[j]); mn = fminf(mn, lam[j]); }
    mm[0] = mx; mm[1] = mn;
  }
  __syncthreads();
  const float mx = mm[0], mn = mm[1];
  const float scale = mx > 0.f ? 1.0f / mx : 1.0f;
  if (blockIdx.x == 0) {
    for (int j = threadIdx.x; j < k; j += blockDim.x) eigvals[(size_t)b * k + j] = lam[j];
    if (threadIdx.x == 0) {
      float* s = scal + (size_t)b * 4;
      s[0] = mx; s[1] = mn;
      s[2] = ridge_epsilon * fmaxf(relative ? mx : 1.0f, 1e-25f);  // absolute ridge, DS:830
      s[3] = scale;
    }
  }
  const size_t nk = (size_t)n * k, nn = (size_t)n * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nk;
       e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / k), j = (int)(e - (size_t)i * k);
    // deflation = l - l_min (DS:806); the scaled problem divides everything by m
    s1[(size_t)b * nk + e] = P[(size_t)i * pd + j] * sqrtf(fmaxf(lam[j] - mn, 0.f) * scale);
  }
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn;
       e += (size_t)gridDim.x * blockDim.x)
    a_scaled[(size_t)b * nn + e] = a[(size_t)b * nn + e] * scale;
}

// (w + alpha)^(-1/p) - (w + beta)^(-1/p), DS:681-699 (the branch with the better log1p argument)
__device__ __forceinline__ float pth_root_difference(float w, float alpha, float beta, float p) {
  const float a = w + alpha, b = w + beta, amb = alpha - beta, ex = -1.0f / p;
  auto stable = [&](float base, float diff) { return powf(base, ex) * expm1f(ex * log1pf(diff / base)); };
  return fabsf(amb / b) < fabsf(amb / a) ? -stable(a, -amb) : stable(b, amb);
}

__global__ void __launch_bounds__(256)
lob_redeflate_prep_kernel(const float* __restrict__ packed, const float* __restrict__ scal,
                          const float* __restrict__ eigvals, const int32_t* __restrict__ ps, int n,
                          int k, float* __restrict__ root, float* __restrict__ s2) {
  const int b = blockIdx.y, pd = k + 2;
  const float* P = packed + (size_t)b * n * pd;
  const float* s = scal + (size_t)b * 4;
  const float mx = s[0], mn = s[1], ridge = s[2];
  const float p = (float)ps[b];
  const float back = mx > 0.f ? powf(mx, -1.0f / p) : 1.0f;  // root(A) = m^(-1/p) root(A / m)
  const size_t nk = (size_t)n * k, nn = (size_t)n * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nk;
       e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / k), j = (int)(e - (size_t)i * k);
    const float d = pth_root_difference(ridge, mn, eigvals[(size_t)b * k + j], p);  // DS:896
    s2[(size_t)b * nk + e] = P[(size_t)i * pd + j] * sqrtf(fmaxf(d, 0.f));
  }
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn;
       e += (size_t)gridDim.x * blockDim.x)
    root[(size_t)b * nn + e] *= back;
}

// out[b] = {max |diag - 1|, mean |diag - 1|, max |offdiag|, mean |offdiag|} of mat_m [n, n]
__global__ void __launch_bounds__(1024)
lob_root_diag_kernel(const float* __restrict__ mat_m, int n, float* __restrict__ out) {
  __shared__ float fs[32];
  __shared__ uint32_t us[32];
  const int b = blockIdx.x;
  const float* M = mat_m + (size_t)b * n * n;
  uint32_t dmax = 0, omax = 0;
  float dsum = 0.f, osum = 0.f;
  for (size_t e = threadIdx.x; e < (size_t)n * n; e += blockDim.x) {
    const int i = (int)(e / n), j = (int)(e - (size_t)i * n);
    const float v = M[e];
    if (i == j) {
      const float d = fabsf(v - 1.0f);
      dsum += d;
      dmax = max(dmax, absbits(d));
    } else {
      osum += fabsf(v);
      omax = max(omax, absbits(v));
    }
  }
  dmax = block_max_u32(dmax, us);
  __syncthreads();
  omax = block_max_u32(omax, us);
  dsum = block_sum(dsum, fs);
  osum = block_sum(osum, fs);
  if (threadIdx.x == 0) {
    float* o = out + (size_t)b * 4;
    o[0] = __uint_as_float(dmax);
    o[1] = dsum / (float)n;
    o[2] = __uint_as_float(omax);
    o[3] = n > 1 ? osum / (float)((size_t)n * n - n) : 0.f;
  }
}

// av = A V [n, k], gram = V^T V [k, k] -> out[b] = {iters, max consistency, mean consistency,
// mean orthogonality error, max eigenvalue, min eigenvalue, k}   (DS:172-195)
__global__ void __launch_bounds__(256)
lob_lobpcg_diag_kernel(const float* __restrict__ packed, const float* __restrict__ av,
                       const float* __restrict__ gram, const float* __restrict__ eigvals, int n,
                       int k, float iters, float* __restrict__ out) {
  __shared__ float fs[32];
  __shared__ float cons[512];
  const int b = blockIdx.x, pd = k + 2;
  const float* P = packed + (size_t)b * n * pd;
  const float* AV = av + (size_t)b * n * k;
  for (int j = 0; j < k; ++j) {
    const float lam = eigvals[(size_t)b * k + j];
    float r = 0.f, a2 = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const float x = AV[(size_t)i * k + j];
      const float d = x - lam * P[(size_t)i * pd + j];
      r = fmaf(d, d, r);
      a2 = fmaf(x, x, a2);
    }
    r = block_sum(r, fs);
    a2 = block_sum(a2, fs);
    if (threadIdx.x == 0) cons[j] = sqrtf(r) / (sqrtf(a2) + lam);
    __syncthreads();
  }
  float osum = 0.f;
  for (int e = threadIdx.x; e < k * k; e += blockDim.x)
    if (e / k != e % k) osum += gram[(size_t)b * k * k + e];  // (signed sum, like the reference)
  osum = block_sum(osum, fs);
  if (threadIdx.x == 0) {
    float cmax = 0.f, csum = 0.f, emax = eigvals[(size_t)b * k], emin = emax;
    for (int j = 0; j < k; ++j) {
      cmax = fmaxf(cmax, cons[j]);
      csum += cons[j];
      emax = fmaxf(emax, eigvals[(size_t)b * k + j]);
      emin = fminf(emin, eigvals[(size_t)b * k + j]);
    }
    float* o = out + (size_t)b * 7;
    o[0] = iters; o[1] = cmax; o[2] = csum / (float)k;
    o[3] = k > 1 ? osum / (float)(k * (k - 1)) : 0.f;
    o[4] = emax; o[5] = emin; o[6] = (float)k;
  }
}

}  // namespace pc

extern "C" {

int pc_lobpcg_deflate_prep(const float* packed, const float* a, int batch, int n, int k,
                           float ridge_epsilon, int relative_matrix_epsilon, float* scalars,
                           float* eigvals, float* s1, float* a_scaled, void* stream) {
  PC_REQUIRE(batch >= 0 && n >= 2 && k >= 1 && k <= 512 && k + 2 < n, "bad sizes (n=%d, k=%d)", n, k);
  if (batch == 0) return PC_OK;
  PC_REQUIRE(packed && a && scalars && eigvals && s1 && a_scaled, "null pointer argument");
  dim3 grid((unsigned)std::min<size_t>(((size_t)n * n + 255) / 256, 256), batch);
  pc::lob_deflate_prep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      packed, a, n, k, ridge_epsilon, relative_matrix_epsilon, scalars, eigvals, s1, a_scaled);
  pc::count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

int pc_lobpcg_redeflate_prep(const float* packed, const float* scalars, const float* eigvals,
                             const int32_t* ps, int batch, int n, int k, float* roots, float* s2,
                             void* stream) {
  PC_REQUIRE(batch >= 0 && n >= 2 && k >= 1 && k <= 512, "bad sizes");
  if (batch == 0) return PC_OK;
  PC_REQUIRE(packed && scalars && eigvals && ps && roots && s2, "null pointer argument");
  dim3 grid((unsigned)std::min<size_t>(((size_t)n * n + 255) / 256, 256), batch);
  pc::lob_redeflate_prep_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(packed, scalars, eigvals,
                                                                       ps, n, k, roots, s2);
  pc::count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

int pc_root_diagnostics(const float* mat_m, int batch, int n, float* out, void* stream) {
  PC_REQUIRE(batch >= 0 && n >= 1, "bad sizes");
  if (batch == 0) return PC_OK;
  PC_REQUIRE(mat_m && out, "null pointer argument");
  pc::lob_root_diag_kernel<<<batch, 1024, 0, (cudaStream_t)stream>>>(mat_m, n, out);
  pc::count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

int pc_lobpcg_diagnostics(const float* packed, const float* av, const float* gram,
                          const float* eigvals, int batch, int n, int k, float iters, float* out,
                          void* stream) {
  PC_REQUIRE(batch >= 0 && n >= 2 && k >= 1 && k <= 512, "bad sizes");
  if (batch == 0) return PC_OK;
  PC_REQUIRE(packed && av && gram && eigvals && out, "null pointer argument");
  pc::lob_lobpcg_diag_kernel<<<batch, 256, 0, (cudaStream_t)stream>>>(packed, av, gram, eigvals, n,
                                                                     k, iters, out);
  pc::count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

}  // extern "C"
