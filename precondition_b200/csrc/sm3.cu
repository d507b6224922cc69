// SM3 (precondition/sm3.py:40-168, "SM3" below) on the device: per-axis accumulators, the
// second-moment estimate of an entry = min over its axes' accumulators, int8 momentum.
// Memory-bound and elementwise apart from the per-axis maxima: one pass reads grad / momentum
// (int8 + per-column bucket) / param and writes the update, the fp32 momentum (requantised by
// pc_quantize_batched afterwards) and the new accumulators through atomicMax on the float bits
// (all values are >= 0, so the integer order is the float order).
#include <algorithm>

#include "common.cuh"

namespace pc {

struct Sm3Args {
  const float* grad; const float* param;
  const float* acc_in[4];
  uint32_t* acc_out[4];     // zero on entry (rank >= 2); plain store for rank 1
  const int8_t* mom_q; const float* mom_bucket;  // bucket over shape[1:] (scalar for rank 1)
  float* mom_f; float* update;
  const float* sumsq_parts; int sumsq_n;  // normalize_grads: partial sums of grad^2 (else null)
  int rank; int dims[4]; int64_t numel;
  float beta1, beta2, w1, w2, eps, wd, lr;
};

__global__ void __launch_bounds__(256) sm3_sumsq_kernel(const float* __restrict__ g, int64_t n,
                                                        float* __restrict__ parts) {
  __shared__ float scratch[32];
  float s = 0.f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n;
       e += (int64_t)gridDim.x * blockDim.x)
    s = fmaf(g[e], g[e], s);
  s = block_sum(s, scratch);
  if (threadIdx.x == 0) parts[blockIdx.x] = s;
}

__global__ void __launch_bounds__(256) sm3_update_kernel(const Sm3Args a) {
  __shared__ float scratch[32];
  float gscale = 1.f;
  if (a.sumsq_parts) {  // SM3:113-115: g / (|g| + 1e-16), same value in every block
    float s = 0.f;
    for (int i = threadIdx.x; i < a.sumsq_n; i += blockDim.x) s += a.sumsq_parts[i];
    s = block_sum(s, scratch);
    gscale = sqrtf(s) + 1e-16f;
  }
  const int64_t cols = a.rank >= 1 ? a.numel / a.dims[0] : 1;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.numel;
       e += (int64_t)gridDim.x * blockDim.x) {
    int idx[4];
    int64_t r = e;
#pragma unroll
    for (int ax = 3; ax >= 0; --ax) {
      if (ax < a.rank) { idx[ax] = (int)(r % a.dims[ax]); r /= a.dims[ax]; }
      else idx[ax] = 0;
    }
    float g = a.grad[e];
    if (a.sumsq_parts) g = g / gscale;
    float mn = a.acc_in[0][idx[0]];
#pragma unroll
    for (int ax = 1; ax < 4; ++ax)
      if (ax < a.rank) mn = fminf(mn, a.acc_in[ax][idx[ax]]);
    const float nu = __fadd_rn(__fmul_rn(a.beta2, mn), __fmul_rn(a.w2, __fmul_rn(g, g)));  // SM3:88-94
    const float pre = 1.0f / sqrtf(__fadd_rn(nu, a.eps));                                  // SM3:133-134
    const float pg = __fmul_rn(g, pre);
    const float m_old = a.mom_q ? __fmul_rn((float)a.mom_q[e], a.mom_bucket[e % cols]) : 0.f;
    const float m_new = __fadd_rn(__fmul_rn(a.beta1, m_old), __fmul_rn(a.w1, pg));         // SM3:96-98
    a.mom_f[e] = m_new;
    const float out = a.wd > 0.f ? __fadd_rn(m_new, __fmul_rn(a.wd, a.param[e])) : m_new; // SM3:154-158
    a.update[e] = __fmul_rn(-a.lr, out);
    if (a.rank <= 1) {
      reinterpret_cast<float*>(a.acc_out[0])[idx[0]] = nu;  // SM3:107-108
    } else {
#pragma unroll
      for (int ax = 0; ax < 4; ++ax)
        if (ax < a.rank) atomicMax(a.acc_out[ax] + idx[ax], __float_as_uint(nu));  // SM3:100-106
    }
  }
}

}  // namespace pc

extern "C" {

size_t pc_sm3_workspace_bytes(int64_t numel) {
  (void)numel;
  return 1024 * sizeof(float) + 256;
}

int pc_sm3_update(const float* grad, const float* param, const float* const* acc_in,
                  float* const* acc_out, const int8_t* momentum_q, const float* momentum_bucket,
                  float* momentum_f, float* update, int rank, const int32_t* dims,
                  const pc_sm3_options* opt, void* workspace, size_t workspace_bytes,
                  void* stream) {
  PC_REQUIRE(rank >= 1 && rank <= 4, "pc_sm3_update supports tensors of rank 1..4 (rank %d)", rank);
  PC_REQUIRE(grad && acc_in && acc_out && momentum_f && update && dims && opt && workspace,
             "null pointer argument");
  PC_REQUIRE(workspace_bytes >= pc_sm3_workspace_bytes(0), "workspace too small");
  PC_REQUIRE(opt->weight_decay <= 0.f || param, "weight decay needs param");
  pc::Sm3Args a{};
  a.grad = grad; a.param = param; a.mom_q = momentum_q; a.mom_bucket = momentum_bucket;
  a.mom_f = momentum_f; a.update = update; a.rank = rank;
  a.numel = 1;
  for (int i = 0; i < 4; ++i) {
    a.dims[i] = i < rank ? dims[i] : 1;
    PC_REQUIRE(a.dims[i] >= 1, "bad dimension");
    a.numel *= a.dims[i];
    a.acc_in[i] = i < rank ? acc_in[i] : nullptr;
    a.acc_out[i] = i < rank ? reinterpret_cast<uint32_t*>(acc_out[i]) : nullptr;
    PC_REQUIRE(i >= rank || (a.acc_in[i] && a.acc_out[i]), "null accumulator");
  }
  a.beta1 = (float)opt->beta1; a.beta2 = (float)opt->beta2;
  a.w1 = opt->beta1 != 1.0 ? (float)(1.0 - opt->beta1) : 1.0f;
  a.w2 = opt->beta2 != 1.0 ? (float)(1.0 - opt->beta2) : 1.0f;
  a.eps = opt->diagonal_epsilon; a.wd = opt->weight_decay; a.lr = opt->learning_rate;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = (int)std::min<int64_t>((a.numel + 255) / 256, 1024);
  if (rank >= 2)
    for (int i = 0; i < rank; ++i)
      PC_CUDA_CHECK(cudaMemsetAsync(a.acc_out[i], 0, sizeof(float) * a.dims[i], st));
  if (opt->normalize_grads) {
    float* parts = reinterpret_cast<float*>(pc::align_up((size_t)workspace, 256));
    pc::sm3_sumsq_kernel<<<blocks, 256, 0, st>>>(grad, a.numel, parts);
    a.sumsq_parts = parts;
    a.sumsq_n = blocks;
    pc::count_launch(1);
  }
  pc::sm3_update_kernel<<<blocks, 256, 0, st>>>(a);
  pc::count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

}  // extern "C"
