// Shared state of the batched coupled-Newton inverse p-th root solver.
//
// Algorithm = matrix_inverse_pth_root of the reference (DS:702-940), in the
// reference's own recurrence (I_m = identity masked by padding_start):
//     M_i = (1 - alpha) I_m + alpha M,   alpha = -1/p           (DS:844)
//     M'  = M_i^p M                      (mat_power, DS:655-678; DS:845)
//     H'  = H M_i                                                 (DS:846)
//     err = max|M' - I_m|                                         (DS:847)
// M_i^p is evaluated with the same LSB-first binary powering and operand order
// as mat_power, minus its no-op products (multiply by I, the discarded last
// squaring): G(p) GEMMs per iteration instead of the 6-8 issued
// upstream.  Every GEMM is a plain product OUT = A*B; the step that produces M'
// also emits M_i' and reduces err in its epilogue.  All operands are symmetric
// polynomials in the input, so the kernels read B by rows ("B^T" == B up to
// rounding).
//
// A measured alternative -- iterating on D = (I_m - M)/p -- reaches lower final
// errors but is ~10x less accurate on ill-conditioned statistics
// (profiles/r01_precision_probe.md), so it is not used.
#pragma once
#include "common.cuh"

namespace pc {

constexpr int kMaxP = 16;
constexpr int kMaxSteps = 8;

// logical buffer ids (resolved per matrix with its ping-pong bit `cur`)
enum : int8_t {
  LB_NONE = -1,
  LB_M = 0,    // current M
  LB_MN = 1,   // next M
  LB_MI = 2,   // current M_i
  LB_MIN = 3,  // next M_i
  LB_H = 4,    // current H
  LB_HN = 5,   // next H
  LB_Q0 = 6,
  LB_Q1 = 7,
  LB_Q2 = 8,
  LB_Q3 = 9,
  kNumBufs = 10
};

struct Step {
  int8_t dst, a, b;
  int8_t emit_mi;  // this step produces M': also write M_i' and reduce err
};

struct Program {
  int nsteps;
  Step steps[kMaxSteps];
};

struct RootCtl {
  int p;
  int pad;        // padding_start
  int active;     // inner Newton loop running
  int need_init;  // (re)start a try with eps * 10^tries
  int done;
  int iter;
  int tries;
  int cur;        // ping-pong bit of M / M_i / H
  int result_h;   // >=0: H ping-pong slot holding the answer; -1 written; -2 zeros
  float err;
  float ratio;
  float max_ev;
  float ridge;    // ridge_epsilon * max(max_ev, 1e-25), DS:830
  float hmul;     // stored H = H / hmul (1 unless the engine keeps H in a scaled format)
  // final metrics (DS:902-907)
  float m_err, m_iters, m_ratio, m_retries;
};

struct RootParams {
  float ridge_epsilon;
  float error_tolerance;
  int num_iters;
  int relative_eps;
};

__host__ __device__ inline int physical_buf(int logical, int cur) {
  switch (logical) {
    case LB_M: return cur;
    case LB_MN: return cur ^ 1;
    case LB_MI: return 2 + cur;
    case LB_MIN: return 2 + (cur ^ 1);
    case LB_H: return 4 + cur;
    case LB_HN: return 4 + (cur ^ 1);
    default: return logical;  // Q buffers map 1:1 (ids 6..9)
  }
}

// lower-triangular tile index t -> (tm, tn), tm >= tn, t = tm (tm + 1) / 2 + tn
__host__ __device__ inline void tri_decode(int t, int& tm, int& tn) {
  int r = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (r * (r + 1) / 2 > t) --r;
  while ((r + 1) * (r + 2) / 2 <= t) ++r;
  tm = r;
  tn = t - r * (r + 1) / 2;
}

// Host: builds the per-exponent step list (false if p is out of range).
bool build_program(int p, Program* out);

// M_i element from an M element, mirroring DS:844 without FMA contraction.
__device__ __forceinline__ float mi_from_m(float m, bool diag_in_mask, float alpha,
                                           float one_minus_alpha) {
  const float t = __fmul_rn(alpha, m);
  return __fadd_rn(diag_in_mask ? one_minus_alpha : 0.f, t);
}

// Device-side end-of-step bookkeeping shared by the init and control kernels;
// mirrors the loop predicates of DS:836-840 and DS:862-885.
__device__ inline void root_after_error_update(RootCtl& c, const RootParams& prm) {
  const bool cont = (c.iter < prm.num_iters) && (c.err > prm.error_tolerance) &&
                    (c.ratio < 1.2f);
  if (cont) {
    c.active = 1;
    return;
  }
  c.active = 0;
  // end of one try (DS:876-882)
  const bool converged = c.ratio < 1.2f;
  c.result_h = converged ? c.cur : (c.cur ^ 1);
  c.m_err = c.err;
  c.m_iters = (float)c.iter;
  c.m_ratio = c.ratio;
  c.tries += 1;
  c.m_retries = (float)c.tries;
  const bool failed = c.err > 0.05f;  // NaN -> false, like the reference
  if (failed && c.tries < 6) {
    c.need_init = 1;
  } else {
    c.done = 1;
  }
}

}  // namespace pc
