// Shared state of the batched coupled-Newton inverse p-th root solver.
//
// Algorithm = matrix_inverse_pth_root of the reference (DS:702-940) in
// "deviation form":  with I_m the masked identity, the solver carries
//     D = (I_m - M) / p        (so that M_i = I_m + D,  DS:844)
//     H                         (DS:846)
// and evaluates  dev(X Y) = dev(X) + dev(Y) + dev(X) dev(Y)  for the binary
// powering chain of mat_power (DS:655-678), finishing with
//     D' = D - Q_p / p + Q_p D          (== (I_m - M_i^p M) / p,  DS:845)
//     H' = H + H D                      (DS:846)
//     err' = p * max|D'|                (== max|M' - I_m|,  DS:847)
// Every GEMM therefore has the shape  OUT = A*B + c1*X1 + c2*X2  and all
// operands are symmetric polynomials in the input, so "B^T" == B.
// The number of GEMMs is the necessary count G(p), not the 6-8 the
// reference issues.
#pragma once
#include "common.cuh"

namespace pc {

constexpr int kMaxP = 16;
constexpr int kMaxSteps = 8;

// logical buffer ids (resolved per matrix with its ping-pong bit `cur`)
enum : int8_t {
  LB_NONE = -1,
  LB_D = 0,   // current D
  LB_DN = 1,  // next D
  LB_H = 2,   // current H
  LB_HN = 3,  // next H
  LB_Q0 = 4,
  LB_Q1 = 5,
  LB_Q2 = 6,
  LB_Q3 = 7,
  kNumBufs = 8
};

struct Step {
  int8_t dst, a, b, x1, x2;
  int8_t reduce_err;  // epilogue reduces max|out| into errbits (the D' step)
  int8_t pad_[2];
  float c1, c2;
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
  int cur;        // ping-pong bit of D / H
  int result_h;   // physical H buffer (0/1) holding the answer
  float err;
  float ratio;
  float max_ev;
  float ridge;    // ridge_epsilon * max(max_ev, 1e-25), DS:830
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
    case LB_D: return cur;
    case LB_DN: return cur ^ 1;
    case LB_H: return 2 + cur;
    case LB_HN: return 2 + (cur ^ 1);
    default: return logical;  // Q buffers map 1:1 (ids 4..7)
  }
}

// Host: builds the per-exponent step list.  Returns false if p needs more than
// kMaxSteps GEMMs or more than 4 scratch buffers.
bool build_program(int p, Program* out);

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
