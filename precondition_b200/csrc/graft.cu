// Grafting + momentum tail of _transform_grad (DS:3496-3625) for one parameter
// tensor: memory-bound, vectorised, two global norms -> reduce-then-apply.
//
// Algorithmic traffic (SGD graft, f32 momenta, no weight decay): read grad,
// precond_grad, both momenta; write update, both momenta = 28 B / element
// (+ one extra read of grad and precond_grad for the norm pass).
#include "common.cuh"

#include <algorithm>

namespace pc {

constexpr int kGraftThreads = 256;
constexpr int kGraftMaxBlocks = 148 * 8;

struct GraftArgs {
  const float* grad;
  const float* param;
  const float* precond;  // may be null -> precond_grad = grafting_update (DS:3561)
  float* diag;           // diagonal_statistics (in/out), null for SGD-like grafts
  float* dmom;           // diagonal_momentum (in/out)
  float* mom;            // momentum (in/out)
  float* update;
  int64_t numel;
  pc_graft_options o;
  float beta1f, beta2f;
  float w2;              // (beta2 == 1 ? beta2 : 1 - beta2), DS:3522
  float lr_mult;         // lr if !decoupled_learning_rate else 1, DS:3549
  // reduction scratch
  float* part_g;   // partial sums of grad^2
  float* part_r;   // partial sums of raw graft^2 (clip)
  float* part_gr;  // partial sums of final graft^2
  float* part_p;   // partial sums of precond^2
  int nblocks;
};

__device__ __forceinline__ bool has_diag(int t) {
  return t == PC_GRAFT_ADAGRAD || t == PC_GRAFT_ADAGRAD_NORMALIZED || t == PC_GRAFT_RMSPROP ||
         t == PC_GRAFT_RMSPROP_NORMALIZED;
}
__device__ __forceinline__ bool is_normalized(int t) {
  return t == PC_GRAFT_ADAGRAD_NORMALIZED || t == PC_GRAFT_RMSPROP_NORMALIZED;
}

// deterministic total of per-block partials (every block recomputes the same value)
__device__ float total_of(const float* part, int n, float* scratch) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += part[i];
  return block_sum(s, scratch);
}

// raw grafting update before clipping / lr multiplier (DS:3502-3543)
// t = the graft type; a compile-time constant at the call site folds the dispatch away
__device__ __forceinline__ float graft_raw(const GraftArgs& a, const int t, float g, float gdenom,
                                           float diag_old, float* diag_new) {
  float sg = g;
  if (is_normalized(t)) sg = g / gdenom;
  if (t == PC_GRAFT_ADAGRAD || t == PC_GRAFT_ADAGRAD_NORMALIZED) {
    const float nd = diag_old + sg * sg;
    *diag_new = nd;
    return sg / (sqrtf(nd) + a.o.diagonal_epsilon);
  }
  if (t == PC_GRAFT_RMSPROP || t == PC_GRAFT_RMSPROP_NORMALIZED) {
    const float nd = a.beta2f * diag_old + a.w2 * (sg * sg);
    *diag_new = nd;
    return sg / (sqrtf(nd) + a.o.diagonal_epsilon);
  }
  *diag_new = diag_old;
  if (t == PC_GRAFT_SQRT_N) return g > 0.f ? 1.f : (g < 0.f ? -1.f : (g == 0.f ? 0.f : g));
  return g;  // SGD, NONE
}

__device__ __forceinline__ float clip_denom_of(const GraftArgs& a, float sum_raw_sq) {
  // DS:3530-3535
  const float norm = sqrtf(sum_raw_sq) / sqrtf((float)a.numel);
  return fmaxf(1.f, norm / a.o.clip_by_scaled_gradient_norm);
}

// Streaming part of a reduction pass, one instance per graft type T: the graft-type dispatch stays
// out of the loop, which otherwise costs ~44 instructions per element and makes the pass
// issue-bound (ncu: 62 % SM throughput at 4.0 TB/s).
template <int STAGE, int T>
__device__ __forceinline__ void graft_reduce_stream(const GraftArgs& a, float gdenom, float cdenom,
                                                    bool clip, float& s0, float& s1) {
  auto accumulate = [&](float g, float diag_e, float pg_e) {
    if (STAGE == 0) {
      s0 = fmaf(g, g, s0);
    } else {
      float nd;
      float r = graft_raw(a, T, g, gdenom, diag_e, &nd);
      if (STAGE == 1) {
        s0 = fmaf(r, r, s0);
      } else {
        if (clip) r = r / cdenom;
        r = r * a.lr_mult;
        s0 = fmaf(r, r, s0);
        const float pg = a.precond ? pg_e : r;
        s1 = fmaf(pg, pg, s1);
      }
    }
  };
  const bool diag = has_diag(T) && STAGE >= 1 && a.diag;
  const bool prec = STAGE == 2 && a.precond;
  // 16-byte loads when every stream is aligned (the optimizer's flat buffers are): a reduction
  // pass moves 8 B / element and has to stay on the HBM roofline
  const bool vec = ((reinterpret_cast<uintptr_t>(a.grad) |
                     reinterpret_cast<uintptr_t>(a.precond ? a.precond : a.grad) |
                     reinterpret_cast<uintptr_t>(a.diag ? a.diag : a.grad)) & 15) == 0;
  const int64_t n4 = vec ? a.numel >> 2 : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(a.grad)[i];
    float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f), p4 = d4;
    if (diag) d4 = reinterpret_cast<const float4*>(a.diag)[i];
    if (prec) p4 = reinterpret_cast<const float4*>(a.precond)[i];
    accumulate(g4.x, d4.x, p4.x);
    accumulate(g4.y, d4.y, p4.y);
    accumulate(g4.z, d4.z, p4.z);
    accumulate(g4.w, d4.w, p4.w);
  }
  for (int64_t e = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.numel;
       e += (int64_t)gridDim.x * blockDim.x)
    accumulate(a.grad[e], diag ? a.diag[e] : 0.f, prec ? a.precond[e] : 0.f);
}

template <int STAGE>  // 0: sum grad^2; 1: sum raw^2; 2: sum graft^2 and precond^2
__global__ void __launch_bounds__(kGraftThreads) graft_reduce_kernel(GraftArgs a) {
  __shared__ float scratch[32];
  float gdenom = 1.f, cdenom = 1.f;
  if (STAGE >= 1 && is_normalized(a.o.graft_type))
    gdenom = sqrtf(total_of(a.part_g, a.nblocks, scratch)) + 1e-25f;
  const bool clip = a.o.clip_by_scaled_gradient_norm > 0.f &&
                    (a.o.graft_type == PC_GRAFT_RMSPROP ||
                     a.o.graft_type == PC_GRAFT_RMSPROP_NORMALIZED);
  if (STAGE == 2 && clip) cdenom = clip_denom_of(a, total_of(a.part_r, a.nblocks, scratch));
  float s0 = 0.f, s1 = 0.f;
  switch (a.o.graft_type) {
#define PC_GRAFT_CASE(T) \
  case T: graft_reduce_stream<STAGE, T>(a, gdenom, cdenom, clip, s0, s1); break;
    PC_GRAFT_CASE(PC_GRAFT_NONE)
    PC_GRAFT_CASE(PC_GRAFT_SGD)
    PC_GRAFT_CASE(PC_GRAFT_ADAGRAD)
    PC_GRAFT_CASE(PC_GRAFT_RMSPROP)
    PC_GRAFT_CASE(PC_GRAFT_RMSPROP_NORMALIZED)
    PC_GRAFT_CASE(PC_GRAFT_SQRT_N)
    PC_GRAFT_CASE(PC_GRAFT_ADAGRAD_NORMALIZED)
#undef PC_GRAFT_CASE
  }
  s0 = block_sum(s0, scratch);
  if (STAGE == 2) s1 = block_sum(s1, scratch);
  if (threadIdx.x == 0) {
    if (STAGE == 0) a.part_g[blockIdx.x] = s0;
    if (STAGE == 1) a.part_r[blockIdx.x] = s0;
    if (STAGE == 2) { a.part_gr[blockIdx.x] = s0; a.part_p[blockIdx.x] = s1; }
  }
}

__global__ void __launch_bounds__(kGraftThreads) graft_apply_kernel(GraftArgs a) {
  __shared__ float scratch[32];
  const pc_graft_options& o = a.o;
  float gdenom = 1.f, cdenom = 1.f;
  if (is_normalized(o.graft_type))
    gdenom = sqrtf(total_of(a.part_g, a.nblocks, scratch)) + 1e-25f;
  const bool clip = o.clip_by_scaled_gradient_norm > 0.f &&
                    (o.graft_type == PC_GRAFT_RMSPROP ||
                     o.graft_type == PC_GRAFT_RMSPROP_NORMALIZED);
  if (clip) cdenom = clip_denom_of(a, total_of(a.part_r, a.nblocks, scratch));
  const float gnorm = sqrtf(total_of(a.part_gr, a.nblocks, scratch));  // DS:3563
  const float pnorm = sqrtf(total_of(a.part_p, a.nblocks, scratch));   // DS:3564
  const float mult = o.graft_type != PC_GRAFT_NONE ? gnorm / (pnorm + 1e-25f) : 1.f;
  const float w = o.moving_average_for_momentum ? (float)(1.0 - o.beta1) : 1.f;
  const float run = o.run_shampoo ? 1.f : 0.f;
  const bool coupled_wd = o.weight_decay != 0.f && !o.decoupled_weight_decay;
  const bool decoupled_wd = o.weight_decay != 0.f && o.decoupled_weight_decay;
  const float wd_lr = o.decoupled_learning_rate ? 1.f : o.learning_rate;
  const float mom_mult = o.decoupled_learning_rate ? o.learning_rate : 1.f;
  // one element of the tail; state values come in by reference and leave updated
  auto element = [&](float g, float prm, float pg_e, float& diag_e, float& dmom_e, float& mom_e) {
    float nd;
    float graft = graft_raw(a, a.o.graft_type, g, gdenom, diag_e, &nd);
    if (clip) graft = graft / cdenom;
    graft = graft * a.lr_mult;
    const float pg = a.precond ? pg_e : graft;
    const float shampoo = pg * mult;                                   // DS:3570
    float shampoo_wd = shampoo, graft_wd = graft;
    if (coupled_wd) {                                                  // DS:3575-3577
      shampoo_wd = shampoo + o.weight_decay * prm;
      graft_wd = graft + o.weight_decay * prm;
    }
    const float shampoo_m = mom_e * a.beta1f + w * shampoo_wd;          // DS:3581-3582
    const float graft_m = dmom_e * a.beta1f + w * graft_wd;             // DS:3584-3586
    const float mom = run * shampoo_m + (1.f - run) * graft_m;         // DS:3591-3593
    const float wdu = run * shampoo_wd + (1.f - run) * graft_wd;       // DS:3595-3597
    float nest = o.nesterov ? (w * wdu + a.beta1f * mom) : mom;         // DS:3601-3602
    if (decoupled_wd) nest = nest + wd_lr * o.weight_decay * prm;      // DS:3604-3608
    dmom_e = graft_m;
    mom_e = shampoo_m;
    diag_e = nd;
    return -1.0f * mom_mult * nest;                                    // DS:3610-3611
  };
  const bool use_prm = coupled_wd || decoupled_wd;
  // 16-byte accesses when every stream is aligned: 28 B / element has to stay on the HBM roofline
  const bool vec = ((reinterpret_cast<uintptr_t>(a.grad) | reinterpret_cast<uintptr_t>(a.update) |
                     reinterpret_cast<uintptr_t>(a.mom) | reinterpret_cast<uintptr_t>(a.dmom) |
                     reinterpret_cast<uintptr_t>(a.precond ? a.precond : a.grad) |
                     reinterpret_cast<uintptr_t>(a.diag ? a.diag : a.mom) |
                     reinterpret_cast<uintptr_t>(use_prm ? a.param : a.grad)) & 15) == 0;
  const int64_t n4 = vec ? a.numel >> 2 : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (int64_t)gridDim.x * blockDim.x) {
    const float4 g4 = reinterpret_cast<const float4*>(a.grad)[i];
    float4 m4 = reinterpret_cast<const float4*>(a.mom)[i];
    float4 dm4 = reinterpret_cast<const float4*>(a.dmom)[i];
    float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f), p4 = d4, r4 = d4, u4;
    if (a.diag) d4 = reinterpret_cast<const float4*>(a.diag)[i];
    if (a.precond) p4 = reinterpret_cast<const float4*>(a.precond)[i];
    if (use_prm) r4 = reinterpret_cast<const float4*>(a.param)[i];
    u4.x = element(g4.x, r4.x, p4.x, d4.x, dm4.x, m4.x);
    u4.y = element(g4.y, r4.y, p4.y, d4.y, dm4.y, m4.y);
    u4.z = element(g4.z, r4.z, p4.z, d4.z, dm4.z, m4.z);
    u4.w = element(g4.w, r4.w, p4.w, d4.w, dm4.w, m4.w);
    reinterpret_cast<float4*>(a.update)[i] = u4;
    reinterpret_cast<float4*>(a.dmom)[i] = dm4;
    reinterpret_cast<float4*>(a.mom)[i] = m4;
    if (a.diag) reinterpret_cast<float4*>(a.diag)[i] = d4;
  }
  for (int64_t e = 4 * n4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < a.numel;
       e += (int64_t)gridDim.x * blockDim.x) {
    float d = a.diag ? a.diag[e] : 0.f, dm = a.dmom[e], m = a.mom[e];
    a.update[e] = element(a.grad[e], use_prm ? a.param[e] : 0.f, a.precond ? a.precond[e] : 0.f,
                          d, dm, m);
    a.dmom[e] = dm;
    a.mom[e] = m;
    if (a.diag) a.diag[e] = d;
  }
}


// ---------------------------------------------------------------------------
// Grouped form: every parameter of the model in a fixed number of launches.
// The optimizer keeps gradients, preconditioned gradients, momenta, diagonal
// statistics and the updates in flat buffers (one segment per parameter, same offset
// in every buffer).  Work = (segment, chunk) pairs of kGroupChunk elements; a reduction
// pass writes one partial per chunk, a tiny pass sums each segment's partials in a fixed
// order (deterministic, like the two-level sums of the per-parameter kernels), the apply
// pass reads the per-segment totals.  3 launches for SGD-like grafts (5 / 7 with the
// normalised / clipped types) instead of 2-4 per parameter.
// ---------------------------------------------------------------------------
constexpr int kGroupChunk = kGraftThreads * 32;  // 8192 elements per work item

struct GraftGroupArgs {
  const float* grad; const float* param; const float* precond;
  float* diag; float* dmom; float* mom; float* update;
  const pc_graft_segment* segs;
  const int32_t* chunk_seg;  // [total_chunks] segment of every chunk
  int total_chunks, nsegs;
  pc_graft_options o;
  float beta1f, beta2f, w2, lr_mult;
  float* part;  // [total_chunks][4]: grad^2, raw^2, graft^2, precond^2
  float* tot;   // [nsegs][4]
};

__device__ __forceinline__ GraftArgs graft_group_view(const GraftGroupArgs& g, int c,
                                                      int64_t* begin, int64_t* end) {
  const int sidx = g.chunk_seg[c];
  const pc_graft_segment sg = g.segs[sidx];
  GraftArgs a;
  a.grad = g.grad + sg.offset;
  a.param = g.param ? g.param + sg.offset : nullptr;
  a.precond = (g.precond && sg.has_precond) ? g.precond + sg.offset : nullptr;
  a.diag = g.diag ? g.diag + sg.offset : nullptr;
  a.dmom = g.dmom + sg.offset;
  a.mom = g.mom + sg.offset;
  a.update = g.update + sg.offset;
  a.numel = sg.numel;
  a.o = g.o;
  a.beta1f = g.beta1f; a.beta2f = g.beta2f; a.w2 = g.w2; a.lr_mult = g.lr_mult;
  a.part_g = g.tot + 4 * (size_t)sidx;  // totals of this segment: [0] g, [1] raw, [2] graft, [3] precond
  a.part_r = a.part_gr = a.part_p = nullptr;
  a.nblocks = 0;
  *begin = (int64_t)(c - sg.first_chunk) * kGroupChunk;
  *end = *begin + kGroupChunk < sg.numel ? *begin + kGroupChunk : sg.numel;
  return a;
}

template <int STAGE, int T>
__device__ __forceinline__ void graft_group_reduce_range(const GraftArgs& a, int64_t begin,
                                                         int64_t end, float gdenom, float cdenom,
                                                         bool clip, float& s0, float& s1) {
  auto accumulate = [&](float g, float diag_e, float pg_e) {
    if (STAGE == 0) {
      s0 = fmaf(g, g, s0);
    } else {
      float nd;
      float r = graft_raw(a, T, g, gdenom, diag_e, &nd);
      if (STAGE == 1) {
        s0 = fmaf(r, r, s0);
      } else {
        if (clip) r = r / cdenom;
        r = r * a.lr_mult;
        s0 = fmaf(r, r, s0);
        const float pg = a.precond ? pg_e : r;
        s1 = fmaf(pg, pg, s1);
      }
    }
  };
  const bool diag = has_diag(T) && STAGE >= 1 && a.diag;
  const bool prec = STAGE == 2 && a.precond;
  // segment offsets are multiples of 32 elements and chunks of 8192: every chunk starts 16-byte
  // aligned; only the last chunk of a segment has a scalar tail
  const int64_t n4 = (end - begin) >> 2;
  const float4* g4p = reinterpret_cast<const float4*>(a.grad + begin);
  const float4* d4p = reinterpret_cast<const float4*>((diag ? a.diag : a.grad) + begin);
  const float4* p4p = reinterpret_cast<const float4*>((prec ? a.precond : a.grad) + begin);
  for (int64_t i = threadIdx.x; i < n4; i += kGraftThreads) {
    const float4 g4 = g4p[i];
    float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f), p4 = d4;
    if (diag) d4 = d4p[i];
    if (prec) p4 = p4p[i];
    accumulate(g4.x, d4.x, p4.x);
    accumulate(g4.y, d4.y, p4.y);
    accumulate(g4.z, d4.z, p4.z);
    accumulate(g4.w, d4.w, p4.w);
  }
  for (int64_t e = begin + 4 * n4 + threadIdx.x; e < end; e += kGraftThreads)
    accumulate(a.grad[e], diag ? a.diag[e] : 0.f, prec ? a.precond[e] : 0.f);
}

template <int STAGE>
__global__ void __launch_bounds__(kGraftThreads) graft_group_reduce_kernel(GraftGroupArgs g) {
  __shared__ float scratch[32];
  const int t = g.o.graft_type;
  const bool clip = g.o.clip_by_scaled_gradient_norm > 0.f &&
                    (t == PC_GRAFT_RMSPROP || t == PC_GRAFT_RMSPROP_NORMALIZED);
  for (int c = blockIdx.x; c < g.total_chunks; c += gridDim.x) {
    int64_t begin, end;
    const GraftArgs a = graft_group_view(g, c, &begin, &end);
    float gdenom = 1.f, cdenom = 1.f;
    if (STAGE >= 1 && is_normalized(t)) gdenom = sqrtf(a.part_g[0]) + 1e-25f;
    if (STAGE == 2 && clip) cdenom = clip_denom_of(a, a.part_g[1]);
    float s0 = 0.f, s1 = 0.f;
    switch (t) {
#define PC_GRAFT_CASE(T) \
  case T: graft_group_reduce_range<STAGE, T>(a, begin, end, gdenom, cdenom, clip, s0, s1); break;
      PC_GRAFT_CASE(PC_GRAFT_NONE)
      PC_GRAFT_CASE(PC_GRAFT_SGD)
      PC_GRAFT_CASE(PC_GRAFT_ADAGRAD)
      PC_GRAFT_CASE(PC_GRAFT_RMSPROP)
      PC_GRAFT_CASE(PC_GRAFT_RMSPROP_NORMALIZED)
      PC_GRAFT_CASE(PC_GRAFT_SQRT_N)
      PC_GRAFT_CASE(PC_GRAFT_ADAGRAD_NORMALIZED)
#undef PC_GRAFT_CASE
    }
    s0 = block_sum(s0, scratch);
    if (STAGE == 2) s1 = block_sum(s1, scratch);
    if (threadIdx.x == 0) {
      float* p = g.part + 4 * (size_t)c;
      if (STAGE == 0) p[0] = s0;
      if (STAGE == 1) p[1] = s0;
      if (STAGE == 2) { p[2] = s0; p[3] = s1; }
    }
  }
}

// one warp per segment: fixed-order sum of its chunk partials, components [k0, k1)
__global__ void __launch_bounds__(256) graft_group_total_kernel(GraftGroupArgs g, int k0, int k1) {
  const int sidx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (sidx >= g.nsegs) return;
  const int lane = threadIdx.x & 31;
  const pc_graft_segment sg = g.segs[sidx];
  for (int k = k0; k < k1; ++k) {
    float s = 0.f;
    for (int i = lane; i < sg.nchunks; i += 32) s += g.part[4 * (size_t)(sg.first_chunk + i) + k];
    s = warp_sum(s);
    if (lane == 0) g.tot[4 * (size_t)sidx + k] = s;
  }
}

__global__ void __launch_bounds__(kGraftThreads) graft_group_apply_kernel(GraftGroupArgs g) {
  const pc_graft_options& o = g.o;
  const int t = o.graft_type;
  const bool clip = o.clip_by_scaled_gradient_norm > 0.f &&
                    (t == PC_GRAFT_RMSPROP || t == PC_GRAFT_RMSPROP_NORMALIZED);
  const float w = o.moving_average_for_momentum ? (float)(1.0 - o.beta1) : 1.f;
  const float run = o.run_shampoo ? 1.f : 0.f;
  const bool coupled_wd = o.weight_decay != 0.f && !o.decoupled_weight_decay;
  const bool decoupled_wd = o.weight_decay != 0.f && o.decoupled_weight_decay;
  const float wd_lr = o.decoupled_learning_rate ? 1.f : o.learning_rate;
  const float mom_mult = o.decoupled_learning_rate ? o.learning_rate : 1.f;
  const bool use_prm = coupled_wd || decoupled_wd;
  for (int c = blockIdx.x; c < g.total_chunks; c += gridDim.x) {
    int64_t begin, end;
    const GraftArgs a = graft_group_view(g, c, &begin, &end);
    float gdenom = 1.f, cdenom = 1.f;
    if (is_normalized(t)) gdenom = sqrtf(a.part_g[0]) + 1e-25f;
    if (clip) cdenom = clip_denom_of(a, a.part_g[1]);
    const float gnorm = sqrtf(a.part_g[2]);  // DS:3563
    const float pnorm = sqrtf(a.part_g[3]);  // DS:3564
    const float mult = t != PC_GRAFT_NONE ? gnorm / (pnorm + 1e-25f) : 1.f;
    auto element = [&](float gr, float prm, float pg_e, float& diag_e, float& dmom_e, float& mom_e) {
      float nd;
      float graft = graft_raw(a, t, gr, gdenom, diag_e, &nd);
      if (clip) graft = graft / cdenom;
      graft = graft * a.lr_mult;
      const float pg = a.precond ? pg_e : graft;
      const float shampoo = pg * mult;                                   // DS:3570
      float shampoo_wd = shampoo, graft_wd = graft;
      if (coupled_wd) {                                                  // DS:3575-3577
        shampoo_wd = shampoo + o.weight_decay * prm;
        graft_wd = graft + o.weight_decay * prm;
      }
      const float shampoo_m = mom_e * a.beta1f + w * shampoo_wd;          // DS:3581-3582
      const float graft_m = dmom_e * a.beta1f + w * graft_wd;             // DS:3584-3586
      const float mom = run * shampoo_m + (1.f - run) * graft_m;         // DS:3591-3593
      const float wdu = run * shampoo_wd + (1.f - run) * graft_wd;       // DS:3595-3597
      float nest = o.nesterov ? (w * wdu + a.beta1f * mom) : mom;         // DS:3601-3602
      if (decoupled_wd) nest = nest + wd_lr * o.weight_decay * prm;      // DS:3604-3608
      dmom_e = graft_m;
      mom_e = shampoo_m;
      diag_e = nd;
      return -1.0f * mom_mult * nest;                                    // DS:3610-3611
    };
    const int64_t n4 = (end - begin) >> 2;
    for (int64_t i = threadIdx.x; i < n4; i += kGraftThreads) {
      const int64_t e = begin + 4 * i;
      const float4 g4 = *reinterpret_cast<const float4*>(a.grad + e);
      float4 m4 = *reinterpret_cast<const float4*>(a.mom + e);
      float4 dm4 = *reinterpret_cast<const float4*>(a.dmom + e);
      float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f), p4 = d4, r4 = d4, u4;
      if (a.diag) d4 = *reinterpret_cast<const float4*>(a.diag + e);
      if (a.precond) p4 = *reinterpret_cast<const float4*>(a.precond + e);
      if (use_prm) r4 = *reinterpret_cast<const float4*>(a.param + e);
      u4.x = element(g4.x, r4.x, p4.x, d4.x, dm4.x, m4.x);
      u4.y = element(g4.y, r4.y, p4.y, d4.y, dm4.y, m4.y);
      u4.z = element(g4.z, r4.z, p4.z, d4.z, dm4.z, m4.z);
      u4.w = element(g4.w, r4.w, p4.w, d4.w, dm4.w, m4.w);
      *reinterpret_cast<float4*>(a.update + e) = u4;
      *reinterpret_cast<float4*>(a.dmom + e) = dm4;
      *reinterpret_cast<float4*>(a.mom + e) = m4;
      if (a.diag) *reinterpret_cast<float4*>(a.diag + e) = d4;
    }
    for (int64_t e = begin + 4 * n4 + threadIdx.x; e < end; e += kGraftThreads) {
      float d = a.diag ? a.diag[e] : 0.f, dm = a.dmom[e], m = a.mom[e];
      a.update[e] = element(a.grad[e], use_prm ? a.param[e] : 0.f, a.precond ? a.precond[e] : 0.f,
                            d, dm, m);
      a.dmom[e] = dm;
      a.mom[e] = m;
      if (a.diag) a.diag[e] = d;
    }
  }
}

}  // namespace pc

extern "C" {

size_t pc_graft_momentum_workspace_bytes(int64_t numel) {
  (void)numel;
  return 4 * pc::kGraftMaxBlocks * sizeof(float) + 256;
}

int pc_graft_momentum(const float* grad, const float* param, const float* precond_grad,
                      float* diagonal_statistics, float* diagonal_momentum, float* momentum,
                      float* update, int64_t numel, const pc_graft_options* opt,
                      void* workspace, size_t workspace_bytes, void* stream) {
  PC_REQUIRE(numel >= 0, "negative numel");
  if (numel == 0) return PC_OK;
  PC_REQUIRE(grad && diagonal_momentum && momentum && update && opt && workspace,
             "null pointer argument");
  PC_REQUIRE(workspace_bytes >= pc_graft_momentum_workspace_bytes(numel), "workspace too small");
  PC_REQUIRE(opt->graft_type >= PC_GRAFT_NONE && opt->graft_type <= PC_GRAFT_ADAGRAD_NORMALIZED,
             "unknown graft_type %d", opt->graft_type);
  const bool diag = opt->graft_type == PC_GRAFT_ADAGRAD || opt->graft_type == PC_GRAFT_RMSPROP ||
                    opt->graft_type == PC_GRAFT_RMSPROP_NORMALIZED ||
                    opt->graft_type == PC_GRAFT_ADAGRAD_NORMALIZED;
  PC_REQUIRE(!diag || diagonal_statistics, "graft type needs diagonal_statistics");
  PC_REQUIRE(opt->weight_decay == 0.f || param, "weight decay needs param");
  pc::GraftArgs a;
  a.grad = grad; a.param = param; a.precond = precond_grad;
  a.diag = diag ? diagonal_statistics : nullptr;
  a.dmom = diagonal_momentum; a.mom = momentum; a.update = update; a.numel = numel;
  a.o = *opt;
  a.beta1f = (float)opt->beta1; a.beta2f = (float)opt->beta2;
  a.w2 = opt->beta2 == 1.0 ? 1.0f : (float)(1.0 - opt->beta2);
  a.lr_mult = opt->decoupled_learning_rate ? 1.0f : opt->learning_rate;
  float* w = reinterpret_cast<float*>(pc::align_up((size_t)workspace, 256));
  a.part_g = w; a.part_r = w + pc::kGraftMaxBlocks; a.part_gr = w + 2 * pc::kGraftMaxBlocks;
  a.part_p = w + 3 * pc::kGraftMaxBlocks;
  // one full wave at most: the grid never exceeds what is resident at once (ncu: 42 registers ->
  // 5 CTAs per SM; the former 8 per SM ran as 1.6 waves with a 60 %-filled tail)
  static const int resident = [] {
    int dev = 0, sms = 148, occ_r = 1, occ_a = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r, pc::graft_reduce_kernel<2>,
                                                  pc::kGraftThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, pc::graft_apply_kernel,
                                                  pc::kGraftThreads, 0);
    const int r = sms * std::max(1, std::min(occ_r, occ_a));
    return r < pc::kGraftMaxBlocks ? r : pc::kGraftMaxBlocks;
  }();
  int64_t want = (numel + pc::kGraftThreads * 4 - 1) / (pc::kGraftThreads * 4);
  a.nblocks = (int)(want < 1 ? 1 : (want > resident ? resident : want));
  cudaStream_t st = (cudaStream_t)stream;
  const bool normalized = opt->graft_type == PC_GRAFT_ADAGRAD_NORMALIZED ||
                          opt->graft_type == PC_GRAFT_RMSPROP_NORMALIZED;
  const bool clip = opt->clip_by_scaled_gradient_norm > 0.f &&
                    (opt->graft_type == PC_GRAFT_RMSPROP ||
                     opt->graft_type == PC_GRAFT_RMSPROP_NORMALIZED);
  if (normalized) pc::graft_reduce_kernel<0><<<a.nblocks, pc::kGraftThreads, 0, st>>>(a);
  if (clip) pc::graft_reduce_kernel<1><<<a.nblocks, pc::kGraftThreads, 0, st>>>(a);
  pc::graft_reduce_kernel<2><<<a.nblocks, pc::kGraftThreads, 0, st>>>(a);
  pc::graft_apply_kernel<<<a.nblocks, pc::kGraftThreads, 0, st>>>(a);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

int64_t pc_graft_group_chunk_elems(void) { return pc::kGroupChunk; }

size_t pc_graft_momentum_grouped_workspace_bytes(int num_segments, int64_t total_chunks) {
  return 4 * sizeof(float) * ((size_t)(total_chunks > 0 ? total_chunks : 0) +
                              (size_t)(num_segments > 0 ? num_segments : 0)) + 512;
}

int pc_graft_momentum_grouped(const pc_graft_segment* segments, const int32_t* chunk_segment,
                              int num_segments, int64_t total_chunks, const float* grad,
                              const float* param, const float* precond_grad,
                              float* diagonal_statistics, float* diagonal_momentum,
                              float* momentum, float* update, const pc_graft_options* opt,
                              void* workspace, size_t workspace_bytes, void* stream) {
  PC_REQUIRE(num_segments >= 0 && total_chunks >= 0 && total_chunks < (1ll << 31),
             "bad segment / chunk count");
  if (num_segments == 0 || total_chunks == 0) return PC_OK;
  PC_REQUIRE(segments && chunk_segment && grad && diagonal_momentum && momentum && update && opt &&
                 workspace, "null pointer argument");
  PC_REQUIRE(workspace_bytes >=
                 pc_graft_momentum_grouped_workspace_bytes(num_segments, total_chunks),
             "workspace too small");
  PC_REQUIRE(opt->graft_type >= PC_GRAFT_NONE && opt->graft_type <= PC_GRAFT_ADAGRAD_NORMALIZED,
             "unknown graft_type %d", opt->graft_type);
  const bool diag = opt->graft_type == PC_GRAFT_ADAGRAD || opt->graft_type == PC_GRAFT_RMSPROP ||
                    opt->graft_type == PC_GRAFT_RMSPROP_NORMALIZED ||
                    opt->graft_type == PC_GRAFT_ADAGRAD_NORMALIZED;
  PC_REQUIRE(!diag || diagonal_statistics, "graft type needs diagonal_statistics");
  PC_REQUIRE(opt->weight_decay == 0.f || param, "weight decay needs param");
  pc::GraftGroupArgs g;
  g.grad = grad; g.param = param; g.precond = precond_grad;
  g.diag = diag ? diagonal_statistics : nullptr;
  g.dmom = diagonal_momentum; g.mom = momentum; g.update = update;
  g.segs = segments; g.chunk_seg = chunk_segment;
  g.total_chunks = (int)total_chunks; g.nsegs = num_segments;
  g.o = *opt;
  g.beta1f = (float)opt->beta1; g.beta2f = (float)opt->beta2;
  g.w2 = opt->beta2 == 1.0 ? 1.0f : (float)(1.0 - opt->beta2);
  g.lr_mult = opt->decoupled_learning_rate ? 1.0f : opt->learning_rate;
  float* w = reinterpret_cast<float*>(pc::align_up((size_t)workspace, 256));
  g.part = w;
  g.tot = w + 4 * (size_t)total_chunks;
  static const int resident = [] {
    int dev = 0, sms = 148, occ_r = 1, occ_a = 1;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r, pc::graft_group_reduce_kernel<2>,
                                                  pc::kGraftThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, pc::graft_group_apply_kernel,
                                                  pc::kGraftThreads, 0);
    return sms * std::max(1, std::min(occ_r, occ_a));
  }();
  const int grid = (int)std::min<int64_t>(total_chunks, resident);
  const int tgrid = (num_segments + 7) / 8;
  cudaStream_t st = (cudaStream_t)stream;
  const bool normalized = opt->graft_type == PC_GRAFT_ADAGRAD_NORMALIZED ||
                          opt->graft_type == PC_GRAFT_RMSPROP_NORMALIZED;
  const bool clip = opt->clip_by_scaled_gradient_norm > 0.f &&
                    (opt->graft_type == PC_GRAFT_RMSPROP ||
                     opt->graft_type == PC_GRAFT_RMSPROP_NORMALIZED);
  int launches = 3;
  if (normalized) {
    pc::graft_group_reduce_kernel<0><<<grid, pc::kGraftThreads, 0, st>>>(g);
    pc::graft_group_total_kernel<<<tgrid, 256, 0, st>>>(g, 0, 1);
    launches += 2;
  }
  if (clip) {
    pc::graft_group_reduce_kernel<1><<<grid, pc::kGraftThreads, 0, st>>>(g);
    pc::graft_group_total_kernel<<<tgrid, 256, 0, st>>>(g, 1, 2);
    launches += 2;
  }
  pc::graft_group_reduce_kernel<2><<<grid, pc::kGraftThreads, 0, st>>>(g);
  pc::graft_group_total_kernel<<<tgrid, 256, 0, st>>>(g, 2, 4);
  pc::graft_group_apply_kernel<<<grid, pc::kGraftThreads, 0, st>>>(g);
  pc::count_launch(launches);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

}  // extern "C"
