// tearfree tail: grafting -> momentum / weight decay -> learning rate for every parameter of
// the model in three launches (pc_tearfree_transform, include/precond_b200.h).
// Reference: precondition/tearfree/grafting.py:190-300, momentum.py:81-139, optimizer.py:91-99.
#include <algorithm>
#include "common.cuh"

namespace pc {

constexpr int kTfThreads = 256;
constexpr int kTfChunk = kTfThreads * 32;  // == pc_graft_group_chunk_elems()

struct TfArgs {
  const pc_tearfree_segment* segs;
  const int32_t* chunk_seg;
  int total_chunks, nsegs;
  pc_tearfree_options o;
  float w_new;  // weight of g^2 in the RMSProp accumulator: (1 - decay), or 1 for a sum
  float w_old;  // weight of the old accumulator
  float* part;  // [total_chunks][2]: graft^2, precond^2
  float* tot;   // [nsegs][2]
};

__device__ __forceinline__ float tf_graft_elem(const TfArgs& a, float g, float acc, float* acc_new) {
  if (a.o.graft_type == PC_TF_GRAFT_RMSPROP) {
    // snew * (1 - decay) + decay * prev  (snew + prev if decay == 1), TF/grafting.py:200-206
    const float sq = g * g;
    const float an = a.o.graft_decay == 1.0f ? sq + acc : __fadd_rn(__fmul_rn(sq, a.w_new), __fmul_rn(a.w_old, acc));
    *acc_new = an;
    return g * (1.0f / sqrtf(an + a.o.graft_epsilon));  // g * rsqrt(acc + eps), TF/grafting.py:210
  }
  *acc_new = acc;
  return g;
}

// pass 1: per chunk, sum of squares of the graft update and of the direction
__global__ void __launch_bounds__(kTfThreads) tf_reduce_kernel(TfArgs a) {
  __shared__ float scratch[32];
  for (int c = blockIdx.x; c < a.total_chunks; c += gridDim.x) {
    const pc_tearfree_segment sg = a.segs[a.chunk_seg[c]];
    float s0 = 0.f, s1 = 0.f;
    if (sg.precond && a.o.graft_type != PC_TF_GRAFT_NONE) {
      const int64_t begin = (int64_t)(c - sg.first_chunk) * kTfChunk;
      const int64_t end = begin + kTfChunk < sg.numel ? begin + kTfChunk : sg.numel;
      const bool rms = a.o.graft_type == PC_TF_GRAFT_RMSPROP;
      const int64_t n4 = (end - begin) >> 2;
      const float4* g4 = reinterpret_cast<const float4*>(sg.grad + begin);
      const float4* p4 = reinterpret_cast<const float4*>(sg.precond + begin);
      const float4* a4 = reinterpret_cast<const float4*>((rms ? sg.acc : sg.grad) + begin);
      float dummy;
      for (int64_t i = threadIdx.x; i < n4; i += kTfThreads) {
        const float4 g = g4[i], p = p4[i];
        float4 ac = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rms) ac = a4[i];
        float u;
        u = tf_graft_elem(a, g.x, ac.x, &dummy); s0 = fmaf(u, u, s0); s1 = fmaf(p.x, p.x, s1);
        u = tf_graft_elem(a, g.y, ac.y, &dummy); s0 = fmaf(u, u, s0); s1 = fmaf(p.y, p.y, s1);
        u = tf_graft_elem(a, g.z, ac.z, &dummy); s0 = fmaf(u, u, s0); s1 = fmaf(p.z, p.z, s1);
        u = tf_graft_elem(a, g.w, ac.w, &dummy); s0 = fmaf(u, u, s0); s1 = fmaf(p.w, p.w, s1);
      }
      for (int64_t e = begin + 4 * n4 + threadIdx.x; e < end; e += kTfThreads) {
        const float u = tf_graft_elem(a, sg.grad[e], rms ? sg.acc[e] : 0.f, &dummy);
        s0 = fmaf(u, u, s0);
        s1 = fmaf(sg.precond[e], sg.precond[e], s1);
      }
    }
    s0 = block_sum(s0, scratch);
    s1 = block_sum(s1, scratch);
    if (threadIdx.x == 0) {
      a.part[2 * (size_t)c] = s0;
      a.part[2 * (size_t)c + 1] = s1;
    }
  }
}

// one warp per segment: fixed-order sum of its chunk partials
__global__ void __launch_bounds__(256) tf_total_kernel(TfArgs a) {
  const int sidx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (sidx >= a.nsegs) return;
  const int lane = threadIdx.x & 31;
  const pc_tearfree_segment sg = a.segs[sidx];
  float s0 = 0.f, s1 = 0.f;
  for (int i = lane; i < sg.nchunks; i += 32) {
    s0 += a.part[2 * (size_t)(sg.first_chunk + i)];
    s1 += a.part[2 * (size_t)(sg.first_chunk + i) + 1];
  }
  s0 = warp_sum(s0);
  s1 = warp_sum(s1);
  if (lane == 0) {
    a.tot[2 * (size_t)sidx] = s0;
    a.tot[2 * (size_t)sidx + 1] = s1;
  }
}

__device__ __forceinline__ float tf_apply_elem(const TfArgs& a, float g, float p, float w,
                                               float acc, float vel, bool has_p, float mult,
                                               float m_in, float* acc_new, float* vel_new) {
  const pc_tearfree_options& o = a.o;
  float x;
  if (o.graft_type == PC_TF_GRAFT_NONE) {
    x = has_p ? p : g;
    *acc_new = acc;
  } else {
    const float u = tf_graft_elem(a, g, acc, acc_new);
    x = (has_p && o.use_precond) ? p * mult : u;  // TF/grafting.py:262-270
  }
  if (o.weight_decay > 0.f && !o.weight_decay_after_momentum)
    x = __fadd_rn(x, __fmul_rn(o.weight_decay, w));  // g + weight_decay * p, unfused like the reference's array ops
  if (o.momentum_decay != 0.f) {
    if (o.ema) x *= m_in;                      // optax.scale(1 - decay), TF/momentum.py:88-89
    const float v = __fadd_rn(x, __fmul_rn(o.momentum_decay, vel));  // trace: g + decay * t
    *vel_new = v;
    x = o.nesterov ? __fadd_rn(x, __fmul_rn(o.momentum_decay, v)) : v;
  }
  if (o.weight_decay > 0.f && o.weight_decay_after_momentum)
    x = __fadd_rn(x, __fmul_rn(o.weight_decay, w));
  return o.scale * x;
}

// pass 2: the element-wise chain, state updated in place
__global__ void __launch_bounds__(kTfThreads) tf_apply_kernel(TfArgs a) {
  const pc_tearfree_options& o = a.o;
  const float m_in = (float)(1.0 - (double)o.momentum_decay);
  for (int c = blockIdx.x; c < a.total_chunks; c += gridDim.x) {
    const int sidx = a.chunk_seg[c];
    const pc_tearfree_segment sg = a.segs[sidx];
    const int64_t begin = (int64_t)(c - sg.first_chunk) * kTfChunk;
    const int64_t end = begin + kTfChunk < sg.numel ? begin + kTfChunk : sg.numel;
    const bool has_p = sg.precond != nullptr;
    const bool rms = o.graft_type == PC_TF_GRAFT_RMSPROP;
    const bool mom = o.momentum_decay != 0.f;
    const bool wd = o.weight_decay > 0.f;
    float mult = 0.f;
    if (has_p && o.graft_type != PC_TF_GRAFT_NONE) {
      // jnp.linalg.norm of both; 0 if the direction vanished (TF/grafting.py:264-267)
      const float gn = sqrtf(a.tot[2 * (size_t)sidx]), bn = sqrtf(a.tot[2 * (size_t)sidx + 1]);
      mult = bn > 0.f ? gn / bn : 0.f;
    }
    const int64_t n4 = (end - begin) >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(sg.grad + begin);
    const float4* p4 = reinterpret_cast<const float4*>((has_p ? sg.precond : sg.grad) + begin);
    const float4* w4 = reinterpret_cast<const float4*>((wd ? sg.param : sg.grad) + begin);
    float4* a4 = reinterpret_cast<float4*>((rms ? sg.acc : sg.update) + begin);
    float4* v4 = reinterpret_cast<float4*>((mom ? sg.velocity : sg.update) + begin);
    float4* u4 = reinterpret_cast<float4*>(sg.update + begin);
    for (int64_t i = threadIdx.x; i < n4; i += kTfThreads) {
      const float4 g = g4[i];
      float4 p = g, w = g, ac = make_float4(0.f, 0.f, 0.f, 0.f), ve = ac;
      if (has_p) p = p4[i];
      if (wd) w = w4[i];
      if (rms) ac = a4[i];
      if (mom) ve = v4[i];
      float4 an, vn, out;
      out.x = tf_apply_elem(a, g.x, p.x, w.x, ac.x, ve.x, has_p, mult, m_in, &an.x, &vn.x);
      out.y = tf_apply_elem(a, g.y, p.y, w.y, ac.y, ve.y, has_p, mult, m_in, &an.y, &vn.y);
      out.z = tf_apply_elem(a, g.z, p.z, w.z, ac.z, ve.z, has_p, mult, m_in, &an.z, &vn.z);
      out.w = tf_apply_elem(a, g.w, p.w, w.w, ac.w, ve.w, has_p, mult, m_in, &an.w, &vn.w);
      if (rms) a4[i] = an;
      if (mom) v4[i] = vn;
      u4[i] = out;
    }
    for (int64_t e = begin + 4 * n4 + threadIdx.x; e < end; e += kTfThreads) {
      float an, vn;
      const float out = tf_apply_elem(a, sg.grad[e], has_p ? sg.precond[e] : 0.f,
                                      wd ? sg.param[e] : 0.f, rms ? sg.acc[e] : 0.f,
                                      mom ? sg.velocity[e] : 0.f, has_p, mult, m_in, &an, &vn);
      if (rms) sg.acc[e] = an;
      if (mom) sg.velocity[e] = vn;
      sg.update[e] = out;
    }
  }
}

}  // namespace pc

extern "C" {

size_t pc_tearfree_transform_workspace_bytes(int num_segments, int64_t total_chunks) {
  if (num_segments <= 0 || total_chunks <= 0) return 0;
  return pc::align_up(sizeof(float) * 2 * (size_t)total_chunks, 256) +
         pc::align_up(sizeof(float) * 2 * (size_t)num_segments, 256) + 256;
}

int pc_tearfree_transform(const pc_tearfree_segment* segments, const int32_t* chunk_segment,
                          int num_segments, int64_t total_chunks, const pc_tearfree_options* opt,
                          void* workspace, size_t workspace_bytes, void* stream) {
  PC_REQUIRE(num_segments >= 0 && total_chunks >= 0, "bad segment counts");
  if (num_segments == 0 || total_chunks == 0) return PC_OK;
  PC_REQUIRE(segments && chunk_segment && opt && workspace, "null pointer argument");
  PC_REQUIRE(total_chunks < (1ll << 31), "too many chunks");
  PC_REQUIRE(opt->graft_type >= PC_TF_GRAFT_NONE && opt->graft_type <= PC_TF_GRAFT_RMSPROP,
             "unknown tearfree graft type %d", opt->graft_type);
  if (workspace_bytes < pc_tearfree_transform_workspace_bytes(num_segments, total_chunks)) {
    pc::set_error("tearfree workspace too small: %zu < %zu", workspace_bytes,
                  pc_tearfree_transform_workspace_bytes(num_segments, total_chunks));
    return PC_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  pc::TfArgs a{};
  a.segs = segments;
  a.chunk_seg = chunk_segment;
  a.total_chunks = (int)total_chunks;
  a.nsegs = num_segments;
  a.o = *opt;
  // Python floats: 1 - decay is formed in double, then rounds to the f32 the array op sees
  a.w_new = (float)(1.0 - (double)opt->graft_decay);
  a.w_old = opt->graft_decay;
  char* w = reinterpret_cast<char*>(pc::align_up((size_t)workspace, 256));
  a.part = reinterpret_cast<float*>(w);
  a.tot = reinterpret_cast<float*>(w + pc::align_up(sizeof(float) * 2 * (size_t)total_chunks, 256));
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)std::min<int64_t>(total_chunks, (int64_t)sms * 8);
  if (opt->graft_type != PC_TF_GRAFT_NONE) {
    pc::tf_reduce_kernel<<<grid, pc::kTfThreads, 0, st>>>(a);
    pc::tf_total_kernel<<<(num_segments + 7) / 8, 256, 0, st>>>(a);
    pc::count_launch(2);
  }
  pc::tf_apply_kernel<<<grid, pc::kTfThreads, 0, st>>>(a);
  pc::count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

}  // extern "C"
