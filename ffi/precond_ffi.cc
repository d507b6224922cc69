// XLA FFI (jax.ffi) adapter over the C ABI of include/precond_b200.h.
//
// One handler per seam of precondition/distributed_shampoo.py (DS) that hands device work
// to libprecond_b200.so.  Compiled ONLY where the XLA FFI headers exist:
//
//   g++ -std=c++17 -fPIC -shared -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I include -I /usr/local/cuda/include ffi/precond_ffi.cc \
//       -L precondition_b200 -lprecond_b200 -Wl,-rpath,'$ORIGIN/../precondition_b200' \
//       -o ffi/libprecond_b200_ffi.so
//
// (`__graft_entry__.build()` does exactly that when `import jax.ffi` works and skips the
// adapter otherwise -- this image has no JAX, so the file is syntax-checked against the stub
// header ffi/xla_ffi_stub.h instead; see tests/test_abi.py.)  ffi/precond_ffi.py registers
// the handlers and shows the call sites in the reference.
//
// Conventions: buffers arrive as device pointers in row-major layout; every handler takes
// the platform stream, enqueues and returns (the root solver is a CUDA graph with a
// device-driven loop, so nothing blocks the XLA host thread); scratch memory is a U8 result
// buffer whose size the Python side gets from the matching pc_*_workspace_bytes query.
#ifdef PC_FFI_SYNTAX_CHECK
#include "xla_ffi_stub.h"
#else
#include "xla/ffi/api/ffi.h"
#endif

#include <cuda_runtime_api.h>

#include <cstdint>
#include <vector>

#include "precond_b200.h"

namespace ffi = xla::ffi;

namespace {

inline ffi::Error Status(int rc) {
  return rc == PC_OK ? ffi::Error::Success() : ffi::Error::Internal(pc_last_error());
}
template <typename B>
inline int Dim(const B& b, int i) { return static_cast<int>(b.dimensions()[i]); }

// ---- _matrix_inverse_pth_root_vmap (DS:2742-2744) ---------------------------------------
ffi::Error InverseRootImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> xs, ffi::Buffer<ffi::S32> ps,
                           ffi::Buffer<ffi::S32> pads, float ridge_epsilon, float error_tolerance,
                           int32_t num_iters, int32_t relative_matrix_epsilon, int32_t engine,
                           ffi::ResultBuffer<ffi::F32> roots, ffi::ResultBuffer<ffi::F32> metrics,
                           ffi::ResultBuffer<ffi::U8> workspace) {
  pc_root_options opt;
  pc_root_options_default(&opt);
  opt.ridge_epsilon = ridge_epsilon;
  opt.error_tolerance = error_tolerance;
  opt.num_iters = num_iters;
  opt.relative_matrix_epsilon = relative_matrix_epsilon;
  opt.engine = engine;
  return Status(pc_inverse_pth_root_batched(
      xs.typed_data(), ps.typed_data(), pads.typed_data(), Dim(xs, 0), Dim(xs, 1), &opt,
      roots->typed_data(), metrics->typed_data(), workspace->typed_data(),
      workspace->element_count(), stream));
}

// ---- matrix_inverse_pth_root_eigh (DS:943-1030), `eigh=True` ---------------------------------
ffi::Error InverseRootEighImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> xs,
                               ffi::Buffer<ffi::S32> ps, ffi::Buffer<ffi::S32> pads,
                               float ridge_epsilon, float error_tolerance,
                               int32_t relative_matrix_epsilon, ffi::ResultBuffer<ffi::F32> roots,
                               ffi::ResultBuffer<ffi::F32> metrics,
                               ffi::ResultBuffer<ffi::U8> workspace) {
  return Status(pc_inverse_pth_root_eigh_batched(
      xs.typed_data(), ps.typed_data(), pads.typed_data(), Dim(xs, 0), Dim(xs, 1), ridge_epsilon,
      error_tolerance, relative_matrix_epsilon, roots->typed_data(), metrics->typed_data(),
      workspace->typed_data(), workspace->element_count(), stream));
}

// ---- _low_rank_root (DS:1033-1120) ----------------------------------------------------------
ffi::Error LowRankRootImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> xs, ffi::Buffer<ffi::S32> ps,
                           ffi::Buffer<ffi::S32> pads, int32_t compression_rank,
                           float ridge_epsilon, float error_tolerance,
                           int32_t relative_matrix_epsilon, ffi::ResultBuffer<ffi::F32> packed,
                           ffi::ResultBuffer<ffi::F32> metrics,
                           ffi::ResultBuffer<ffi::U8> workspace) {
  return Status(pc_low_rank_root_batched(
      xs.typed_data(), ps.typed_data(), pads.typed_data(), Dim(xs, 0), Dim(xs, 1),
      compression_rank, ridge_epsilon, error_tolerance, relative_matrix_epsilon,
      packed->typed_data(), metrics->typed_data(), workspace->typed_data(),
      workspace->element_count(), stream));
}

// ---- power_iteration (DS:595-652) -------------------------------------------------------------
ffi::Error PowerIterationImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> xs,
                              ffi::Buffer<ffi::S32> pads, int32_t num_iters, float error_tolerance,
                              ffi::ResultBuffer<ffi::F32> lambdas,
                              ffi::ResultBuffer<ffi::S32> iters) {
  return Status(pc_power_iteration_batched(xs.typed_data(), pads.typed_data(), Dim(xs, 0),
                                           Dim(xs, 1), num_iters, error_tolerance,
                                           lambdas->typed_data(), iters->typed_data(), stream));
}

// ---- _fd_update_root via new_mi_pth_root (DS:2706-2738) --------------------------------------
ffi::Error FdUpdateImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> new_grad,
                        ffi::Buffer<ffi::F32> prev, ffi::Buffer<ffi::S32> ps,
                        ffi::Buffer<ffi::S32> pads, float ridge_epsilon, float error_tolerance,
                        int32_t relative_matrix_epsilon, float decay, int32_t input_is_gram,
                        ffi::ResultBuffer<ffi::F32> out, ffi::ResultBuffer<ffi::F32> metrics,
                        ffi::ResultBuffer<ffi::U8> workspace) {
  pc_fd_options opt;
  pc_fd_options_default(&opt);
  opt.ridge_epsilon = ridge_epsilon;
  opt.error_tolerance = error_tolerance;
  opt.relative_matrix_epsilon = relative_matrix_epsilon;
  opt.decay = decay;
  opt.input_is_gram = input_is_gram;
  return Status(pc_fd_update_batched(
      new_grad.typed_data(), prev.typed_data(), ps.typed_data(), pads.typed_data(),
      Dim(new_grad, 0), Dim(new_grad, 1), Dim(new_grad, 2), Dim(prev, 2) - 2, &opt,
      out->typed_data(), metrics->typed_data(), workspace->typed_data(),
      workspace->element_count(), stream));
}

// ---- low-rank branch of _precondition_block (DS:1690-1705): packed sketch -> dense operator
ffi::Error LowRankToDenseImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> packed,
                              ffi::ResultBuffer<ffi::F32> dense,
                              ffi::ResultBuffer<ffi::U8> workspace) {
  return Status(pc_low_rank_to_dense(packed.typed_data(), Dim(packed, 0), Dim(packed, 1),
                                     Dim(packed, 2) - 2, dense->typed_data(),
                                     workspace->typed_data(), workspace->element_count(), stream));
}

// ---- the same operator in factored form (c, W = V diag(lambda^- - c)) -----------------------
ffi::Error LowRankFactorsImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> packed,
                              ffi::ResultBuffer<ffi::F32> w, ffi::ResultBuffer<ffi::F32> c) {
  return Status(pc_low_rank_factors(packed.typed_data(), Dim(packed, 0), Dim(packed, 1),
                                    Dim(packed, 2) - 2, w->typed_data(), c->typed_data(), stream));
}

// ---- tearfree _pth_inv_root (TF/shampoo.py:440-448) -----------------------------------------
ffi::Error PinvRootEighImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> xs, ffi::Buffer<ffi::S32> ps,
                            float rel_cutoff, ffi::ResultBuffer<ffi::F32> roots,
                            ffi::ResultBuffer<ffi::U8> workspace) {
  return Status(pc_pinv_pth_root_eigh_batched(
      xs.typed_data(), ps.typed_data(), Dim(xs, 0), Dim(xs, 1), rel_cutoff, roots->typed_data(),
      workspace->typed_data(), workspace->element_count(), stream));
}

// ---- _select_preconditioner (DS:2936-2950) ---------------------------------------------------
// `old` is aliased to the result by the caller (input_output_aliases={2: 0}); rows whose root
// failed keep it.
ffi::Error SelectImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> roots,
                      ffi::Buffer<ffi::F32> metrics, ffi::Buffer<ffi::F32> old, float threshold,
                      ffi::ResultBuffer<ffi::F32> out) {
  if (out->typed_data() != old.typed_data()) {
    cudaError_t e = cudaMemcpyAsync(out->typed_data(), old.typed_data(),
                                    old.element_count() * sizeof(float), cudaMemcpyDeviceToDevice,
                                    stream);
    if (e != cudaSuccess) return ffi::Error::Internal(cudaGetErrorString(e));
  }
  return Status(pc_select_preconditioners(roots.typed_data(), metrics.typed_data(), threshold,
                                          out->typed_data(), Dim(roots, 0), Dim(roots, 1),
                                          Dim(roots, 2), Dim(old, 1), Dim(old, 2), stream));
}

// ---- QuantizedValue.quantize / to_float (QU:49-113), int16 and int8 --------------------------
template <typename QBuf, int kQ>
ffi::Error QuantizeImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> x, int32_t extract_diagonal,
                        ffi::Result<QBuf> q, ffi::ResultBuffer<ffi::F32> diag,
                        ffi::ResultBuffer<ffi::F32> bucket) {
  return Status(pc_quantize_batched(x.typed_data(), Dim(x, 0), Dim(x, 1), Dim(x, 2), kQ,
                                    extract_diagonal, q->untyped_data(), diag->typed_data(),
                                    bucket->typed_data(), stream));
}
template <typename QBuf, int kQ>
ffi::Error DequantizeImpl(cudaStream_t stream, QBuf q, ffi::Buffer<ffi::F32> diag,
                          ffi::Buffer<ffi::F32> bucket, int32_t extract_diagonal,
                          ffi::ResultBuffer<ffi::F32> x) {
  return Status(pc_dequantize_batched(q.untyped_data(), diag.typed_data(), bucket.typed_data(),
                                      Dim(q, 0), Dim(q, 1), Dim(q, 2), kQ, extract_diagonal,
                                      x->typed_data(), stream));
}

// ---- gram_weighted_update (DS:1440-1470) for one batch of equally shaped blocks --------------
// g [b, m, k] (the unfolding with the preconditioned axis first), old statistics [b, m, m]:
//   S <- w1 * S + w2 * g g^T.  The descriptors are built on the host; pc_grouped_gemm_tc keeps
// the plan in `workspace` (reuse_plan = 0 here: XLA may move buffers between calls).
ffi::Error GramUpdateImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> g,
                          ffi::Buffer<ffi::F32> old_stats, float w1, float w2,
                          ffi::ResultBuffer<ffi::F32> new_stats,
                          ffi::ResultBuffer<ffi::U8> workspace) {
  const int b = Dim(g, 0), m = Dim(g, 1), k = Dim(g, 2);
  if (m % 4 != 0)  // (sizes that are no multiples of 128 run through zero-filled edge tiles)
    return ffi::Error::InvalidArgument("pc_gram_update: the tcgen05 path needs m % 4 == 0 "
                                       "(use pc_grouped_gemm with device descriptors otherwise)");
  std::vector<pc_gemm_desc> descs(b);
  for (int i = 0; i < b; ++i) {
    pc_gemm_desc& d = descs[i];
    d = pc_gemm_desc{};
    d.a = d.b = g.typed_data() + (size_t)i * m * k;
    d.c_in = old_stats.typed_data() + (size_t)i * m * m;
    d.c = new_stats->typed_data() + (size_t)i * m * m;
    d.a_iinner = m; d.a_sio = 0; d.a_si = k;
    d.a_kinner = d.b_kinner = k; d.a_sko = d.b_sko = 0; d.a_ski = d.b_ski = 1;
    d.b_sj = k;
    d.c_iinner = m; d.c_sio = 0; d.c_sii = m;
    d.m = d.n = m; d.k = k;
    d.alpha = w2; d.beta = w1;
  }
  return Status(pc_grouped_gemm_tc(descs.data(), b, workspace->typed_data(),
                                   workspace->element_count(), /*reuse_plan=*/0, stream));
}

// ---- tail of _transform_grad (DS:3496-3625) for one parameter --------------------------------
ffi::Error GraftMomentumImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> grad,
                             ffi::Buffer<ffi::F32> param, ffi::Buffer<ffi::F32> precond_grad,
                             ffi::Buffer<ffi::F32> diag, ffi::Buffer<ffi::F32> dmom,
                             ffi::Buffer<ffi::F32> mom, double beta1, double beta2,
                             int32_t graft_type, float diagonal_epsilon, float weight_decay,
                             float learning_rate, int32_t nesterov, int32_t moving_average,
                             int32_t decoupled_lr, int32_t decoupled_wd, int32_t run_shampoo,
                             int32_t has_precond, float clip,
                             ffi::ResultBuffer<ffi::F32> update,
                             ffi::ResultBuffer<ffi::F32> new_diag,
                             ffi::ResultBuffer<ffi::F32> new_dmom,
                             ffi::ResultBuffer<ffi::F32> new_mom,
                             ffi::ResultBuffer<ffi::U8> workspace) {
  const int64_t numel = static_cast<int64_t>(grad.element_count());
  // state arrives as inputs and leaves as results (aliased by the caller when donated)
  struct { const float* src; float* dst; } copies[3] = {{diag.typed_data(), new_diag->typed_data()},
                                                        {dmom.typed_data(), new_dmom->typed_data()},
                                                        {mom.typed_data(), new_mom->typed_data()}};
  for (auto& c : copies)
    if (c.src != c.dst) {
      cudaError_t e = cudaMemcpyAsync(c.dst, c.src, numel * sizeof(float),
                                      cudaMemcpyDeviceToDevice, stream);
      if (e != cudaSuccess) return ffi::Error::Internal(cudaGetErrorString(e));
    }
  pc_graft_options o{};
  o.beta1 = beta1; o.beta2 = beta2; o.graft_type = graft_type;
  o.diagonal_epsilon = diagonal_epsilon; o.weight_decay = weight_decay;
  o.learning_rate = learning_rate; o.nesterov = nesterov;
  o.moving_average_for_momentum = moving_average; o.decoupled_learning_rate = decoupled_lr;
  o.decoupled_weight_decay = decoupled_wd; o.run_shampoo = run_shampoo;
  o.clip_by_scaled_gradient_norm = clip;
  return Status(pc_graft_momentum(grad.typed_data(), param.typed_data(),
                                  has_precond ? precond_grad.typed_data() : nullptr,
                                  new_diag->typed_data(), new_dmom->typed_data(),
                                  new_mom->typed_data(), update->typed_data(), numel, &o,
                                  workspace->typed_data(), workspace->element_count(), stream));
}

}  // namespace

#define PC_STREAM ffi::Ffi::Bind().Ctx<ffi::PlatformStream<cudaStream_t>>()
using F32 = ffi::Buffer<ffi::F32>;
using S32 = ffi::Buffer<ffi::S32>;
using S16 = ffi::Buffer<ffi::S16>;
using S8 = ffi::Buffer<ffi::S8>;
using U8 = ffi::Buffer<ffi::U8>;

XLA_FFI_DEFINE_HANDLER_SYMBOL(PcInverseRoot, InverseRootImpl,
    PC_STREAM.Arg<F32>().Arg<S32>().Arg<S32>()
        .Attr<float>("ridge_epsilon").Attr<float>("error_tolerance").Attr<int32_t>("num_iters")
        .Attr<int32_t>("relative_matrix_epsilon").Attr<int32_t>("engine")
        .Ret<F32>().Ret<F32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcInverseRootEigh, InverseRootEighImpl,
    PC_STREAM.Arg<F32>().Arg<S32>().Arg<S32>()
        .Attr<float>("ridge_epsilon").Attr<float>("error_tolerance")
        .Attr<int32_t>("relative_matrix_epsilon").Ret<F32>().Ret<F32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcLowRankRoot, LowRankRootImpl,
    PC_STREAM.Arg<F32>().Arg<S32>().Arg<S32>().Attr<int32_t>("compression_rank")
        .Attr<float>("ridge_epsilon").Attr<float>("error_tolerance")
        .Attr<int32_t>("relative_matrix_epsilon").Ret<F32>().Ret<F32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcPowerIteration, PowerIterationImpl,
    PC_STREAM.Arg<F32>().Arg<S32>().Attr<int32_t>("num_iters").Attr<float>("error_tolerance")
        .Ret<F32>().Ret<S32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcFdUpdate, FdUpdateImpl,
    PC_STREAM.Arg<F32>().Arg<F32>().Arg<S32>().Arg<S32>()
        .Attr<float>("ridge_epsilon").Attr<float>("error_tolerance")
        .Attr<int32_t>("relative_matrix_epsilon").Attr<float>("decay")
        .Attr<int32_t>("input_is_gram").Ret<F32>().Ret<F32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcLowRankToDense, LowRankToDenseImpl,
    PC_STREAM.Arg<F32>().Ret<F32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcLowRankFactors, LowRankFactorsImpl,
    PC_STREAM.Arg<F32>().Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcPinvRootEigh, PinvRootEighImpl,
    PC_STREAM.Arg<F32>().Arg<S32>().Attr<float>("rel_cutoff").Ret<F32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcSelectPreconditioners, SelectImpl,
    PC_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Attr<float>("threshold").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcQuantizeInt16, (QuantizeImpl<S16, PC_QDTYPE_INT16>),
    PC_STREAM.Arg<F32>().Attr<int32_t>("extract_diagonal").Ret<S16>().Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcQuantizeInt8, (QuantizeImpl<S8, PC_QDTYPE_INT8>),
    PC_STREAM.Arg<F32>().Attr<int32_t>("extract_diagonal").Ret<S8>().Ret<F32>().Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcDequantizeInt16, (DequantizeImpl<S16, PC_QDTYPE_INT16>),
    PC_STREAM.Arg<S16>().Arg<F32>().Arg<F32>().Attr<int32_t>("extract_diagonal").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcDequantizeInt8, (DequantizeImpl<S8, PC_QDTYPE_INT8>),
    PC_STREAM.Arg<S8>().Arg<F32>().Arg<F32>().Attr<int32_t>("extract_diagonal").Ret<F32>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcGramUpdate, GramUpdateImpl,
    PC_STREAM.Arg<F32>().Arg<F32>().Attr<float>("w1").Attr<float>("w2").Ret<F32>().Ret<U8>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcGraftMomentum, GraftMomentumImpl,
    PC_STREAM.Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>().Arg<F32>()
        .Attr<double>("beta1").Attr<double>("beta2").Attr<int32_t>("graft_type")
        .Attr<float>("diagonal_epsilon").Attr<float>("weight_decay").Attr<float>("learning_rate")
        .Attr<int32_t>("nesterov").Attr<int32_t>("moving_average_for_momentum")
        .Attr<int32_t>("decoupled_learning_rate").Attr<int32_t>("decoupled_weight_decay")
        .Attr<int32_t>("run_shampoo").Attr<int32_t>("has_precond")
        .Attr<float>("clip_by_scaled_gradient_norm")
        .Ret<F32>().Ret<F32>().Ret<F32>().Ret<F32>().Ret<U8>());
