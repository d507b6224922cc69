// tcgen05 split-bf16 engine for the Newton chain (implemented in tc_gemm.cu).
#pragma once
#include "root_common.cuh"

namespace pc {

struct TcEngine {
  int batch, n, passes;  // passes: 6 (3 planes, ~fp32) or 3 (2 planes)
  int fmt;               // 0: bf16 planes, 1: fp16 + 2^11-scaled fp16 residual (3 passes)
  // 8 logical matrices x 3 bf16 planes, each [batch, n, n]
  uint16_t* planes[kNumBufs][3];
  void* tmaps;  // device copy of the CUtensorMap table (one per buffer x plane)
  void* host_state;
};

bool tc_engine_available();
int tc_engine_variant_mask();
// one-time per-device setup (kernel attributes, constants); call before any stream capture
int tc_engine_prepare();
size_t tc_engine_bytes(int batch, int n, int planes);
int tc_engine_init(TcEngine* e, void* mem, int batch, int n, int passes, int fmt,
                   cudaStream_t stream);
// first_init_scratch != nullptr: first iteration of a call -- initialise with the whole-GPU
// strip kernels (scratch of batch * (n / 32 + 2) floats)
int tc_engine_iteration(TcEngine* e, const float* xs, RootCtl* ctl, uint32_t* errbits,
                        RootParams prm, float* roots, int max_steps, float* first_init_scratch,
                        bool do_init, cudaStream_t stream);
int tc_engine_final(TcEngine* e, const RootCtl* ctl, float* roots, float* metrics,
                    cudaStream_t stream);

}  // namespace pc
