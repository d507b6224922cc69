// tcgen05 split-bf16 GEMM engine for the Newton chain -- placeholder until the
// tensor-core kernels land; the engine reports itself unavailable so that
// PC_ENGINE_AUTO resolves to the CUDA-core fp32 path.
#include "tc_engine.cuh"

namespace pc {
bool tc_engine_available() { return false; }
size_t tc_engine_bytes(int, int) { return 0; }
int tc_engine_init(TcEngine*, void*, int, int, int) {
  set_error("tcgen05 engine not built");
  return PC_ERR_UNSUPPORTED;
}
int tc_engine_iteration(TcEngine*, const float*, RootCtl*, uint32_t*, RootParams, float*, int,
                        cudaStream_t) {
  return PC_ERR_UNSUPPORTED;
}
int tc_engine_final(TcEngine*, const RootCtl*, float*, float*, cudaStream_t) {
  return PC_ERR_UNSUPPORTED;
}
}  // namespace pc
