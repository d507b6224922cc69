// Copy-engine all-gather over NVLink peer memory (one process per GPU, CUDA IPC).
//
// Why not NCCL here: the Newton-chain GEMMs are persistent kernels that own every SM (one
// CTA per SM, ~200 KB of shared memory, the whole register file).  An NCCL all-gather that
// runs beside them needs SMs of its own, so the next GEMM launch waits for them (measured:
// a 2-GPU step got SLOWER when the gather of one sub-batch overlapped the solve of the next).
// This exchange uses no SM at all: every rank PUSHES its payload into the peers' receive
// buffers with cudaMemcpyAsync (copy engines, NVLink), then pushes a 4-byte epoch flag behind
// it on the same stream; consumers wait for the flags with cuStreamWaitValue32.  Everything
// is stream-ordered -- no host synchronisation, no kernel.
//
//   recv buffer of rank r : [world][bytes]            (slot s = payload of rank s)
//   flags of rank r       : data[world], ack[world]   (uint32 epochs), epoch word
// all_gather(epoch e) on `stream`:
//   wait ack[p] >= e-1 for every peer p   (p has consumed what I pushed last time)
//   epoch word <- e; local slot <- payload; for every peer p: p.recv[me] <- payload,
//   p.data[me] <- e;  wait data[p] >= e for every peer p.
// release(epoch e): p.ack[me] <- e for every peer p (enqueue after the consumers of recv).
#include <cuda.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>

#include "common.cuh"

namespace pc {

typedef CUresult (*WaitValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*WriteValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*AddrRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);

template <typename F>
static F driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<F>(p);
}

constexpr int kPushStreams = 4;
struct PushStreams {
  cudaStream_t s[kPushStreams];
  cudaEvent_t fork, join[kPushStreams];
};
// helper streams of the calling host thread on the current device (created once)
static PushStreams* push_streams() {
  static thread_local PushStreams* per_dev[64] = {nullptr};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  PushStreams*& p = per_dev[dev & 63];
  if (p) return p;
  PushStreams* n = new PushStreams();
  bool ok = cudaEventCreateWithFlags(&n->fork, cudaEventDisableTiming) == cudaSuccess;
  for (int i = 0; i < kPushStreams && ok; ++i)
    ok = cudaStreamCreateWithFlags(&n->s[i], cudaStreamNonBlocking) == cudaSuccess &&
         cudaEventCreateWithFlags(&n->join[i], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    set_error("cannot create the push streams: %s", cudaGetErrorString(cudaGetLastError()));
    delete n;
    return nullptr;
  }
  p = n;
  return p;
}

static std::mutex g_ipc_mu;
static std::map<std::string, void*> g_ipc_open;  // handle bytes -> mapped base

}  // namespace pc

extern "C" {

int pc_ipc_export(const void* dev_ptr, pc_ipc_handle* out) {
  PC_REQUIRE(dev_ptr && out, "null pointer argument");
  static pc::AddrRangeFn range = pc::driver_fn<pc::AddrRangeFn>("cuMemGetAddressRange");
  if (!range) {
    pc::set_error("cuMemGetAddressRange not available from the driver");
    return PC_ERR_UNSUPPORTED;
  }
  CUdeviceptr base = 0;
  size_t size = 0;
  CUresult r = range(&base, &size, (CUdeviceptr)dev_ptr);
  if (r != CUDA_SUCCESS) {
    pc::set_error("cuMemGetAddressRange failed with %d", (int)r);
    return PC_ERR_CUDA;
  }
  memset(out, 0, sizeof(*out));
  cudaIpcMemHandle_t h;
  PC_CUDA_CHECK(cudaIpcGetMemHandle(&h, (void*)base));
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out->handle, &h, 64);
  out->offset = (int64_t)((CUdeviceptr)dev_ptr - base);
  out->size = (int64_t)size;
  int dev = 0;
  cudaGetDevice(&dev);
  out->device = dev;
  return PC_OK;
}

int pc_ipc_open(const pc_ipc_handle* h, void** out_ptr) {
  PC_REQUIRE(h && out_ptr, "null pointer argument");
  std::lock_guard<std::mutex> lock(pc::g_ipc_mu);
  std::string key(reinterpret_cast<const char*>(h->handle), 64);
  auto it = pc::g_ipc_open.find(key);
  void* base = nullptr;
  if (it == pc::g_ipc_open.end()) {
    cudaIpcMemHandle_t mh;
    memcpy(&mh, h->handle, 64);
    PC_CUDA_CHECK(cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess));
    pc::g_ipc_open[key] = base;
  } else {
    base = it->second;
  }
  *out_ptr = reinterpret_cast<char*>(base) + h->offset;
  return PC_OK;
}

int pc_peer_all_gather(const pc_peer_group* g, const void* send, size_t bytes, uint32_t epoch,
                       void* stream) {
  PC_REQUIRE(g && send && g->world >= 1 && g->world <= PC_MAX_PEERS && g->rank >= 0 &&
                 g->rank < g->world && epoch >= 1, "bad peer group / epoch");
  PC_REQUIRE(g->slot_bytes >= 0 && bytes <= (size_t)g->slot_bytes, "payload (%zu B) exceeds the slot size (%zu B)", bytes,
             (size_t)g->slot_bytes);
  static pc::WaitValueFn wait = pc::driver_fn<pc::WaitValueFn>("cuStreamWaitValue32");
  static pc::WriteValueFn write = pc::driver_fn<pc::WriteValueFn>("cuStreamWriteValue32");
  if (!wait || !write) {
    pc::set_error("cuStreamWaitValue32 / cuStreamWriteValue32 not available from the driver");
    return PC_ERR_UNSUPPORTED;
  }
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* flags = reinterpret_cast<uint32_t*>(g->flags[g->rank]);
  uint32_t* data = flags;                     // [world]
  uint32_t* ack = flags + PC_MAX_PEERS;       // [world]
  uint32_t* word = flags + 2 * PC_MAX_PEERS;  // epoch word
#define PC_DRV(expr)                                                     \
  do {                                                                   \
    CUresult _r = (expr);                                                \
    if (_r != CUDA_SUCCESS) {                                            \
      pc::set_error("%s failed with %d", #expr, (int)_r);                \
      return PC_ERR_CUDA;                                                \
    }                                                                    \
  } while (0)
  for (int p = 0; p < g->world; ++p)
    if (p != g->rank)
      PC_DRV(wait((CUstream)st, (CUdeviceptr)(ack + p), epoch - 1, CU_STREAM_WAIT_VALUE_GEQ));
  PC_DRV(write((CUstream)st, (CUdeviceptr)word, epoch, CU_STREAM_WRITE_VALUE_DEFAULT));
  char* mine = reinterpret_cast<char*>(g->recv[g->rank]) + (size_t)g->rank * g->slot_bytes;
  if (mine != send)
    PC_CUDA_CHECK(cudaMemcpyAsync(mine, send, bytes, cudaMemcpyDeviceToDevice, st));
  // The pushes to the peers go out on up to kPushStreams helper streams (several copy engines
  // and NVLink targets at once); every helper pushes payload then flag, in order.
  pc::PushStreams* ps = pc::push_streams();
  if (!ps) return PC_ERR_CUDA;
  const int nhelp = g->world - 1 < pc::kPushStreams ? g->world - 1 : pc::kPushStreams;
  if (nhelp > 0) PC_CUDA_CHECK(cudaEventRecord(ps->fork, st));
  for (int i = 0; i < nhelp; ++i) PC_CUDA_CHECK(cudaStreamWaitEvent(ps->s[i], ps->fork, 0));
  for (int k = 1; k < g->world; ++k) {  // start with the next rank: spreads the NVLink targets
    const int p = (g->rank + k) % g->world;
    cudaStream_t hs = ps->s[(k - 1) % nhelp];
    char* dst = reinterpret_cast<char*>(g->recv[p]) + (size_t)g->rank * g->slot_bytes;
    PC_CUDA_CHECK(cudaMemcpyAsync(dst, send, bytes, cudaMemcpyDefault, hs));
    uint32_t* pflag = reinterpret_cast<uint32_t*>(g->flags[p]) + g->rank;
    PC_CUDA_CHECK(cudaMemcpyAsync(pflag, word, sizeof(uint32_t), cudaMemcpyDefault, hs));
  }
  for (int i = 0; i < nhelp; ++i) {  // `send` and the epoch word may be reused after the call
    PC_CUDA_CHECK(cudaEventRecord(ps->join[i], ps->s[i]));
    PC_CUDA_CHECK(cudaStreamWaitEvent(st, ps->join[i], 0));
  }
  for (int p = 0; p < g->world; ++p)
    if (p != g->rank)
      PC_DRV(wait((CUstream)st, (CUdeviceptr)(data + p), epoch, CU_STREAM_WAIT_VALUE_GEQ));
  return PC_OK;
}

int pc_peer_release(const pc_peer_group* g, uint32_t epoch, void* stream) {
  PC_REQUIRE(g && g->world >= 1 && g->world <= PC_MAX_PEERS && g->rank >= 0 &&
                 g->rank < g->world, "bad peer group");
  (void)epoch;  // the epoch word still holds it: nothing on this stream has rewritten it
  cudaStream_t st = (cudaStream_t)stream;
  uint32_t* word = reinterpret_cast<uint32_t*>(g->flags[g->rank]) + 2 * PC_MAX_PEERS;
  for (int k = 1; k < g->world; ++k) {
    const int p = (g->rank + k) % g->world;
    uint32_t* pack = reinterpret_cast<uint32_t*>(g->flags[p]) + PC_MAX_PEERS + g->rank;
    PC_CUDA_CHECK(cudaMemcpyAsync(pack, word, sizeof(uint32_t), cudaMemcpyDefault, st));
  }
#undef PC_DRV
  return PC_OK;
}

}  // extern "C"
