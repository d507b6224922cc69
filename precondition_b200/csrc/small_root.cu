// Persistent per-matrix inverse p-th root solver for small statistics (n <= 128; the
// reference's DEFAULT block size is 128, DS:1917-1920): ONE CTA runs the whole
// matrix_inverse_pth_root of one matrix (DS:702-940) -- power iteration, ridge, every coupled
// Newton iteration, retries, convergence test -- with all iterates resident on the SM.  No
// host involvement, no per-iteration launches, one launch per batch.
//
//   shared memory   M and the running chain matrix (M_i -> M_i^2 -> ... -> M_i^p), each as three
//                   bf16 planes X0 + X1 + X2 == X (exact split of fp32) in the K-major
//                   SWIZZLE_128B layout tcgen05.mma reads: 2 x 96 KB
//   tensor memory   two 128 x 128 fp32 chunk accumulators (one per 64-column K-chunk, summed in
//                   fp32 registers with round-to-nearest, as in the large engine) and H as the
//                   A operand of H' = H M_i (three bf16 planes, 64 columns each): 448 columns
//   global memory   H' in fp32, ping-pong (`roots` itself and a scratch slot): the answer is
//                   H or the previous H depending on the final error ratio (DS:878-880)
//
// Products are the exact 6-term bf16 split (sum_{i+j<=2} A_i B_j, dropped terms <= 2^-23|a||b|),
// smallest terms first.  Supported exponents: p = 2^s (1, 2, 4, 8, 16) -- M_i^p is then a pure
// squaring chain that runs IN PLACE in the chain buffer; other exponents use the generic
// engines.  Every iterate of the M chain is bitwise symmetric (squarings are symmetric by
// construction; M' = M_i^p M is mirrored from its lower triangle), which the recurrence needs
// (see root.cu); the root is symmetrised from its lower triangle when it is written out.
#include <cuda.h>
#include <cuda_bf16.h>

#include <vector>

#include "root_common.cuh"

namespace pc {

namespace small {

constexpr int kN = 128;                    // padded problem size
constexpr int kThreads = 256;              // 8 warps: lane quadrant = warp % 4, column half = warp / 4
constexpr int kTileBytes = 128 * 64 * 2;   // one plane, one 64-column K-tile: 16 KiB
constexpr int kMatBytes = 6 * kTileBytes;  // 3 planes x 2 K-tiles: 96 KiB
constexpr uint32_t kAcc0 = 0, kAcc1 = 128, kHCols = 256;
constexpr int kLdA = kN + 1;  // row stride of the fp32 input in shared memory (bank-conflict free)  // TMEM column map (H: 3 x 64 columns)

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
               : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]   (A: lane = row, two bf16 per 32-bit column along K)
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc,
                                        uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(
          taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
      "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]),
      "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]),
      "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]),
      "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() {
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"), SBO = 1024 B
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// c = F32, a = b = BF16, K-major both, N = 128, M = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(128 >> 3) << 17) |
                            ((uint32_t)(128 >> 4) << 24);

// byte offset of the 16-byte chunk holding columns [8 ch, 8 ch + 8) of row r inside one
// matrix buffer (plane pl): K-tile = ch / 8, SWIZZLE_128B inside the tile
__device__ __forceinline__ uint32_t plane_chunk_off(int pl, int r, int ch) {
  const int kt = ch >> 3, c = ch & 7;
  return (uint32_t)((pl * 2 + kt) * kTileBytes + r * 128 + ((c ^ (r & 7)) << 4));
}
__device__ __forceinline__ uint32_t plane_elem_off(int pl, int r, int col) {
  return plane_chunk_off(pl, r, col >> 3) + (uint32_t)((col & 7) * 2);
}

// exact 3-way bf16 split of 2 fp32 values -> packed pairs (low half = first value)
__device__ __forceinline__ void split2(float a, float b, uint32_t& p0, uint32_t& p1, uint32_t& p2) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  p0 = *reinterpret_cast<uint32_t*>(&h);
  a -= __uint_as_float(p0 << 16);
  b -= __uint_as_float(p0 & 0xffff0000u);
  h = __floats2bfloat162_rn(a, b);
  p1 = *reinterpret_cast<uint32_t*>(&h);
  a -= __uint_as_float(p1 << 16);
  b -= __uint_as_float(p1 & 0xffff0000u);
  h = __floats2bfloat162_rn(a, b);
  p2 = *reinterpret_cast<uint32_t*>(&h);
}

// Stores this thread's row segment x[0..64) (columns c0 .. c0 + 63 of row r) as planes: full
// rows, 16-byte stores (for matrices that are symmetric by construction).
__device__ __forceinline__ void store_planes(uint8_t* buf, int r, int c0, const float (&x)[64]) {
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {
    uint32_t w[3][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      split2(x[ch * 8 + 2 * k], x[ch * 8 + 2 * k + 1], w[0][k], w[1][k], w[2][k]);
    const int col = c0 + ch * 8;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl)
      *reinterpret_cast<uint4*>(buf + plane_chunk_off(pl, r, col >> 3)) =
          make_uint4(w[pl][0], w[pl][1], w[pl][2], w[pl][3]);
  }
}

__device__ __forceinline__ float block_sum_256(float v, float* scratch) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) r += scratch[w];  // same order in every thread
  return r;
}
// two sums at once (one barrier pair): a -> r.x, b -> r.y
__device__ __forceinline__ float2 block_sum2_256(float a, float b, float2* scratch) {
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = make_float2(a, b);
  __syncthreads();
  float2 r = make_float2(0.f, 0.f);
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) {
    r.x += scratch[w].x;
    r.y += scratch[w].y;
  }
  return r;
}
__device__ __forceinline__ uint32_t block_max_256(uint32_t v, uint32_t* scratch) {
  v = warp_max_u32(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  uint32_t r = 0;
#pragma unroll
  for (int w = 0; w < kThreads / 32; ++w) r = scratch[w] > r ? scratch[w] : r;
  return r;
}

struct SmallShared {
  alignas(16) float v[kN];
  alignas(16) float nv[kN];
  alignas(16) float y[2][kN];
  float red[8];
  float2 red2[8];
  uint32_t ured[8];
  RootCtl ctl;
  uint64_t bar[2];
  uint32_t tmem_base;
  long long prof[12];  // PC_SMALL_PROF: cycles per phase (thread 0)
};

// 6-term split product of one 64-column K-chunk into accumulator `acc`; A from shared memory
// (a_buf) or from tensor memory (a_tmem, column of plane 0 of this chunk), B from b_buf
template <bool kATmem, int kMaxSum>
__device__ __forceinline__ void issue_chunk(uint32_t acc, uint32_t a_buf, uint32_t a_tmem,
                                            uint32_t b_buf, int kt) {
  bool first = true;
#pragma unroll
  for (int sum = kMaxSum; sum >= 0; --sum) {  // smallest terms first
#pragma unroll
    for (int i = 0; i <= 2; ++i) {
      const int j = sum - i;
      if (j < 0 || j > 2) continue;
      const uint64_t bd = make_desc(b_buf + (j * 2 + kt) * kTileBytes);
      const uint64_t ad = make_desc(a_buf + (i * 2 + kt) * kTileBytes);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (kATmem)
          umma_ts(acc, a_tmem + i * 64 + kt * 32 + k * 8, bd + 2u * k, kIdesc, first ? 0u : 1u);
        else
          umma_ss(acc, ad + 2u * k, bd + 2u * k, kIdesc, first ? 0u : 1u);
        first = false;
      }
    }
  }
}

}  // namespace small

using namespace small;

// xs [batch, n, n]; hslot [batch, 2, 128 * 128] scratch (H and the previous H, DS:878-880)
template <int kMaxSum>  // plane products A_i B_j with i + j <= kMaxSum (2: six terms, 3: eight)
__global__ void __launch_bounds__(kThreads, 1)
small_root_kernel(const float* __restrict__ xs, const int32_t* __restrict__ ps,
                  const int32_t* __restrict__ pads, int batch, int n, RootParams prm,
                  const float* __restrict__ v0, float* __restrict__ roots,
                  float* __restrict__ hslot, float* __restrict__ metrics,
                  long long* __restrict__ prof) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~(uintptr_t)1023);
  uint8_t* bufM = base;              // M  (3 planes x 2 K-tiles)
  uint8_t* bufX = base + kMatBytes;  // M_i, then M_i^2, M_i^4, ...; fp32 A during setup
  SmallShared& S = *reinterpret_cast<SmallShared*>(base + 2 * kMatBytes);
  float* Af = reinterpret_cast<float*>(bufX);  // [128][128] fp32, only while no planes live there
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int q = warp & 3, half = warp >> 2;
  const int row = q * 32 + lane, c0 = half * 64;
  const uint32_t lane_off = (uint32_t)(q * 32) << 16;
  const uint32_t bar0 = smem_u32(&S.bar[0]), bar1 = smem_u32(&S.bar[1]);

  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar1, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&S.tmem_base), 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = S.tmem_base;
  uint32_t phase = 0;  // parity of both MMA barriers (one completion each per product)
  long long t_prev = 0;
  if (prof && tid == 0) {
    for (int i = 0; i < 12; ++i) S.prof[i] = 0;
    t_prev = clock64();
  }
  // diagnostics: cycles since the previous mark go to slot `k`
#define PC_MARK(k)                                   \
  do {                                               \
    if (prof && tid == 0) {                          \
      const long long t_now = clock64();             \
      S.prof[k] += t_now - t_prev;                   \
      t_prev = t_now;                                \
    }                                                \
  } while (0)

  // one product: OUT = A * B with 2 K-chunks -> acc0, acc1; returns after the MMAs were issued
  auto issue_product = [&](bool a_tmem, uint8_t* a_buf, uint8_t* b_buf) {
    fence_proxy_async_smem();  // operand planes written with st.shared -> visible to the MMA
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) {
      tcgen05_fence_after();
      if (lane == 0) {
        const uint32_t a32 = smem_u32(a_buf), b32 = smem_u32(b_buf);
        if (a_tmem) issue_chunk<true, kMaxSum>(tmem + kAcc0, a32, tmem + kHCols, b32, 0);
        else issue_chunk<false, kMaxSum>(tmem + kAcc0, a32, 0, b32, 0);
        umma_commit(bar0);
        if (a_tmem) issue_chunk<true, kMaxSum>(tmem + kAcc1, a32, tmem + kHCols, b32, 1);
        else issue_chunk<false, kMaxSum>(tmem + kAcc1, a32, 0, b32, 1);
        umma_commit(bar1);
      }
      __syncwarp();
    }
  };
  // x[0..64) = acc0 + acc1 (round to nearest) for (row, c0 ..); ends with all MMAs retired
  auto pull = [&](float (&x)[64]) {
    mbar_wait(bar0, phase);
    tcgen05_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem + lane_off + kAcc0 + c0 + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) x[c * 32 + i] = __uint_as_float(r[i]);
    }
    mbar_wait(bar1, phase);
    tcgen05_fence_after();
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t r[32];
      tmem_ld_32x32(tmem + lane_off + kAcc1 + c0 + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) x[c * 32 + i] = __fadd_rn(x[c * 32 + i], __uint_as_float(r[i]));
    }
    phase ^= 1;
    tcgen05_fence_before();
    __syncthreads();  // every warp has drained the accumulators: operands / TMEM may be rewritten
  };

  for (int b = blockIdx.x; b < batch; b += gridDim.x) {
    const float* A = xs + (size_t)b * n * n;
    // H ping-pong in fp32, "fragment" layout: the value of (row, c0 + k) of thread tid sits at
    // [(k / 4) * 256 + tid][k % 4] -- every warp store is one contiguous 512-byte run
    float* slots[2] = {hslot + (size_t)blockIdx.x * 2 * kN * kN,
                       hslot + ((size_t)blockIdx.x * 2 + 1) * kN * kN};
    // ---- setup (thread 0 owns the control block; mirrors root_setup_kernel) ----
    if (tid == 0) {
      RootCtl c;
      memset(&c, 0, sizeof(c));
      c.p = ps[b];
      int pad = pads ? pads[b] : n;
      c.pad = pad < 0 ? 0 : (pad > n ? n : pad);
      c.max_ev = 1.0f; c.ratio = 1.0f; c.err = 1000.0f;
      c.m_err = 1000.0f; c.m_iters = 100.f; c.m_ratio = 1.0f;
      const bool pow2 = c.p >= 1 && c.p <= kMaxP && (c.p & (c.p - 1)) == 0;
      if (c.pad == 0 || !pow2) {
        c.done = 1;
        c.result_h = -2;
        if (c.pad != 0) c.m_err = __int_as_float(0x7fc00000);
      } else {
        c.need_init = 1;
      }
      S.ctl = c;
    }
    __syncthreads();
    const int pad = S.ctl.pad, p = S.ctl.p;
    if (!S.ctl.done) {
      // masked input, lower triangle authoritative, zero-padded to 128 x 128
      for (int e = tid; e < kN * kN; e += kThreads) {
        const int i = e >> 7, j = e & 127;
        Af[i * kLdA + j] = (i < pad && j < pad) ? __ldg(A + (size_t)max(i, j) * n + min(i, j)) : 0.f;
      }
      __syncthreads();
      if (n == 1) {  // DS:850-855
        if (tid == 0) {
          const float a = Af[0];
          const float ev = prm.relative_eps ? a : 1.0f;
          RootCtl& c = S.ctl;
          c.max_ev = ev;
          c.ridge = prm.ridge_epsilon * fmaxf(ev, 1e-25f);
          roots[(size_t)b] = powf(a + c.ridge, -1.0f / (float)p);
          c.need_init = 0; c.done = 1;
          c.m_err = 0.f; c.m_iters = 0.f; c.m_ratio = 0.f; c.m_retries = 0.f;
          c.result_h = -1;
        }
        __syncthreads();
      } else if (prm.relative_eps) {
        // ---- power iteration (DS:595-652), y = A nv through the symmetric columns ----
        for (int i = tid; i < kN; i += kThreads) S.v[i] = i < pad ? v0[i] : 0.f;
        __syncthreads();
        float s = 0.f;
        int it = 0;
        bool run = true;
        const int col = tid & 127, part = tid >> 7;  // two threads per output element
        const int j0 = part * 64;
        float areg[64];  // this thread's 64 entries of column `col` stay in registers
#pragma unroll
        for (int j = 0; j < 64; ++j) areg[j] = Af[(j0 + j) * kLdA + col];
        float nrm2 = 0.f;  // |v|^2 of the current iterate
        {
          const float ss = tid < kN ? S.v[tid] * S.v[tid] : 0.f;
          nrm2 = block_sum_256(ss, S.red);
        }
        while (it < 100 && run) {
          const float norm = sqrtf(nrm2);
          if (tid < kN) S.nv[tid] = S.v[tid] / norm;  // DS:634
          __syncthreads();
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
          for (int j = 0; j < 64; j += 4) {
            const float4 nv4 = *reinterpret_cast<const float4*>(&S.nv[j0 + j]);  // broadcast
            a0 = fmaf(areg[j], nv4.x, a0);
            a1 = fmaf(areg[j + 1], nv4.y, a1);
            a2 = fmaf(areg[j + 2], nv4.z, a2);
            a3 = fmaf(areg[j + 3], nv4.w, a3);
          }
          S.y[part][col] = (a0 + a1) + (a2 + a3);
          __syncthreads();
          float dot = 0.f, yy = 0.f;
          if (tid < kN) {
            const float yi = S.y[0][tid] + S.y[1][tid];  // DS:636
            S.v[tid] = yi;
            dot = S.nv[tid] * yi;
            yy = yi * yi;
          }
          const float2 r2 = block_sum2_256(dot, yy, S.red2);  // DS:637 and the next |v|^2
          const float s_new = r2.x;
          nrm2 = r2.y;
          run = fabsf(s_new - s) > 1e-6f;                 // DS:639 (NaN -> stop)
          s = s_new;
          ++it;
        }
        __syncthreads();
        if (tid == 0) S.ctl.max_ev = s;
        __syncthreads();
      }
    }
    PC_MARK(0);

    // ---- tries (DS:858-885) ----
    while (!S.ctl.done) {
      if (S.ctl.need_init) {
        if (S.ctl.tries > 0) {  // the fp32 input was overwritten by the planes: reload it
          __syncthreads();
          for (int e = tid; e < kN * kN; e += kThreads) {
            const int i = e >> 7, j = e & 127;
            Af[i * kLdA + j] = (i < pad && j < pad) ? __ldg(A + (size_t)max(i, j) * n + min(i, j)) : 0.f;
          }
          __syncthreads();
        }
        RootCtl c = S.ctl;
        if (c.tries == 0) {
          const float ev = prm.relative_eps ? c.max_ev : 1.0f;
          c.max_ev = ev;
          c.ridge = prm.ridge_epsilon * fmaxf(ev, 1e-25f);  // DS:830
        }
        float tenpow = 1.f;
        for (int t = 0; t < c.tries; ++t) tenpow *= 10.f;
        const float eps = c.ridge * tenpow;  // DS:869
        const float alpha = -1.0f / (float)p, oma = 1.0f - alpha;
        // this thread's row segment of the damped matrix
        float x[64];
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          float a = Af[row * kLdA + c0 + k];
          if (row == c0 + k && row < pad) a += eps;
          x[k] = a;
          ss = fmaf(a, a, ss);
        }
        const float norm = sqrtf(block_sum_256(ss, S.red));  // DS:870
        const float z = (float)(1 + p) / (2.0f * norm);
        const float h0 = powf(z, (float)(1.0 / (double)p));  // DS:873
        uint32_t emax = 0;
        float mi[64];
#pragma unroll
        for (int k = 0; k < 64; ++k) {
          const bool in = row < pad && c0 + k < pad;
          const bool dg = row == c0 + k;
          const float m0 = in ? x[k] * z : 0.f;  // DS:871
          if (in) {
            const uint32_t ab = absbits(m0 - (dg ? 1.f : 0.f));
            emax = ab > emax ? ab : emax;
          }
          x[k] = m0;
          mi[k] = in ? mi_from_m(m0, dg, alpha, oma) : 0.f;
        }
        emax = block_max_256(emax, S.ured);  // (barrier: every thread has read its part of Af)
        store_planes(bufM, row, c0, x);
        store_planes(bufX, row, c0, mi);
        // H0 = z^(1/p) I_m as the TMEM A operand (lane = row, bf16 pairs along the columns)
        {
          uint32_t h0p[3] = {0u, 0u, 0u};
          float hv = row < pad ? h0 : 0.f;
          uint32_t d0, d1, d2;
          split2(hv, 0.f, d0, d1, d2);
          h0p[0] = d0 & 0xffffu; h0p[1] = d1 & 0xffffu; h0p[2] = d2 & 0xffffu;
#pragma unroll
          for (int pl = 0; pl < 3; ++pl) {
            uint32_t w[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int cpair = c0 + 2 * k;  // columns cpair, cpair + 1
              w[k] = (cpair == row) ? h0p[pl] : ((cpair + 1 == row) ? (h0p[pl] << 16) : 0u);
            }
            tmem_st_32x32(tmem + lane_off + kHCols + pl * 64 + half * 32, w);
          }
          tmem_st_wait();
          // fp32 H0 into slot 0 (the answer if the very first step diverges)
#pragma unroll
          for (int k = 0; k < 64; k += 4)
            *reinterpret_cast<float4*>(slots[0] + ((size_t)(k >> 2) * kThreads + tid) * 4) =
                make_float4(row == c0 + k ? hv : 0.f, row == c0 + k + 1 ? hv : 0.f,
                            row == c0 + k + 2 ? hv : 0.f, row == c0 + k + 3 ? hv : 0.f);
        }
        if (tid == 0) {
          c.need_init = 0;
          c.iter = 0;
          c.cur = 0;
          c.err = __uint_as_float(emax);  // DS:872
          c.ratio = 1.0f;
          c.hmul = 1.0f;
          root_after_error_update(c, prm);
          S.ctl = c;
        }
        __syncthreads();
        PC_MARK(1);
        if (S.ctl.done || S.ctl.need_init) continue;
      }
      // ---- one try: coupled Newton iterations (DS:836-848) ----
      while (S.ctl.active) {
        const int cur = S.ctl.cur;
        float x[64];
        // H' = H M_i  (DS:846): A = H from tensor memory, B = M_i
        issue_product(true, bufX, bufX);
        pull(x);
        PC_MARK(2);
        {
          float* out = slots[cur ^ 1];
#pragma unroll
          for (int k = 0; k < 64; k += 4)
            *reinterpret_cast<float4*>(out + ((size_t)(k >> 2) * kThreads + tid) * 4) =
                make_float4(x[k], x[k + 1], x[k + 2], x[k + 3]);
          {
            uint32_t w0[32], w1[32], w2[32];
#pragma unroll
            for (int k = 0; k < 32; ++k) split2(x[2 * k], x[2 * k + 1], w0[k], w1[k], w2[k]);
            tmem_st_32x32(tmem + lane_off + kHCols + 0 * 64 + half * 32, w0);
            tmem_st_32x32(tmem + lane_off + kHCols + 1 * 64 + half * 32, w1);
            tmem_st_32x32(tmem + lane_off + kHCols + 2 * 64 + half * 32, w2);
          }
          tmem_st_wait();
        }
        PC_MARK(3);
        // M_i^p by in-place squarings (p = 2^s), DS:655-678 without its no-op products
        for (int sq = p; sq > 1; sq >>= 1) {
          issue_product(false, bufX, bufX);
          pull(x);
          PC_MARK(4);
          store_planes(bufX, row, c0, x);
          PC_MARK(5);
        }
        // M' = M_i^p M (DS:845), M_i' (DS:844), err = max|M' - I_m| (DS:847).  Q M is not bitwise
        // symmetric; the lower triangle is authoritative.  The transposed values come from the
        // product in the other order: (M Q)(i, j) = sum_k M(i,k) Q(j,k) == (Q M)(j, i) bit for
        // bit, so every thread gets the mirror of its row in registers -- no data exchange.
        issue_product(false, bufX, bufM);
        pull(x);
        PC_MARK(6);
        issue_product(false, bufM, bufX);
        {
          float y[64];
          pull(y);
#pragma unroll
          for (int k = 0; k < 64; ++k)
            if (c0 + k > row) x[k] = y[k];
        }
        uint32_t emax = 0;
        {
          const float alpha = -1.0f / (float)p, oma = 1.0f - alpha;
          float mi[64];
#pragma unroll
          for (int k = 0; k < 64; ++k) {
            const bool dg = (row == c0 + k) && row < pad;
            const uint32_t ab = absbits(x[k] - (dg ? 1.f : 0.f));
            emax = ab > emax ? ab : emax;
            mi[k] = mi_from_m(x[k], dg, alpha, oma);
          }
          store_planes(bufM, row, c0, x);
          store_planes(bufX, row, c0, mi);
        }
        PC_MARK(7);
        emax = block_max_256(emax, S.ured);
        if (tid == 0) {
          RootCtl c = S.ctl;
          const float new_err = __uint_as_float(emax);
          c.ratio = new_err / c.err;  // DS:848
          c.err = new_err;
          c.iter += 1;
          c.cur ^= 1;
          root_after_error_update(c, prm);
          S.ctl = c;
        }
        __syncthreads();
        PC_MARK(8);
      }
    }

    // ---- result: H or the previous H (DS:878-880), symmetrised from its lower triangle ----
    __syncthreads();
    const RootCtl c = S.ctl;
    if (c.result_h != -1) {
      const bool zero = c.result_h == -2 || c.pad == 0;  // DS:930-937
      const float* src = slots[c.result_h >= 0 ? c.result_h : 0];
      float* dst = roots + (size_t)b * n * n;
      if (row < n) {
#pragma unroll 4
        for (int k = 0; k < 64; ++k) {
          const int colk = c0 + k;
          if (colk <= row) {  // lower triangle authoritative, mirrored
            float v = 0.f;
            if (!zero && row < c.pad) v = src[((size_t)(k >> 2) * kThreads + tid) * 4 + (k & 3)];
            dst[(size_t)row * n + colk] = v;
            if (colk != row) dst[(size_t)colk * n + row] = v;
          }
        }
      }
    }
    if (tid == 0) {
      float* m = metrics + (size_t)b * PC_NUM_METRICS;
      m[PC_METRIC_ERROR] = c.pad == 0 ? 0.f : c.m_err;
      m[PC_METRIC_ITERS] = c.m_iters;
      m[PC_METRIC_ERROR_RATIO] = c.m_ratio;
      m[PC_METRIC_MAX_EV] = c.max_ev;
      m[PC_METRIC_RETRIES] = c.m_retries;
    }
    __syncthreads();
    PC_MARK(9);
  }
  if (prof && tid == 0)
    for (int i = 0; i < 12; ++i) prof[(size_t)blockIdx.x * 12 + i] = S.prof[i];
#undef PC_MARK
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ---------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------
// start vector + two fp32 H slots per resident CTA (not per matrix: a CTA reuses its slots)
size_t small_root_workspace_bytes(int batch, int n) {
  (void)n;
  const int ctas = batch < 256 ? batch : 256;
  return (size_t)ctas * 2 * kN * kN * sizeof(float) + 1024 + 512;
}

bool small_root_supported_exponents(const int32_t* ps_host, int batch) {
  if (!ps_host) return false;
  for (int b = 0; b < batch; ++b) {
    const int p = ps_host[b];
    if (p < 1 || p > kMaxP || (p & (p - 1)) != 0) return false;
  }
  return true;
}

int run_small_root(const float* xs, const int32_t* ps, const int32_t* pads, int batch, int n,
                   const pc_root_options* opt, const float* v0_device, float* roots,
                   float* metrics, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PC_REQUIRE(n >= 1 && n <= kN, "persistent small-block solver needs n <= 128 (n=%d)", n);
  PC_REQUIRE(workspace_bytes >= small_root_workspace_bytes(batch, n), "workspace too small");
  char* w = reinterpret_cast<char*>(align_up((size_t)workspace, 256));
  const float* v0 = v0_device;  // resident start vector (prepare_power_iteration)
  float* hslot = reinterpret_cast<float*>(w + 512);
  constexpr size_t smem = 2 * (size_t)kMatBytes + sizeof(SmallShared) + 1024;
  static bool configured[64] = {false};
  int dev = 0, sms = 148;
  PC_CUDA_CHECK(cudaGetDevice(&dev));
  if (!configured[dev & 63]) {
    PC_CUDA_CHECK(cudaFuncSetAttribute(small_root_kernel<2>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PC_CUDA_CHECK(cudaFuncSetAttribute(small_root_kernel<3>,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[dev & 63] = true;
  }
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  RootParams prm{opt->ridge_epsilon, opt->error_tolerance, opt->num_iters,
                 opt->relative_matrix_epsilon};
  const int grid = batch < sms ? batch : (sms < 256 ? sms : 256);
  // PC_SMALL_PROF=1: per-CTA phase cycle counts, printed after the launch (diagnostics; syncs)
  long long* prof = nullptr;
  const char* pf = getenv("PC_SMALL_PROF");
  if (pf && pf[0] == '1') {
    PC_CUDA_CHECK(cudaMalloc(&prof, sizeof(long long) * 12 * grid));
    PC_CUDA_CHECK(cudaMemsetAsync(prof, 0, sizeof(long long) * 12 * grid, stream));
  }
  const char* terms = getenv("PC_SMALL_TERMS");
  if (terms && terms[0] == '8')
    small_root_kernel<3><<<grid, kThreads, smem, stream>>>(xs, ps, pads, batch, n, prm, v0, roots,
                                                          hslot, metrics, prof);
  else
    small_root_kernel<2><<<grid, kThreads, smem, stream>>>(xs, ps, pads, batch, n, prm, v0, roots,
                                                          hslot, metrics, prof);
  if (prof) {
    std::vector<long long> h(12 * (size_t)grid);
    cudaMemcpyAsync(h.data(), prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost, stream);
    cudaStreamSynchronize(stream);
    long long worst = 0;
    int wi = 0;
    for (int c = 0; c < grid; ++c) {
      long long t = 0;
      for (int i = 0; i < 12; ++i) t += h[c * 12 + i];
      if (t > worst) { worst = t; wi = c; }
    }
    static const char* names[10] = {"load+PI", "init", "H mma+pull", "H epilogue", "sq mma+pull",
                                    "sq epilogue", "M' mma+pull", "M' epilogue", "control", "final"};
    fprintf(stderr, "[small_root] slowest CTA %d: %lld cycles;", wi, worst);
    for (int i = 0; i < 10; ++i) fprintf(stderr, " %s %lld", names[i], h[wi * 12 + i]);
    fprintf(stderr, "\n");
    cudaFree(prof);
  }
  count_launch(1);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

}  // namespace pc
