// tcgen05 / TMEM / TMA engine for the Newton-chain GEMMs of the inverse p-th root
// solver (sm_100a).
//
// Number format: every fp32 matrix X is stored as three bf16 planes X0+X1+X2 == X
// (exact: 8+8+8 mantissa bits).  A product A*B is evaluated on the tensor cores
// as  sum_{i+j<=2} A_i B_j  (6 bf16 MMAs, "BF16x6", dropped terms <= 2^-23 |a||b|;
// PC_ENGINE_TC_BF16X3 keeps i+j<=1).  The tensor core accumulates in fp32 with
// truncating adds; over K=1024 that costs ~10x in accuracy on ill-conditioned
// statistics (profiles/r01_precision_probe.md).  Therefore only a short K-chunk
// (kChunkKB * 64 columns) is accumulated in TMEM; the epilogue warps pull each
// chunk with tcgen05.ld and add it to an fp32 register accumulator with
// round-to-nearest, overlapped with the MMAs of the next chunk (2 TMEM stages).
//
// Kernel anatomy (one persistent CTA per SM, 256 threads):
//   warp 0   TMA producer: 2 * kLP plane tiles (128 x 64 bf16, SWIZZLE_128B) per
//            k-block into a kStages-deep smem ring, mbarrier complete_tx
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma.cta_group::1
//            .kind::f16 (M=128, N=128, K=16), tcgen05.commit frees smem slots and
//            publishes TMEM chunks
//   warp 2   TMEM allocator (256 columns = 2 chunk accumulators)
//   warps 4-7  chunk accumulation + epilogue: split to 3 bf16 planes, optional
//            M_i' emission and max|M' - I_m| reduction (DS:844-847)
// Operands are symmetric, so both A and B tiles are K-major row blocks of the
// same row-major planes (no transposes anywhere).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "root_kernels.cuh"
#include "tc_engine.cuh"

namespace pc {

constexpr int TC_BM = 128, TC_BN = 128, TC_BK = 64, TC_UMMA_K = 16;
constexpr int TC_TILE_BYTES = TC_BM * TC_BK * 2;  // one plane tile: 16 KiB

// Plane formats.
//   TC_FMT_BF16   X = X0 + X1 + X2, bf16 planes (8+8+8 mantissa bits, exact), 6 (or 3) MMAs
//   TC_FMT_FP16S  X ~ X0 + 2^-11 X1, fp16 planes (11+11 mantissa bits, |error| <= 2^-23 |X|):
//                 the residual plane is stored scaled by 2^11 so that it lives in fp16's
//                 normal range.  A*B = A0 B0 + 2^-11 (A0 B1 + A1 B0) takes THREE MMAs: the
//                 cross terms are accumulated first and the leading term is issued with the
//                 instruction's scale-input-d = 11 (D = A0 B0 + D * 2^-11).  Same 11-bit
//                 operand mantissa as 3xTF32, at the f16 MMA rate (2x TF32) and 4 B / element.
constexpr int TC_FMT_BF16 = 0, TC_FMT_FP16S = 1;
constexpr float TC_FP16_SCALE = 2048.0f;

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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
// true on exactly one lane of a converged warp; unlike `lane == 0` the compiler knows the
// branch holds a single thread, so TMA / MMA operands go to uniform registers directly
// instead of through a per-lane readback loop (113 instructions per TMA load before)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}"
      : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// L2 eviction-priority policies (same encodings as cute::TMA::CacheHintSm90)
constexpr uint64_t kPolicyEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kPolicyEvictLast = 0x14F0000000000000ull;

// EVICT_NORMAL by default; PC_TC_HINTS=1 switches operand loads to evict_last and result
// stores to evict_first (measured neutral-to-worse on B200 at batch 32-64, kept as a knob)
constexpr uint64_t kPolicyEvictNormal = 0x1000000000000000ull;
__constant__ uint64_t g_load_policy = kPolicyEvictNormal;
__constant__ uint64_t g_store_policy = kPolicyEvictNormal;

// loads storage tile (tile_row, tile_col) of matrix `mat` (= buffer * batch + b): 16 KiB
__device__ __forceinline__ void tma_load_tile(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                              int tile_col, int tile_row, int mat) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2], %8;" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(0), "r"(0), "r"(tile_col),
      "r"(tile_row), "r"(mat), "l"(g_load_policy)
      : "memory");
}
// results are consumed by a later launch: do not let them displace operands
__device__ __forceinline__ void st_global_v4_stream(void* p, uint4 v) {
  asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "r"(v.x),
               "r"(v.y), "r"(v.z), "r"(v.w), "l"(g_store_policy)
               : "memory");
}
__device__ __forceinline__ void st_global_u16_stream(void* p, uint16_t v) {
  asm volatile("st.global.L2::cache_hint.u16 [%0], %1, %2;" ::"l"(p), "h"(v),
               "l"(g_store_policy)
               : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
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
// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> fp32
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
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
// D[tmem] = A * B + D * 2^-11 (scale-input-d; needs the explicit disable-output-lane vector)
__device__ __forceinline__ void umma_f16_scaled_d(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                                  uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p, 11;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// arrives on the mbarrier once all previously issued tcgen05.mma have completed
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

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1"):
//   bits [0,14) start address >> 4, [16,30) LBO >> 4 (unused for swizzled K-major),
//   [32,46) SBO >> 4 = 1024 B between 8-row groups, [46,48) version = 1,
//   [61,64) layout type = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor: c=F32 (1<<4), a=b=BF16 (1<<7, 1<<10), K-major both,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdescBf16M128N128 =
    (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TC_BN >> 3) << 17) |
    ((uint32_t)(TC_BM >> 4) << 24);
// a = b = F16 (format 0)
constexpr uint32_t kIdescF16M128N128 =
    (1u << 4) | ((uint32_t)(TC_BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

// ---------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------
struct TcParams {
  uint16_t* plane[3];   // plane p: [kNumBufs][batch][n][n] bf16
  size_t buf_stride;    // batch * n * n
  size_t mat_stride;    // n * n
  const RootCtl* ctl;
  uint32_t* errbits;
  int n, batch, tiles;  // tiles per dimension (n / 128)
  // Group scheduling: the `sync_group` consecutive work items that read the same operand
  // matrices start their tiles together (a counter per group, see group_sync), so every
  // operand block is pulled from HBM once and shared through L2.  0 = no synchronisation.
  uint32_t* sync_ctr;       // this launch's counters [2 * batch] (zero on entry)
  uint32_t* sync_ctr_next;  // the next launch's counters: zeroed by this launch
  int sync_group;           // items per group: per_mat, or per (matrix, op)
  int sync_arrivals;        // arrivals that release a group (= sync_group x CTAs per item)
  // PC_TC_ABLATE (diagnostics only, results are garbage): bit 0 = the epilogue computes but
  // neither stages nor stores; bit 1 = the accumulate warps skip the TMEM pull and the adds.
  int ablate;
};

struct TcWork {
  int b, tm, tn, kblocks, cur, p, pad, op;
  Step st;
};

// Work items of launch `s`: step 0 runs two products per matrix (chain step 0 and the
// H update), later steps one.
__device__ __forceinline__ int tc_ops_per_matrix(int s) { return s == 0 ? 2 : 1; }

// Start-of-tile rendezvous of the producers that work on one group (relaxed: it orders
// nothing, it only keeps the CTAs that share operands within a tile of each other so the
// shared blocks are still in L2).  Bounded spin: a missing partner costs time, never a hang.
__device__ __forceinline__ void group_sync(uint32_t* ctr, uint32_t expected) {
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
  uint32_t v;
  int spins = 0;
  do {
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
  } while (v < expected && ++spins < (1 << 14));
}
__device__ __forceinline__ void tc_group_sync(const TcParams& P, const TcWork& wk, int ntri, int s) {
  if (P.sync_group <= 1) return;
  const int g = P.sync_group == ntri * tc_ops_per_matrix(s) ? wk.b : wk.b * 2 + wk.op;
  group_sync(P.sync_ctr + g, (uint32_t)P.sync_arrivals);
}

__device__ __forceinline__ bool tc_get_work(const TcParams& P, const Program* progs, int s,
                                            int w, TcWork& out) {
  const int ntri = P.tiles * (P.tiles + 1) / 2;  // lower-triangular tiles only
  const int per_mat = tc_ops_per_matrix(s) * ntri;
  const int b = w / per_mat;
  int r = w - b * per_mat;
  const int op = r / ntri;
  r -= op * ntri;
  const RootCtl& c = P.ctl[b];
  if (!c.active) return false;
  out.op = op;
  if (op == 1) {
    out.st = Step{LB_HN, LB_H, LB_MI, 0};
  } else {
    const Program& pr = progs[c.p];
    if (s >= pr.nsteps) return false;
    out.st = pr.steps[s];
  }
  out.b = b;
  tri_decode(r, out.tm, out.tn);  // tm >= tn; the epilogue mirrors the tile
  out.cur = c.cur;
  out.p = c.p;
  out.pad = c.pad;
  out.kblocks = (c.pad + TC_BK - 1) / TC_BK;  // columns >= pad are zero
  return true;
}

__device__ __forceinline__ uint32_t pack_bf16x2(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}

// Per-warp staging buffer for coalesced plane stores: 32 rows x 32 bf16 (2 KiB), 16-byte
// chunks XOR-swizzled so that both the row writes and the 8-rows-x-64-B reads are
// bank-conflict free.
constexpr int TC_STAGE_BYTES_PER_WARP = 2048;
__device__ __forceinline__ uint32_t stage_addr(uint32_t base, int r, int ch) {
  return base + r * 64 + ((ch ^ ((r >> 1) & 3)) << 4);
}

// ---------------------------------------------------------------------------
// CTA-pair kernel (cta_group::2): one cluster of two CTAs owns a 256 x 128 output
// tile.  CTA r holds A rows [r*128, r*128+128) and half of B's rows
// [r*64, r*64+64); the leader's single thread issues tcgen05.mma.cta_group::2
// (M = 256, N = 128) that reads both CTAs' shared memory.  Per SM this loads
// 72 KB per k-block instead of 96 KB and leaves room for a third stage.
// ---------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-pair bit of a smem address

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t"
      ".reg .b32 remAddr32;\n\t"
      "mapa.shared::cluster.u32 remAddr32, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [remAddr32];\n\t"
      "}" ::"r"(bar),
      "r"(cta)
      : "memory");
}
// both CTAs issue; the transaction bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_tile_2sm(uint32_t dst, const CUtensorMap* map,
                                                  uint32_t bar, int tile_col, int tile_row,
                                                  int mat) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerBitMask), "r"(0), "r"(0),
      "r"(tile_col), "r"(tile_row), "r"(mat)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_scaled_d_2sm(uint32_t tmem_d, uint64_t adesc,
                                                      uint64_t bdesc, uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, "
      "p, 11;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u),
      "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// arrive (once the MMAs retire) on the same-offset barrier of BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(bar),
      "h"((uint16_t)3)
      : "memory");
}

constexpr uint32_t kIdescBf16M256N256 =
    (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
constexpr uint32_t kIdescF16M256N256 =
    (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);


// ---------------------------------------------------------------------------
// MMA issue helpers shared by the 1-CTA and the CTA-pair kernels.  One "k-block" is a
// 64-column slab: kLP plane tiles of A at a0 and of B at b0 (K-major, SWIZZLE_128B).
// ---------------------------------------------------------------------------
template <bool k2sm>
__device__ __forceinline__ void umma_any(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc,
                                         uint32_t acc) {
  if (k2sm) umma_bf16_2sm(d, a, b, idesc, acc);
  else umma_bf16(d, a, b, idesc, acc);
}
// scaled fp16 planes, cross terms of one k-block: D (+)= A0 B1 + A1 B0   (both x 2^11)
template <bool k2sm>
__device__ __forceinline__ void issue_fp16_cross(uint32_t d, uint32_t a0, uint32_t b0,
                                                 uint32_t idesc, bool first,
                                                 uint32_t b_plane = TC_TILE_BYTES) {
  const uint64_t a0d = make_kmajor_sw128_desc(a0), a1d = make_kmajor_sw128_desc(a0 + TC_TILE_BYTES);
  const uint64_t b0d = make_kmajor_sw128_desc(b0), b1d = make_kmajor_sw128_desc(b0 + b_plane);
#pragma unroll
  for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
    umma_any<k2sm>(d, a0d + 2u * k, b1d + 2u * k, idesc, (first && k == 0) ? 0u : 1u);
#pragma unroll
  for (int k = 0; k < TC_BK / TC_UMMA_K; ++k)
    umma_any<k2sm>(d, a1d + 2u * k, b0d + 2u * k, idesc, 1u);
}
// leading term of one k-block: D += A0 B0; with `rescale` the first MMA applies
// scale-input-d = 11, i.e. D = A0 B0 + D * 2^-11 (all cross terms of the chunk are in D)
template <bool k2sm>
__device__ __forceinline__ void issue_fp16_main(uint32_t d, uint32_t a0, uint32_t b0,
                                                uint32_t idesc, bool rescale) {
  const uint64_t a0d = make_kmajor_sw128_desc(a0), b0d = make_kmajor_sw128_desc(b0);
  if (rescale) {
    if (k2sm) umma_f16_scaled_d_2sm(d, a0d, b0d, idesc);
    else umma_f16_scaled_d(d, a0d, b0d, idesc);
  } else {
    umma_any<k2sm>(d, a0d, b0d, idesc, 1u);
  }
#pragma unroll
  for (int k = 1; k < TC_BK / TC_UMMA_K; ++k)
    umma_any<k2sm>(d, a0d + 2u * k, b0d + 2u * k, idesc, 1u);
}
// bf16 planes, all products i + j <= kLP - 1 of one k-block, smallest terms first
template <int kLP, bool k2sm>
__device__ __forceinline__ void issue_bf16_block(uint32_t d, uint32_t a0, uint32_t b0,
                                                 uint32_t idesc, bool first_in) {
  bool first = first_in;
#pragma unroll
  for (int sum = kLP - 1; sum >= 0; --sum) {
#pragma unroll
    for (int i = 0; i <= sum; ++i) {
      const int j = sum - i;
      const uint64_t ad = make_kmajor_sw128_desc(a0 + i * TC_TILE_BYTES);
      const uint64_t bd = make_kmajor_sw128_desc(b0 + j * TC_TILE_BYTES);
#pragma unroll
      for (int k = 0; k < TC_BK / TC_UMMA_K; ++k) {
        umma_any<k2sm>(d, ad + 2u * k, bd + 2u * k, idesc, (first && k == 0) ? 0u : 1u);
        if (k == 0) first = false;
      }
    }
  }
}

// ---------------------------------------------------------------------------
// Warp-specialised 1-CTA kernel with dedicated epilogue warpgroups (512 threads):
//   warpgroup 0  warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator
//   warpgroup 1  chunk accumulation: pulls every K-chunk from TMEM, adds it to fp32
//                registers (round-to-nearest) and parks the finished 128 x 128 tile
//                in one of two TMEM "output" stages with tcgen05.st
//   warpgroup 2  epilogue: tcgen05.ld from the output stage, 3-plane split, staged
//                coalesced stores + mirror, M_i' / err
// so tile i's global-memory epilogue overlaps tile i+1's MMAs.  TMEM: 2 chunk
// accumulators + 2 output stages = 512 columns.  Registers are rebalanced with
// setmaxnreg (40 / 200 / 136 / 136); the two epilogue warpgroups split each tile's columns.
// ---------------------------------------------------------------------------
constexpr int TC_WS_THREADS = 512;

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


// ---- TMA store epilogue helpers (warp-specialised kernel) ----
// stores a 32 x 32 block whose top-left element is (row, col) of matrix `mat`
__device__ __forceinline__ void tma_store_block(const CUtensorMap* map, uint32_t src, int col,
                                                int row, int mat) {
  asm volatile(
      "cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::
          "l"(reinterpret_cast<uint64_t>(map)),
      "r"(src), "r"(col & 63), "r"(row & 127), "r"(col >> 6), "r"(row >> 7), "r"(mat)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read0() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// Per-thread staging addresses, computed once: row `lane` of buffer D (4 x 16-byte chunks)
// and column `lane` of buffer T (rows 2k / 2k+1 are 128 / 64 bytes apart, the XOR swizzle
// term only depends on k & 3).
struct EpiAddr {
  uint32_t d[4];  // D(lane, ch)
  uint32_t t[4];  // T(0, lane) with the swizzle term of k & 3 == j
};
__device__ __forceinline__ EpiAddr make_epi_addr(uint32_t bufD, uint32_t bufT, int lane) {
  EpiAddr e;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) e.d[ch] = stage_addr(bufD, lane, ch);
#pragma unroll
  for (int j = 0; j < 4; ++j)
    e.t[j] = bufT + ((((uint32_t)lane >> 3) ^ (uint32_t)j) << 4) + ((uint32_t)lane & 7) * 2;
  return e;
}

// Diagonal 32 x 32 sub-block already staged row-major in `buf`: overwrite the strictly
// upper triangle with the transposed lower one, in place (lower triangle authoritative).
// Out of line and loop-rolled on purpose: it is rare and must not bloat the hot epilogue.
__device__ __noinline__ void symmetrise_diag_block(uint32_t buf, int lane) {
  __syncwarp();
  uint16_t col[32];
#pragma unroll 1
  for (int i = 0; i < 32; ++i) {  // lane reads element (i, lane): one contiguous row per step
    const uint32_t a = stage_addr(buf, i, lane >> 3) + (lane & 7) * 2;
    uint16_t x;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(x) : "r"(a) : "memory");
    col[i] = x;
  }
  __syncwarp();
#pragma unroll 1
  for (int i = 0; i < 32; ++i) {  // element (lane, i), i > lane, <- value of (i, lane)
    if (i > lane) {
      const uint32_t a = stage_addr(buf, lane, i >> 3) + (i & 7) * 2;
      asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"(col[i]) : "memory");
    }
  }
  __syncwarp();
}

// One bf16 plane of a 32 x 32 sub-block, one row per lane (hp[k] = columns 2k, 2k+1):
// the row-major block goes to staging buffer D, its transpose to buffer T (both in the
// SWIZZLE_64B layout of the store tensor map), then one lane issues two bulk tensor
// stores: D -> (row0, col0) and T -> (col0, row0).  No LDS / STG on the warp's critical
// path; the buffers are recycled after cp.async.bulk.wait_group.read.
__device__ __forceinline__ void store_plane_block_tma(uint32_t bufD, uint32_t bufT,
                                                      const EpiAddr& ea, int lane,
                                                      uint32_t (&hp)[16], bool diag_sub,
                                                      bool do_mirror, const CUtensorMap* map,
                                                      int row0, int col0, int mat, int ablate) {
  if (ablate & 1) return;
  if (lane == 0) tma_store_wait_read0();  // previous stores have consumed the buffers
  __syncwarp();
#pragma unroll
  for (int ch = 0; ch < 4; ++ch)
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ea.d[ch]), "r"(hp[4 * ch]),
                 "r"(hp[4 * ch + 1]), "r"(hp[4 * ch + 2]), "r"(hp[4 * ch + 3])
                 : "memory");
  if (diag_sub) {
    symmetrise_diag_block(bufD, lane);  // rare: 1 of 4 sub-blocks of the 4-8 diagonal tiles
  } else if (do_mirror) {
    // transpose: element (lane, i) -> T(i, lane); per row i the lanes write 64 contiguous bytes
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const uint32_t a0 = ea.t[k & 3] + k * 128;  // row 2k
      asm volatile(
          "{\n\t"
          ".reg .b16 lo, hi;\n\t"
          "mov.b32 {lo, hi}, %1;\n\t"
          "st.shared.b16 [%0], lo;\n\t"
          "st.shared.b16 [%0 + 64], hi;\n\t"
          "}" ::"r"(a0),
          "r"(hp[k])
          : "memory");
    }
  }
  fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    tma_store_block(map, bufD, col0, row0, mat);
    if (do_mirror && !diag_sub) tma_store_block(map, bufT, row0, col0, mat);
    tma_store_commit();
  }
}

// fp32 row segment (consumed) -> planes (3 x bf16, or fp16 + 2^11-scaled fp16 residual), each
// staged and bulk-stored.
template <int kFmt>
__device__ __forceinline__ void store_block_3planes_tma(const TcParams& P, uint32_t stage,
                                                        const EpiAddr& ea, int lane, float (&x)[32],
                                                        bool diag_sub, bool do_mirror,
                                                        const CUtensorMap* const (&smaps)[3],
                                                        int row0, int col0, int mat) {
  constexpr int kPlanes = kFmt == TC_FMT_FP16S ? 2 : 3;
#pragma unroll 1
  for (int pl = 0; pl < kPlanes; ++pl) {  // rolled: keeps the kernel inside the instruction cache
    uint32_t hp[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (kFmt == TC_FMT_FP16S) {
        uint32_t w;  // one F2FP; saturates instead of overflowing to inf, NaN stays NaN
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(x[2 * k + 1]), "f"(x[2 * k]));
        hp[k] = w;
        if (pl == 0) {  // exact residual, scaled into fp16's normal range
          const __half2 h2 = *reinterpret_cast<const __half2*>(&w);
          const float2 f2 = __half22float2(h2);
          x[2 * k] = (x[2 * k] - f2.x) * TC_FP16_SCALE;
          x[2 * k + 1] = (x[2 * k + 1] - f2.y) * TC_FP16_SCALE;
        }
      } else {
        const __nv_bfloat162 pr = __floats2bfloat162_rn(x[2 * k], x[2 * k + 1]);  // one F2FP
        const uint32_t w = *reinterpret_cast<const uint32_t*>(&pr);
        hp[k] = w;
        if (pl < 2) {  // exact residuals feed the next plane
          x[2 * k] -= __uint_as_float(w << 16);
          x[2 * k + 1] -= __uint_as_float(w & 0xffff0000u);
        }
      }
    }
    store_plane_block_tma(stage, stage + TC_STAGE_BYTES_PER_WARP, ea, lane, hp, diag_sub,
                          do_mirror, smaps[pl], row0, col0, mat, P.ablate);
  }
}

// Same as tc_epilogue_tile but the tile is read from a TMEM output stage.
template <int kFmt>
__device__ __forceinline__ void tc_epilogue_tile_tmem(const TcParams& P, const TcWork& wk, int tm,
                                                      int row_in_tile, int lane, uint32_t stage,
                                                      uint32_t taddr, uint32_t oempty_bar,
                                                      int c_begin, int c_end,
                                                      const CUtensorMap* const (&smaps)[3]) {
  const int q = row_in_tile >> 5;
  const int row = tm * TC_BM + row_in_tile;
  const int row0 = tm * TC_BM + q * 32;
  const bool diag_tile = tm == wk.tn;
  const size_t mat_off = (size_t)wk.b * P.mat_stride;
  const size_t out_base = (size_t)physical_buf(wk.st.dst, wk.cur) * P.buf_stride + mat_off;
  const size_t mi_base = (size_t)physical_buf(LB_MIN, wk.cur) * P.buf_stride + mat_off;
  const float alpha = -1.0f / (float)wk.p, oma = 1.0f - alpha;
  const bool mirror_on = true;
  uint32_t emax = 0;
  const EpiAddr ea = make_epi_addr(stage, stage + TC_STAGE_BYTES_PER_WARP, lane);
  auto write_group = [&](int buf, float (&x)[32], int col0, bool diag_sub) {
    store_block_3planes_tma<kFmt>(P, stage, ea, lane, x, diag_sub, mirror_on, smaps, row0, col0,
                            buf * P.batch + wk.b);
  };
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {  // rolled: g is addressed in TMEM, not in registers
    const bool skip = diag_tile && c > q;  // strictly upper sub-block: its mirror writes it
    const bool diag_sub = diag_tile && c == q;
    const int col0 = wk.tn * TC_BN + c * 32;
    float g[32];
    if (!skip) {
      uint32_t r[32];
      tmem_ld_32x32(taddr + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) g[i] = __uint_as_float(r[i]);
    }
    if (c == c_end - 1) {  // last TMEM read of this tile by this warp: recycle the stage
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(oempty_bar);
    }
    if (skip) continue;
#pragma unroll 1
    for (int o = wk.st.emit_mi ? 0 : 1; o < 2; ++o) {  // o = 0: M_i' (DS:844), o = 1: OUT
      float x[32];
      if (o == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const bool dg = (col0 + i == row) && (row < wk.pad);
          if (!diag_sub || col0 + i <= row) {
            const uint32_t ab = absbits(g[i] - (dg ? 1.f : 0.f));  // DS:847
            emax = ab > emax ? ab : emax;
          }
          x[i] = mi_from_m(g[i], dg, alpha, oma);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = g[i];
      }
      write_group(physical_buf(o == 0 ? (int)LB_MIN : (int)wk.st.dst, wk.cur), x, col0, diag_sub);
    }
  }
  if (wk.st.emit_mi) {
    emax = warp_max_u32(emax);
    if (lane == 0 && emax) atomicMax(P.errbits + wk.b, emax);
  }
}

template <int kLP, int kStages, int kFmt, int kChunkKB>
__global__ void __launch_bounds__(TC_WS_THREADS, 1)
tc_phase_kernel_ws(const __grid_constant__ CUtensorMap tmap0,
                   const __grid_constant__ CUtensorMap tmap1,
                   const __grid_constant__ CUtensorMap tmap2,
                   const __grid_constant__ CUtensorMap smap0,
                   const __grid_constant__ CUtensorMap smap1,
                   const __grid_constant__ CUtensorMap smap2, const TcParams P,
                   const Program* __restrict__ progs, int s, int total_work) {
  constexpr int kStageBytes = 2 * kLP * TC_TILE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int i) { return bar_base + 8u * i; };
  auto empty_bar = [&](int i) { return bar_base + 8u * (kStages + i); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * kStages + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * kStages + 2 + i); };
  auto ofull_bar = [&](int i) { return bar_base + 8u * (2 * kStages + 4 + i); };
  auto oempty_bar = [&](int i) { return bar_base + 8u * (2 * kStages + 6 + i); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 8);
  const uint32_t stage_base = bar_base + 1024;  // 8 epilogue warps x (D + T) x 2 KiB, 1 KiB aligned
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap0);
    tma_prefetch_desc(&tmap1);
    if (kLP > 2) tma_prefetch_desc(&tmap2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 4);
      mbar_init(ofull_bar(i), 4);
      mbar_init(oempty_bar(i), 8);  // 2 epilogue warpgroups x 4 warps
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  if (warp == 3 && blockIdx.x == 0 && P.sync_ctr_next)
    for (int i = lane; i < 2 * P.batch; i += 32) P.sync_ctr_next[i] = 0u;
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && elect_one()) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
#pragma unroll 1
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        TcWork wk;
        if (!tc_get_work(P, progs, s, w, wk)) continue;
        tc_group_sync(P, wk, P.tiles * (P.tiles + 1) / 2, s);
        const int pa = physical_buf(wk.st.a, wk.cur), pb = physical_buf(wk.st.b, wk.cur);
#pragma unroll 1
        for (int kb = 0; kb < wk.kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t dst = smem_base + stage * kStageBytes;
          mbar_expect_tx(full_bar(stage), kStageBytes);
          const CUtensorMap* maps[3] = {&tmap0, &tmap1, &tmap2};
#pragma unroll
          for (int pl = 0; pl < kLP; ++pl) {
            tma_load_tile(dst + pl * TC_TILE_BYTES, maps[pl], full_bar(stage), kb, wk.tm,
                          pa * P.batch + wk.b);
            tma_load_tile(dst + (kLP + pl) * TC_TILE_BYTES, maps[pl], full_bar(stage), kb, wk.tn,
                          pb * P.batch + wk.b);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && elect_one()) {
      // ===================== MMA issuer =====================
      int stage = 0;
      uint32_t phase = 0;
      int chunk = 0;
#pragma unroll 1
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        TcWork wk;
        if (!tc_get_work(P, progs, s, w, wk)) continue;
#pragma unroll 1
        for (int kb = 0; kb < wk.kblocks; kb += kChunkKB, ++chunk) {
          const int nkb = min(kChunkKB, wk.kblocks - kb);
          const int acc = chunk & 1;
          mbar_wait(tempty_bar(acc), ((chunk >> 1) & 1) ^ 1);
          const uint32_t tmem_d = tmem_base + acc * TC_BN;
          constexpr uint32_t idesc = kFmt == TC_FMT_FP16S ? kIdescF16M128N128 : kIdescBf16M128N128;
          int st[kChunkKB];
          // phase 1 (per k-block, as its operands land): cross terms / all bf16 products
#pragma unroll
          for (int j = 0; j < kChunkKB; ++j) {
            if (j < nkb) {
              st[j] = stage;
              mbar_wait(full_bar(stage), phase);
              tcgen05_fence_after();
              const uint32_t a0 = smem_base + stage * kStageBytes;
              const uint32_t b0 = a0 + kLP * TC_TILE_BYTES;
              if (kFmt == TC_FMT_FP16S) {
                issue_fp16_cross<false>(tmem_d, a0, b0, idesc, j == 0);
              } else {
                issue_bf16_block<kLP, false>(tmem_d, a0, b0, idesc, j == 0);
                umma_commit(empty_bar(stage));
              }
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
          // phase 2 (scaled fp16 only): D = A0 B0 + D * 2^-11, then the other leading terms
          if (kFmt == TC_FMT_FP16S) {
#pragma unroll
            for (int j = 0; j < kChunkKB; ++j) {
              if (j < nkb) {
                const uint32_t a0 = smem_base + st[j] * kStageBytes;
                issue_fp16_main<false>(tmem_d, a0, a0 + kLP * TC_TILE_BYTES, idesc, j == 0);
                umma_commit(empty_bar(st[j]));
              }
            }
          }
          umma_commit(tfull_bar(acc));
        }
      }
    }
  } else if (warp < 8) {
    // ============ warpgroup 1: chunk accumulation -> TMEM output stage ============
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int chunk = 0, tile = 0;
#pragma unroll 1
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      TcWork wk;
      if (!tc_get_work(P, progs, s, w, wk)) continue;
      float sum[TC_BN];
#pragma unroll
      for (int i = 0; i < TC_BN; ++i) sum[i] = 0.f;
#pragma unroll 1
      for (int kb = 0; kb < wk.kblocks; kb += kChunkKB, ++chunk) {
        const int acc = chunk & 1;
        mbar_wait(tfull_bar(acc), (chunk >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + lane_off + acc * TC_BN;
        if (!(P.ablate & 2)) {
#pragma unroll
          for (int c = 0; c < TC_BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) sum[c * 32 + i] += __uint_as_float(r[i]);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      // park the finished tile in output stage o for the epilogue warpgroup
      const int o = tile & 1;
      mbar_wait(oempty_bar(o), ((tile >> 1) & 1) ^ 1);
      tcgen05_fence_after();
      const uint32_t oaddr = tmem_base + lane_off + 256 + o * TC_BN;
#pragma unroll
      for (int c = 0; c < TC_BN / 32; ++c) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(sum[c * 32 + i]);
        tmem_st_32x32(oaddr + c * 32, r);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ofull_bar(o));
      ++tile;
    }
  } else {
    // ============ warpgroups 2, 3: epilogue (two column halves of every tile) ============
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    const int q = warp & 3;
    const int half = (warp >> 2) - 2;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int tile = 0;
#pragma unroll 1
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      TcWork wk;
      if (!tc_get_work(P, progs, s, w, wk)) continue;
      const int o = tile & 1;
      mbar_wait(ofull_bar(o), (tile >> 1) & 1);
      tcgen05_fence_after();
      const CUtensorMap* const smaps[3] = {&smap0, &smap1, &smap2};
      tc_epilogue_tile_tmem<kFmt>(P, wk, wk.tm, row_in_tile, lane,
                            stage_base + (warp - 8) * 2 * TC_STAGE_BYTES_PER_WARP,
                            tmem_base + lane_off + 256 + o * TC_BN, oempty_bar(o), 2 * half,
                            2 * half + 2, smaps);
      ++tile;
    }
    if (lane == 0) tma_store_wait_all();  // all bulk stores landed before the CTA retires
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// CTA-pair kernel with the 1-CTA kernel's role structure (scaled-fp16 planes only):
// cta_group::2, M = 256, N = 128.  A cluster of two CTAs owns a 256 x 128 block of the
// result; CTA r holds A rows [r*128, +128) and HALF of the B tile (rows [r*64, +64)), the
// leader's elected thread issues the MMAs for both, and every CTA keeps the 1-CTA kernel's
// private pipeline for ITS 128 x 128 accumulator: chunk-accumulation warpgroup -> two TMEM
// output stages -> two epilogue warpgroups (so the epilogue stays hidden, unlike pair256).
// Why: the ablation (profiles/r01s2_ncu_full_fp16x3.md) shows the 1-CTA kernel is paced by
// shared-memory operand fetch (8 KB read + 5.3 KB TMA-written per 64-cycle MMA).  Here each
// SM reads 6 KB and receives 4 KB per MMA: 4 stages of 48 KB.
// Tiles: (tm2, tn) with tn <= 2 tm2 + 1; the strictly-upper half of a diagonal pair is
// computed but not written (20 pair tiles = 40 of 64 tiles at n = 1024).
// ---------------------------------------------------------------------------
constexpr uint32_t kIdescF16M256N128 =
    (1u << 4) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

// loads rows [row_off, row_off + 64) of storage tile (tile_row, tile_col): 8 KiB (box 64 x 64)
__device__ __forceinline__ void tma_load_half_tile_2sm(uint32_t dst, const CUtensorMap* map,
                                                       uint32_t bar, int tile_col, int tile_row,
                                                       int row_off, int mat) {
  asm volatile(
      "cp.async.bulk.tensor.5d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & kPeerBitMask), "r"(0), "r"(row_off),
      "r"(tile_col), "r"(tile_row), "r"(mat)
      : "memory");
}

// pair-tile index t -> (tm2, tn), tn <= 2 tm2 + 1, t = tm2 (tm2 + 1) + tn
__device__ __forceinline__ void pair_decode(int t, int& tm2, int& tn) {
  int r = (int)((sqrtf(4.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
  while (r * (r + 1) > t) --r;
  while ((r + 1) * (r + 2) <= t) ++r;
  tm2 = r;
  tn = t - r * (r + 1);
}

__device__ __forceinline__ bool tc_get_work_ws2(const TcParams& P, const Program* progs, int s,
                                                int w, int cta_rank, TcWork& out) {
  const int t2 = P.tiles / 2;
  const int npair = t2 * (t2 + 1);
  const int per_mat = tc_ops_per_matrix(s) * npair;
  const int b = w / per_mat;
  int r = w - b * per_mat;
  const int op = r / npair;
  r -= op * npair;
  const RootCtl& c = P.ctl[b];
  if (!c.active) return false;
  out.op = op;
  if (op == 1) {
    out.st = Step{LB_HN, LB_H, LB_MI, 0};
  } else {
    const Program& pr = progs[c.p];
    if (s >= pr.nsteps) return false;
    out.st = pr.steps[s];
  }
  int tm2;
  pair_decode(r, tm2, out.tn);
  out.b = b;
  out.tm = 2 * tm2 + cta_rank;
  out.cur = c.cur;
  out.p = c.p;
  out.pad = c.pad;
  out.kblocks = (c.pad + TC_BK - 1) / TC_BK;
  return true;
}

template <int kStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_WS_THREADS, 1)
tc_phase_kernel_ws2(const __grid_constant__ CUtensorMap tmap0,
                    const __grid_constant__ CUtensorMap tmap1,
                    const __grid_constant__ CUtensorMap hmap0,
                    const __grid_constant__ CUtensorMap hmap1,
                    const __grid_constant__ CUtensorMap smap0,
                    const __grid_constant__ CUtensorMap smap1, const TcParams P,
                    const Program* __restrict__ progs, int s, int total_work) {
  constexpr int kATile = TC_TILE_BYTES;      // 128 rows x 64 k
  constexpr int kBTile = TC_TILE_BYTES / 2;  // this CTA's 64 rows of the B tile
  constexpr int kStageBytes = 2 * (kATile + kBTile);  // A0 A1 B0h B1h = 48 KiB
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int i) { return bar_base + 8u * i; };                         // leader's
  auto empty_bar = [&](int i) { return bar_base + 8u * (kStages + i); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * kStages + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * kStages + 2 + i); };   // leader's
  auto ofull_bar = [&](int i) { return bar_base + 8u * (2 * kStages + 4 + i); };
  auto oempty_bar = [&](int i) { return bar_base + 8u * (2 * kStages + 6 + i); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 8);
  const uint32_t stage_base = bar_base + 1024;
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(full_bar(i), 2);   // one arrival per CTA of the pair (+ the tx bytes of both)
      mbar_init(empty_bar(i), 1);  // tcgen05.commit, multicast to both CTAs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);   // tcgen05.commit, multicast
      mbar_init(tempty_bar(i), 8);  // 4 accumulate warps x 2 CTAs (leader's barrier)
      mbar_init(ofull_bar(i), 4);
      mbar_init(oempty_bar(i), 8);
    }
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && elect_one()) {
      // ===================== TMA producer (both CTAs) =====================
      int stage = 0;
      uint32_t phase = 0;
#pragma unroll 1
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        TcWork wk;
        if (!tc_get_work_ws2(P, progs, s, w, (int)cta_rank, wk)) continue;
        const int pa = physical_buf(wk.st.a, wk.cur), pb = physical_buf(wk.st.b, wk.cur);
#pragma unroll 1
        for (int kb = 0; kb < wk.kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t dst = smem_base + stage * kStageBytes;
          if (leader) mbar_expect_tx(full_bar(stage), 2 * kStageBytes);
          else mbar_arrive_remote(full_bar(stage), 0);
          tma_load_tile_2sm(dst, &tmap0, full_bar(stage), kb, wk.tm, pa * P.batch + wk.b);
          tma_load_tile_2sm(dst + kATile, &tmap1, full_bar(stage), kb, wk.tm, pa * P.batch + wk.b);
          tma_load_half_tile_2sm(dst + 2 * kATile, &hmap0, full_bar(stage), kb, wk.tn,
                                 (int)cta_rank * 64, pb * P.batch + wk.b);
          tma_load_half_tile_2sm(dst + 2 * kATile + kBTile, &hmap1, full_bar(stage), kb, wk.tn,
                                 (int)cta_rank * 64, pb * P.batch + wk.b);
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && leader && elect_one()) {
      // ===================== MMA issuer (leader CTA) =====================
      int stage = 0;
      uint32_t phase = 0;
      int chunk = 0;
#pragma unroll 1
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        TcWork wk;
        if (!tc_get_work_ws2(P, progs, s, w, 0, wk)) continue;
#pragma unroll 1
        for (int kb = 0; kb < wk.kblocks; ++kb, ++chunk) {
          const int acc = chunk & 1;
          mbar_wait(tempty_bar(acc), ((chunk >> 1) & 1) ^ 1);
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + acc * TC_BN;
          const uint32_t a0 = smem_base + stage * kStageBytes;
          const uint32_t b0 = a0 + 2 * kATile;
          issue_fp16_cross<true>(tmem_d, a0, b0, kIdescF16M256N128, true, kBTile);
          issue_fp16_main<true>(tmem_d, a0, b0, kIdescF16M256N128, true);
          umma_commit_2sm(empty_bar(stage));  // frees the slot in both CTAs
          umma_commit_2sm(tfull_bar(acc));    // chunk accumulator ready in both CTAs
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 8) {
    // ============ warpgroup 1: chunk accumulation -> TMEM output stage (own 128 rows) ============
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int chunk = 0, tile = 0;
#pragma unroll 1
    for (int w = cluster_id; w < total_work; w += num_clusters) {
      TcWork wk;
      if (!tc_get_work_ws2(P, progs, s, w, (int)cta_rank, wk)) continue;
      float sum[TC_BN];
#pragma unroll
      for (int i = 0; i < TC_BN; ++i) sum[i] = 0.f;
#pragma unroll 1
      for (int kb = 0; kb < wk.kblocks; ++kb, ++chunk) {
        const int acc = chunk & 1;
        mbar_wait(tfull_bar(acc), (chunk >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + lane_off + acc * TC_BN;
        if (!(P.ablate & 2)) {
#pragma unroll
          for (int c = 0; c < TC_BN / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + c * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) sum[c * 32 + i] += __uint_as_float(r[i]);
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(tempty_bar(acc));
          else mbar_arrive_remote(tempty_bar(acc), 0);
        }
      }
      const int o = tile & 1;
      mbar_wait(oempty_bar(o), ((tile >> 1) & 1) ^ 1);
      tcgen05_fence_after();
      const uint32_t oaddr = tmem_base + lane_off + 256 + o * TC_BN;
#pragma unroll
      for (int c = 0; c < TC_BN / 32; ++c) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(sum[c * 32 + i]);
        tmem_st_32x32(oaddr + c * 32, r);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ofull_bar(o));
      ++tile;
    }
  } else {
    // ============ warpgroups 2, 3: epilogue of this CTA's 128 x 128 tile ============
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    const int q = warp & 3;
    const int half = (warp >> 2) - 2;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int tile = 0;
#pragma unroll 1
    for (int w = cluster_id; w < total_work; w += num_clusters) {
      TcWork wk;
      if (!tc_get_work_ws2(P, progs, s, w, (int)cta_rank, wk)) continue;
      const int o = tile & 1;
      mbar_wait(ofull_bar(o), (tile >> 1) & 1);
      tcgen05_fence_after();
      if (wk.tm < wk.tn) {
        // strictly-upper tile of a diagonal pair: its mirror owns it; just recycle the stage
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(oempty_bar(o));
      } else {
        const CUtensorMap* const smaps[3] = {&smap0, &smap1, &smap1};
        tc_epilogue_tile_tmem<TC_FMT_FP16S>(P, wk, wk.tm, row_in_tile, lane,
                              stage_base + (warp - 8) * 2 * TC_STAGE_BYTES_PER_WARP,
                              tmem_base + lane_off + 256 + o * TC_BN, oempty_bar(o), 2 * half,
                              2 * half + 2, smaps);
      }
      ++tile;
    }
    if (lane == 0) tma_store_wait_all();
  }
  tcgen05_fence_before();
  cluster_sync_all();  // no CTA may exit while its peer can still signal it
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// CTA-pair kernel, 256 x 256 output tile per cluster (cta_group::2, M = 256, N = 256).
// This is the configuration that un-saturates shared memory: a 1-CTA M=128/N=128 MMA
// reads 8 KB of operands per 64 cycles = the full 128 B/cycle smem bandwidth, so TMA
// fills and epilogue staging starve it (measured: tensor pipe <= 69 %).  Here each CTA
// feeds 8 KB per 128 cycles (64 B/cycle) and TMA adds 31 B/cycle.
//   warpgroup 0   warp 0 TMA producer (both CTAs), warp 1 MMA issuer (leader), warp 2 TMEM
//   warpgroup 1/2 chunk accumulation of columns [0,128) / [128,256) of this CTA's 128
//                 rows in fp32 registers, then the epilogue of that 128 x 128 sub-tile
//                 (3-plane split, swizzled staging, bulk tensor stores, mirror, M_i', err)
// TMEM: 2 chunk accumulators x 256 columns.  setmaxnreg 40 / 232 / 232.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool tc_get_work_pair256(const TcParams& P, const Program* progs, int s,
                                                    int w, TcWork& out, int& tm2, int& tn2) {
  const int t2 = P.tiles / 2;
  const int ntri = t2 * (t2 + 1) / 2;
  const int per_mat = tc_ops_per_matrix(s) * ntri;
  const int b = w / per_mat;
  int r = w - b * per_mat;
  const int op = r / ntri;
  r -= op * ntri;
  const RootCtl& c = P.ctl[b];
  if (!c.active) return false;
  out.op = op;
  if (op == 1) {
    out.st = Step{LB_HN, LB_H, LB_MI, 0};
  } else {
    const Program& pr = progs[c.p];
    if (s >= pr.nsteps) return false;
    out.st = pr.steps[s];
  }
  tri_decode(r, tm2, tn2);
  out.b = b;
  out.cur = c.cur;
  out.p = c.p;
  out.pad = c.pad;
  out.kblocks = (c.pad + TC_BK - 1) / TC_BK;
  return true;
}

// Epilogue of one 128 x 128 sub-tile whose fp32 values sit in this thread's registers.
template <int kFmt>
__device__ __forceinline__ void tc_epilogue_regs(const TcParams& P, const TcWork& wk, int tm, int tn,
                                                 int row_in_tile, int lane, uint32_t stage,
                                                 const float (&sum)[TC_BN],
                                                 const CUtensorMap* const (&smaps)[3]) {
  const int q = row_in_tile >> 5;
  const int row = tm * TC_BM + row_in_tile;
  const int row0 = tm * TC_BM + q * 32;
  const bool diag_tile = tm == tn;
  const float alpha = -1.0f / (float)wk.p, oma = 1.0f - alpha;
  const bool mirror_on = true;
  uint32_t emax = 0;
  const EpiAddr ea = make_epi_addr(stage, stage + TC_STAGE_BYTES_PER_WARP, lane);
  auto write_group = [&](int buf, float (&x)[32], int col0, bool diag_sub) {
    store_block_3planes_tma<kFmt>(P, stage, ea, lane, x, diag_sub, mirror_on, smaps, row0, col0,
                            buf * P.batch + wk.b);
  };
#pragma unroll
  for (int c = 0; c < TC_BN / 32; ++c) {
    if (diag_tile && c > q) continue;  // strictly upper sub-block: its mirror writes it
    const bool diag_sub = diag_tile && c == q;
    const int col0 = tn * TC_BN + c * 32;
    float g[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) g[i] = sum[c * 32 + i];
#pragma unroll 1
    for (int o = wk.st.emit_mi ? 0 : 1; o < 2; ++o) {  // o = 0: M_i' (DS:844), o = 1: OUT
      float x[32];
      if (o == 0) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const bool dg = (col0 + i == row) && (row < wk.pad);
          if (!diag_sub || col0 + i <= row) {
            const uint32_t ab = absbits(g[i] - (dg ? 1.f : 0.f));  // DS:847
            emax = ab > emax ? ab : emax;
          }
          x[i] = mi_from_m(g[i], dg, alpha, oma);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = g[i];
      }
      write_group(physical_buf(o == 0 ? (int)LB_MIN : (int)wk.st.dst, wk.cur), x, col0, diag_sub);
    }
  }
  if (wk.st.emit_mi) {
    emax = warp_max_u32(emax);
    if (lane == 0 && emax) atomicMax(P.errbits + wk.b, emax);
  }
}

constexpr int TC_P256_THREADS = 384;

template <int kLP, int kStages, int kFmt, int kChunkKB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC_P256_THREADS, 1)
tc_phase_kernel_pair256(const __grid_constant__ CUtensorMap tmap0,
                        const __grid_constant__ CUtensorMap tmap1,
                        const __grid_constant__ CUtensorMap tmap2,
                        const __grid_constant__ CUtensorMap smap0,
                        const __grid_constant__ CUtensorMap smap1,
                        const __grid_constant__ CUtensorMap smap2, const TcParams P,
                        const Program* __restrict__ progs, int s, int total_work) {
  constexpr int kStageBytes = 2 * kLP * TC_TILE_BYTES;  // A: 128 rows, B: this CTA's 128 rows
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + kStages * kStageBytes;
  auto full_bar = [&](int i) { return bar_base + 8u * i; };            // leader's is used
  auto empty_bar = [&](int i) { return bar_base + 8u * (kStages + i); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * kStages + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * kStages + 2 + i); };  // leader's
  const uint32_t tmem_slot = bar_base + 8u * (2 * kStages + 4);
  const uint32_t stage_base = bar_base + 1024;
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(full_bar(i), 2);
      mbar_init(empty_bar(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 16);  // 8 accumulate warps x 2 CTAs
    }
    fence_barrier_init();
  }
  cluster_sync_all();
  if (warp == 2) tmem_alloc_2sm(tmem_slot, 512);
  if (warp == 3 && blockIdx.x == 0 && P.sync_ctr_next)
    for (int i = lane; i < 2 * P.batch; i += 32) P.sync_ctr_next[i] = 0u;
  tcgen05_fence_before();
  cluster_sync_all();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && elect_one()) {
      // ===================== TMA producer (both CTAs) =====================
      int stage = 0;
      uint32_t phase = 0;
#pragma unroll 1
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        TcWork wk;
        int tm2, tn2;
        if (!tc_get_work_pair256(P, progs, s, w, wk, tm2, tn2)) continue;
        tc_group_sync(P, wk, (P.tiles / 2) * (P.tiles / 2 + 1) / 2, s);
        const int pa = physical_buf(wk.st.a, wk.cur), pb = physical_buf(wk.st.b, wk.cur);
        const int atile = 2 * tm2 + (int)cta_rank, btile = 2 * tn2 + (int)cta_rank;
#pragma unroll 1
        for (int kb = 0; kb < wk.kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t dst = smem_base + stage * kStageBytes;
          if (leader) mbar_expect_tx(full_bar(stage), 2 * kStageBytes);
          else mbar_arrive_remote(full_bar(stage), 0);
          const CUtensorMap* maps[3] = {&tmap0, &tmap1, &tmap2};
#pragma unroll
          for (int pl = 0; pl < kLP; ++pl) {
            tma_load_tile_2sm(dst + pl * TC_TILE_BYTES, maps[pl], full_bar(stage), kb, atile,
                              pa * P.batch + wk.b);
            tma_load_tile_2sm(dst + (kLP + pl) * TC_TILE_BYTES, maps[pl], full_bar(stage), kb,
                              btile, pb * P.batch + wk.b);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && leader && elect_one()) {
      // ===================== MMA issuer (leader CTA) =====================
      int stage = 0;
      uint32_t phase = 0;
      int chunk = 0;
#pragma unroll 1
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        TcWork wk;
        int tm2, tn2;
        if (!tc_get_work_pair256(P, progs, s, w, wk, tm2, tn2)) continue;
#pragma unroll 1
        for (int kb = 0; kb < wk.kblocks; kb += kChunkKB, ++chunk) {
          const int nkb = min(kChunkKB, wk.kblocks - kb);
          const int acc = chunk & 1;
          mbar_wait(tempty_bar(acc), ((chunk >> 1) & 1) ^ 1);
          const uint32_t tmem_d = tmem_base + acc * 256;
          constexpr uint32_t idesc = kFmt == TC_FMT_FP16S ? kIdescF16M256N256 : kIdescBf16M256N256;
          int st[kChunkKB];
#pragma unroll
          for (int j = 0; j < kChunkKB; ++j) {
            if (j < nkb) {
              st[j] = stage;
              mbar_wait(full_bar(stage), phase);
              tcgen05_fence_after();
              const uint32_t a0 = smem_base + stage * kStageBytes;
              const uint32_t b0 = a0 + kLP * TC_TILE_BYTES;
              if (kFmt == TC_FMT_FP16S) {
                issue_fp16_cross<true>(tmem_d, a0, b0, idesc, j == 0);
              } else {
                issue_bf16_block<kLP, true>(tmem_d, a0, b0, idesc, j == 0);
                umma_commit_2sm(empty_bar(stage));
              }
              if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
          }
          if (kFmt == TC_FMT_FP16S) {
#pragma unroll
            for (int j = 0; j < kChunkKB; ++j) {
              if (j < nkb) {
                const uint32_t a0 = smem_base + st[j] * kStageBytes;
                issue_fp16_main<true>(tmem_d, a0, a0 + kLP * TC_TILE_BYTES, idesc, j == 0);
                umma_commit_2sm(empty_bar(st[j]));
              }
            }
          }
          umma_commit_2sm(tfull_bar(acc));
        }
      }
    }
  } else {
    // ======= warpgroups 1, 2: accumulate columns [half*128, +128) and write that sub-tile =======
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    const int q = warp & 3;
    const int half = (warp >> 2) - 1;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const CUtensorMap* const smaps[3] = {&smap0, &smap1, &smap2};
    int chunk = 0;
#pragma unroll 1
    for (int w = cluster_id; w < total_work; w += num_clusters) {
      TcWork wk;
      int tm2, tn2;
      if (!tc_get_work_pair256(P, progs, s, w, wk, tm2, tn2)) continue;
      float sum[TC_BN];
#pragma unroll
      for (int i = 0; i < TC_BN; ++i) sum[i] = 0.f;
#pragma unroll 1
      for (int kb = 0; kb < wk.kblocks; kb += kChunkKB, ++chunk) {
        const int acc = chunk & 1;
        mbar_wait(tfull_bar(acc), (chunk >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + lane_off + acc * 256 + half * TC_BN;
#pragma unroll
        for (int c = 0; c < TC_BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) sum[c * 32 + i] += __uint_as_float(r[i]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (leader) mbar_arrive(tempty_bar(acc));
          else mbar_arrive_remote(tempty_bar(acc), 0);
        }
      }
      const int tm = 2 * tm2 + (int)cta_rank, tn = 2 * tn2 + half;
      if (tm >= tn)  // (2 tm2, 2 tm2 + 1) of a diagonal pair-tile is upper: its mirror owns it
        tc_epilogue_regs<kFmt>(P, wk, tm, tn, row_in_tile, lane,
                         stage_base + (warp - 4) * 2 * TC_STAGE_BYTES_PER_WARP, sum, smaps);
    }
    if (lane == 0) tma_store_wait_all();
  }
  tcgen05_fence_before();
  cluster_sync_all();
  if (warp == 2) tmem_dealloc_2sm(tmem_base, 512);
}

// ---------------------------------------------------------------------------
// plane store/load policy for the shared init / final kernels
// ---------------------------------------------------------------------------
struct PlaneStore {
  uint16_t* plane[3];
  size_t buf_stride, mat_elems;
  int fmt;  // TC_FMT_*
  // H is kept as H / hmul with hmul = z^(1/p) / s, s a power of two <= 1 chosen so that
  // every entry stays inside fp16's range: (z eps)^(-1/p) bounds the largest eigenvalue
  // of H / z^(1/p) (eps = the damping of this try, a lower bound of lambda_min).
  __device__ __forceinline__ float h_init_scale(float h0, float bound, float* hmul) const {
    if (fmt != TC_FMT_FP16S) { *hmul = 1.0f; return h0; }
    float s = 1.0f;
    if (bound > 8192.0f) {
      int e = 0;
      frexpf(bound / 8192.0f, &e);  // bound / 8192 in [2^(e-1), 2^e)
      s = ldexpf(1.0f, -(e < 100 ? e : 100));
    }
    *hmul = h0 / s;
    return s;
  }
  // blocked layout: 128 x 64 tiles, each contiguous
  __device__ __forceinline__ size_t elem(int phys, int b, int i, int j, int n) const {
    return (size_t)phys * buf_stride + (size_t)b * mat_elems +
           ((size_t)(i >> 7) * (n >> 6) + (j >> 6)) * 8192 + (size_t)(i & 127) * 64 + (j & 63);
  }
  __device__ __forceinline__ void store(int phys, int b, int i, int j, int n, float v) const {
    const size_t off = elem(phys, b, i, j, n);
    if (fmt == TC_FMT_FP16S) {
      const float c = fabsf(v) > 65504.0f ? copysignf(65504.0f, v) : v;  // NaN stays NaN
      const __half h0 = __float2half_rn(c);
      const float r1 = (c - __half2float(h0)) * TC_FP16_SCALE;
      plane[0][off] = __half_as_ushort(h0);
      plane[1][off] = __half_as_ushort(__float2half_rn(r1));
      return;
    }
    const __nv_bfloat16 a0 = __float2bfloat16_rn(v);
    const float r1 = v - __bfloat162float(a0);
    const __nv_bfloat16 a1 = __float2bfloat16_rn(r1);
    const float r2 = r1 - __bfloat162float(a1);
    plane[0][off] = __bfloat16_as_ushort(a0);
    plane[1][off] = __bfloat16_as_ushort(a1);
    plane[2][off] = __bfloat16_as_ushort(__float2bfloat16_rn(r2));
  }
  __device__ __forceinline__ float load(int phys, int b, int i, int j, int n) const {
    const size_t off = elem(phys, b, i, j, n);
    if (fmt == TC_FMT_FP16S) {
      const float x0 = __half2float(__ushort_as_half(plane[0][off]));
      const float x1 = __half2float(__ushort_as_half(plane[1][off]));
      return fmaf(x1, 1.0f / TC_FP16_SCALE, x0);
    }
    const float x0 = __bfloat162float(__ushort_as_bfloat16(plane[0][off]));
    const float x1 = __bfloat162float(__ushort_as_bfloat16(plane[1][off]));
    const float x2 = __bfloat162float(__ushort_as_bfloat16(plane[2][off]));
    return (x2 + x1) + x0;
  }
};

// ---------------------------------------------------------------------------
// First initialisation of M0 / M_i0 / H0 (DS:871-874) straight into the plane tiles of the
// scaled-fp16 engine: one CTA per 128 x 64 plane tile, a thread owns 32 consecutive columns of
// one row, so every plane row leaves as four 16-byte stores (the strip kernel stores 2 bytes
// at a time).  The lower triangle of the input is authoritative: tiles above the diagonal read
// the mirrored tile coalesced and transpose it through shared memory; the two tiles per row
// block that straddle the diagonal take element-wise max / min indexing.  The last CTA of a
// matrix publishes err0 and runs the loop-predicate bookkeeping, like root_init_strip_kernel.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void split_fp16s(float v, uint16_t* hi, uint16_t* lo) {
  const float c = fabsf(v) > 65504.0f ? copysignf(65504.0f, v) : v;  // NaN stays NaN
  const __half h0 = __float2half_rn(c);
  *hi = __half_as_ushort(h0);
  *lo = __half_as_ushort(__float2half_rn((c - __half2float(h0)) * TC_FP16_SCALE));
}

__global__ void __launch_bounds__(256, 3)
root_init_tile_kernel(const float* __restrict__ xs, RootCtl* ctl, PlaneStore bufs, int batch, int n,
                      int strips, RootParams prm, float* __restrict__ scratch) {
  __shared__ float tr[64][129];
  __shared__ uint32_t ured[32];
  __shared__ int is_last;
  const int b = blockIdx.z, ti = blockIdx.y, tj = blockIdx.x;
  RootCtl c = ctl[b];
  if (!c.need_init) return;
  const int pad = c.pad, p = c.p;
  const float eps = root_try_eps(c, prm);
  const int t = threadIdx.x, r = t >> 1, half = t & 1;
  // |A + eps I|_F^2 from the strip partials: one warp, the same tree in every CTA (a serial
  // loop over the partials cost 5 us of dependent loads per CTA, 32 waves of CTAs per call)
  __shared__ float ssum_s;
  if (t < 32) {
    float v = 0.f;
    for (int q = t; q < strips; q += 32) v += scratch[(size_t)b * strips + q];
    v = warp_sum(v);
    if (t == 0) ssum_s = v;
  }
  __syncthreads();
  const float norm = sqrtf(ssum_s);
  const float alpha = -1.0f / (float)p, one_minus_alpha = 1.0f - alpha;
  const float z = (float)(1 + p) / (2.0f * norm);
  const float h0 = powf(z, (float)(1.0 / (double)p));  // DS:873
  float hmul = 1.0f;
  const float hdiag = bufs.h_init_scale(h0, powf(fmaxf(z * eps, 1e-37f), alpha), &hmul);
  const float* A = xs + (size_t)b * n * n;
  const int i = ti * 128 + r, j0 = tj * 64 + half * 32;
  float a[32];
  if (tj * 64 + 63 <= ti * 128) {  // whole tile on or below the diagonal: rows as stored
    const float4* src = reinterpret_cast<const float4*>(A + (size_t)i * n + j0);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = __ldg(src + q);
      a[4 * q] = v.x; a[4 * q + 1] = v.y; a[4 * q + 2] = v.z; a[4 * q + 3] = v.w;
    }
  } else if (tj * 64 >= ti * 128 + 128) {  // whole tile above: the mirrored tile, transposed
    const int mr = t >> 2, mc = (t & 3) * 32;  // row of the mirrored tile (a column here)
    const float4* src = reinterpret_cast<const float4*>(A + (size_t)(tj * 64 + mr) * n + ti * 128 + mc);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = __ldg(src + q);
      tr[mr][mc + 4 * q] = v.x; tr[mr][mc + 4 * q + 1] = v.y;
      tr[mr][mc + 4 * q + 2] = v.z; tr[mr][mc + 4 * q + 3] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 32; ++k) a[k] = tr[half * 32 + k][r];
  } else {  // straddles the diagonal
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int j = j0 + k;
      a[k] = __ldg(A + (size_t)max(i, j) * n + min(i, j));
    }
  }
  uint32_t emax = 0;
  // eight columns at a time: split, then the six 16-byte stores of the chunk (keeps the packed
  // words of only one chunk live: 3 CTAs per SM instead of 2)
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    uint32_t m_hi[4], m_lo[4], mi_hi[4], mi_lo[4], h_hi[4], h_lo[4];
#pragma unroll
    for (int kq = 0; kq < 4; ++kq) {
      const int k = ch * 8 + 2 * kq;
      uint16_t w[2][6];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int j = j0 + k + u;
        float m0 = 0.f, mi0 = 0.f, h = 0.f;
        if (i < pad && j < pad) {
          float av = a[k + u];
          if (i == j) av += eps;
          m0 = av * z;  // DS:871
          const uint32_t ab = absbits(m0 - (i == j ? 1.f : 0.f));
          emax = ab > emax ? ab : emax;
          mi0 = mi_from_m(m0, i == j, alpha, one_minus_alpha);
          h = (i == j) ? hdiag : 0.f;
        }
        split_fp16s(m0, &w[u][0], &w[u][1]);
        split_fp16s(mi0, &w[u][2], &w[u][3]);
        split_fp16s(h, &w[u][4], &w[u][5]);
      }
      m_hi[kq] = w[0][0] | ((uint32_t)w[1][0] << 16);  m_lo[kq] = w[0][1] | ((uint32_t)w[1][1] << 16);
      mi_hi[kq] = w[0][2] | ((uint32_t)w[1][2] << 16); mi_lo[kq] = w[0][3] | ((uint32_t)w[1][3] << 16);
      h_hi[kq] = w[0][4] | ((uint32_t)w[1][4] << 16);  h_lo[kq] = w[0][5] | ((uint32_t)w[1][5] << 16);
    }
    auto put = [&](int phys, int plane, const uint32_t (&v)[4]) {
      uint4* dst = reinterpret_cast<uint4*>(bufs.plane[plane] + bufs.elem(phys, b, i, j0, n));
      dst[ch] = make_uint4(v[0], v[1], v[2], v[3]);
    };
    put(0, 0, m_hi); put(0, 1, m_lo);
    put(2, 0, mi_hi); put(2, 1, mi_lo);
    put(4, 0, h_hi); put(4, 1, h_lo);
  }
  emax = block_max_u32(emax, ured);
  uint32_t* slots = reinterpret_cast<uint32_t*>(scratch + (size_t)batch * strips);
  if (threadIdx.x == 0) {
    atomicMax(slots + b, emax);
    __threadfence();
    is_last = atomicAdd(slots + batch + b, 1u) == (uint32_t)(gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    const uint32_t e = atomicMax(slots + b, 0u);  // atomic read of the final maximum
    if (c.tries == 0) {
      const float ev = prm.relative_eps ? c.max_ev : 1.0f;
      c.max_ev = ev;
      c.ridge = prm.ridge_epsilon * fmaxf(ev, 1e-25f);
    }
    c.need_init = 0;
    c.iter = 0;
    c.cur = 0;
    c.err = __uint_as_float(e);  // DS:872
    c.ratio = 1.0f;
    c.hmul = hmul;
    root_after_error_update(c, prm);
    ctl[b] = c;
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// kernel-variant knobs read from the environment at every call (part of the graph-cache key)
int tc_engine_variant_mask() {
  int m = 0;
  const char* v;
  if ((v = getenv("PC_TC_PAIR256")) && v[0] == '1') m |= 1;
  if ((v = getenv("PC_TC_WS2")) && v[0] == '1') m |= 2;
  if ((v = getenv("PC_TC_CHUNK")) && v[0] == '2') m |= 4;
  if ((v = getenv("PC_TC_ABLATE"))) m |= atoi(v) << 4;
  return m;
}

bool tc_engine_available() {
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess)
    return false;
  return major == 10;
}

static size_t tc_plane_bytes(int batch, int n, int planes) {
  return (size_t)planes * kNumBufs * batch * n * n * sizeof(uint16_t);
}
size_t tc_engine_bytes(int batch, int n, int planes) {
  return tc_plane_bytes(batch, n, planes) + 1024 + 4 * (size_t)batch * sizeof(uint32_t) + 256;
}

struct TcHostState {
  CUtensorMap maps[3];     // operand loads: one 128 x 64 storage tile (16 KiB, contiguous)
  CUtensorMap maps_st[3];  // epilogue bulk stores: box {32, 32}, SWIZZLE_64B
  CUtensorMap maps_half[2];  // B half tiles of the ws2 kernel: box {64, 64}, SWIZZLE_128B
  bool use_ws2;            // cta_group::2 M=256/N=128 kernel with output stages (fp16 planes)
  bool use_pair256;        // cta_group::2, 256 x 256 cluster tiles (PC_TC_PAIR256=1)
  int fmt, planes;         // TC_FMT_*, stored planes per matrix
  int chunk_kb;            // 64-column k-blocks accumulated in TMEM before the fp32 register add
  TcParams prm;
  Program* progs_dev;
  uint32_t* sync_mem;  // 2 x [2 * batch] group counters, used alternately by successive launches
  unsigned launch_seq;
  bool group_sync;     // PC_TC_SYNC=1 enables the rendezvous (off by default)
  int sms;
};

// Grid and synchronisation group of one launch.  `units` = CTAs (ws kernel) or clusters
// (pair kernel) the device can hold, `ntri` = work items per (matrix, product), `cpi` =
// producer threads per item.  A group must fit in one round (group <= units) and rounds
// must not split groups (grid % group == 0), otherwise a CTA would wait on itself.
static void plan_launch(TcHostState* hs, int s, int ntri, int units, int cpi, int* grid_units,
                        int* total_work) {
  const int per_mat = (s == 0 ? 2 : 1) * ntri;
  *total_work = hs->prm.batch * per_mat;
  int group = 0, best = 0;
  if (hs->group_sync) {
    for (int cand : {per_mat, ntri}) {
      if (cand <= 1 || cand > units) continue;
      const int g = (units / cand) * cand;
      if (g > best) { best = g; group = cand; }
    }
  }
  uint32_t* cur = hs->sync_mem + (size_t)(hs->launch_seq & 1) * 2 * hs->prm.batch;
  uint32_t* nxt = hs->sync_mem + (size_t)((hs->launch_seq + 1) & 1) * 2 * hs->prm.batch;
  ++hs->launch_seq;
  hs->prm.sync_ctr = cur;
  hs->prm.sync_ctr_next = nxt;
  hs->prm.sync_group = group;
  hs->prm.sync_arrivals = group * cpi;
  int g = group ? best : units;
  if (!group && *total_work < g) g = *total_work;
  *grid_units = g;
}

static Program* g_progs_dev[64] = {nullptr};

constexpr size_t kPhaseSmemBytes(int stages, int lp) {
  return (size_t)stages * 2 * lp * TC_TILE_BYTES + 1024 + 1024 + 16 * TC_STAGE_BYTES_PER_WARP;
}

// One-time, per-device setup that must not run inside a stream capture: dynamic shared-memory
// attributes of every kernel variant, the L2 policy constants, the device copy of the step
// programs.
int tc_engine_prepare() {
  static std::mutex mu;
  static bool done[64] = {false};
  int dev = 0;
  PC_CUDA_CHECK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  if (dev < 64 && done[dev]) return PC_OK;
#define PC_TC_ATTR(kernel, bytes) \
  PC_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)))
  PC_TC_ATTR((tc_phase_kernel_ws<2, 3, TC_FMT_FP16S, 1>), kPhaseSmemBytes(3, 2));
  PC_TC_ATTR((tc_phase_kernel_ws<2, 3, TC_FMT_FP16S, 2>), kPhaseSmemBytes(3, 2));
  PC_TC_ATTR((tc_phase_kernel_ws<3, 2, TC_FMT_BF16, 1>), kPhaseSmemBytes(2, 3));
  PC_TC_ATTR((tc_phase_kernel_ws<2, 3, TC_FMT_BF16, 1>), kPhaseSmemBytes(3, 2));
  PC_TC_ATTR((tc_phase_kernel_pair256<2, 3, TC_FMT_FP16S, 1>), kPhaseSmemBytes(3, 2));
  PC_TC_ATTR((tc_phase_kernel_pair256<2, 3, TC_FMT_FP16S, 2>), kPhaseSmemBytes(3, 2));
  PC_TC_ATTR((tc_phase_kernel_pair256<3, 2, TC_FMT_BF16, 1>), kPhaseSmemBytes(2, 3));
  PC_TC_ATTR((tc_phase_kernel_pair256<2, 3, TC_FMT_BF16, 1>), kPhaseSmemBytes(3, 2));
  PC_TC_ATTR((tc_phase_kernel_ws2<4>),
             (size_t)4 * 3 * TC_TILE_BYTES + 1024 + 1024 + 16 * TC_STAGE_BYTES_PER_WARP);
#undef PC_TC_ATTR
  {
    const char* h = getenv("PC_TC_HINTS");
    const int want = (h && h[0] == '1') ? 1 : 0;
    const uint64_t lp = want ? kPolicyEvictLast : kPolicyEvictNormal;
    const uint64_t sp = want ? kPolicyEvictFirst : kPolicyEvictNormal;
    PC_CUDA_CHECK(cudaMemcpyToSymbol(g_load_policy, &lp, sizeof(lp)));
    PC_CUDA_CHECK(cudaMemcpyToSymbol(g_store_policy, &sp, sizeof(sp)));
  }
  if (dev < 64 && !g_progs_dev[dev]) {
    Program hp[kMaxP + 1];
    memset(hp, 0, sizeof(hp));
    for (int p = 1; p <= kMaxP; ++p) build_program(p, &hp[p]);
    PC_CUDA_CHECK(cudaMalloc(&g_progs_dev[dev], sizeof(hp)));
    PC_CUDA_CHECK(cudaMemcpy(g_progs_dev[dev], hp, sizeof(hp), cudaMemcpyHostToDevice));
  }
  if (dev < 64) done[dev] = true;
  return PC_OK;
}

int tc_engine_init(TcEngine* e, void* mem, int batch, int n, int passes, int fmt,
                   cudaStream_t stream) {
  PC_REQUIRE(n % TC_BM == 0, "tcgen05 engine needs n %% 128 == 0");
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return PC_ERR_UNSUPPORTED;
  }
  e->batch = batch; e->n = n; e->passes = passes; e->fmt = fmt;
  auto* hs = new TcHostState();
  e->host_state = hs;
  hs->fmt = fmt;
  hs->planes = fmt == TC_FMT_FP16S ? 2 : 3;
  for (int pl = 0; pl < 3; ++pl) hs->prm.plane[pl] = nullptr;
  uint16_t* base = reinterpret_cast<uint16_t*>(align_up((size_t)mem, 1024));
  const size_t buf_stride = (size_t)batch * n * n;
  for (int pl = 0; pl < hs->planes; ++pl) {
    uint16_t* plane = base + (size_t)pl * kNumBufs * buf_stride;
    hs->prm.plane[pl] = plane;
    for (int k = 0; k < kNumBufs; ++k) e->planes[k][pl] = plane + (size_t)k * buf_stride;
    // Blocked storage: a matrix is a grid of 128 x 64 tiles, each contiguous (16 KiB),
    // so one TMA box = one DRAM-contiguous chunk (no 2 KiB power-of-two row strides).
    // 5-D view: [col in tile (64), row in tile (128), tile col (n/64), tile row (n/128), matrix]
    cuuint64_t dims[5] = {64, 128, (cuuint64_t)(n / 64), (cuuint64_t)(n / 128),
                          (cuuint64_t)batch * kNumBufs};
    cuuint64_t strides[4] = {64 * 2, 8192 * 2, (cuuint64_t)8192 * (n / 64) * 2,
                             (cuuint64_t)n * n * 2};
    cuuint32_t box[5] = {64, 128, 1, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUtensorMapDataType dt =
        fmt == TC_FMT_FP16S ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    CUresult r = enc(&hs->maps[pl], dt, 5, plane, dims, strides, box,
                     estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint32_t box_h[5] = {64, 64, 1, 1, 1};
    if (r == CUDA_SUCCESS && pl < 2)
      r = enc(&hs->maps_half[pl], dt, 5, plane, dims, strides, box_h, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint32_t box_s[5] = {32, 32, 1, 1, 1};
    if (r == CUDA_SUCCESS)
      r = enc(&hs->maps_st[pl], dt, 5, plane, dims, strides, box_s,
              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B,
              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled failed with %d", (int)r);
      delete hs;
      e->host_state = nullptr;
      return PC_ERR_CUDA;
    }
  }
  for (int pl = hs->planes; pl < 3; ++pl) {  // unused slots: valid descriptors, never dereferenced
    hs->maps[pl] = hs->maps[0];
    hs->maps_st[pl] = hs->maps_st[0];
  }
  {
    const char* p256 = getenv("PC_TC_PAIR256");
    // measured on B200 (batch 74, n = 1024): the 1-CTA kernel is 3-7 % faster for both
    // plane formats (36 of 64 tiles instead of 40 of 64, hidden epilogue), so the CTA-pair
    // kernel is opt-in: PC_TC_PAIR256=1
    hs->use_pair256 = (n % 256 == 0) && (p256 && p256[0] == '1');
    const char* w2 = getenv("PC_TC_WS2");
    hs->use_ws2 = fmt == TC_FMT_FP16S && (n % 256 == 0) && (w2 && w2[0] == '1');
    const char* ab = getenv("PC_TC_ABLATE");
    hs->prm.ablate = ab ? atoi(ab) : 0;
    const char* ck = getenv("PC_TC_CHUNK");
    // measured on B200: 128-column chunks are not faster (the TMEM pull is not the limiter)
    // and cost accuracy, so 64 stays the default; PC_TC_CHUNK=2 selects 128
    hs->chunk_kb = (fmt == TC_FMT_FP16S && ck && ck[0] == '2') ? 2 : 1;
  }
  {
    const char* gs = getenv("PC_TC_SYNC");
    hs->group_sync = gs && gs[0] == '1';  // measured: DRAM reads -60 %, but slower
    hs->sync_mem = reinterpret_cast<uint32_t*>(
        align_up((size_t)(reinterpret_cast<char*>(base) + tc_plane_bytes(batch, n, hs->planes)),
                 256));
    hs->launch_seq = 0;
    PC_CUDA_CHECK(cudaMemsetAsync(hs->sync_mem, 0, 4 * (size_t)batch * sizeof(uint32_t), stream));
    int dev = 0;
    hs->sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&hs->sms, cudaDevAttrMultiProcessorCount, dev);
  }
  hs->prm.buf_stride = buf_stride;
  hs->prm.mat_stride = (size_t)n * n;
  hs->prm.n = n; hs->prm.batch = batch; hs->prm.tiles = n / TC_BM;
  int dev = 0;
  PC_CUDA_CHECK(cudaGetDevice(&dev));
  int rc_prep = tc_engine_prepare();  // no-op when already done (must precede any capture)
  if (rc_prep != PC_OK) return rc_prep;
  hs->progs_dev = g_progs_dev[dev < 64 ? dev : 0];
  return PC_OK;
}

template <int kLP, int kStages, int kFmt, int kChunkKB>
static int launch_phase_ws(TcHostState* hs, int s, cudaStream_t stream) {
  constexpr size_t smem = (size_t)kStages * 2 * kLP * TC_TILE_BYTES + 1024 + 1024 +
                          16 * TC_STAGE_BYTES_PER_WARP;
  int grid, total_work;
  plan_launch(hs, s, hs->prm.tiles * (hs->prm.tiles + 1) / 2, hs->sms, 1, &grid, &total_work);
  tc_phase_kernel_ws<kLP, kStages, kFmt, kChunkKB><<<grid, TC_WS_THREADS, smem, stream>>>(
      hs->maps[0], hs->maps[1], hs->maps[2], hs->maps_st[0], hs->maps_st[1], hs->maps_st[2],
      hs->prm, hs->progs_dev, s, total_work);
  return PC_OK;
}

template <int kLP, int kStages, int kFmt, int kChunkKB>
static int launch_phase_pair256(TcHostState* hs, int s, cudaStream_t stream) {
  constexpr size_t smem = (size_t)kStages * 2 * kLP * TC_TILE_BYTES + 1024 + 1024 +
                          16 * TC_STAGE_BYTES_PER_WARP;
  const int t2 = hs->prm.tiles / 2;
  int clusters, total_work;
  plan_launch(hs, s, t2 * (t2 + 1) / 2, hs->sms / 2, 2, &clusters, &total_work);
  tc_phase_kernel_pair256<kLP, kStages, kFmt, kChunkKB><<<2 * clusters, TC_P256_THREADS, smem, stream>>>(
      hs->maps[0], hs->maps[1], hs->maps[2], hs->maps_st[0], hs->maps_st[1], hs->maps_st[2],
      hs->prm, hs->progs_dev, s, total_work);
  return PC_OK;
}

static int launch_phase_ws2(TcHostState* hs, int s, cudaStream_t stream) {
  constexpr int kStages = 4;
  constexpr size_t smem = (size_t)kStages * 3 * TC_TILE_BYTES + 1024 + 1024 +
                          16 * TC_STAGE_BYTES_PER_WARP;
  const int t2 = hs->prm.tiles / 2;
  int clusters, total_work;
  plan_launch(hs, s, t2 * (t2 + 1), hs->sms / 2, 2, &clusters, &total_work);
  hs->prm.sync_group = 0;
  tc_phase_kernel_ws2<kStages><<<2 * clusters, TC_WS_THREADS, smem, stream>>>(
      hs->maps[0], hs->maps[1], hs->maps_half[0], hs->maps_half[1], hs->maps_st[0],
      hs->maps_st[1], hs->prm, hs->progs_dev, s, total_work);
  return PC_OK;
}

static int launch_phase(TcHostState* hs, int passes, int s, cudaStream_t stream) {
  if (hs->fmt == TC_FMT_FP16S && hs->use_ws2) return launch_phase_ws2(hs, s, stream);
  if (hs->fmt == TC_FMT_FP16S) {
    // K-chunk of 128 columns (24 MMAs per TMEM accumulation, like bf16x6 at 64): the
    // chunk pull from TMEM (64 B/clk) then stays shorter than the chunk's MMAs
    if (hs->chunk_kb == 2)
      return hs->use_pair256 ? launch_phase_pair256<2, 3, TC_FMT_FP16S, 2>(hs, s, stream)
                             : launch_phase_ws<2, 3, TC_FMT_FP16S, 2>(hs, s, stream);
    return hs->use_pair256 ? launch_phase_pair256<2, 3, TC_FMT_FP16S, 1>(hs, s, stream)
                           : launch_phase_ws<2, 3, TC_FMT_FP16S, 1>(hs, s, stream);
  }
  if (hs->use_pair256)
    return passes == 6 ? launch_phase_pair256<3, 2, TC_FMT_BF16, 1>(hs, s, stream)
                       : launch_phase_pair256<2, 3, TC_FMT_BF16, 1>(hs, s, stream);
  return passes == 6 ? launch_phase_ws<3, 2, TC_FMT_BF16, 1>(hs, s, stream)
                     : launch_phase_ws<2, 3, TC_FMT_BF16, 1>(hs, s, stream);
}

int tc_engine_iteration(TcEngine* e, const float* xs, RootCtl* ctl, uint32_t* errbits,
                        RootParams prm, float* roots, int max_steps, float* first_init_scratch,
                        bool do_init, cudaStream_t stream) {
  auto* hs = static_cast<TcHostState*>(e->host_state);
  hs->prm.ctl = ctl;
  hs->prm.errbits = errbits;
  PlaneStore ps;
  for (int pl = 0; pl < 3; ++pl) ps.plane[pl] = hs->prm.plane[pl];
  ps.fmt = hs->fmt;
  ps.buf_stride = hs->prm.buf_stride;
  ps.mat_elems = hs->prm.mat_stride;
  if (first_init_scratch && e->n >= 256) {
    const int strips = (e->n + 31) / 32;
    root_norm_strip_kernel<<<dim3(strips, e->batch), 256, 0, stream>>>(
        xs, ctl, e->n, strips, prm, first_init_scratch, e->batch);
    static const bool strip_only = [] {
      const char* v = getenv("PC_INIT_STRIP");
      return v && v[0] == '1';
    }();
    if (hs->fmt == TC_FMT_FP16S && e->n % 128 == 0 && ((uintptr_t)xs & 15) == 0 &&
        e->batch <= 65535 && !strip_only)
      root_init_tile_kernel<<<dim3(e->n / 64, e->n / 128, e->batch), 256, 0, stream>>>(
          xs, ctl, ps, e->batch, e->n, strips, prm, first_init_scratch);
    else
      root_init_strip_kernel<PlaneStore><<<dim3(strips, e->batch), 1024, 0, stream>>>(
          xs, ctl, ps, e->batch, e->n, strips, prm, first_init_scratch);
    count_launch(2);
  }
  if (do_init) {  // (re)initialise whoever needs it: first iteration, or a retry was reported
    root_init_kernel<PlaneStore><<<e->batch, 1024, 0, stream>>>(xs, ctl, ps, e->batch, e->n, prm,
                                                               roots);
    count_launch(1);
  }
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  if (gemm_timing_enabled() && max_steps > 0) {
    cudaEventCreate(&ev0); cudaEventCreate(&ev1);
    cudaEventRecord(ev0, stream);
  }
  for (int s = 0; s < max_steps; ++s) {
    int rc;
    rc = launch_phase(hs, e->passes, s, stream);
    if (rc != PC_OK) return rc;
  }
  if (ev0) { cudaEventRecord(ev1, stream); gemm_timing_record(ev0, ev1); }
  count_launch(max_steps);
  gemm_count(max_steps);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

// Final gather of the scaled-fp16 engine (root_final_kernel's arithmetic, by plane tile): a
// thread turns 8 consecutive elements of a tile row -- one 16-byte load per plane -- into two
// 16-byte stores of the fp32 root (the generic kernel moves 2 + 2 + 4 bytes per thread and step
// and divides a 64-bit index per element).
__global__ void __launch_bounds__(256)
root_final_tile_kernel(const RootCtl* __restrict__ ctl, PlaneStore ps, int n,
                       float* __restrict__ roots, float* __restrict__ metrics) {
  const int b = blockIdx.y;
  const RootCtl& c = ctl[b];
  float* out = roots + (size_t)b * n * n;
  if (c.result_h != -1) {  // -1: already written by the n == 1 closed form
    const bool zero = (c.result_h == -2) || (c.pad == 0) || !c.done;  // DS:930-937
    const int tiles_j = n >> 6, tiles = (n >> 7) * tiles_j;
    const float hmul = c.hmul;
    const size_t mat = zero ? 0 : (size_t)(4 + c.result_h) * ps.buf_stride + (size_t)b * ps.mat_elems;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
      const int ti = t / tiles_j, tj = t - ti * tiles_j;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = threadIdx.x + 256 * u;
        const int r = e >> 3, g = (e & 7) * 8;
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = 0.f;
        if (!zero) {
          const size_t off = mat + (size_t)t * 8192 + (size_t)r * 64 + g;
          const uint4 w0 = *reinterpret_cast<const uint4*>(ps.plane[0] + off);
          const uint4 w1 = *reinterpret_cast<const uint4*>(ps.plane[1] + off);
          const uint32_t a0[4] = {w0.x, w0.y, w0.z, w0.w}, a1[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&a0[q]));
            const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&a1[q]));
            v[2 * q] = fmaf(x1.x, 1.0f / TC_FP16_SCALE, x0.x) * hmul;
            v[2 * q + 1] = fmaf(x1.y, 1.0f / TC_FP16_SCALE, x0.y) * hmul;
          }
        }
        float* o = out + (size_t)(ti * 128 + r) * n + tj * 64 + g;
        *reinterpret_cast<float4*>(o) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float* m = metrics + (size_t)b * PC_NUM_METRICS;
    m[PC_METRIC_ERROR] = c.pad == 0 ? 0.f : c.m_err;
    m[PC_METRIC_ITERS] = c.m_iters;
    m[PC_METRIC_ERROR_RATIO] = c.m_ratio;
    m[PC_METRIC_MAX_EV] = c.max_ev;
    m[PC_METRIC_RETRIES] = c.m_retries;
  }
}

int tc_engine_final(TcEngine* e, const RootCtl* ctl, float* roots, float* metrics,
                    cudaStream_t stream) {
  auto* hs = static_cast<TcHostState*>(e->host_state);
  PlaneStore ps;
  for (int pl = 0; pl < 3; ++pl) ps.plane[pl] = hs->prm.plane[pl];
  ps.fmt = hs->fmt;
  ps.buf_stride = hs->prm.buf_stride;
  ps.mat_elems = hs->prm.mat_stride;
  if (hs->fmt == TC_FMT_FP16S && e->n % 128 == 0 && ((uintptr_t)roots & 15) == 0 &&
      e->batch <= 65535) {
    const int tiles = (e->n / 128) * (e->n / 64);
    root_final_tile_kernel<<<dim3((unsigned)std::min(tiles, 256), e->batch), 256, 0, stream>>>(
        ctl, ps, e->n, roots, metrics);
  } else {
    dim3 fgrid((unsigned)std::min<size_t>(((size_t)e->n * e->n + 255) / 256, 64), e->batch);
    root_final_kernel<PlaneStore><<<fgrid, 256, 0, stream>>>(ctl, ps, e->n, roots, metrics);
  }
  PC_CUDA_CHECK(cudaGetLastError());
  delete hs;
  e->host_state = nullptr;
  return PC_OK;
}

// ---------------------------------------------------------------------------
// debug / test hook: C = A * B^T on the tensor-core engine (one program step)
// ---------------------------------------------------------------------------
__global__ void tc_debug_fill_kernel(const float* a, const float* b, PlaneStore ps, int n) {
  const int bt = blockIdx.y;
  const size_t nn = (size_t)n * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn;
       e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e - (size_t)i * n);
    ps.store(0, bt, i, j, n, a[(size_t)bt * nn + e]);  // M   <- A
    ps.store(2, bt, i, j, n, b[(size_t)bt * nn + e]);  // M_i <- B
    ps.store(4, bt, i, j, n, 0.f);                     // H   <- 0
  }
}
__global__ void tc_debug_read_kernel(PlaneStore ps, int phys, int n, float* c) {
  const int bt = blockIdx.y;
  const size_t nn = (size_t)n * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < nn;
       e += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(e / n), j = (int)(e - (size_t)i * n);
    c[(size_t)bt * nn + e] = ps.load(phys, bt, i, j, n);
  }
}
__global__ void tc_debug_ctl_kernel(RootCtl* ctl, int batch, int n, uint32_t* errbits) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  RootCtl c;
  memset(&c, 0, sizeof(c));
  c.p = 1; c.pad = n; c.active = 1;
  ctl[b] = c;
  errbits[b] = 0;
}

int tc_debug_gemm(const float* a, const float* b, float* c, int batch, int n, int passes,
                  void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  PC_REQUIRE(n % TC_BM == 0 && n >= TC_BM, "n must be a multiple of 128");
  if (!tc_engine_available()) {
    set_error("tcgen05 engine requested but device is not sm_100");
    return PC_ERR_UNSUPPORTED;
  }
  const int fmt = passes < 0 ? TC_FMT_FP16S : TC_FMT_BF16;
  const size_t need = tc_engine_bytes(batch, n, fmt == TC_FMT_FP16S ? 2 : 3) + 4096 +
                      sizeof(RootCtl) * batch + 4 * batch;
  if (workspace_bytes < need) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, need);
    return PC_ERR_WORKSPACE;
  }
  char* w = reinterpret_cast<char*>(align_up((size_t)workspace, 256));
  RootCtl* ctl = reinterpret_cast<RootCtl*>(w); w += align_up(sizeof(RootCtl) * batch, 256);
  uint32_t* errbits = reinterpret_cast<uint32_t*>(w); w += align_up(4 * batch, 256);
  TcEngine e;
  int rc = tc_engine_init(&e, w, batch, n, passes < 0 ? 3 : passes, fmt, stream);
  if (rc != PC_OK) return rc;
  auto* hs = static_cast<TcHostState*>(e.host_state);
  // private one-step program table: p = 1 -> Q0 = M * M_i^T
  Program hp[kMaxP + 1];
  memset(hp, 0, sizeof(hp));
  hp[1].nsteps = 1;
  hp[1].steps[0] = Step{LB_Q0, LB_M, LB_MI, 0};
  Program* dprog = nullptr;
  PC_CUDA_CHECK(cudaMallocAsync(&dprog, sizeof(hp), stream));
  PC_CUDA_CHECK(cudaMemcpyAsync(dprog, hp, sizeof(hp), cudaMemcpyHostToDevice, stream));
  PC_CUDA_CHECK(cudaStreamSynchronize(stream));
  hs->progs_dev = dprog;
  hs->prm.ctl = ctl;
  hs->prm.errbits = errbits;
  PlaneStore ps;
  for (int pl = 0; pl < 3; ++pl) ps.plane[pl] = hs->prm.plane[pl];
  ps.fmt = hs->fmt;
  ps.buf_stride = hs->prm.buf_stride;
  ps.mat_elems = hs->prm.mat_stride;
  tc_debug_ctl_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(ctl, batch, n, errbits);
  dim3 g(64, batch);
  tc_debug_fill_kernel<<<g, 256, 0, stream>>>(a, b, ps, n);
  rc = launch_phase(hs, passes < 0 ? 3 : passes, 0, stream);
  if (rc == PC_OK) tc_debug_read_kernel<<<g, 256, 0, stream>>>(ps, LB_Q0, n, c);
  cudaFreeAsync(dprog, stream);
  delete hs;
  PC_CUDA_CHECK(cudaGetLastError());
  return rc;
}


// ===========================================================================
// tcgen05 grouped GEMM on fp32 operands:  C = alpha * A B^T-view + beta * C_in
// (the statistics update L <- b2 L + (1-b2) G G^T, DS:1440-1470, and the preconditioner
// application, DS:1676-1708, for blocks whose output sizes are multiples of 128).
//   1. tc_pack_kernel: every operand view (strided, possibly transposed -- the same
//      two-level addressing as pc_gemm_desc) is converted ONCE into scaled-fp16 plane tiles
//      (128 x 64, contiguous): X ~ (X0 + 2^-11 X1) / s with a per-TILE power-of-two s that
//      brings the tile's max|X| to ~2^10 (block floating point), so gradients of any
//      magnitude keep 22 mantissa bits (fp16 alone would flush 1e-6 gradients).
//   2. tc_ggemm_kernel: the Newton engine's pipeline (TMA producer, single-thread
//      tcgen05.mma issuer with the scale-input-d three-pass scheme, 64-column K-chunks
//      summed in fp32 registers, TMEM output stages) with an fp32 epilogue:
//      every K-chunk is multiplied by 1 / (sA sB) of its two tiles as it is added, then
//      C = alpha * acc + beta * C_in.  Symmetric products (A == B: the Gram
//      update) compute lower tiles only and write the mirror from the same registers.
// ===========================================================================
struct TcGgItem {
  const float* a; const float* b; const float* c_in; float* c;
  int64_t a_sio, a_si, a_sko, a_ski, b_sj, b_sko, b_ski, c_sio, c_sii;
  int a_iinner, a_kinner, b_kinner, c_iinner;
  int m, n, k, kblocks;
  int a_tile0, b_tile0;  // first packed tile of each operand (b_tile0 == a_tile0 if symmetric)
  int symmetric;
  float alpha, beta;
  const float* beta_dev;  // device scalar replacing beta (pc_gemm_desc.beta_dev)
  // optional fused (de)quantisation of a square symmetric C (QuantizedValue, QU:49-113):
  // C_in = q_in * bucket_in[col] + diag_in on the diagonal; colmax receives the bit
  // patterns of max |off-diagonal| per column of the RESULT for the requantisation
  const void* q_in; const float* diag_in; const float* bucket_in; uint32_t* colmax;
  int qdtype;
};
struct TcGgWork { int z, tm, tn, pad; };
struct TcGgOperand {  // one packed operand: view + where its tiles go
  const float* base;
  int64_t s_io, s_i, s_ko, s_ki;
  int i_inner, k_inner, rows, k, kblocks, tile0;
  int is_b, reserved;  // second operand of a general product (its pack can be skipped, see below)
  // quantised source (pc_gemm_quant.b_q): element offset `off` of the view addresses q, the
  // value is q[off] * bucket[off % ld] (+ diag[off / ld] on the diagonal)
  const void* q; const float* q_diag; const float* q_bucket; int q_ld, q_dtype;
};

__device__ __forceinline__ float tc_gg_view(const TcGgOperand& o, int i, int kk) {
  if (i >= o.rows || kk >= o.k) return 0.f;
  const int io = i / o.i_inner, ii = i - io * o.i_inner;
  const int ko = kk / o.k_inner, ki = kk - ko * o.k_inner;
  return __ldg(o.base + io * o.s_io + ii * o.s_i + ko * o.s_ko + ki * o.s_ki);
}

// Power-of-two scale that brings max|x| of a tile to [2^9, 2^10): fp16 then holds the tile's
// largest entries with full precision and the 2^11-scaled residual stays in range.
__device__ __forceinline__ float tc_gg_scale(uint32_t maxbits) {
  const float mx = __uint_as_float(maxbits);
  if (!(mx > 0.f) || !(mx < 3.0e38f)) return 1.0f;  // zero, inf or NaN tile: leave as is
  int e = 0;
  frexpf(mx, &e);  // mx in [2^(e-1), 2^e)
  int sh = 10 - e;
  sh = sh > 120 ? 120 : (sh < -120 ? -120 : sh);
  return ldexpf(1.0f, sh);
}

// Element (i, kk) of an operand view, any addressing mode (zero outside the view); the
// quantised form is to_float of a square QuantizedValue on the fly (QU:97-113).
__device__ __forceinline__ float tc_pack_elem(const TcGgOperand& o, bool simple, int i, int kk) {
  if (i >= o.rows || kk >= o.k) return 0.f;
  if (o.q) {
    // (a square matrix of up to 46340 rows: the offset fits 32 bits)
    const uint32_t off = (uint32_t)((int64_t)i * o.s_i + (int64_t)kk * o.s_ki);
    const uint32_t qr = off / (uint32_t)o.q_ld, qc = off - qr * (uint32_t)o.q_ld;
    const float qv = o.q_dtype == PC_QDTYPE_INT16
                         ? (float)reinterpret_cast<const int16_t*>(o.q)[off]
                         : (float)reinterpret_cast<const int8_t*>(o.q)[off];
    float v = qv * __ldg(o.q_bucket + qc);
    if (qr == qc) v += __ldg(o.q_diag + qr);
    return v;
  }
  if (simple) return __ldg(o.base + (int64_t)i * o.s_i + (int64_t)kk * o.s_ki);
  return tc_gg_view(o, i, kk);
}

// Four consecutive elements along the view's fast axis at element offset `off` of the source
// as one 16 / 8 / 4-byte load (fp32 / int16 / int8).  Quantised sources: the four are columns
// col .. col+3 of one row of q; diag_at = position of the diagonal element among them or -1.
// kSrc: 0 fp32, 1 int16, 2 int8 (a template parameter: no branch between the loads of a loop).
template <int kSrc>
__device__ __forceinline__ float4 tc_pack_load4(const TcGgOperand& o, int64_t off, bool ok,
                                                int col, int diag_at, int diag_idx) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kSrc == 0) {
    const float4* p = reinterpret_cast<const float4*>(o.base + off);
    if (ok) v = __ldg(p);
    return v;
  }
  if (kSrc == 1) {
    const uint2* p = reinterpret_cast<const uint2*>(reinterpret_cast<const int16_t*>(o.q) + off);
    uint2 w = make_uint2(0u, 0u);
    if (ok) w = __ldg(p);
    v = make_float4((float)(int16_t)(w.x & 0xffffu), (float)(int16_t)(w.x >> 16),
                    (float)(int16_t)(w.y & 0xffffu), (float)(int16_t)(w.y >> 16));
  } else {
    const uint32_t* p = reinterpret_cast<const uint32_t*>(reinterpret_cast<const int8_t*>(o.q) + off);
    uint32_t w = 0u;
    if (ok) w = __ldg(p);
    v = make_float4((float)(int8_t)(w & 0xffu), (float)(int8_t)((w >> 8) & 0xffu),
                    (float)(int8_t)((w >> 16) & 0xffu), (float)(int8_t)(w >> 24));
  }
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  if (ok) b = __ldg(reinterpret_cast<const float4*>(o.q_bucket + col));
  v.x *= b.x; v.y *= b.y; v.z *= b.z; v.w *= b.w;
  if (ok && diag_at >= 0) {
    const float dg = __ldg(o.q_diag + diag_idx);
    if (diag_at == 0) v.x += dg;
    else if (diag_at == 1) v.y += dg;
    else if (diag_at == 2) v.z += dg;
    else v.w += dg;
  }
  return v;
}

// two fp32 values -> the fp16 pair of plane 0 and the 2^11-scaled fp16 residual pair of plane 1
__device__ __forceinline__ void tc_pack_split2(float v0, float v1, uint32_t& w0, uint32_t& w1) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w0) : "f"(v1), "f"(v0));
  const float2 f2 = __half22float2(*reinterpret_cast<const __half2*>(&w0));
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w1)
      : "f"((v1 - f2.y) * TC_FP16_SCALE), "f"((v0 - f2.x) * TC_FP16_SCALE));
}

// Vector staging of one tile (modes 1 / 2 of tc_pack_kernel; the mode is a template parameter so
// that the four loads of a half sit in one basic block and are all in flight together).
template <int kMode, int kSrc>
__device__ __forceinline__ uint32_t tc_pack_stage_vec(const TcGgOperand& o, bool simple, int tr,
                                                      int kb, float (&tile)[64][129], int lane,
                                                      int warp) {
  uint32_t mx = 0;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    float4 v[4];
    bool ok[4];
    int ri[4], ci[4];  // tile coordinates of element 0 of each group of four
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      const int s8 = half * 4 + s4;
      if (kMode == 1) {
        // four consecutive k per load; a warp instruction covers 4 rows x 32 k (128 B runs)
        ri[s4] = 16 * warp + 4 * (s8 >> 1) + (lane >> 3);
        ci[s4] = 4 * (lane & 7) + 32 * (s8 & 1);
        const int i = tr * 128 + ri[s4], kk = kb * 64 + ci[s4];
        ok[s4] = i < o.rows && kk + 3 < o.k;
        // q row = i, columns kk .. kk+3; the diagonal is element i - kk of the four
        const int da = (kSrc != 0 && i >= kk && i < kk + 4) ? i - kk : -1;
        v[s4] = tc_pack_load4<kSrc>(o, (int64_t)i * o.s_i + kk, ok[s4], kk, da, i);
      } else {
        // four consecutive rows per load; a warp owns 8 k-columns, a warp instruction covers
        // 8 k x (4 groups of 4 rows, 32 B apart)
        ci[s4] = (lane & 7) + 8 * warp;
        ri[s4] = 4 * (2 * (lane >> 3) + (s8 & 1) + 8 * (s8 >> 1));
        const int i = tr * 128 + ri[s4], kk = kb * 64 + ci[s4];
        ok[s4] = i + 3 < o.rows && kk < o.k;
        // q row = kk, columns i .. i+3; the diagonal is element kk - i of the four
        const int da = (kSrc != 0 && kk >= i && kk < i + 4) ? kk - i : -1;
        v[s4] = tc_pack_load4<kSrc>(o, (int64_t)kk * o.s_ki + i, ok[s4], i, da, kk);
      }
    }
#pragma unroll
    for (int s4 = 0; s4 < 4; ++s4) {
      const int i = tr * 128 + ri[s4], kk = kb * 64 + ci[s4];
      float x[4] = {v[s4].x, v[s4].y, v[s4].z, v[s4].w};
      if (!ok[s4] && i < o.rows && kk < o.k) {  // the view's last, partial group of four
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4)
          x[q4] = kMode == 1 ? tc_pack_elem(o, simple, i, kk + q4)
                             : tc_pack_elem(o, simple, i + q4, kk);
      }
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4) {
        if (kMode == 1) tile[ci[s4] + q4][ri[s4]] = x[q4];
        else tile[ci[s4]][ri[s4] + q4] = x[q4];
        const uint32_t ab = absbits(x[q4]);
        mx = ab > mx ? ab : mx;
      }
    }
  }
  return mx;
}

// One CTA per 128 x 64 tile of an operand view: stage the tile in shared memory, reduce its
// max |x|, then write the two fp16 planes scaled by the TILE's power of two (block floating
// point; the GEMM multiplies each K-chunk by 1 / (sA sB) when it adds it in fp32).
// One-level views whose fast axis is k (mode 1) or i (mode 2) with 16-byte aligned rows are read
// with one vector load per four elements through lane mappings whose transposing shared-memory
// stores are bank-conflict free; everything else (two-level views, unaligned sources) goes
// element by element (mode 0).  The planes leave as 16-byte stores (8 k of one row per thread).
__global__ void __launch_bounds__(256, 4)
tc_pack_kernel(const TcGgOperand* __restrict__ ops, float* __restrict__ inv_scale,
               uint16_t* __restrict__ plane0, uint16_t* __restrict__ plane1, int skip_b) {
  __shared__ float tile[64][129];
  __shared__ uint32_t red[32];
  const TcGgOperand o = ops[blockIdx.y];
  if (skip_b && o.is_b) return;  // its planes from the previous call are still valid
  const int tiles = ((o.rows + 127) / 128) * o.kblocks;  // rows past the view are zero-filled
  const bool kfast = o.s_ki == 1;
  const bool simple = o.i_inner >= o.rows && o.k_inner >= o.k;  // one-level addressing
  int mode = 0;
  if (simple) {
    const int64_t slow = kfast ? o.s_i : o.s_ki;
    bool aligned;
    if (!o.q) {
      aligned = (slow & 3) == 0 && (reinterpret_cast<uintptr_t>(o.base) & 15) == 0;
    } else {
      // off = i * s_i + kk * s_ki addresses q: the slow stride must be q's row stride
      const int es = o.q_dtype == PC_QDTYPE_INT16 ? 2 : 1;
      aligned = slow == o.q_ld && (o.q_ld & 3) == 0 &&
                (reinterpret_cast<uintptr_t>(o.q) & (4 * es - 1)) == 0 &&
                (reinterpret_cast<uintptr_t>(o.q_bucket) & 15) == 0;
    }
    if (aligned && kfast && o.s_i != 1) mode = 1;
    else if (aligned && o.s_i == 1 && !kfast) mode = 2;
  }
  const int src = !o.q ? 0 : (o.q_dtype == PC_QDTYPE_INT16 ? 1 : 2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
    const int tr = t / o.kblocks, kb = t - tr * o.kblocks;
    uint32_t mx = 0;
    __syncthreads();
    if (mode == 1) {
      mx = src == 0 ? tc_pack_stage_vec<1, 0>(o, simple, tr, kb, tile, lane, warp)
         : src == 1 ? tc_pack_stage_vec<1, 1>(o, simple, tr, kb, tile, lane, warp)
                    : tc_pack_stage_vec<1, 2>(o, simple, tr, kb, tile, lane, warp);
    } else if (mode == 2) {
      mx = src == 0 ? tc_pack_stage_vec<2, 0>(o, simple, tr, kb, tile, lane, warp)
         : src == 1 ? tc_pack_stage_vec<2, 1>(o, simple, tr, kb, tile, lane, warp)
                    : tc_pack_stage_vec<2, 2>(o, simple, tr, kb, tile, lane, warp);
    } else {
      for (int e = threadIdx.x; e < 128 * 64; e += blockDim.x) {  // coalesced along the fast axis
        const int r = kfast ? e >> 6 : e & 127, c = kfast ? e & 63 : e >> 7;
        const float v = tc_pack_elem(o, simple, tr * 128 + r, kb * 64 + c);
        tile[c][r] = v;
        const uint32_t ab = absbits(v);
        mx = ab > mx ? ab : mx;
      }
    }
    mx = block_max_u32(mx, red);  // (contains the barrier that publishes the tile)
    const float sc = tc_gg_scale(mx);
    if (threadIdx.x == 0) inv_scale[o.tile0 + t] = 1.0f / sc;
    const size_t base = (size_t)(o.tile0 + t) * 8192;
#pragma unroll
    for (int u = 0; u < 4; ++u) {  // eight k of one row per thread: one 16-byte store per plane
      const int e = threadIdx.x + 256 * u;
      const int r = e >> 3, c8 = (e & 7) * 8;
      uint32_t w0[4], w1[4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
        tc_pack_split2(tile[c8 + 2 * q][r] * sc, tile[c8 + 2 * q + 1][r] * sc, w0[q], w1[q]);
      *reinterpret_cast<uint4*>(plane0 + base + (size_t)r * 64 + c8) =
          make_uint4(w0[0], w0[1], w0[2], w0[3]);
      *reinterpret_cast<uint4*>(plane1 + base + (size_t)r * 64 + c8) =
          make_uint4(w1[0], w1[1], w1[2], w1[3]);
    }
  }
}

// loads packed tile `t` (16 KiB) of a plane: 3-D map [64, 128, tiles]
__device__ __forceinline__ void tma_load_packed_tile(uint32_t dst, const CUtensorMap* map,
                                                     uint32_t bar, int t) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(0), "r"(0), "r"(t)
      : "memory");
}

constexpr int TC_GG_THREADS = 512;
constexpr int TC_GG_STAGES = 3;

__global__ void __launch_bounds__(TC_GG_THREADS, 1)
tc_ggemm_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1,
                const TcGgItem* __restrict__ items, const TcGgWork* __restrict__ work,
                const float* __restrict__ inv_scale, int total_work) {
  constexpr int kStageBytes = 4 * TC_TILE_BYTES;  // A0 A1 B0 B1
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + TC_GG_STAGES * kStageBytes;
  auto full_bar = [&](int i) { return bar_base + 8u * i; };
  auto empty_bar = [&](int i) { return bar_base + 8u * (TC_GG_STAGES + i); };
  auto tfull_bar = [&](int i) { return bar_base + 8u * (2 * TC_GG_STAGES + i); };
  auto tempty_bar = [&](int i) { return bar_base + 8u * (2 * TC_GG_STAGES + 2 + i); };
  auto ofull_bar = [&](int i) { return bar_base + 8u * (2 * TC_GG_STAGES + 4 + i); };
  auto oempty_bar = [&](int i) { return bar_base + 8u * (2 * TC_GG_STAGES + 6 + i); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * TC_GG_STAGES + 8);
  uint32_t* tmem_slot_ptr =
      reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&map0);
    tma_prefetch_desc(&map1);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < TC_GG_STAGES; ++i) {
      mbar_init(full_bar(i), 1);
      mbar_init(empty_bar(i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(tfull_bar(i), 1);
      mbar_init(tempty_bar(i), 4);
      mbar_init(ofull_bar(i), 4);
      mbar_init(oempty_bar(i), 8);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp == 0 && elect_one()) {
      // ===================== TMA producer =====================
      int stage = 0;
      uint32_t phase = 0;
#pragma unroll 1
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const TcGgWork wk = work[w];
        const TcGgItem& it = items[wk.z];
        const int kblocks = it.kblocks;
        const int at = it.a_tile0 + wk.tm * kblocks, bt = it.b_tile0 + wk.tn * kblocks;
#pragma unroll 1
        for (int kb = 0; kb < kblocks; ++kb) {
          mbar_wait(empty_bar(stage), phase ^ 1);
          const uint32_t dst = smem_base + stage * kStageBytes;
          mbar_expect_tx(full_bar(stage), kStageBytes);
          tma_load_packed_tile(dst, &map0, full_bar(stage), at + kb);
          tma_load_packed_tile(dst + TC_TILE_BYTES, &map1, full_bar(stage), at + kb);
          tma_load_packed_tile(dst + 2 * TC_TILE_BYTES, &map0, full_bar(stage), bt + kb);
          tma_load_packed_tile(dst + 3 * TC_TILE_BYTES, &map1, full_bar(stage), bt + kb);
          if (++stage == TC_GG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1 && elect_one()) {
      // ===================== MMA issuer =====================
      int stage = 0;
      uint32_t phase = 0;
      int chunk = 0;
#pragma unroll 1
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const int kblocks = items[work[w].z].kblocks;
#pragma unroll 1
        for (int kb = 0; kb < kblocks; ++kb, ++chunk) {
          const int acc = chunk & 1;
          mbar_wait(tempty_bar(acc), ((chunk >> 1) & 1) ^ 1);
          mbar_wait(full_bar(stage), phase);
          tcgen05_fence_after();
          const uint32_t tmem_d = tmem_base + acc * TC_BN;
          const uint32_t a0 = smem_base + stage * kStageBytes;
          const uint32_t b0 = a0 + 2 * TC_TILE_BYTES;
          issue_fp16_cross<false>(tmem_d, a0, b0, kIdescF16M128N128, true);
          issue_fp16_main<false>(tmem_d, a0, b0, kIdescF16M128N128, true);
          umma_commit(empty_bar(stage));
          umma_commit(tfull_bar(acc));
          if (++stage == TC_GG_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp < 8) {
    // ============ warpgroup 1: chunk accumulation -> TMEM output stage ============
    asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    const int q = warp & 3;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int chunk = 0, tile = 0;
#pragma unroll 1
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const TcGgWork wk = work[w];
      const TcGgItem& it = items[wk.z];
      const int kblocks = it.kblocks;
      const float* sa = inv_scale + it.a_tile0 + wk.tm * kblocks;
      const float* sb = inv_scale + it.b_tile0 + wk.tn * kblocks;
      float sum[TC_BN];
#pragma unroll
      for (int i = 0; i < TC_BN; ++i) sum[i] = 0.f;
#pragma unroll 1
      for (int kb = 0; kb < kblocks; ++kb, ++chunk) {
        const float f = __ldg(sa + kb) * __ldg(sb + kb);  // 1 / (sA sB) of this K-chunk's tiles
        const int acc = chunk & 1;
        mbar_wait(tfull_bar(acc), (chunk >> 1) & 1);
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + lane_off + acc * TC_BN;
#pragma unroll
        for (int c = 0; c < TC_BN / 32; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i)
            sum[c * 32 + i] = fmaf(__uint_as_float(r[i]), f, sum[c * 32 + i]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
      const int o = tile & 1;
      mbar_wait(oempty_bar(o), ((tile >> 1) & 1) ^ 1);
      tcgen05_fence_after();
      const uint32_t oaddr = tmem_base + lane_off + 256 + o * TC_BN;
#pragma unroll
      for (int c = 0; c < TC_BN / 32; ++c) {
        uint32_t r[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(sum[c * 32 + i]);
        tmem_st_32x32(oaddr + c * 32, r);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ofull_bar(o));
      ++tile;
    }
  } else {
    // ============ warpgroups 2, 3: fp32 epilogue (two column halves of every tile) ============
    asm volatile("setmaxnreg.inc.sync.aligned.u32 136;");
    const int q = warp & 3;
    const int half = (warp >> 2) - 2;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    int tile = 0;
#pragma unroll 1
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const TcGgWork wk = work[w];
      const TcGgItem it = items[wk.z];
      const int o = tile & 1;
      mbar_wait(ofull_bar(o), (tile >> 1) & 1);
      tcgen05_fence_after();
      const float alpha = it.alpha;
      const float beta = it.beta_dev ? __ldg(it.beta_dev) : it.beta;
      const int row = wk.tm * TC_BM + q * 32 + lane;
      const bool row_ok = row < it.m;  // ragged edge tiles: rows / columns past the block are masked
      const int io = row / it.c_iinner, ii = row - io * it.c_iinner;
      const int64_t rowoff = io * it.c_sio + ii * it.c_sii;
      const bool diag_tile = it.symmetric && wk.tm == wk.tn;
#pragma unroll 1
      for (int c = 2 * half; c < 2 * half + 2; ++c) {
        uint32_t r[32];
        tmem_ld_32x32(tmem_base + lane_off + 256 + o * TC_BN + c * 32, r);
        tmem_ld_wait();
        if (c == 2 * half + 1) {  // last TMEM read of this tile by this warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(oempty_bar(o));
        }
        const int col0 = wk.tn * TC_BN + c * 32;
        if (diag_tile && c > q) continue;  // strictly upper sub-block: written by its mirror
        if (col0 >= it.n) continue;        // sub-block past the last column (warp-uniform)
        float x[32];
        const bool diag_sub = diag_tile && c == q;
        if (it.q_in) {
          // to_float fused into the load (QU:97-113): q * bucket[col] (+ diag on the diagonal)
          const size_t qoff = (size_t)row * it.n + col0;
          float oldv[32];
          if (it.qdtype == PC_QDTYPE_INT16) {
            const int16_t* qp = reinterpret_cast<const int16_t*>(it.q_in) + qoff;
#pragma unroll
            for (int i = 0; i < 32; i += 8) {
              const uint4 w = *reinterpret_cast<const uint4*>(qp + i);
              const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int h = 0; h < 4; ++h) {
                oldv[i + 2 * h] = (float)(int16_t)(ws[h] & 0xffffu);
                oldv[i + 2 * h + 1] = (float)(int16_t)(ws[h] >> 16);
              }
            }
          } else {
            const int8_t* qp = reinterpret_cast<const int8_t*>(it.q_in) + qoff;
#pragma unroll
            for (int i = 0; i < 32; i += 16) {
              const uint4 w = *reinterpret_cast<const uint4*>(qp + i);
              const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
              for (int h = 0; h < 4; ++h)
#pragma unroll
                for (int bb = 0; bb < 4; ++bb)
                  oldv[i + 4 * h + bb] = (float)(int8_t)((ws[h] >> (8 * bb)) & 0xffu);
            }
          }
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float o = oldv[i] * __ldg(it.bucket_in + col0 + i);
            if (col0 + i == row) o += __ldg(it.diag_in + row);
            x[i] = fmaf(beta, o, alpha * __uint_as_float(r[i]));
          }
        } else {
          const float* cin = (it.c_in && row_ok) ? it.c_in + rowoff + col0 : nullptr;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            float4 old = make_float4(0.f, 0.f, 0.f, 0.f);
            if (cin && col0 + i < it.n) old = *reinterpret_cast<const float4*>(cin + i);
            x[i] = fmaf(beta, old.x, alpha * __uint_as_float(r[i]));
            x[i + 1] = fmaf(beta, old.y, alpha * __uint_as_float(r[i + 1]));
            x[i + 2] = fmaf(beta, old.z, alpha * __uint_as_float(r[i + 2]));
            x[i + 3] = fmaf(beta, old.w, alpha * __uint_as_float(r[i + 3]));
          }
        }
        if (it.colmax) {
          // column max of |off-diagonal| for the requantisation (QU:86), fused into the
          // epilogue: an element (row, col) of the lower triangle also stands for (col, row)
          uint32_t rmax = 0, mine = 0;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const bool owned = diag_sub ? i < lane : true;  // strictly lower part of a diagonal block
            const uint32_t ab = (owned && col0 + i != row) ? absbits(x[i]) : 0u;
            rmax = ab > rmax ? ab : rmax;
            const uint32_t cm = __reduce_max_sync(0xffffffffu, ab);
            if (lane == i) mine = cm;
          }
          if (rmax) atomicMax(it.colmax + row, rmax);
          if (mine) atomicMax(it.colmax + col0 + lane, mine);
        }
        float* crow = it.c + rowoff + col0;
        if (!diag_sub) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            if (row_ok && col0 + i < it.n)  // (n % 4 == 0: a group of four is in or out as a whole)
              *reinterpret_cast<float4*>(crow + i) = make_float4(x[i], x[i + 1], x[i + 2], x[i + 3]);
        } else {  // lower triangle of the diagonal sub-block only (its mirror fills the rest)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i <= lane && row_ok) crow[i] = x[i];
        }
        if (it.symmetric) {
          // mirror: element (row, col0 + i) -> (col0 + i, row); lanes write consecutive floats
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int mr = col0 + i;
            if (diag_sub && i >= lane) continue;  // strictly lower elements only
            if (!row_ok || mr >= it.m) continue;
            const int mio = mr / it.c_iinner, mii = mr - mio * it.c_iinner;
            it.c[mio * it.c_sio + mii * it.c_sii + row] = x[i];
          }
        }
      }
      ++tile;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

// ---- host ----
static bool tc_gg_supported(const pc_gemm_desc& d) {
  // (any m; n % 4 == 0: the epilogue moves groups of four columns; edge tiles are zero-filled
  // by the pack and masked by the epilogue)
  return d.m > 0 && d.n > 0 && d.k > 0 && d.n % 4 == 0 &&
         (reinterpret_cast<uintptr_t>(d.c) % 16 == 0) && d.c_sii % 4 == 0 && d.c_sio % 4 == 0 &&
         (!d.c_in || (reinterpret_cast<uintptr_t>(d.c_in) % 16 == 0));
}
static bool tc_gg_symmetric(const pc_gemm_desc& d) {
  return d.a == d.b && d.m == d.n && d.a_sio == 0 && d.a_iinner >= d.m && d.a_si == d.b_sj &&
         d.a_sko == d.b_sko && d.a_ski == d.b_ski && d.a_kinner == d.b_kinner;
}

struct TcGgPlan {
  std::vector<TcGgItem> items;
  std::vector<TcGgOperand> ops;
  std::vector<TcGgWork> work;
  int total_tiles = 0;
};

// Two-level addressing whose outer stride continues the inner one (an unpartitioned contiguous
// tensor described axis by axis) is one-level addressing: the pack then takes its div / mod-free
// path.
static void tc_gg_collapse(TcGgOperand* o) {
  if (o->k_inner < o->k && o->s_ko == (int64_t)o->k_inner * o->s_ki) o->k_inner = 0x7fffffff;
  if (o->i_inner < o->rows && o->s_io == (int64_t)o->i_inner * o->s_i) o->i_inner = 0x7fffffff;
}

static void tc_gg_plan(const pc_gemm_desc* descs, const pc_gemm_quant* quant, int count,
                       TcGgPlan* pl) {
  for (int z = 0; z < count; ++z) {
    const pc_gemm_desc& d = descs[z];
    TcGgItem it{};
    if (quant) {
      it.q_in = quant[z].q_in; it.diag_in = quant[z].diag_in; it.bucket_in = quant[z].bucket_in;
      it.colmax = quant[z].colmax_out; it.qdtype = quant[z].qdtype;
    }
    it.a = d.a; it.b = d.b; it.c_in = d.c_in; it.c = d.c;
    it.c_sio = d.c_sio; it.c_sii = d.c_sii; it.c_iinner = d.c_iinner > 0 ? d.c_iinner : d.m;
    it.m = d.m; it.n = d.n; it.k = d.k; it.kblocks = (d.k + TC_BK - 1) / TC_BK;
    it.alpha = d.alpha; it.beta = d.beta; it.beta_dev = d.beta_dev;
    it.symmetric = tc_gg_symmetric(d) ? 1 : 0;
    TcGgOperand oa{};
    oa.base = d.a; oa.s_io = d.a_sio; oa.s_i = d.a_si; oa.s_ko = d.a_sko; oa.s_ki = d.a_ski;
    oa.i_inner = d.a_iinner > 0 ? d.a_iinner : d.m; oa.k_inner = d.a_kinner > 0 ? d.a_kinner : d.k;
    oa.rows = d.m; oa.k = d.k; oa.kblocks = it.kblocks; oa.tile0 = pl->total_tiles;
    tc_gg_collapse(&oa);
    it.a_tile0 = oa.tile0;
    pl->total_tiles += ((d.m + 127) / 128) * it.kblocks;
    pl->ops.push_back(oa);
    if (it.symmetric) {
      it.b_tile0 = it.a_tile0;
    } else {
      TcGgOperand ob{};
      ob.base = d.b; ob.s_io = 0; ob.s_i = d.b_sj; ob.s_ko = d.b_sko; ob.s_ki = d.b_ski;
      ob.i_inner = d.n; ob.k_inner = d.b_kinner > 0 ? d.b_kinner : d.k;
      ob.rows = d.n; ob.k = d.k; ob.kblocks = it.kblocks; ob.tile0 = pl->total_tiles;
      ob.is_b = 1;
      tc_gg_collapse(&ob);
      if (quant && quant[z].b_q) {
        ob.q = quant[z].b_q; ob.q_diag = quant[z].b_diag; ob.q_bucket = quant[z].b_bucket;
        ob.q_ld = quant[z].b_ld; ob.q_dtype = quant[z].b_qdtype;
      }
      it.b_tile0 = ob.tile0;
      pl->total_tiles += ((d.n + 127) / 128) * it.kblocks;
      pl->ops.push_back(ob);
    }
    pl->items.push_back(it);
    for (int tm = 0; tm < (d.m + 127) / 128; ++tm)
      for (int tn = 0; tn < (d.n + 127) / 128; ++tn)
        if (!it.symmetric || tn <= tm) pl->work.push_back(TcGgWork{z, tm, tn, 0});
  }
}

static size_t tc_gg_bytes(const TcGgPlan& pl) {
  size_t s = 0;
  s += align_up(pl.items.size() * sizeof(TcGgItem), 256);
  s += align_up(pl.ops.size() * sizeof(TcGgOperand), 256);
  s += align_up(pl.work.size() * sizeof(TcGgWork), 256);
  s += align_up((size_t)pl.total_tiles * sizeof(float), 256);
  s += 1024 + 2 * align_up((size_t)pl.total_tiles * TC_TILE_BYTES, 1024);
  return s + 1024;
}

size_t tc_grouped_gemm_workspace_bytes(const pc_gemm_desc* descs, int count) {
  TcGgPlan pl;
  tc_gg_plan(descs, nullptr, count, &pl);
  return tc_gg_bytes(pl);
}

int tc_grouped_gemm(const pc_gemm_desc* descs, const pc_gemm_quant* quant, int count,
                    void* workspace, size_t workspace_bytes, int reuse_plan, cudaStream_t stream) {
  for (int z = 0; z < count; ++z) {
    PC_REQUIRE(tc_gg_supported(descs[z]),
               "descriptor %d is not eligible for the tcgen05 grouped GEMM (n %% 4, alignment)", z);
    if (quant && (quant[z].q_in || quant[z].colmax_out)) {
      PC_REQUIRE(tc_gg_symmetric(descs[z]) && descs[z].c_sii == descs[z].n && descs[z].c_sio == 0 &&
                     descs[z].m % 128 == 0,
                 "descriptor %d: fused (de)quantisation needs a symmetric product into a contiguous "
                 "square matrix whose size is a multiple of 128", z);
      PC_REQUIRE(!quant[z].q_in || ((quant[z].qdtype == PC_QDTYPE_INT16 ||
                                    quant[z].qdtype == PC_QDTYPE_INT8) &&
                                   quant[z].diag_in && quant[z].bucket_in),
                 "descriptor %d: quantised C_in needs int16 / int8 data, diagonal and buckets", z);
    }
  }
  for (int z = 0; z < count && quant; ++z) {
    if (!quant[z].b_q) continue;
    const pc_gemm_desc& d = descs[z];
    PC_REQUIRE(!tc_gg_symmetric(d) && quant[z].b_diag && quant[z].b_bucket && quant[z].b_ld > 0 &&
                   (quant[z].b_qdtype == PC_QDTYPE_INT16 || quant[z].b_qdtype == PC_QDTYPE_INT8) &&
                   (d.b_kinner <= 0 || d.b_kinner >= d.k),
               "descriptor %d: a quantised B operand needs diagonal, buckets, its row stride and "
               "one-level addressing", z);
  }
  if (!tc_engine_available()) {
    set_error("tcgen05 grouped GEMM requested but device is not sm_100");
    return PC_ERR_UNSUPPORTED;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return PC_ERR_UNSUPPORTED;
  }
  TcGgPlan pl;
  tc_gg_plan(descs, quant, count, &pl);
  if (workspace_bytes < tc_gg_bytes(pl)) {
    set_error("workspace too small: %zu < %zu", workspace_bytes, tc_gg_bytes(pl));
    return PC_ERR_WORKSPACE;
  }
  char* w = reinterpret_cast<char*>(align_up((size_t)workspace, 256));
  TcGgItem* d_items = reinterpret_cast<TcGgItem*>(w); w += align_up(pl.items.size() * sizeof(TcGgItem), 256);
  TcGgOperand* d_ops = reinterpret_cast<TcGgOperand*>(w); w += align_up(pl.ops.size() * sizeof(TcGgOperand), 256);
  TcGgWork* d_work = reinterpret_cast<TcGgWork*>(w); w += align_up(pl.work.size() * sizeof(TcGgWork), 256);
  float* d_inv = reinterpret_cast<float*>(w); w += align_up((size_t)pl.total_tiles * sizeof(float), 256);
  uint16_t* plane0 = reinterpret_cast<uint16_t*>(align_up((size_t)w, 1024));
  uint16_t* plane1 = plane0 + align_up((size_t)pl.total_tiles * TC_TILE_BYTES, 1024) / 2;
  // reuse_plan bit 0: the uploaded plan is still in the workspace; bit 1 (with bit 0): so are the
  // packed planes of every B operand (the caller guarantees the B views did not change)
  if (!(reuse_plan & 1)) {
    // the plan lives in heap vectors: wait for the copies before they go out of scope
    PC_CUDA_CHECK(cudaMemcpyAsync(d_items, pl.items.data(), pl.items.size() * sizeof(TcGgItem),
                                  cudaMemcpyHostToDevice, stream));
    PC_CUDA_CHECK(cudaMemcpyAsync(d_ops, pl.ops.data(), pl.ops.size() * sizeof(TcGgOperand),
                                  cudaMemcpyHostToDevice, stream));
    PC_CUDA_CHECK(cudaMemcpyAsync(d_work, pl.work.data(), pl.work.size() * sizeof(TcGgWork),
                                  cudaMemcpyHostToDevice, stream));
    PC_CUDA_CHECK(cudaStreamSynchronize(stream));
  }
  CUtensorMap maps[2];
  for (int plx = 0; plx < 2; ++plx) {
    cuuint64_t dims[3] = {64, 128, (cuuint64_t)pl.total_tiles};
    cuuint64_t strides[2] = {64 * 2, 8192 * 2};
    cuuint32_t box[3] = {64, 128, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&maps[plx], CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, plx ? plane1 : plane0, dims,
                     strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("cuTensorMapEncodeTiled failed with %d", (int)r);
      return PC_ERR_CUDA;
    }
  }
  const int nops = (int)pl.ops.size();
  int max_tiles = 1;
  for (const TcGgOperand& o : pl.ops)
    max_tiles = std::max(max_tiles, ((o.rows + 127) / 128) * o.kblocks);
  tc_pack_kernel<<<dim3((unsigned)std::min(max_tiles, 1024), nops), 256, 0, stream>>>(
      d_ops, d_inv, plane0, plane1, (reuse_plan & 3) == 3 ? 1 : 0);
  constexpr size_t smem = (size_t)TC_GG_STAGES * 4 * TC_TILE_BYTES + 1024 + 1024;
  static bool configured = false;
  if (!configured) {
    PC_CUDA_CHECK(cudaFuncSetAttribute(tc_ggemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)smem));
    configured = true;
  }
  int sms = 148, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int total_work = (int)pl.work.size();
  const int grid = total_work < sms ? total_work : sms;
  tc_ggemm_kernel<<<grid, TC_GG_THREADS, smem, stream>>>(maps[0], maps[1], d_items, d_work, d_inv,
                                                        total_work);
  count_launch(2);
  PC_CUDA_CHECK(cudaGetLastError());
  return PC_OK;
}

}  // namespace pc

extern "C" int pc_debug_tc_gemm(const float* a, const float* b, float* c, int batch, int n,
                                int passes, void* workspace, size_t workspace_bytes,
                                void* stream) {
  PC_REQUIRE(a && b && c && workspace && batch > 0, "bad arguments");
  PC_REQUIRE(passes == 6 || passes == 3 || passes == -3,
             "passes must be 6, 3 (bf16 planes) or -3 (scaled fp16 planes)");
  return pc::tc_debug_gemm(a, b, c, batch, n, passes, workspace, workspace_bytes,
                           (cudaStream_t)stream);
}

extern "C" size_t pc_grouped_gemm_tc_workspace_bytes(const pc_gemm_desc* descs_host, int count) {
  if (!descs_host || count <= 0) return 0;
  return pc::tc_grouped_gemm_workspace_bytes(descs_host, count);
}

extern "C" int pc_grouped_gemm_tc(const pc_gemm_desc* descs_host, int count, void* workspace,
                                  size_t workspace_bytes, int reuse_plan, void* stream) {
  PC_REQUIRE(count >= 0, "bad count");
  if (count == 0) return PC_OK;
  PC_REQUIRE(descs_host && workspace, "null pointer argument");
  return pc::tc_grouped_gemm(descs_host, nullptr, count, workspace, workspace_bytes, reuse_plan,
                             (cudaStream_t)stream);
}

extern "C" int pc_grouped_gemm_tc_quant(const pc_gemm_desc* descs_host,
                                        const pc_gemm_quant* quant_host, int count,
                                        void* workspace, size_t workspace_bytes, int reuse_plan,
                                        void* stream) {
  PC_REQUIRE(count >= 0, "bad count");
  if (count == 0) return PC_OK;
  PC_REQUIRE(descs_host && quant_host && workspace, "null pointer argument");
  return pc::tc_grouped_gemm(descs_host, quant_host, count, workspace, workspace_bytes, reuse_plan,
                             (cudaStream_t)stream);
}
