// opp_common.cuh — sm_100a PTX wrappers shared by every kernel of the 2D-3D matching hot path.
//
// Everything here is inline PTX for Blackwell (tcgen05 / TMEM / TMA / mbarrier). There is no
// fallback for other architectures: the translation units that include this header are
// compiled with -gencode arch=compute_100a,code=sm_100a only.
#pragma once

#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace opp {

// ---------------------------------------------------------------------------------------------
// error plumbing (C-ABI functions return int status; 0 = ok)
// ---------------------------------------------------------------------------------------------
enum Status : int {
  OPP_OK = 0,
  OPP_ERR_INVALID = 1,   // bad argument (shape / alignment / null pointer)
  OPP_ERR_CUDA = 2,      // a CUDA runtime / driver call failed (see opp_last_error)
  OPP_ERR_UNSUPPORTED = 3
};

void set_last_error(const char* fmt, ...);

#define OPP_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::opp::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                            cudaGetErrorString(_e));                                      \
      return ::opp::OPP_ERR_CUDA;                                                         \
    }                                                                                     \
  } while (0)

#define OPP_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      ::opp::set_last_error(__VA_ARGS__);                                                 \
      return ::opp::OPP_ERR_INVALID;                                                      \
    }                                                                                     \
  } while (0)

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of the library is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization (launch_pdl below / launch<> in opp_gemm.cu):
// the NEXT kernel of the stream may start while this one is still running, so its launch latency
// and its prologue (barrier init, TMEM allocation, descriptor prefetch) overlap with our tail.
// Contract: each kernel calls pdl_wait() on every path before it touches global memory — it
// returns once the preceding kernel has completed and its writes are visible — and then
// pdl_trigger() so that its own successor may be scheduled.  Because every kernel waits before it
// completes, completion stays transitive along the stream.  Both are no-ops for a kernel that
// was launched without the attribute.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// OPP_PDL_TRIGGER: 0 = no explicit trigger: the successor is released when our CTAs exit, i.e. its
// launch overlaps with the end-of-grid memory flush only; 1 = trigger right after the wait (the
// successor parks early on free SMs); 2 = trigger when the main work of the CTA is done.
// Measured at batch 1 in CUDA-graph mode (gpurun_out/lat2, B200): attribute off 1.654 ms,
// 0: 1.620 ms, 1: 1.818 ms, 2: 1.804 ms — a grid released by launch_dependents pays more in
// griddepcontrol.wait than the overlapped prologue saves, so 0 is the default.
#ifndef OPP_PDL_TRIGGER
#define OPP_PDL_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_sync() {
  pdl_wait();
#if OPP_PDL_TRIGGER == 1
  pdl_trigger();
#endif
}
__device__ __forceinline__ void pdl_done() {
#if OPP_PDL_TRIGGER == 2
  pdl_trigger();
#endif
}

// $OPP_PDL=0 launches everything fully serialised (debugging / A-B timing)
inline int pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("OPP_PDL");
    v = e ? atoi(e) : 1;
  }
  return v;
}

// kernel<<<grid, block, smem, stream>>>(args...) with the programmatic-serialization attribute
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                              cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a mis-programmed pipeline traps (launch failure reported to the host) instead of
// hanging the GPU. ~4e9 cycles is about two seconds at B200 clocks.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("opp: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// Lean wait for the hot MMA-issue loop: no clock reads; the watchdog counts polls instead (every
// try_wait already suspends the thread for a hardware-defined interval when the phase is pending).
__device__ __forceinline__ void mbar_wait_hot(uint64_t* bar, uint32_t parity) {
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++polls > (1u << 24)) {
      printf("opp: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) loads; tensor maps are passed as __grid_constant__ kernel params
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* smem,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* smem,
                                            int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* smem,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}

// multicast variant: the box lands at the same smem offset of every CTA in `cta_mask` and performs
// complete_tx on the mbarrier at the same offset in each of them
__device__ __forceinline__ void tma_load_3d_mc(const CUtensorMap* map, uint64_t* bar, void* smem,
                                               int c0, int c1, int c2, uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%4, %5, %6}], [%2], %3;" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(cta_mask), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}

// cta_group::2 loads: executed by both CTAs of a pair, each into its own shared memory; the
// transaction bytes are credited to the LEADER CTA's mbarrier (peer bit of the address cleared).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ void tma_load_3d_2sm(const CUtensorMap* map, uint64_t* bar, void* smem,
                                                int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(const CUtensorMap* map, uint64_t* bar, void* smem,
                                                int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1),
      "r"(c2), "r"(c3)
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// thread-block clusters
// ---------------------------------------------------------------------------------------------
// arrive on the mbarrier at the same shared-memory offset in CTA `cta_rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta_rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta_rank)
      : "memory");
}
// the same with release semantics at cluster scope: data written to the peer's shared memory
// (st_cluster_f32x2) before the arrive is visible to a peer thread that acquires the phase
__device__ __forceinline__ void mbar_arrive_cluster_release(uint64_t* bar, uint32_t cta_rank) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(cta_rank)
      : "memory");
}
// two floats into CTA `cta_rank`'s shared memory, at the offset `local` has in this CTA
__device__ __forceinline__ void st_cluster_f32x2(void* local, uint32_t cta_rank, float a, float b) {
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "st.shared::cluster.v2.f32 [ra], {%2, %3};\n\t"
      "}\n" ::"r"(smem_u32(local)),
      "r"(cta_rank), "f"(a), "f"(b)
      : "memory");
}
// bounded wait with acquire at cluster scope (pairs with mbar_arrive_cluster_release)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  const long long t0 = clock64();
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (clock64() - t0 > 4000000000LL) __trap();
  }
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.aligned;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// cta_group::2 (CTA pair) variants
__device__ __forceinline__ void tmem_alloc2(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_commit2_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows: 128 per CTA] * B[N rows: N/2 per CTA]^T; leader issues
__device__ __forceinline__ void tc_mma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma2_f16_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                                uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.u32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// tcgen05.commit: arrive on an mbarrier once every previously issued MMA of this thread retired.
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// same, arriving on the barrier at the same smem offset of every CTA in `cta_mask`
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
      "[%0], %1;" ::"r"(smem_u32(bar)),
      "h"(cta_mask)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, fp16/bf16 operands, fp32 accumulate. One thread issues.
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                           uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// accumulate variant without the predicate set-up (every MMA of a tile but the first)
__device__ __forceinline__ void tc_mma_f16_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                               uint32_t idesc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.eq.u32 p, 0, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc)
      : "memory");
}

// K-major operand tile in shared memory, 128-byte swizzle (what TMA SWIZZLE_128B produces for a
// box whose inner extent is 64 fp16): rows of 128 B, 8-row groups 1024 B apart.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused: 1)
//   bits [32,46) stride byte offset >> 4   bits [46,48) descriptor version = 1 (sm_100)
//   bits [61,64) layout type: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor for kind::f16, fp16 A/B (K-major both), fp32 accumulator.
//   [4,6) c_format (1 = f32)  [7,10) a_format (0 = f16, 1 = bf16)  [10,13) b_format
//   [15] a_major (0 = K)  [16] b_major (0 = K)  [17,23) N >> 3  [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
  return (1u << 4) | (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp reads TMEM lane (base_lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// Issue-only variant + explicit wait, so the next chunk's TMEM load overlaps the math on the
// current one.  The wait names the registers as in/out operands so the compiler cannot hoist
// their uses above it.
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
        "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
        "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]),
                 "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]),
                 "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]),
                 "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]),
                 "+r"(r[30]), "+r"(r[31])::"memory");
}
// Walk the accumulator row of this thread in 32-column chunks with one chunk of lookahead:
// f(col, v[32]) is called for col = first, first+step, ... < ncols (all warp-uniform).  Two
// epilogue warp groups share a tile by taking alternate chunks (first = 32*group, step = 64).
template <class F>
__device__ __forceinline__ void tmem_foreach32(uint32_t tbase, int ncols, int first, int step,
                                               F&& f) {
  if (first >= ncols) return;
  uint32_t ra[32], rb[32];
  float v[32];
  tmem_ld32_issue(tbase + first, ra);
  tmem_ld_wait(ra);
  for (int col = first; col < ncols; col += 2 * step) {
    const bool has_b = col + step < ncols;
    if (has_b) tmem_ld32_issue(tbase + col + step, rb);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(ra[i]);
    f(col, v);
    if (has_b) {
      tmem_ld_wait(rb);
      const bool has_a = col + 2 * step < ncols;
      if (has_a) tmem_ld32_issue(tbase + col + 2 * step, ra);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(rb[i]);
      f(col + step, v);
      if (has_a) tmem_ld_wait(ra);
    }
  }
}
template <class F>
__device__ __forceinline__ void tmem_foreach32(uint32_t tbase, int ncols, F&& f) {
  tmem_foreach32(tbase, ncols, 0, 32, f);
}
// Same walk without the lookahead chunk (32 fewer live registers).  Meant for epilogues that run
// two warp groups: 320 threads cap a thread at 168 registers, and the second warp resident on
// each SM sub-partition hides the TMEM load latency instead.
template <class F>
__device__ __forceinline__ void tmem_foreach32_lean(uint32_t tbase, int ncols, int first, int step,
                                                    F&& f) {
  float v[32];
  for (int col = first; col < ncols; col += step) {
    tmem_ld32(tbase + col, v);
    f(col, v);
  }
}
template <bool LOOKAHEAD, class F>
__device__ __forceinline__ void tmem_foreach32_sel(uint32_t tbase, int ncols, int first, int step,
                                                   F&& f) {
  if constexpr (LOOKAHEAD) tmem_foreach32(tbase, ncols, first, step, f);
  else tmem_foreach32_lean(tbase, ncols, first, step, f);
}
// explicit shared-space accesses (32-bit shared addresses): pointers kept in structs decay to
// generic LD/ST, which showed up as long-scoreboard stalls all over the epilogues
__device__ __forceinline__ void sts128(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ float lds32f(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts32f(uint32_t addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// named barrier among a subset of warps (id 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// packing helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// store 8 consecutive fp16 (16 bytes) converted from fp32
__device__ __forceinline__ void store_half8(__half* dst, const float* v) {
  uint4 u;
  u.x = pack_half2(v[0], v[1]);
  u.y = pack_half2(v[2], v[3]);
  u.z = pack_half2(v[4], v[5]);
  u.w = pack_half2(v[6], v[7]);
  *reinterpret_cast<uint4*>(dst) = u;
}
__device__ __forceinline__ void load_half8(const __half* src, float* v) {
  uint4 u = *reinterpret_cast<const uint4*>(src);
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 f = __half22float2(h[i]);
    v[2 * i] = f.x;
    v[2 * i + 1] = f.y;
  }
}

__device__ __forceinline__ float elu_plus_one(float x) { return x > 0.f ? x + 1.f : expf(x); }
// Epilogue variant on the SFU (ex2.approx): |rel err| <= ~(2 + |1.44 x|) ulp, i.e. < 2e-6 for the
// arguments that matter (x in (-10, 0]); one instruction pair instead of ~25.
// exp(x) as one FMUL + MUFU.EX2 (flush-to-zero): relative error ~2^-22 from ex2.approx plus
// |x| * 6e-8 from the log2(e) product, i.e. <= 1e-5 for |x| <= 100.
__device__ __forceinline__ float fast_exp(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
  return y;
}
__device__ __forceinline__ float elu_plus_one_fast(float x) { return x > 0.f ? x + 1.f : fast_exp(x); }

// ---------------------------------------------------------------------------------------------
// split-precision activations.  Every tensor that feeds a tensor-core GEMM is stored as a pair of
// fp16 planes along its channel axis: [hi(C) | lo(C)], hi = fp16(x), lo = fp16(x - hi), so that
// hi + lo carries ~22 mantissa bits.  `lo_off` is the element offset of the lo plane inside a row
// (0 = plain fp16 storage, no lo plane).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void store_split8(__half* row, int col, const float* v, int lo_off) {
  store_half8(row + col, v);
  if (lo_off) {
    float lo[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) lo[j] = v[j] - __half2float(__float2half_rn(v[j]));
    store_half8(row + lo_off + col, lo);
  }
}
__device__ __forceinline__ void load_split8(const __half* row, int col, float* v, int lo_off) {
  load_half8(row + col, v);
  if (lo_off) {
    float lo[8];
    load_half8(row + lo_off + col, lo);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += lo[j];
  }
}
__device__ __forceinline__ float load_split1(const __half* row, int col, int lo_off) {
  float v = __half2float(row[col]);
  if (lo_off) v += __half2float(row[lo_off + col]);
  return v;
}
__device__ __forceinline__ void store_split1(__half* row, int col, float v, int lo_off) {
  const __half h = __float2half_rn(v);
  row[col] = h;
  if (lo_off) row[lo_off + col] = __float2half_rn(v - __half2float(h));
}

}  // namespace opp
