// Blackwell (sm_100a) device-side primitives used by every kernel in this library:
// mbarrier, TMA (cp.async.bulk.tensor), tcgen05 MMA / TMEM, UMMA descriptors.
// All inline PTX; no CUTLASS/CuTe dependency.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace stswin {

#ifndef STSWIN_SPIN_LIMIT
// Bounded mbarrier waits: a descriptor or pipeline bug traps instead of hanging the GPU.
#define STSWIN_SPIN_LIMIT (1u << 26)
#endif

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// First 1024-byte aligned address inside a dynamic shared-memory array.  Written as base + offset
// (not an integer round trip) so the compiler keeps the shared address space and emits LDS / STS.
__device__ __forceinline__ uint8_t* align_smem_1024(uint8_t* raw) {
  return raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ----------------------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > STSWIN_SPIN_LIMIT) __trap();
  }
}

// ----------------------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {
  // generic-proxy writes to smem -> visible to the async proxy (TMA store, tcgen05.mma operands)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ----------------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// multicast variant: the box lands at the same smem offset in every CTA of `cta_mask`, and each of
// those CTAs' mbarrier (same offset) receives the complete_tx
__device__ __forceinline__ void tma_load_2d_mcast(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                  uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, "
      "%4}], [%2], %5;" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::
          "r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// fp32 add-reduction of an smem tile into global memory (split-K weight gradients)
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* m, const void* smem, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap* m, const void* smem, int c0, int c1, int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void tma_wait_group() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ----------------------------------------------------------------------------- TMEM
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {   // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {     // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}

#define STSWIN_TMEM_LD_X32(taddr, v)                                                                                 \
  asm volatile(                                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                      \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                      \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                      \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),  \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),       \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),      \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                    \
      : "r"(taddr)                                                                                                   \
      : "memory")

// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32*(warp%4) + laneid)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) { STSWIN_TMEM_LD_X32(taddr, v); }
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ----------------------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   K-major  operand: rows of 64 bf16 (128 B); 8-row groups SBO bytes apart; LBO unused.
//   MN-major operand: 64 contiguous MN elements (128 B) per k; 8-k groups SBO bytes apart;
//                     further 64-element MN chunks LBO bytes apart.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;   // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;   // SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16, bf16 x bf16 -> fp32.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                      // D format: f32
         | (1u << 7)                    // A format: bf16
         | (1u << 10)                   // B format: bf16
         | (uint32_t(a_mn_major) << 15) // A major
         | (uint32_t(b_mn_major) << 16) // B major
         | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// same, arriving on the mbarrier at this offset in every CTA of `cta_mask` (cluster multicast)
__device__ __forceinline__ void umma_commit_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}

// ---- CTA-pair (cta_group::2) variants: one tcgen05.mma spans the tensor cores of both SMs of a
// cluster pair (M = 256: 128 rows per CTA); each CTA keeps only ITS half of the B operand in shared
// memory.  Issued by the leader CTA (cluster rank 0); completion is multicast to both CTAs.
template <int COLS>
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem) {   // one full warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm_mcast(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"(cta_mask)
               : "memory");
}
// shared::cluster address of `p` (an address in this CTA's shared memory) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// arrive (+ expect_tx) on an mbarrier given by its shared::cluster address (may be the peer CTA's)
__device__ __forceinline__ void mbar_arrive_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster_addr),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA load into THIS CTA's shared memory whose completion is signalled on an mbarrier of the CTA pair
// (shared::cluster address, normally the leader's barrier)
__device__ __forceinline__ void tma_load_2d_2sm(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}

__device__ __forceinline__ void tma_load_3d_2sm(void* smem, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, "
      "%5}], [%2];" ::"r"(smem_u32(smem)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// plain (non-tensor) bulk copy global -> this CTA's shared memory; addresses and size multiples of 16 bytes
__device__ __forceinline__ void bulk_load_1d(void* smem, const void* gptr, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem)),
               "l"(reinterpret_cast<uint64_t>(gptr)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// ----------------------------------------------------------------------------- programmatic dependent launch
// A kernel launched with the programmatic-stream-serialization attribute (host_util.h: launch_pdl) may start while its
// predecessor on the stream is still running: everything before pdl_wait() (barrier initialisation, TMEM allocation,
// descriptor prefetch) overlaps the predecessor's tail; pdl_wait() returns once the predecessor has completed and its
// memory is visible.  pdl_launch_dependents() lets the NEXT kernel begin its own prologue early.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ----------------------------------------------------------------------------- clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ----------------------------------------------------------------------------- small math helpers
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// 256-bit global store (STG.256 on sm_100): a full 32-byte sector per lane and instruction
__device__ __forceinline__ void st_global_v8(void* p, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// Column sums over the 32 lanes of a warp for NV (32 or 64) per-lane values (lane = row): a butterfly
// that halves the live values at every step (NV - NV/32 shuffles).  On return lane l holds the sums
// of columns l*NV/32 .. in a[0 .. NV/32).
template <int NV>
__device__ __forceinline__ void warp_colsum(float (&a)[NV], int lane) {
#pragma unroll
  for (int n = NV / 2, off = 16; off >= 1; n >>= 1, off >>= 1) {
    const bool hi = (lane & off) != 0;
#pragma unroll
    for (int k = 0; k < n; ++k) {
      const float send = hi ? a[k] : a[k + n];
      const float recv = __shfl_xor_sync(0xffffffffu, send, off);
      a[k] = (hi ? a[k + n] : a[k]) + recv;
    }
  }
}
// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2 on sm_100): two lanes of work per issue slot
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// erf-GELU (nn.GELU default, swin_512.py:13) and its derivative for two values at once, packed fp32x2 arithmetic: u = acc + bias, returns the packed bf16
// pairs of GELU(u) and GELU'(u).  The GELU epilogue is bound by FP32 issue slots (ncu, profiles/r2u_gemm_gelu_roles.txt:
// the eight epilogue warps issue in 53 % of all cycles and the tensor pipe idles at 52 %), so this version spends
// 12 packed operations + 2 MUFU per pair instead of 17 + 2:
//   Phi(u) ~ 0.5 + 0.5 tanh(u (c0 + c1 u^2))   -- cubic argument fitted to the exact normal CDF: |error| 2.7e-4 on GELU,
//            8.7e-4 on GELU' (the quartic fit of round 1: 5.6e-5 / 1.8e-4, at 17 operations and a clamp; tanh.approx itself is good to ~2.4e-4 on Phi, and the
//            stored bf16 results round at 2e-3 relative) -- positive for every u, so no clamp of u^2 is needed;
//   GELU'(u) = Phi + u (1 - Phi) * 2 Phi a'(u) = Phi + Phi * (u - h) * 2 a'(u)   with h = u Phi   (1 - t^2 = 4 Phi (1 - Phi))
constexpr float kGeluC0 = 0.80015708f, kGeluC1 = 0.03470089f;
__device__ __forceinline__ void gelu_and_grad_x2(float acc0, float acc1, float b0, float b1, uint32_t& h_bf16x2,
                                                 uint32_t& g_bf16x2) {
  const uint64_t u = f2_add(f2_pack(acc0, acc1), f2_pack(b0, b1));
  const uint64_t s = f2_mul(u, u);
  float a0, a1, t0, t1;
  f2_unpack(f2_mul(u, f2_fma(s, f2_pack(kGeluC1, kGeluC1), f2_pack(kGeluC0, kGeluC0))), a0, a1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
  const uint64_t cdf = f2_fma(f2_pack(t0, t1), f2_pack(0.5f, 0.5f), f2_pack(0.5f, 0.5f));
  const uint64_t h = f2_mul(u, cdf);
  const uint64_t da2 = f2_fma(s, f2_pack(6.0f * kGeluC1, 6.0f * kGeluC1), f2_pack(2.0f * kGeluC0, 2.0f * kGeluC0));   // 2 a'(u)
  const uint64_t d = f2_fma(h, f2_pack(-1.0f, -1.0f), u);                       // u - h = u (1 - Phi)
  const uint64_t g = f2_fma(f2_mul(cdf, d), da2, cdf);
  float h0, h1, g0, g1;
  f2_unpack(h, h0, h1);
  f2_unpack(g, g0, g1);
  h_bf16x2 = pack_bf16(h0, h1);
  g_bf16x2 = pack_bf16(g0, g1);
}
// The same with GELU' quantised to one byte (STSWIN_EPI_BIAS_GELU_Q8): q = round((g + 0.14) * 255 / 1.28) through the
// 2^23 trick -- the low byte of the returned words is q (g is bounded by [-0.13, 1.13], so no clamp).
constexpr float kGeluQ8Scale = 255.0f / 1.28f, kGeluQ8Off = 0.14f;
__device__ __forceinline__ void gelu_and_grad_q8_x2(float acc0, float acc1, float b0, float b1, uint32_t& h_bf16x2,
                                                    uint32_t& q0, uint32_t& q1) {
  const uint64_t u = f2_add(f2_pack(acc0, acc1), f2_pack(b0, b1));
  const uint64_t s = f2_mul(u, u);
  float a0, a1, t0, t1;
  f2_unpack(f2_mul(u, f2_fma(s, f2_pack(kGeluC1, kGeluC1), f2_pack(kGeluC0, kGeluC0))), a0, a1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
  const uint64_t cdf = f2_fma(f2_pack(t0, t1), f2_pack(0.5f, 0.5f), f2_pack(0.5f, 0.5f));
  const uint64_t h = f2_mul(u, cdf);
  const uint64_t da2 = f2_fma(s, f2_pack(6.0f * kGeluC1, 6.0f * kGeluC1), f2_pack(2.0f * kGeluC0, 2.0f * kGeluC0));
  const uint64_t d = f2_fma(h, f2_pack(-1.0f, -1.0f), u);
  const uint64_t g = f2_fma(f2_mul(cdf, d), da2, cdf);
  constexpr float kMagic = 8388608.0f + kGeluQ8Off * kGeluQ8Scale;
  const uint64_t t = f2_fma(g, f2_pack(kGeluQ8Scale, kGeluQ8Scale), f2_pack(kMagic, kMagic));
  float h0, h1, f0, f1;
  f2_unpack(h, h0, h1);
  f2_unpack(t, f0, f1);
  h_bf16x2 = pack_bf16(h0, h1);
  q0 = __float_as_uint(f0);
  q1 = __float_as_uint(f1);
}
// two quantised GELU' values (bytes k0, k1 of w) back to fp32, packed
__device__ __forceinline__ uint64_t gelu_q8_dequant_x2(uint32_t w, int k0, int k1) {
  const float f0 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + k0));     // 2^23 + byte, exactly
  const float f1 = __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + k1));
  const uint64_t x = f2_add(f2_pack(f0, f1), f2_pack(-8388608.0f, -8388608.0f));
  return f2_fma(x, f2_pack(1.0f / kGeluQ8Scale, 1.0f / kGeluQ8Scale), f2_pack(-kGeluQ8Off, -kGeluQ8Off));
}
// forward-only GELU of two values (inference: no derivative output)
__device__ __forceinline__ uint32_t gelu_x2(float acc0, float acc1, float b0, float b1) {
  const uint64_t u = f2_add(f2_pack(acc0, acc1), f2_pack(b0, b1));
  float a0, a1, t0, t1;
  f2_unpack(f2_mul(u, f2_fma(f2_mul(u, u), f2_pack(kGeluC1, kGeluC1), f2_pack(kGeluC0, kGeluC0))), a0, a1);
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(a0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(a1));
  const uint64_t h = f2_mul(u, f2_fma(f2_pack(t0, t1), f2_pack(0.5f, 0.5f), f2_pack(0.5f, 0.5f)));
  float h0, h1;
  f2_unpack(h, h0, h1);
  return pack_bf16(h0, h1);
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a 128B-swizzled tile whose
// rows are 128 bytes and whose base is 1024-byte aligned (the pattern TMA and UMMA both use)
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return row * 128u + ((chunk ^ (row & 7u)) << 4);
}

}  // namespace stswin
