// Thin inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma / commit / ld).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace vpk {
namespace ptx {

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-stream-serialization attribute may become
// resident while its predecessor in the stream is still draining.  `pdl_launch_dependents` lets the NEXT kernel's CTAs be
// scheduled as soon as every CTA of this grid has issued it (or exited); `pdl_wait` blocks until the PREVIOUS grid has
// completed and its memory is visible -- everything before it may only touch data no earlier kernel writes (plans,
// packed weights, biases, barrier / TMEM setup).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, %1;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}

__device__ __forceinline__ uint2 lds_u2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}

__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// ---- mbarrier --------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x989680u)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch fails with an error) instead of hanging the device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 8000000000ll) {   // ~4 s at 2 GHz
      printf("vpk: mbarrier wait timed out (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar,
             parity);
      __trap();
    }
  }
}

// Hot-loop wait for the tcgen05 pipelines.  No suspend-time hint: with a hint a failed try_wait parks the thread
// (NANOSLEEP.SYNCS) and the wake-up costs far more than the barrier round trip it waits for -- measured: the bare
// barrier skeleton of the gate GEMM (no MMA, no TMA, no epilogue math) took 5.3 us per tile that way.  The default
// try_wait returns after a short hardware-defined window, so this is a polite spin.  No clock reads or printf (no stack
// frame).  ~2^26 failed tries (seconds) mean a protocol bug: trap, so the launch fails instead of hanging the device.
__device__ __forceinline__ bool mbar_try_wait_nohint(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_fast(uint32_t bar, uint32_t parity) {
  uint32_t tries = 0;
  while (!mbar_try_wait_nohint(bar, parity)) {
    if (++tries > (1u << 26)) __trap();
  }
}

// Pure spin on mbarrier.test_wait (never parks the thread): for the single-thread producer / MMA-issue loops, where
// the wake-up latency of a parked try_wait would sit on the critical path of every pipeline hand-off.
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
  uint32_t tries = 0;
  while (!mbar_test_wait(bar, parity)) {
    if (++tries > (1u << 28)) __trap();
  }
}

// Register re-allocation between warpgroups (4 consecutive warps): producers give registers to the epilogue.
template <int N> __device__ __forceinline__ void setmaxnreg_dec() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N> __device__ __forceinline__ void setmaxnreg_inc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// ---- TMA -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tensormap(const void* desc) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(desc)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const void* desc, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(const void* desc, uint32_t bar, uint32_t dst, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// bulk tensor STORE shared -> global (bulk async-group completion); out-of-bounds parts of the box are not written
__device__ __forceinline__ void tma_store_4d(const void* desc, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(desc)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void sts_u4(uint32_t addr, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- tcgen05 ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, one CTA
__device__ __forceinline__ void mma_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}

// ---- 2-CTA (cta_group::2) forms -----------------------------------------------------------------------------------
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-parity bit of a shared::cluster address

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive (no tx) on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n"
      ".reg .b32 ra;\n"
      "mapa.shared::cluster.u32 ra, %0, %1;\n"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n"
      "}\n" ::"r"(bar),
      "r"(cta)
      : "memory");
}
// TMA loads issued by either CTA of a pair; completion bytes are signalled on the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(const void* desc, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
// the same, multicast: the box lands at the same shared-memory offset of every CTA in `cta_mask`, and each copy signals
// the barrier at this offset in ITS pair's leader CTA
__device__ __forceinline__ void tma_load_2d_pair_mc(const void* desc, uint32_t bar, uint32_t dst, int c0, int c1,
                                                    uint16_t cta_mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "h"(cta_mask)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_pair(const void* desc, uint32_t bar, uint32_t dst, int c0, int c1, int c2,
                                                 int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(desc)), "r"(bar & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows per CTA] * B[N/2 rows per CTA]^T : M = 256 across the pair, issued by the leader
__device__ __forceinline__ void mma_bf16_ss_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                 uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives on the barrier at this offset in every CTA of `cta_mask` once the previously issued MMAs have completed
__device__ __forceinline__ void mma_commit_pair(uint32_t bar, uint16_t cta_mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(cta_mask)
      : "memory");
}

// One tap = up to four K=16 slices issued back to back from one asm block.  The start-address field of a descriptor
// counts 16-byte units: +2 = 32 bytes = 16 bf16 along K (no carry out of the 14-bit field: shared memory < 256 KB).
__device__ __forceinline__ void mma_bf16_ss_tap(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                uint32_t accumulate, uint32_t nk) {
  if (nk == 4) {      // the common case, no predicates: four tcgen05.mma and six 64-bit immediate adds
    asm volatile(
        "{\n"
        ".reg .pred pacc;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 pacc, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pacc;\n"
        "add.s64 da, %1, 2;\n"
        "add.s64 db, %2, 2;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n"
        "add.s64 da, %1, 4;\n"
        "add.s64 db, %2, 4;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n"
        "add.s64 da, %1, 6;\n"
        "add.s64 db, %2, 6;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred pacc, p0, p1, p2;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 pacc, %4, 0;\n"
        "setp.gt.u32 p0, %5, 0;\n"
        "setp.gt.u32 p1, %5, 1;\n"
        "setp.gt.u32 p2, %5, 2;\n"
        "@p0 tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, pacc;\n"
        "add.s64 da, %1, 2;\n"
        "add.s64 db, %2, 2;\n"
        "@p1 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n"
        "add.s64 da, %1, 4;\n"
        "add.s64 db, %2, 4;\n"
        "@p2 tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, 1;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(nk)
        : "memory");
  }
}
__device__ __forceinline__ void mma_bf16_ss_tap_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                                     uint32_t accumulate, uint32_t nk) {
  if (nk == 4) {      // the common case, no predicates: four tcgen05.mma and six 64-bit immediate adds
    asm volatile(
        "{\n"
        ".reg .pred pacc;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 pacc, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, pacc;\n"
        "add.s64 da, %1, 2;\n"
        "add.s64 db, %2, 2;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n"
        "add.s64 da, %1, 4;\n"
        "add.s64 db, %2, 4;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n"
        "add.s64 da, %1, 6;\n"
        "add.s64 db, %2, 6;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n"
        ".reg .pred pacc, p0, p1, p2;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 pacc, %4, 0;\n"
        "setp.gt.u32 p0, %5, 0;\n"
        "setp.gt.u32 p1, %5, 1;\n"
        "setp.gt.u32 p2, %5, 2;\n"
        "@p0 tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, pacc;\n"
        "add.s64 da, %1, 2;\n"
        "add.s64 db, %2, 2;\n"
        "@p1 tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n"
        "add.s64 da, %1, 4;\n"
        "add.s64 db, %2, 4;\n"
        "@p2 tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, 1;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(nk)
        : "memory");
  }
}

// two full taps (2 x 4 K-slices) from one asm block
__device__ __forceinline__ void mma_bf16_ss_tap2(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint64_t a1, uint64_t b1,
                                                 uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred pacc;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 pacc, %6, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %5, pacc;\n"
      "add.s64 da, %1, 2;\n"
      "add.s64 db, %2, 2;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n"
      "add.s64 da, %1, 4;\n"
      "add.s64 db, %2, 4;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n"
      "add.s64 da, %1, 6;\n"
      "add.s64 db, %2, 6;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %4, %5, 1;\n"
      "add.s64 da, %3, 2;\n"
      "add.s64 db, %4, 2;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n"
      "add.s64 da, %3, 4;\n"
      "add.s64 db, %4, 4;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n"
      "add.s64 da, %3, 6;\n"
      "add.s64 db, %4, 6;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, 1;\n"
      "}\n" ::"r"(tmem_d),
      "l"(a0), "l"(b0), "l"(a1), "l"(b1), "r"(idesc), "r"(accumulate)
      : "memory");
}
// two full taps (2 x 4 K-slices) from one asm block
__device__ __forceinline__ void mma_bf16_ss_tap2_pair(uint32_t tmem_d, uint64_t a0, uint64_t b0, uint64_t a1, uint64_t b1,
                                                      uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred pacc;\n"
      ".reg .b64 da, db;\n"
      "setp.ne.b32 pacc, %6, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %5, pacc;\n"
      "add.s64 da, %1, 2;\n"
      "add.s64 db, %2, 2;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, 1;\n"
      "add.s64 da, %1, 4;\n"
      "add.s64 db, %2, 4;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, 1;\n"
      "add.s64 da, %1, 6;\n"
      "add.s64 db, %2, 6;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, 1;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %3, %4, %5, 1;\n"
      "add.s64 da, %3, 2;\n"
      "add.s64 db, %4, 2;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, 1;\n"
      "add.s64 da, %3, 4;\n"
      "add.s64 db, %4, 4;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, 1;\n"
      "add.s64 da, %3, 6;\n"
      "add.s64 db, %4, 6;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, 1;\n"
      "}\n" ::"r"(tmem_d),
      "l"(a0), "l"(b0), "l"(a1), "l"(b1), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Shared-memory matrix descriptor, K-major operand tile stored as rows of 128 B with the 128-byte swizzle
// (what TMA writes for a {64 x bf16, rows...} box with CU_TENSOR_MAP_SWIZZLE_128B): 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr >> 4) & 0x3FFF);      // start address        bits [0,14)
  d |= static_cast<uint64_t>(0) << 16;                        // leading byte offset  bits [16,30) (unused: 1 atom in K)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                // stride byte offset   bits [32,46)
  d |= static_cast<uint64_t>(1) << 46;                        // descriptor version   bits [46,48) = 1 on sm_100
  d |= static_cast<uint64_t>(2) << 61;                        // layout type          bits [61,64) = SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16 with bf16 (or, f16 = true, fp16) A/B (both K-major), fp32 accumulate, M x N tile.
__device__ __host__ __forceinline__ uint32_t idesc_bf16_f32(int M, int N, bool f16 = false) {
  uint32_t d = 0;
  d |= 1u << 4;                                   // D format  = F32
  d |= (f16 ? 0u : 1u) << 7;                      // A format  = BF16 (1) / F16 (0)
  d |= (f16 ? 0u : 1u) << 10;                     // B format  = BF16 (1) / F16 (0)
  d |= static_cast<uint32_t>(N >> 3) << 17;       // N / 8
  d |= static_cast<uint32_t>(M >> 4) << 24;       // M / 16
  return d;
}

}  // namespace ptx
}  // namespace vpk
