// mt_ptx.cuh — thin inline-PTX wrappers for sm_100a (mbarrier, TMA, tcgen05/TMEM, cluster).
// Everything here is hand-written for Blackwell; nothing compiles for older architectures.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda.h>
#include <stdint.h>
#include <stdio.h>

namespace mt {

#define MT_DEVINL __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// addresses
// ---------------------------------------------------------------------------------------------
MT_DEVINL uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
// address of the same shared-memory offset inside CTA `cta` of this cluster
MT_DEVINL uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
MT_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
MT_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
MT_DEVINL uint32_t lane_id() { return threadIdx.x & 31; }

MT_DEVINL uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b32 r;\n\t"
      "elect.sync r|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
MT_DEVINL void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
MT_DEVINL void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes to smem -> visible to the async proxy (TMA store / UMMA operand reads)
MT_DEVINL void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

MT_DEVINL void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
MT_DEVINL void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on the barrier at the same offset in CTA `cta` of the cluster
MT_DEVINL void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(bar), "r"(cta) : "memory");
}
// Suspend-time hint of mbarrier.try_wait: a waiting thread sleeps until the phase completes (or the hint expires) instead
// of coming back every few tens of nanoseconds (without a hint the limit measured here is one poll per ~40 ns: the 16
// epilogue warps of a GEMM CTA and the issuer threads spin through 40-60 % of a kernel).  Same-call A/B, 2000 ns against
// none: C2 step 9.79 / 9.83 ms against 9.94 / 10.20, fc1 220 us against 226, qkv 154 against 156, attention 92.1 against
// 92.9 — fewer wasted issue slots and less power under the cap.  0 = no hint.
#ifndef MT_WAIT_HINT_NS
#define MT_WAIT_HINT_NS 2000
#endif
MT_DEVINL uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
#if MT_WAIT_HINT_NS > 0
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity), "r"(uint32_t(MT_WAIT_HINT_NS)) : "memory");
#else
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
  return ok;
}
// Bounded wait: a protocol bug traps (kernel error) instead of hanging the GPU box.  try_wait itself suspends the
// thread for a hardware-defined interval, so the loop body is kept to a handful of instructions (spinning role
// warps share issue slots with the compute warps of their SM sub-partition).
#ifndef MT_WAIT_MAX_POLLS
#define MT_WAIT_MAX_POLLS (1u << 24)
#endif
MT_DEVINL void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t polls = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++polls > MT_WAIT_MAX_POLLS) {
#ifndef MT_WAIT_QUIET
      printf("[mt_b200] mbarrier wait timeout: block %d thread %d bar 0x%x parity %u\n", blockIdx.x, threadIdx.x, bar,
             parity);
#endif
      __trap();
    }
  }
}

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor)
// ---------------------------------------------------------------------------------------------
MT_DEVINL void tma_prefetch_desc(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
// L2 cache-policy descriptors (same encodings CUTLASS ships as TMA::CacheHintSm90)
constexpr uint64_t kEvictNormal = 0x1000000000000000ull;
constexpr uint64_t kEvictFirst = 0x12F0000000000000ull;
constexpr uint64_t kEvictLast = 0x14F0000000000000ull;

MT_DEVINL void tma_load_2d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
// 2-CTA flavour: data lands in this CTA's smem, completion bytes are signalled on `bar`, which may be
// a shared::cluster address (the leader CTA's barrier).
MT_DEVINL void tma_load_2d_cg2(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
MT_DEVINL void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3,
                           uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(hint)
      : "memory");
}
MT_DEVINL void tma_store_2d(const void* tmap, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
}
MT_DEVINL void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
// element-wise += of a shared-memory tile into global memory by the TMA engine (fp32 add; out-of-range rows are clipped)
MT_DEVINL void tma_reduce_add_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
MT_DEVINL void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
MT_DEVINL void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
MT_DEVINL void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
template <int CG>
MT_DEVINL void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  if constexpr (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
MT_DEVINL void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
MT_DEVINL void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
MT_DEVINL void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
MT_DEVINL void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
MT_DEVINL void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[smem] * B[smem]; bf16/fp16 inputs, fp32 accumulate.  Issued by ONE thread.
template <int CG>
MT_DEVINL void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M x 16, K-major) is read from tensor memory — row m = lane m, two
// 16-bit elements per 32-bit column (8 columns per instruction) — instead of shared memory.  With both operands in
// shared memory an M = 128 instruction occupies the tensor pipe for >= ~60 cycles however small N is (the A tile is
// re-read from shared memory every time); from tensor memory the floor is N / 2 cycles (tools/bench_umma.cu).
MT_DEVINL void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  const uint32_t z = 0u;
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(z)
      : "memory");
}
// same with fp8 (e4m3 / e5m2, chosen in the instruction descriptor) operands: 32 elements of K per instruction
template <int CG>
MT_DEVINL void umma_f8_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  }
}
// tcgen05.commit: arrive on `bar` once every previously issued tcgen05.mma of this thread has finished.
// (implies tcgen05.fence::before_thread_sync)
MT_DEVINL void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 2-CTA: arrive on the barrier at this offset in every CTA of `mask`
MT_DEVINL void umma_commit_cg2_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(bar), "h"(mask)
      : "memory");
}

// 32 lanes x 32 columns of fp32: thread t of the warp gets lane (32*(warp%4)+t), 32 consecutive columns
MT_DEVINL void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// same, writing r[OFF .. OFF+31] of a larger register array (keeps the array in registers: no address taken)
template <int OFF, int LEN>
MT_DEVINL void tmem_ld_32x32b_x32_at(uint32_t taddr, uint32_t (&r)[LEN]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[OFF + 0]), "=r"(r[OFF + 1]), "=r"(r[OFF + 2]), "=r"(r[OFF + 3]), "=r"(r[OFF + 4]), "=r"(r[OFF + 5]),
        "=r"(r[OFF + 6]), "=r"(r[OFF + 7]), "=r"(r[OFF + 8]), "=r"(r[OFF + 9]), "=r"(r[OFF + 10]), "=r"(r[OFF + 11]),
        "=r"(r[OFF + 12]), "=r"(r[OFF + 13]), "=r"(r[OFF + 14]), "=r"(r[OFF + 15]), "=r"(r[OFF + 16]),
        "=r"(r[OFF + 17]), "=r"(r[OFF + 18]), "=r"(r[OFF + 19]), "=r"(r[OFF + 20]), "=r"(r[OFF + 21]),
        "=r"(r[OFF + 22]), "=r"(r[OFF + 23]), "=r"(r[OFF + 24]), "=r"(r[OFF + 25]), "=r"(r[OFF + 26]),
        "=r"(r[OFF + 27]), "=r"(r[OFF + 28]), "=r"(r[OFF + 29]), "=r"(r[OFF + 30]), "=r"(r[OFF + 31])
      : "r"(taddr)
      : "memory");
}
MT_DEVINL void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// same, writing r[OFF .. OFF+15] of a larger register array (keeps the array in registers: no address taken)
template <int OFF, int LEN>
MT_DEVINL void tmem_ld_32x32b_x16_at(uint32_t taddr, uint32_t (&r)[LEN]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[OFF + 0]), "=r"(r[OFF + 1]), "=r"(r[OFF + 2]), "=r"(r[OFF + 3]), "=r"(r[OFF + 4]), "=r"(r[OFF + 5]),
        "=r"(r[OFF + 6]), "=r"(r[OFF + 7]), "=r"(r[OFF + 8]), "=r"(r[OFF + 9]), "=r"(r[OFF + 10]), "=r"(r[OFF + 11]),
        "=r"(r[OFF + 12]), "=r"(r[OFF + 13]), "=r"(r[OFF + 14]), "=r"(r[OFF + 15])
      : "r"(taddr)
      : "memory");
}
MT_DEVINL void tmem_st_32x32b_x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
MT_DEVINL void tmem_st_32x32b_x16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (bit layouts follow the PTX ISA "tcgen05 matrix/instruction descriptor" tables)
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor for a 128B-swizzled tile whose rows are 128 bytes
// (K-major: row = one M/N index, 64 bf16 of K; MN-major: row = one K index, 64 bf16 of M/N).
// 8-row groups are 1024 B apart (SBO). LBO is unused for a single 128B-wide swizzle atom.
MT_DEVINL uint64_t make_smem_desc_sw128(uint32_t saddr, uint32_t sbo_bytes = 1024, uint32_t lbo_bytes = 16) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);        // bits [0,14)  start address >> 4
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;  // bits [16,30) leading byte offset >> 4
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;  // bits [32,46) stride byte offset >> 4
  d |= static_cast<uint64_t>(1) << 46;                       // bits [46,48) descriptor version = 1 (sm_100)
  d |= static_cast<uint64_t>(2) << 61;                       // bits [61,64) layout = SWIZZLE_128B
  return d;
}
// Instruction descriptor for kind::f16: bf16 x bf16 -> fp32.
// a_mn_major / b_mn_major: 0 = K-major operand tile, 1 = MN-major operand tile.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N, int a_mn_major = 0, int b_mn_major = 0) {
  return (1u << 4)                      // c_format = F32
         | (1u << 7)                    // a_format = BF16
         | (1u << 10)                   // b_format = BF16
         | (uint32_t(a_mn_major) << 15) // a_major
         | (uint32_t(b_mn_major) << 16) // b_major
         | (uint32_t(N >> 3) << 17)     // n_dim
         | (uint32_t(M >> 4) << 24);    // m_dim
}

// Instruction descriptor for kind::f8f6f4 with e4m3 x e4m3 -> fp32 (a_format = b_format = 0), K-major operands.
__host__ __device__ constexpr uint32_t make_idesc_e4m3(int M, int N) {
  return (1u << 4) | (uint32_t(N >> 3) << 17) | (uint32_t(M >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// small math helpers
// ---------------------------------------------------------------------------------------------
MT_DEVINL uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
MT_DEVINL float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
MT_DEVINL float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

// Programmatic dependent launch: a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor in the stream is still draining; pdl_wait() blocks until the predecessor has completed and
// its memory is visible (call it before the first dependent global access), pdl_launch_dependents() lets the
// successor's CTAs be scheduled as soon as SMs free up.
MT_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
MT_DEVINL void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// packed fp32x2 arithmetic (sm_100): one issue slot for two lanes of FMA work
MT_DEVINL uint64_t pack_f32x2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
MT_DEVINL void unpack_f32x2(uint64_t v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
MT_DEVINL uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
MT_DEVINL uint64_t mul_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
MT_DEVINL uint64_t add_f32x2(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}

// device timeline (mt_timeline_begin): slot[0] = earliest CTA start, slot[1] = latest CTA end, in %globaltimer ns
MT_DEVINL unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
MT_DEVINL void timeline_start(unsigned long long* tl) {
  if (tl != nullptr && threadIdx.x == 0) atomicMin(tl, globaltimer_ns());
}
MT_DEVINL void timeline_end(unsigned long long* tl) {
  if (tl != nullptr && threadIdx.x == 0) atomicMax(tl + 1, globaltimer_ns());
}

// Attention dropout (Graph fork: attn_drop = 0.1): keep / drop decision of attention-matrix element (q, k) of one
// (batch, head) as a pure function of (seed, batch*H + head, q*N + k), so that the forward kernel (one query row per
// thread), the backward kernel (one key row per thread) and the mask-dump kernel of the tests all agree without storing
// the mask.  lowbias32 integer hash (two multiplies): enough statistical quality for a Bernoulli mask.
MT_DEVINL uint32_t mix32(uint32_t x) {
  x ^= x >> 16;
  x *= 0x7feb352du;
  x ^= x >> 15;
  x *= 0x846ca68bu;
  x ^= x >> 16;
  return x;
}
MT_DEVINL uint32_t attn_drop_seed(uint32_t seed_lo, uint32_t seed_hi, uint32_t bh) {
  return mix32(seed_lo ^ mix32(seed_hi + bh * 0x9e3779b1u));
}
// keep when hash >= thresh, thresh = p * 2^32
MT_DEVINL bool attn_keep(uint32_t bh_seed, uint32_t idx, uint32_t thresh) { return mix32(idx ^ bh_seed) >= thresh; }

MT_DEVINL void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
MT_DEVINL void named_bar_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace mt
