// Thin inline-PTX layer for sm_100a: mbarrier, bulk async copy, tcgen05 (alloc / mma / commit / ld), descriptors.
// Bit layouts follow the PTX ISA "tcgen05 matrix / instruction descriptors" (same fields CUTLASS's
// cute/arch/mma_sm100_desc.hpp encodes); nothing here is copied from a library, the kernels own their pipelines.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait with a suspend-time hint: a waiting thread sleeps in hardware (no issue slots) until the phase completes
// or the hint expires.  Plain try_wait polling by the role warps that wait most of the time (epilogue, loader, MMA)
// was eating ~45 % of the issue bandwidth of the gather warps sharing their SM sub-partitions.
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(0x4000u)
      : "memory");
  return ok != 0;
}
#ifdef PASCO_HANG_TRAP
// debug build (PASCO_NVCC_FLAGS=-DPASCO_HANG_TRAP python -m pasco_b200.build --force): a wait that spins for ~1 s reports
// who waits on what and traps, so that a pipeline deadlock shows up as an error with a location instead of a hang
static __device__ __noinline__ void mbar_wait_report(uint32_t bar, uint32_t parity) {
  printf("HANG block %d thread %d (warp %d) waits on smem barrier 0x%x parity %u\n", blockIdx.x, threadIdx.x, threadIdx.x >> 5, bar, parity);
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned long long spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1ull << 17)) mbar_wait_report(bar, parity);
  }
}
#else
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
#endif

// ---- async-proxy visibility of generic st.shared writes ------------------------------------------------
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- 1-D bulk copy global → shared, completion on an mbarrier (UBLKCP) ---------------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---- TMEM -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one lane of a converged warp (the issue loops run warp-uniform so that descriptors stay in uniform registers; only the
// tcgen05.mma / tcgen05.commit themselves are predicated on the elected lane)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[smem desc] · B[smem desc], bf16 operands, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 32 lanes x 32 columns of fp32: thread t ← lane (base+t), registers ← consecutive columns
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, 128-byte swizzle.  Tiles are stored as rows of 128 bytes (64 bf16), 8 rows =
// one 1024-byte swizzle atom, 16-byte chunk c of row r lives at chunk (c ^ (r & 7)).
//   K-major  operand (rows = M/N index, 128 B = 64 K-elements): SBO = 1024 (next 8 rows), LBO unused (1)
//   MN-major operand (rows = K index,   128 B = 64 M/N-elements): SBO = 1024 (next 8 K-rows),
//                                                                LBO = byte distance to the next 64 M/N elements
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

// Instruction descriptor for kind::f16 with bf16 A/B and fp32 D.  a_mn / b_mn: operand is MN-major.
__host__ __device__ __forceinline__ uint32_t make_idesc_bf16(int M, int N, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;                    // D format  = F32
  d |= 1u << 7;                    // A format  = BF16
  d |= 1u << 10;                   // B format  = BF16
  d |= (uint32_t)(a_mn & 1) << 15; // A major
  d |= (uint32_t)(b_mn & 1) << 16; // B major
  d |= (uint32_t)(N >> 3) << 17;   // N / 8
  d |= (uint32_t)(M >> 4) << 24;   // M / 16
  return d;
}

// ---- bf16 split helpers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// x ≈ hi + lo with hi = bf16(x), lo = bf16(x − hi): 16 mantissa bits in total.  Packed conversions only
// (F2FP.BF16.PACK_AB, full-rate ALU) — the scalar F2F path is quarter rate and dominated the gather warps.
__device__ __forceinline__ void split4(const float4& x, uint2& hi, uint2& lo) {
  const uint32_t h01 = pack_bf16x2(x.x, x.y), h23 = pack_bf16x2(x.z, x.w);
  hi.x = h01;
  hi.y = h23;
  const float r0 = x.x - __uint_as_float(h01 << 16), r1 = x.y - __uint_as_float(h01 & 0xffff0000u);
  const float r2 = x.z - __uint_as_float(h23 << 16), r3 = x.w - __uint_as_float(h23 & 0xffff0000u);
  lo.x = pack_bf16x2(r0, r1);
  lo.y = pack_bf16x2(r2, r3);
}
__device__ __forceinline__ uint2 to_bf16x4(const float4& x) { return make_uint2(pack_bf16x2(x.x, x.y), pack_bf16x2(x.z, x.w)); }

}  // namespace umma
