// tcgen05 / TMEM / UMMA shared-memory descriptor helpers (sm_100a), shared by the layer GEMMs (linear_tc.cu) and the
// snapshot-resident kernels (resident2.cu).
#pragma once
#include <stdint.h>
#include "tma.cuh"

namespace gatres {

__device__ __forceinline__ void tc_cp_async16(uint32_t smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xffffe000u); }
// round-to-nearest (ties away) to TF32 with two integer instructions: add half an ulp of the 10-bit mantissa to the
// magnitude bits, drop the low 13 bits.  cvt.rna.tf32.f32 is emulated by ptxas with a ~5-instruction sequence
// (inf/nan guards), which made the per-tile operand split the bottleneck of the staging threads; the operands
// here are finite residuals of magnitude << FLT_MAX, so the guards are not needed.
__device__ __forceinline__ float rna_tf32_fast(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
// lo part of the 3xTF32 split of x whose hi part is the tensor core's own truncation of x
__device__ __forceinline__ float lo_tf32(float x) { return rna_tf32_fast(x - trunc_tf32(x)); }
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {       // one full warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {          // the allocating warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One lane of a fully active warp (elect.sync).  Issue tcgen05.mma / commit under `if (warp == W) if (elect_one())`:
// with a warp-uniform outer branch ptxas keeps descriptors in uniform registers and emits the UTCHMMAs back to back;
// under a divergent `if (threadIdx.x == 0)` every MMA is wrapped in an ELECT / BRA.U.ANY loop with ~10 setup instructions
// (measured: ~55 cycles per MMA, 0.66 us for the 12 MMAs of one projection).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when they complete
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 columns of fp32 accumulator -> 32 registers per thread (thread = lane = row)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}

// MN-major descriptor for 32-bit (tf32) operands.  tcgen05 reads an MN-major tf32 operand only in the
// SWIZZLE_128B_BASE32B layout (cute: Layout_MN_SW128_32B_Atom, "the only available smem layout" for mn-major tf32):
// the tile is [K rows][32 elements = 128 B along M or N], rows 128 B apart, and the 32-byte chunk index (0..3) of a row is
// XORed with row % 4 (Swizzle<2,5,2> on byte addresses); atoms are 4 rows = 512 B.  LBO = byte distance between
// consecutive groups of 32 elements along M / N, SBO = between groups of 4 rows along K (512: dense rows).
__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)(512 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)1 << 61);
}
// 32 lanes x 16 columns of fp32 accumulator -> 16 registers per thread
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 16; ++k) v[k] = __uint_as_float(r[k]);
}
// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) = 1 | SBO>>4 [32,46) = 1024>>4 | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64)
__device__ __forceinline__ uint64_t umma_desc_k128(uint32_t smem_addr) {
  return (uint64_t)((smem_addr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// byte offset of element (row, k) of a [rows x KK] fp32 operand tile: KK/32 blocks of [rows x 128 B]
__device__ __forceinline__ uint32_t swz_off(uint32_t row, uint32_t k, uint32_t rows) {
  const uint32_t atom = k >> 5, kk = k & 31u;
  return atom * rows * 128u + row * 128u + ((((kk >> 2) ^ (row & 7u))) << 4) + ((kk & 3u) << 2);
}

}  // namespace gatres
