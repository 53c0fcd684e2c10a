// tcgen05 / TMEM / mbarrier PTX wrappers shared by the tensor-core kernels (gemm_tc.cu, ffn_fused.cu).
#pragma once
#include "common.cuh"

namespace b200ctr {

// ---- PTX wrappers -------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// bounded spin: a wrong phase would otherwise hang the GPU; trap instead (the host sees a CUDA error)
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
  }
  asm volatile("trap;");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
// accumulate flag known at compile time: no predicate set-up in the issuer's instruction stream
template <bool ACC>
__device__ __forceinline__ void umma_bf16_c(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc) {
  if (ACC)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, 1, 1;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc) : "memory");
}
// descriptor of the same operand `bytes` further on (start-address field only: bits 0-13 of the low word, no carry for
// addresses inside the 227 KB of shared memory)
__device__ __forceinline__ uint64_t desc_advance(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// SWIZZLE_NONE smem matrix descriptor: start>>4 | LBO>>4 @16 | SBO>>4 @32 | version=1 @46
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}


__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// instruction descriptor for kind::f16: D = f32, A = B = bf16; a_mn / b_mn = operand is MN-major in shared memory
__device__ __forceinline__ uint32_t umma_idesc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// same with A = B = fp16 (format field 0)
__device__ __forceinline__ uint32_t umma_idesc_f16(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}


// two 16-column loads in flight, one wait (the second load does not queue behind the first one's completion)
__device__ __forceinline__ void tmem_ld16_pair(uint32_t taddr_a, uint32_t* a, uint32_t taddr_b, uint32_t* b) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]), "=r"(a[4]), "=r"(a[5]), "=r"(a[6]), "=r"(a[7]), "=r"(a[8]), "=r"(a[9]),
        "=r"(a[10]), "=r"(a[11]), "=r"(a[12]), "=r"(a[13]), "=r"(a[14]), "=r"(a[15])
      : "r"(taddr_a) : "memory");
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(b[0]), "=r"(b[1]), "=r"(b[2]), "=r"(b[3]), "=r"(b[4]), "=r"(b[5]), "=r"(b[6]), "=r"(b[7]), "=r"(b[8]), "=r"(b[9]),
        "=r"(b[10]), "=r"(b[11]), "=r"(b[12]), "=r"(b[13]), "=r"(b[14]), "=r"(b[15])
      : "r"(taddr_b) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Row-per-lane epilogues hold one output row per thread; storing it directly makes every STG.128 of a warp hit 32
// different rows (32 half-used sectors per instruction).  warp_store_rows transposes a 32-row x NC-column fp32 block
// through a warp-private shared-memory scratch (32 x (NC+4) floats, conflict-free both ways) so that the global stores
// cover whole rows: NC/4 lanes x 16 bytes contiguous.  v = this lane's NC values; rows_valid = rows of the block to write.
template <int NC>
__device__ __forceinline__ void warp_store_rows(float* scratch, const uint32_t* v, float* gbase, int64_t ld, int rows_valid, int lane) {
  constexpr int SR = NC + 4, LPR = NC / 4, RPI = 32 / LPR;
#pragma unroll
  for (int j = 0; j < NC / 4; ++j)
    *reinterpret_cast<uint4*>(scratch + lane * SR + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int rsub = lane / LPR, c = lane % LPR;
#pragma unroll
  for (int r0 = 0; r0 < 32; r0 += RPI) {
    const int row = r0 + rsub;
    if (row < rows_valid)
      *reinterpret_cast<uint4*>(gbase + (int64_t)row * ld + c * 4) = *reinterpret_cast<const uint4*>(scratch + row * SR + c * 4);
  }
  __syncwarp();
}

// warp_store_rows with a residual: gbase[row] = v[row] + addbase[row], the residual read in the same coalesced pattern
// as the store (row-per-lane 16-byte loads cost one L1 wavefront per lane, whole-row loads one per 128 bytes)
template <int NC>
__device__ __forceinline__ void warp_store_rows_add(float* scratch, const uint32_t* v, float* gbase, int64_t ld,
                                                    const float* __restrict__ addbase, int64_t ld_add, int rows_valid, int lane) {
  constexpr int SR = NC + 4, LPR = NC / 4, RPI = 32 / LPR;
#pragma unroll
  for (int j = 0; j < NC / 4; ++j)
    *reinterpret_cast<uint4*>(scratch + lane * SR + 4 * j) = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int rsub = lane / LPR, c = lane % LPR;
  float4 x[32 / RPI];
#pragma unroll
  for (int i = 0; i < 32 / RPI; ++i) {
    const int row = i * RPI + rsub;
    x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (addbase && row < rows_valid) x[i] = __ldg(reinterpret_cast<const float4*>(addbase + (int64_t)row * ld_add + c * 4));
  }
#pragma unroll
  for (int i = 0; i < 32 / RPI; ++i) {
    const int row = i * RPI + rsub;
    const float4 a = *reinterpret_cast<const float4*>(scratch + row * SR + c * 4);
    if (row < rows_valid)
      *reinterpret_cast<float4*>(gbase + (int64_t)row * ld + c * 4) = make_float4(a.x + x[i].x, a.y + x[i].y, a.z + x[i].z, a.w + x[i].w);
  }
  __syncwarp();
}

}  // namespace b200ctr
