// Packed-half (f16x2) helpers on raw 32-bit registers, shared by the kernels that run their per-element work on half2 pairs
// (attn_h16.cu, the FFN epilogues of ffn_fused.cu).  A value is a uint32_t holding (lo, hi) halves.
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>

namespace b200ctr {

// log2(e) / sqrt(head_dim = 8): q is carried times this factor (scores in log2 units); one constant for the kernels that form it
constexpr float H2_QSCALE = 0.51006972789764404297f;

__device__ __forceinline__ uint32_t h2_pack(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t h2_add(uint32_t a, uint32_t b) { uint32_t r; asm("add.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_mul(uint32_t a, uint32_t b) { uint32_t r; asm("mul.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_max(uint32_t a, uint32_t b) { uint32_t r; asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_ex2(uint32_t a) { uint32_t r; asm("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a)); return r; }
__device__ __forceinline__ float h2_lo(uint32_t a) { return __half2float(__ushort_as_half((unsigned short)(a & 0xffffu))); }
__device__ __forceinline__ float h2_hi(uint32_t a) { return __half2float(__ushort_as_half((unsigned short)(a >> 16))); }
__device__ __forceinline__ uint32_t h2_swap(uint32_t a) { return __byte_perm(a, a, 0x1032); }
__device__ __forceinline__ uint32_t h2_dup_lo(uint32_t a) { return __byte_perm(a, a, 0x1010); }
__device__ __forceinline__ uint32_t h2_dup_hi(uint32_t a) { return __byte_perm(a, a, 0x3232); }
// 0xffff per half from the flags at bits 15 and 31
__device__ __forceinline__ uint32_t h2_mask_from_signs(uint32_t y) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(y), "r"(0u), "r"(0xBB99u));
  return r;
}

__device__ __forceinline__ uint32_t h2_fma(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
__device__ __forceinline__ uint32_t h2_min(uint32_t a, uint32_t b) { uint32_t r; asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t h2_tanh(uint32_t a) { uint32_t r; asm("tanh.approx.f16x2 %0, %1;" : "=r"(r) : "r"(a)); return r; }
// both halves = c (folded at compile time for a literal)
__device__ __forceinline__ uint32_t h2_const(float c) {
  const __half2 v = __float2half2_rn(c);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// both halves = a run-time value
__device__ __forceinline__ uint32_t h2_const_rt(float c) { return h2_pack(c, c); }
// erf-GELU on a pair: x * Phi(x), Phi(x) ~= 0.5 + 0.5 tanh(x (a + b u + c u^2)), u = min(x^2, 64) -- the sigmoid fit of
// ffn_fused.cu (|error| < 4.2e-5 in fp32) with its coefficients halved for the tanh form; in half arithmetic the result is
// good to ~5e-4 relative for |f| > 0.1 and ~1e-3 absolute in the negative tail (bf16, its consumer's format, resolves 4e-3)
__device__ __forceinline__ uint32_t h2_gelu(uint32_t x) {
  const uint32_t u = h2_min(h2_mul(x, x), h2_const(64.f));
  uint32_t q = h2_fma(u, h2_const(-3.2625e-4f), h2_const(0.036835f));
  q = h2_fma(u, q, h2_const(0.797687f));
  const uint32_t t = h2_tanh(h2_mul(x, q));
  const uint32_t hx = h2_mul(x, h2_const(0.5f));
  return h2_fma(hx, t, hx);
}

}  // namespace b200ctr
