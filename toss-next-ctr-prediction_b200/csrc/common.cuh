// Shared device/host helpers for libb200ctr (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/b200ctr.h"

namespace b200ctr {

// ---- error plumbing (thread-local message, never throws, never exits) ------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define B200_REQUIRE(cond, code, ...)                 \
  do {                                                \
    if (!(cond)) {                                    \
      ::b200ctr::set_error(__VA_ARGS__);              \
      return (code);                                  \
    }                                                 \
  } while (0)

// launch-config errors are picked up without a device sync
#define B200_LAUNCH_CHECK(name)                                        \
  do {                                                                 \
    cudaError_t e__ = cudaPeekAtLastError();                           \
    if (e__ != cudaSuccess) return ::b200ctr::cuda_fail(e__, name);    \
  } while (0)

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }
__host__ __device__ static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
int num_sms();

// ---- counter-based dropout RNG -----------------------------------------------------------------------
// u32 hash of (seed, stream_id, element index): two rounds of a multiply-xorshift mixer (lowbias32 constants)
// around the key injections.  Stateless, so the backward pass regenerates the mask.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t rng_u32(const b200ctr_dropout_t& d, uint64_t idx) {
  uint32_t lo = (uint32_t)idx, hi = (uint32_t)(idx >> 32);
  uint32_t x = mix32(lo ^ d.seed_lo);
  x = mix32(x + 0x9E3779B9U * (d.stream_id + 1u) + (hi ^ d.seed_hi) * 0x85EBCA6BU);
  return x;
}
// keep-scale: 1/(1-p) if kept, 0 if dropped.  p == 0 short-circuits (exact identity).
// One 32-bit hash serves the two elements of an aligned pair (idx >> 1): bits 0-14 decide the even element, bits 16-30 the
// odd one, each against a 15-bit threshold (p is honoured to 2^-15).  Bits 15 and 31 are unused on purpose: a 15-bit
// field is a NON-NEGATIVE fp16 bit pattern, whose integer order is its float order, so the packed-half kernels
// (attn_h16.cu) turn a pair of fields into a pair of keep masks with one AND and one HSET2 (drop_keep_mask2).
__device__ __forceinline__ uint32_t drop_thr15(const b200ctr_dropout_t& d) { return min(__float2uint_rn(d.p * 32768.f), 0x7c00u); }
__device__ __forceinline__ bool drop_keep_lo(uint32_t h, uint32_t thr) { return (h & 0x7fffu) >= thr; }
__device__ __forceinline__ bool drop_keep_hi(uint32_t h, uint32_t thr) { return (h & 0x7fff0000u) >= (thr << 16); }
// 0xffff in each half whose field is >= its threshold (thr2 = thr | thr << 16).  Fields 0x7c01..0x7fff are fp16 NaNs:
// the unordered-or-greater compare keeps them (thr <= 0x7c00 always), exactly like the integer compare.
__device__ __forceinline__ uint32_t drop_keep_mask2(uint32_t h, uint32_t thr2) {
  uint32_t m;
  asm("set.geu.u32.f16x2 %0, %1, %2;" : "=r"(m) : "r"(h & 0x7fff7fffu), "r"(thr2));
  return m;
}
__device__ __forceinline__ float drop_scale(const b200ctr_dropout_t& d, uint64_t idx) {
  if (d.p <= 0.f) return 1.f;
  const uint32_t h = rng_u32(d, idx >> 1), thr = drop_thr15(d);
  return ((idx & 1) ? drop_keep_hi(h, thr) : drop_keep_lo(h, thr)) ? 1.f / (1.f - d.p) : 0.f;
}
// both elements of the pair starting at the EVEN index idx0 (same values as two drop_scale calls)
__device__ __forceinline__ void drop_scale2(const b200ctr_dropout_t& d, uint64_t idx0, float& s0, float& s1) {
  const uint32_t h = rng_u32(d, idx0 >> 1), thr = drop_thr15(d);
  const float inv = 1.f / (1.f - d.p);
  s0 = drop_keep_lo(h, thr) ? inv : 0.f;
  s1 = drop_keep_hi(h, thr) ? inv : 0.f;
}

// Row/column form used by the 2-D hot loops (attention probabilities, FFN / MLP activations): a full-strength
// key per row (rng_u32 of the 64-bit row id, computed once per row) and ONE mixer round per aligned column pair.
struct DropCtx {
  uint32_t thr;     // 15-bit keep threshold
  float inv;        // 1/(1-p)
  bool on;
};
__device__ __forceinline__ DropCtx drop_ctx(const b200ctr_dropout_t& d) {
  DropCtx c;
  c.on = d.p > 0.f;
  c.thr = c.on ? drop_thr15(d) : 0u;
  c.inv = c.on ? 1.f / (1.f - d.p) : 1.f;
  return c;
}
__device__ __forceinline__ uint32_t drop_row_key(const b200ctr_dropout_t& d, uint64_t row) { return rng_u32(d, row); }
// the random bits of one aligned column pair: bits 0-14 -> column 2*colpair, bits 16-30 -> column 2*colpair+1.
// The row key already is a full-strength hash, so the per-pair step is ONE multiply-xorshift-multiply round of
// key ^ (colpair * golden) instead of the full mix32 (6 instead of 10 instructions in every dropout hot loop).  Keep rates
// and all pairwise column / half / row correlations of the masks were measured at the sampling-noise level (3e5 rows x
// 384 columns, p = 0.1 / 0.2 / 0.5); a single multiply-xorshift or a Weyl ADD in front of the round is NOT enough
// (1.5 - 7 % correlations between specific columns).
// (_pre: the caller supplies colpair * golden, e.g. a per-thread term plus a compile-time constant per unrolled tile)
__device__ __forceinline__ uint32_t drop_pair_hash_pre(uint32_t key, uint32_t colpair_golden) {
  uint32_t x = key ^ colpair_golden;
  x *= 0x7feb352dU;
  x ^= x >> 15;
  return x * 0x846ca68bU;
}
__device__ __forceinline__ uint32_t drop_pair_hash(uint32_t key, uint32_t colpair) { return drop_pair_hash_pre(key, colpair * 0x9E3779B9U); }
// keep-scales of columns 2*colpair and 2*colpair+1 of the row with key `key`
__device__ __forceinline__ void drop_pair(const DropCtx& c, uint32_t key, uint32_t colpair, float& s0, float& s1) {
  const uint32_t h = drop_pair_hash(key, colpair);
  s0 = drop_keep_lo(h, c.thr) ? c.inv : 0.f;
  s1 = drop_keep_hi(h, c.thr) ? c.inv : 0.f;
}
__device__ __forceinline__ float drop_rc(const DropCtx& c, uint32_t key, uint32_t col) {
  const uint32_t h = drop_pair_hash(key, col >> 1);
  return ((col & 1) ? drop_keep_hi(h, c.thr) : drop_keep_lo(h, c.thr)) ? c.inv : 0.f;
}

// ---- math ------------------------------------------------------------------------------------------------
// erf-GELU (nn.GELU() default) and its derivative.  erf(|x|/sqrt2) = 1 - poly(t) * exp(-x^2/2), t = 1/(1 + p|x|/sqrt2)
// (Abramowitz-Stegun 7.1.26, |error| < 1.5e-7): one exp shared between the cdf and the pdf, no erff() call.
__device__ __forceinline__ void gelu_parts(float x, float& cdf, float& ex) {
  const float ax = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.f, fmaf(0.3275911f, ax, 1.f));
  ex = __expf(-0.5f * x * x);
  float poly = fmaf(t, 1.061405429f, -1.453152027f);
  poly = fmaf(t, poly, 1.421413741f);
  poly = fmaf(t, poly, -0.284496736f);
  poly = fmaf(t, poly, 0.254829592f);
  const float erf_abs = 1.f - poly * t * ex;
  cdf = 0.5f * (1.f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float gelu_erf(float x) {
  float cdf, ex;
  gelu_parts(x, cdf, ex);
  return x * cdf;
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float cdf, ex;
  gelu_parts(x, cdf, ex);
  return fmaf(x * 0.39894228040143268f, ex, cdf);
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
// torch softplus(beta=1, threshold=20)
__device__ __forceinline__ float softplusf_(float x) { return x > 20.f ? x : log1pf(expf(x)); }


// Cross-CTA partial sums are reduced by (32, 8) blocks: x = 32 consecutive outputs (coalesced 128-byte reads per partial
// row), y = 8 interleaved slices of the partial index, summed through shared memory.  One thread per output with a serial
// loop over ~150 partials left these tiny kernels latency-bound at 18-30 us each (0.36 ms per training step in total).
constexpr int RED_TX = 32, RED_TY = 8;
__device__ __forceinline__ float red_finish_32x8(float a) {          // returns the total in the threads with threadIdx.y == 0
  __shared__ float red_[RED_TY][RED_TX + 1];
  red_[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0) {
#pragma unroll
    for (int y = 1; y < RED_TY; ++y) a += red_[y][threadIdx.x];
  }
  return a;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `red` is >= 32 floats of shared memory; result valid in every thread
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.f;
  if (w == 0) t = warp_sum(t);
  if (threadIdx.x == 0) red[0] = t;
  __syncthreads();
  return red[0];
}

// ---- order-independent accumulation ------------------------------------------------------------------------------
// The squared-gradient-norm accumulator is an 8-byte slot that many blocks add to.  Floating-point atomics would make the sum
// depend on the order in which the blocks retire; instead every block rounds its partial to a 2^-32 grid and adds it as a
// 64-bit INTEGER (associative), so the global norm -- and with it the clip coefficient and every parameter after the step --
// is bit-identical from run to run.  Range 2^31 (a squared norm of 2e9), resolution 2.3e-10.  The slot is read back with
// norm_acc_value(); across ranks it is all-reduced as int64.
constexpr double NORM_ACC_SCALE = 4294967296.0;
__device__ __forceinline__ void norm_acc_add(double* acc, double partial) {
  atomicAdd(reinterpret_cast<unsigned long long*>(acc), (unsigned long long)__double2ll_rn(partial * NORM_ACC_SCALE));
}
__device__ __forceinline__ double norm_acc_value(const double* acc) {
  return (double)(*reinterpret_cast<const long long*>(acc)) * (1.0 / NORM_ACC_SCALE);
}

// 128-bit streaming loads (read-only path, no L1 allocation) for gather rows
__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

}  // namespace b200ctr
