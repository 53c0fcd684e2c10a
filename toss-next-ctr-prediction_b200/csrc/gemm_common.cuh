// Shared by gemm.cu (fp32 CUDA-core path) and gemm_tc.cu (tcgen05 bf16 path): epilogue descriptor + epilogue math.
#pragma once
#include "common.cuh"

namespace b200ctr {

struct EpiDev {
  float alpha;
  const float* bias;
  int act;
  float* pre_out; int64_t ld_pre;
  b200ctr_dropout_t drop;
  int mask_mode; const float* mask_src; int64_t ld_mask; float mask_scale;
  const float* residual; int64_t ld_res;
  const float* accum; int64_t ld_accum;
};

__device__ __forceinline__ float apply_epi(float acc, int64_t m, int64_t n, int64_t N, const EpiDev& e) {
  float v = e.alpha * acc;
  if (e.accum) v += e.accum[m * e.ld_accum + n];
  if (e.bias) v += e.bias[n];
  if (e.pre_out) e.pre_out[m * e.ld_pre + n] = v;
  if (e.act == B200CTR_ACT_RELU) v = fmaxf(v, 0.f);
  else if (e.act == B200CTR_ACT_GELU) v = gelu_erf(v);
  if (e.drop.p > 0.f) v *= drop_rc(drop_ctx(e.drop), drop_row_key(e.drop, (uint64_t)m), (uint32_t)n);
  if (e.mask_mode == B200CTR_MASK_NONZERO) v = (e.mask_src[m * e.ld_mask + n] != 0.f) ? v * e.mask_scale : 0.f;
  else if (e.mask_mode == B200CTR_MASK_GELU_GRAD) v *= gelu_erf_grad(e.mask_src[m * e.ld_mask + n]);
  if (e.residual) v += e.residual[m * e.ld_res + n];
  return v;
}

// 4-wide epilogue (n multiple of 4, all row pointers 16-byte aligned: checked by the caller via epi_vec_ok)
struct EpiAux4 { float4 accum, mask, res; };
__device__ __forceinline__ void epi_load4(EpiAux4& a, int64_t m, int64_t n, const EpiDev& e) {
  if (e.accum) a.accum = *reinterpret_cast<const float4*>(e.accum + m * e.ld_accum + n);
  if (e.mask_mode != B200CTR_MASK_NONE) a.mask = *reinterpret_cast<const float4*>(e.mask_src + m * e.ld_mask + n);
  if (e.residual) a.res = *reinterpret_cast<const float4*>(e.residual + m * e.ld_res + n);
}
__device__ __forceinline__ float4 apply_epi4(float4 acc, const EpiAux4& a, int64_t m, int64_t n, int64_t N, const EpiDev& e) {
  float v[4] = {e.alpha * acc.x, e.alpha * acc.y, e.alpha * acc.z, e.alpha * acc.w};
  if (e.accum) { v[0] += a.accum.x; v[1] += a.accum.y; v[2] += a.accum.z; v[3] += a.accum.w; }
  if (e.bias) {
    float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n));
    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
  }
  if (e.pre_out) *reinterpret_cast<float4*>(e.pre_out + m * e.ld_pre + n) = make_float4(v[0], v[1], v[2], v[3]);
  if (e.act == B200CTR_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = fmaxf(v[j], 0.f);
  } else if (e.act == B200CTR_ACT_GELU) {
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
  }
  if (e.drop.p > 0.f) {          // n % 4 == 0: two aligned column pairs of row m
    const DropCtx dc = drop_ctx(e.drop);
    const uint32_t key = drop_row_key(e.drop, (uint64_t)m);
    float s0, s1, s2, s3;
    drop_pair(dc, key, (uint32_t)n >> 1, s0, s1);
    drop_pair(dc, key, ((uint32_t)n >> 1) + 1, s2, s3);
    v[0] *= s0; v[1] *= s1; v[2] *= s2; v[3] *= s3;
  }
  if (e.mask_mode == B200CTR_MASK_NONZERO) {
    const float mk[4] = {a.mask.x, a.mask.y, a.mask.z, a.mask.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = (mk[j] != 0.f) ? v[j] * e.mask_scale : 0.f;
  } else if (e.mask_mode == B200CTR_MASK_GELU_GRAD) {
    const float mk[4] = {a.mask.x, a.mask.y, a.mask.z, a.mask.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] *= gelu_erf_grad(mk[j]);
  }
  if (e.residual) { v[0] += a.res.x; v[1] += a.res.y; v[2] += a.res.z; v[3] += a.res.w; }
  return make_float4(v[0], v[1], v[2], v[3]);
}
__host__ __device__ inline bool ptr_al16(const void* p, int64_t ld) { return p == nullptr || (((uintptr_t)p & 15) == 0 && (ld & 3) == 0); }
inline bool epi_vec_ok(const EpiDev& e, const float* C, int64_t ldc, int64_t N) {
  return (N & 3) == 0 && ptr_al16(C, ldc) && ptr_al16(e.bias, 0) && ptr_al16(e.pre_out, e.ld_pre) &&
         ptr_al16(e.mask_src, e.ld_mask) && ptr_al16(e.residual, e.ld_res) && ptr_al16(e.accum, e.ld_accum);
}

// splits written by either contraction kernel -> C with the fused epilogue (defined in gemm.cu)
int launch_splitk_reduce(const float* partial, int split_k, int64_t M, int64_t N, float* C, int64_t ldc, const EpiDev& epi,
                         cudaStream_t st);

}  // namespace b200ctr
