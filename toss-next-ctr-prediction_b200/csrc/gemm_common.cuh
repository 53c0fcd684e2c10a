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
  if (e.drop.p > 0.f) v *= drop_scale(e.drop, (uint64_t)(m * N + n));
  if (e.mask_mode == B200CTR_MASK_NONZERO) v = (e.mask_src[m * e.ld_mask + n] != 0.f) ? v * e.mask_scale : 0.f;
  else if (e.mask_mode == B200CTR_MASK_GELU_GRAD) v *= gelu_erf_grad(e.mask_src[m * e.ld_mask + n]);
  if (e.residual) v += e.residual[m * e.ld_res + n];
  return v;
}

// splits written by either contraction kernel -> C with the fused epilogue (defined in gemm.cu)
int launch_splitk_reduce(const float* partial, int split_k, int64_t M, int64_t N, float* C, int64_t ldc, const EpiDev& epi,
                         cudaStream_t st);

}  // namespace b200ctr
