// K4: fused class-balanced BCE (+aux) with its gradient, global-norm clipping, fused AdamW + EMA, and the
// bookkeeping for row-sparse gradients of the big sequence tables.
// Reference: src/train.py:71-90 (bce_wll_style), 165-168, 193-199; torch.optim.AdamW; src/utils/ema.py:91-131;
//            src/infer.py:122-158 + src/utils/metrics.py:68-71 (logit_mean).
#include "common.cuh"
#include <algorithm>

namespace b200ctr {

// single CTA: counts, the two class-conditional means, gradients, probabilities.
__global__ void __launch_bounds__(1024)
bce_wll_kernel(const float* __restrict__ logits, const float* __restrict__ aux, const float* __restrict__ y, float aux_w,
               int64_t B, const int32_t* __restrict__ counts_override, float* __restrict__ loss, int32_t* __restrict__ counts,
               float* __restrict__ dlogits, float* __restrict__ daux, float* __restrict__ prob) {
  __shared__ float red[32];
  float np = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) np += (y[i] > 0.5f) ? 1.f : 0.f;
  np = block_sum(np, red);
  int n_pos = (int)(np + 0.5f), n_neg = (int)B - n_pos;
  if (counts && threadIdx.x == 0) { counts[0] = n_pos; counts[1] = n_neg; }
  if (counts_override) { n_pos = counts_override[0]; n_neg = counts_override[1]; }
  const float ip = n_pos > 0 ? 1.f / (float)n_pos : 0.f, in = n_neg > 0 ? 1.f / (float)n_neg : 0.f;
  float lp = 0.f, ln = 0.f, ap = 0.f, an = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    const bool pos = y[i] > 0.5f;
    const float z = logits[i];
    if (pos) { lp += softplusf_(-z); dlogits[i] = -0.5f * sigmoidf_(-z) * ip; }
    else     { ln += softplusf_(z);  dlogits[i] = 0.5f * sigmoidf_(z) * in; }
    if (prob) prob[i] = sigmoidf_(z);
    if (aux && daux) {
      if (aux_w > 0.f) {
        const float a = aux[i];
        if (pos) { ap += softplusf_(-a); daux[i] = -0.5f * aux_w * sigmoidf_(-a) * ip; }
        else     { an += softplusf_(a);  daux[i] = 0.5f * aux_w * sigmoidf_(a) * in; }
      } else daux[i] = 0.f;
    }
  }
  lp = block_sum(lp, red); ln = block_sum(ln, red); ap = block_sum(ap, red); an = block_sum(an, red);
  if (threadIdx.x == 0) loss[0] = 0.5f * (lp * ip + ln * in) + aux_w * 0.5f * (ap * ip + an * in);
}

__global__ void __launch_bounds__(256)
sqnorm_kernel(const float* __restrict__ x, int64_t n, double* __restrict__ acc) {
  __shared__ double red[8];
  double s = 0.0;
  int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  const bool al = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  for (; i < n; i += stride) {
    float p = 0.f;
    if (al && i + 3 < n) { float4 v = *reinterpret_cast<const float4*>(x + i); p = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w; }
    else for (int k = 0; k < 4 && i + k < n; ++k) p = fmaf(x[i + k], x[i + k], p);
    s += (double)p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    norm_acc_add(acc, t);
  }
}

__global__ void clip_coef_kernel(const double* __restrict__ acc, float max_norm, float* __restrict__ coef, float* __restrict__ total) {
  double tn = sqrt(norm_acc_value(acc));
  if (total) total[0] = (float)tn;
  float c = 1.f;
  if (max_norm > 0.f) { c = max_norm / ((float)tn + 1e-6f); c = fminf(c, 1.f); }
  coef[0] = c;
}

__global__ void __launch_bounds__(256)
adamw_ema_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                 float* __restrict__ ema, int64_t n, float lr, float wd, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                 const float* __restrict__ coef, float ema_decay) {
  const float c = coef ? coef[0] : 1.f;
  const float step_size = lr / bc1;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    float gi = (g ? g[i] : 0.f) * c;
    float pi = p[i] * (1.f - lr * wd);
    float mi = m[i] + (gi - m[i]) * (1.f - b1);        // lerp_(g, 1-b1)
    float vi = v[i] * b2 + (1.f - b2) * gi * gi;
    float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * (mi / denom);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (ema) ema[i] = ema[i] * ema_decay + pi * (1.f - ema_decay);
  }
}

// Multi-tensor form for the UNMODIFIED reference loop (Tier A: 138 separate parameter tensors, torch.optim.AdamW semantics,
// utils/ema.py:91-131): one launch walks a chunk table instead of 3-4 launches per tensor.  A tensor record is six 64-bit
// words {p, g, m, v, ema, n}; mode 0 = AdamW (+ EMA when the record has a shadow), mode 1 = EMA only (p is read).
struct MtRec { float* p; const float* g; float* m; float* v; float* ema; int64_t n; };
__global__ void __launch_bounds__(256)
multi_adamw_ema_kernel(const MtRec* __restrict__ recs, const int32_t* __restrict__ chunk_tensor, const int64_t* __restrict__ chunk_off,
                       int64_t n_chunks, int64_t chunk, float lr, float wd, float omb1, float b2, float omb2, float eps, float bc1,
                       float bc2_sqrt, const float* __restrict__ coef, float ema_decay, int mode) {
  const float c = coef ? coef[0] : 1.f;
  const float step_size = lr / bc1;
  for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
    const MtRec r = recs[chunk_tensor[ch]];
    const int64_t i0 = chunk_off[ch], i1 = min(i0 + chunk, r.n);
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
      float pi = r.p[i];
      if (mode == 0) {
        const float gi = r.g[i] * c;
        pi *= (1.f - lr * wd);
        // torch: exp_avg.lerp_(g, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(g, g, value=1 - beta2) with the weights formed in
        // double on the host (1.f - 0.999f is 1.3e-5 away from float(1 - 0.999))
        const float mi = r.m[i] + (gi - r.m[i]) * omb1;
        const float vi = r.v[i] * b2 + omb2 * gi * gi;
        pi -= step_size * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        r.p[i] = pi; r.m[i] = mi; r.v[i] = vi;
      }
      if (r.ema) r.ema[i] = r.ema[i] * ema_decay + pi * (1.f - ema_decay);
    }
  }
}

// ---- row-sparse bookkeeping: one representative slot per distinct token ---------------------------------
__global__ void rows_claim_kernel(const int32_t* __restrict__ tok, int64_t n, int pad_id, int32_t* __restrict__ owner) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int t = tok[i];
  if (t != pad_id) owner[t] = (int32_t)i;           // racing writers: exactly one value survives
}
__global__ void __launch_bounds__(256)
rows_sqnorm_kernel(const int32_t* __restrict__ tok, int64_t n, int pad_id, const int32_t* __restrict__ owner,
                   const float* __restrict__ g, int D, double* __restrict__ acc) {
  __shared__ double red[8];
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  double s = 0.0;
  for (; warp < n; warp += nw) {
    int t = tok[warp];
    if (t == pad_id || owner[t] != (int32_t)warp) continue;
    float p = 0.f;
    for (int d = lane; d < D; d += 32) { float v = g[(int64_t)t * D + d]; p = fmaf(v, v, p); }
    s += (double)p;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    norm_acc_add(acc, t);
  }
}
__global__ void rows_zero_kernel(const int32_t* __restrict__ tok, int64_t n, int pad_id, const int32_t* __restrict__ owner,
                                 float* __restrict__ g, int D) {
  const int lane = threadIdx.x & 31;
  int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (warp >= n) return;
  int t = tok[warp];
  if (t == pad_id || owner[t] != (int32_t)warp) return;
  for (int d = lane; d < D; d += 32) g[(int64_t)t * D + d] = 0.f;
}

__global__ void logit_mean_kernel(const float* __restrict__ probs, int n_models, int64_t B, float eps, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float s = 0.f;
  for (int m = 0; m < n_models; ++m) {
    float p = fminf(fmaxf(probs[(int64_t)m * B + i], eps), 1.f - eps);
    s += logf(p) - log1pf(-p);
  }
  out[i] = sigmoidf_(s / (float)n_models);
}

// infer.py:108-158 for logit_mean: calibrate (temperature), clip, logit, mean over folds, sigmoid -- one pass over (M, B)
__global__ void ensemble_logit_mean_kernel(const float* __restrict__ logits, const float* __restrict__ inv_T, int n_models,
                                           int64_t B, float eps, float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  float s = 0.f;
  for (int m = 0; m < n_models; ++m) {
    float z = logits[(int64_t)m * B + i];
    if (inv_T) z = fminf(fmaxf(z * inv_T[m], -50.f), 50.f);            // calibration._sigmoid_stable_numpy clips the scaled logit
    float p = fminf(fmaxf(sigmoidf_(z), eps), 1.f - eps);
    s += logf(p) - log1pf(-p);
  }
  out[i] = sigmoidf_(s / (float)n_models);
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_ensemble_logit_mean(const float* logits, const float* inv_T, int n_models, int64_t B, float eps, float* out,
                                           void* stream) {
  B200_REQUIRE(logits && out && n_models >= 1, B200CTR_E_INVALID_ARG, "ensemble_logit_mean: bad argument");
  if (B == 0) return B200CTR_OK;
  ensemble_logit_mean_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, as_stream(stream)>>>(logits, inv_T, n_models, B, eps, out);
  B200_LAUNCH_CHECK("ensemble_logit_mean_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_bce_wll(const float* logits, const float* aux, const float* y, float aux_w, int64_t B,
                               const int32_t* counts_override, float* loss, int32_t* counts, float* dlogits, float* daux,
                               float* prob, void* stream) {
  B200_REQUIRE(logits && y && loss && dlogits, B200CTR_E_INVALID_ARG, "bce_wll: null pointer");
  B200_REQUIRE(B < (1LL << 24), B200CTR_E_UNSUPPORTED_SHAPE, "bce_wll: batch too large for the single-CTA kernel");
  bce_wll_kernel<<<1, 1024, 0, as_stream(stream)>>>(logits, aux, y, aux_w, B, counts_override, loss, counts, dlogits, daux, prob);
  B200_LAUNCH_CHECK("bce_wll_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_sqnorm(const float* x, int64_t n, double* acc, void* stream) {
  B200_REQUIRE(x && acc, B200CTR_E_INVALID_ARG, "sqnorm: null pointer");
  if (n == 0) return B200CTR_OK;
  int grid = (int)std::min<int64_t>(ceil_div(n, 1024), (int64_t)num_sms() * 8);
  sqnorm_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, n, acc);
  B200_LAUNCH_CHECK("sqnorm_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_clip_coef(const double* acc, float max_norm, float* coef, float* total_norm, void* stream) {
  B200_REQUIRE(acc && coef, B200CTR_E_INVALID_ARG, "clip_coef: null pointer");
  clip_coef_kernel<<<1, 1, 0, as_stream(stream)>>>(acc, max_norm, coef, total_norm);
  B200_LAUNCH_CHECK("clip_coef_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_adamw_ema(float* p, const float* g, float* m, float* v, float* ema, int64_t n, float lr, float wd,
                                 float beta1, float beta2, float eps, int64_t step, const float* coef, float ema_decay,
                                 void* stream) {
  B200_REQUIRE(p && m && v, B200CTR_E_INVALID_ARG, "adamw_ema: null pointer");
  B200_REQUIRE(step >= 1, B200CTR_E_INVALID_ARG, "adamw_ema: step must be >= 1");
  if (n == 0) return B200CTR_OK;
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  int grid = (int)std::min<int64_t>(ceil_div(n, 256), (int64_t)num_sms() * 16);
  adamw_ema_kernel<<<grid, 256, 0, as_stream(stream)>>>(p, g, m, v, ema, n, lr, wd, beta1, beta2, eps, (float)bc1,
                                                        (float)sqrt(bc2), coef, ema_decay);
  B200_LAUNCH_CHECK("adamw_ema_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_multi_adamw_ema(const void* recs, const int32_t* chunk_tensor, const int64_t* chunk_off, int64_t n_chunks,
                                       int64_t chunk, float lr, float wd, double beta1, double beta2, float eps, int64_t step,
                                       const float* coef, float ema_decay, int mode, void* stream) {
  B200_REQUIRE(recs && chunk_tensor && chunk_off && chunk > 0, B200CTR_E_INVALID_ARG, "multi_adamw_ema: null pointer");
  B200_REQUIRE(mode == 0 || mode == 1, B200CTR_E_INVALID_ARG, "multi_adamw_ema: mode %d", mode);
  B200_REQUIRE(mode == 1 || step >= 1, B200CTR_E_INVALID_ARG, "multi_adamw_ema: step must be >= 1");
  if (n_chunks == 0) return B200CTR_OK;
  const double bc1 = mode == 0 ? 1.0 - pow(beta1, (double)step) : 1.0;
  const double bc2 = mode == 0 ? 1.0 - pow(beta2, (double)step) : 1.0;
  const int grid = (int)std::min<int64_t>(n_chunks, (int64_t)num_sms() * 8);
  multi_adamw_ema_kernel<<<grid, 256, 0, as_stream(stream)>>>((const MtRec*)recs, chunk_tensor, chunk_off, n_chunks, chunk, lr, wd,
                                                              (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), eps,
                                                              (float)bc1, (float)sqrt(bc2), coef, ema_decay, mode);
  B200_LAUNCH_CHECK("multi_adamw_ema_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_rows_claim(const int32_t* tok, int64_t n_slots, int pad_id, int32_t* owner, void* stream) {
  B200_REQUIRE(tok && owner, B200CTR_E_INVALID_ARG, "rows_claim: null pointer");
  if (n_slots == 0) return B200CTR_OK;
  rows_claim_kernel<<<(unsigned)ceil_div(n_slots, 256), 256, 0, as_stream(stream)>>>(tok, n_slots, pad_id, owner);
  B200_LAUNCH_CHECK("rows_claim_kernel");
  return B200CTR_OK;
}
extern "C" int b200ctr_rows_sqnorm(const int32_t* tok, int64_t n_slots, int pad_id, const int32_t* owner, const float* g_dense,
                                   int D, double* acc, void* stream) {
  B200_REQUIRE(tok && owner && g_dense && acc, B200CTR_E_INVALID_ARG, "rows_sqnorm: null pointer");
  if (n_slots == 0) return B200CTR_OK;
  int grid = (int)std::min<int64_t>(ceil_div(n_slots, 8), (int64_t)num_sms() * 16);
  rows_sqnorm_kernel<<<grid, 256, 0, as_stream(stream)>>>(tok, n_slots, pad_id, owner, g_dense, D, acc);
  B200_LAUNCH_CHECK("rows_sqnorm_kernel");
  return B200CTR_OK;
}
extern "C" int b200ctr_rows_zero(const int32_t* tok, int64_t n_slots, int pad_id, const int32_t* owner, float* g_dense, int D,
                                 void* stream) {
  B200_REQUIRE(tok && owner && g_dense, B200CTR_E_INVALID_ARG, "rows_zero: null pointer");
  if (n_slots == 0) return B200CTR_OK;
  rows_zero_kernel<<<(unsigned)ceil_div(n_slots, 8), 256, 0, as_stream(stream)>>>(tok, n_slots, pad_id, owner, g_dense, D);
  B200_LAUNCH_CHECK("rows_zero_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_logit_mean(const float* probs, int n_models, int64_t B, float eps, float* out, void* stream) {
  B200_REQUIRE(probs && out && n_models >= 1, B200CTR_E_INVALID_ARG, "logit_mean: bad argument");
  if (B == 0) return B200CTR_OK;
  logit_mean_kernel<<<(unsigned)ceil_div(B, 256), 256, 0, as_stream(stream)>>>(probs, n_models, B, eps, out);
  B200_LAUNCH_CHECK("logit_mean_kernel");
  return B200CTR_OK;
}
