// Gated pooling over the K retrieved (and encoded) tokens + DARE aux head.  Reference: src/models/dare.py:151-161.
//   w = softmax(vals) | relu(vals)/(sum+1e-12);  u = dropout(sum_k w_k x_k);  aux = aux_head(u)
// One CTA (128 threads) per sample; the K x 64 tile is read once, coalesced.
#include "common.cuh"

namespace b200ctr {

constexpr int PL_D = 64;
constexpr int PL_THREADS = 128;
constexpr int PL_MAXK = 1024;

__global__ void __launch_bounds__(PL_THREADS)
pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ vals, const float* __restrict__ aux_w,
                const float* __restrict__ aux_b, int gating_relu, int K, float* __restrict__ u, int64_t ld_u,
                float* __restrict__ aux, float* __restrict__ w_out, b200ctr_dropout_t drop) {
  __shared__ float ws[PL_MAXK];
  __shared__ float red[32];
  __shared__ float part[2][PL_D];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  float mx = -INFINITY;
  for (int k = tid; k < K; k += PL_THREADS) { float v = vals[b * K + k]; ws[k] = v; mx = fmaxf(mx, v); }
  // block max
  mx = warp_max(mx);
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
  __syncthreads();
  float sum = 0.f;
  for (int k = tid; k < K; k += PL_THREADS) {
    float e = gating_relu ? fmaxf(ws[k], 0.f) : expf(ws[k] - mx);
    ws[k] = e;
    sum += e;
  }
  sum = block_sum(sum, red);
  const float inv = gating_relu ? 1.f / (sum + 1e-12f) : 1.f / sum;
  for (int k = tid; k < K; k += PL_THREADS) { float w = ws[k] * inv; ws[k] = w; w_out[b * K + k] = w; }
  __syncthreads();
  const int d = tid & 63, half = tid >> 6;
  float acc = 0.f;
  const float* xb = x + b * (int64_t)K * PL_D;
  for (int k = half; k < K; k += 2) acc = fmaf(ws[k], xb[(int64_t)k * PL_D + d], acc);
  part[half][d] = acc;
  __syncthreads();
  float a = 0.f;
  if (tid < PL_D) {
    float uv = (part[0][tid] + part[1][tid]) * drop_scale(drop, (uint64_t)b * PL_D + tid);
    u[b * ld_u + tid] = uv;
    a = uv * aux_w[tid];
  }
  a = block_sum(a, red);
  if (tid == 0) aux[b] = a + aux_b[0];
}

__global__ void __launch_bounds__(PL_THREADS)
pool_bwd_kernel(const float* __restrict__ x, const float* __restrict__ vals, const float* __restrict__ w,
                const float* __restrict__ aux_w, const float* __restrict__ u, int64_t ld_u, const float* __restrict__ du,
                int64_t ld_du, const float* __restrict__ daux, int gating_relu, int K, float* __restrict__ dx,
                float* __restrict__ dvals, float* __restrict__ daux_w, float* __restrict__ daux_b, b200ctr_dropout_t drop) {
  __shared__ float du0[PL_D];
  __shared__ float dws[PL_MAXK];
  __shared__ float red[32];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float da = daux ? daux[b] : 0.f;
  if (tid < PL_D) {
    float g = (du ? du[b * ld_du + tid] : 0.f) + da * aux_w[tid];
    du0[tid] = g * drop_scale(drop, (uint64_t)b * PL_D + tid);
    if (daux_w && daux) daux_w[b * PL_D + tid] = da * u[b * ld_u + tid];     // per-sample partial rows, summed in sample order
  }                                                                          // afterwards (b200ctr_colsum): no float atomics
  if (tid == 0 && daux_b && daux) daux_b[b] = da;
  __syncthreads();
  const float* xb = x + b * (int64_t)K * PL_D;
  float* dxb = dx + b * (int64_t)K * PL_D;
  const float g0 = du0[lane], g1 = du0[lane + 32];
  for (int k = warp; k < K; k += PL_THREADS / 32) {
    float wk = w[b * K + k];
    float x0 = xb[(int64_t)k * PL_D + lane], x1 = xb[(int64_t)k * PL_D + lane + 32];
    dxb[(int64_t)k * PL_D + lane] = wk * g0;
    dxb[(int64_t)k * PL_D + lane + 32] = wk * g1;
    float dwk = warp_sum(x0 * g0 + x1 * g1);
    if (lane == 0) dws[k] = dwk;
  }
  __syncthreads();
  if (gating_relu) {
    // w = wr/s, wr = relu(v), s = sum(wr)+1e-12 : dwr_k = dw_k/s - sum_j(dw_j*wr_j)/s^2
    float ps = 0.f, pdot = 0.f;
    for (int k = tid; k < K; k += PL_THREADS) { float wr = fmaxf(vals[b * K + k], 0.f); ps += wr; pdot = fmaf(dws[k], wr, pdot); }
    float s = block_sum(ps, red) + 1e-12f;
    float dot = block_sum(pdot, red);
    for (int k = tid; k < K; k += PL_THREADS) {
      float v = vals[b * K + k];
      dvals[b * K + k] = v > 0.f ? (dws[k] / s - dot / (s * s)) : 0.f;
    }
  } else {
    float pdot = 0.f;
    for (int k = tid; k < K; k += PL_THREADS) pdot = fmaf(dws[k], w[b * K + k], pdot);
    float dot = block_sum(pdot, red);
    for (int k = tid; k < K; k += PL_THREADS) dvals[b * K + k] = w[b * K + k] * (dws[k] - dot);
  }
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_pool_fwd(const float* x, const float* vals, const float* aux_w, const float* aux_b, int gating_relu,
                                int64_t B, int K, int D, float* u, int64_t ld_u, float* aux, float* w_out,
                                b200ctr_dropout_t drop, void* stream) {
  B200_REQUIRE(x && vals && aux_w && aux_b && u && aux && w_out, B200CTR_E_INVALID_ARG, "pool_fwd: null pointer");
  B200_REQUIRE(D == PL_D, B200CTR_E_UNSUPPORTED_SHAPE, "pool_fwd: D=%d unsupported (64 only)", D);
  B200_REQUIRE(K >= 1 && K <= PL_MAXK, B200CTR_E_UNSUPPORTED_SHAPE, "pool_fwd: K=%d unsupported", K);
  if (B == 0) return B200CTR_OK;
  pool_fwd_kernel<<<(unsigned)B, PL_THREADS, 0, as_stream(stream)>>>(x, vals, aux_w, aux_b, gating_relu, K, u, ld_u, aux, w_out, drop);
  B200_LAUNCH_CHECK("pool_fwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_pool_bwd(const float* x, const float* vals, const float* w, const float* aux_w, const float* u,
                                int64_t ld_u, const float* du, int64_t ld_du, const float* daux, int gating_relu, int64_t B,
                                int K, int D, float* dx, float* dvals, float* daux_w, float* daux_b, b200ctr_dropout_t drop,
                                void* stream) {
  B200_REQUIRE(x && vals && w && aux_w && u && dx && dvals, B200CTR_E_INVALID_ARG, "pool_bwd: null pointer");
  B200_REQUIRE(D == PL_D, B200CTR_E_UNSUPPORTED_SHAPE, "pool_bwd: D=%d unsupported (64 only)", D);
  B200_REQUIRE(K >= 1 && K <= PL_MAXK, B200CTR_E_UNSUPPORTED_SHAPE, "pool_bwd: K=%d unsupported", K);
  if (B == 0) return B200CTR_OK;
  pool_bwd_kernel<<<(unsigned)B, PL_THREADS, 0, as_stream(stream)>>>(x, vals, w, aux_w, u, ld_u, du, ld_du, daux, gating_relu, K,
                                                                      dx, dvals, daux_w, daux_b, drop);
  B200_LAUNCH_CHECK("pool_bwd_kernel");
  return B200CTR_OK;
}
