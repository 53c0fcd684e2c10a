// K2b pieces: tiny-head attention (head_dim 8) with Toeplitz relative-position bias, residual + RMSNorm,
// and their backward.  Reference: src/models/dare.py:29-37 (PositionalBias), 53-70 (DAREEncoderLayer),
// torch.nn.MultiheadAttention math (q scaled by 1/sqrt(hd), additive float mask, dropout on probabilities).
//
// Attention runs one CTA per sample with K and V of all heads resident in shared memory; each thread owns one
// (query row, head) pair, so the K x K probabilities never leave registers.  Only the per-row log-sum-exp is
// saved; the backward pass recomputes probabilities (the reference stores B*H*K*K floats per layer instead).
#include "common.cuh"
#include <algorithm>

namespace b200ctr {

constexpr int AT_D = 64;
constexpr int AT_HD = 8;
constexpr int AT_THREADS = 256;

__device__ __forceinline__ float dot8(const float* a, const float4& b0, const float4& b1) {
  float s = a[0] * b0.x;
  s = fmaf(a[1], b0.y, s); s = fmaf(a[2], b0.z, s); s = fmaf(a[3], b0.w, s);
  s = fmaf(a[4], b1.x, s); s = fmaf(a[5], b1.y, s); s = fmaf(a[6], b1.z, s); s = fmaf(a[7], b1.w, s);
  return s;
}

// Each thread owns TWO query rows (i, i + ceil(K/2)) of one head: every K/V row read from shared memory feeds
// 32 FMAs instead of 16, and K=60 x 8 heads fits one pass of the 256-thread CTA.
__global__ void __launch_bounds__(AT_THREADS)
attn_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ relmean, int max_len, int K, int H,
                float* __restrict__ o, float* __restrict__ lse, b200ctr_dropout_t drop) {
  extern __shared__ __align__(16) float sm[];
  float* ks = sm;                         // K x 64
  float* vs = ks + (size_t)K * AT_D;      // K x 64
  float* bias = vs + (size_t)K * AT_D;    // 2K-1 : bias[(j-i)+K-1]
  const int64_t b = blockIdx.x;
  const float* base = qkv + b * (int64_t)K * 3 * AT_D;
  for (int i = threadIdx.x; i < K * (AT_D / 4); i += AT_THREADS) {
    int r = i / (AT_D / 4), c = i % (AT_D / 4);
    reinterpret_cast<float4*>(ks)[i] = *reinterpret_cast<const float4*>(base + (int64_t)r * 3 * AT_D + AT_D + c * 4);
    reinterpret_cast<float4*>(vs)[i] = *reinterpret_cast<const float4*>(base + (int64_t)r * 3 * AT_D + 2 * AT_D + c * 4);
  }
  for (int t = threadIdx.x; t < 2 * K - 1; t += AT_THREADS) {
    int d = t - (K - 1);                                   // j - i
    int di = min(max(d, -max_len), max_len) + max_len;
    bias[t] = relmean ? relmean[di] : 0.f;
  }
  __syncthreads();
  const float scale = rsqrtf((float)AT_HD);
  const DropCtx dctx = drop_ctx(drop);
  const int half = (K + 1) >> 1;
  for (int item = threadIdx.x; item < half * H; item += AT_THREADS) {
    const int h = item / half, ia = item % half;
    const int ib_raw = ia + half;
    const bool has_b = ib_raw < K;
    const int ib = has_b ? ib_raw : ia;
    float qa[AT_HD], qb[AT_HD];
    {
      const float4* qp = reinterpret_cast<const float4*>(base + (int64_t)ia * 3 * AT_D + h * AT_HD);
      float4 a = qp[0], c = qp[1];
      qa[0] = a.x * scale; qa[1] = a.y * scale; qa[2] = a.z * scale; qa[3] = a.w * scale;
      qa[4] = c.x * scale; qa[5] = c.y * scale; qa[6] = c.z * scale; qa[7] = c.w * scale;
      qp = reinterpret_cast<const float4*>(base + (int64_t)ib * 3 * AT_D + h * AT_HD);
      a = qp[0]; c = qp[1];
      qb[0] = a.x * scale; qb[1] = a.y * scale; qb[2] = a.z * scale; qb[3] = a.w * scale;
      qb[4] = c.x * scale; qb[5] = c.y * scale; qb[6] = c.z * scale; qb[7] = c.w * scale;
    }
    const float* ba = bias + (K - 1) - ia;                 // ba[j] = bias[j-ia+K-1]
    const float* bb = bias + (K - 1) - ib;
    float ma = -INFINITY, mb = -INFINITY;
    for (int j = 0; j < K; ++j) {
      const float4* kp = reinterpret_cast<const float4*>(ks + (size_t)j * AT_D + h * AT_HD);
      const float4 k0 = kp[0], k1 = kp[1];
      ma = fmaxf(ma, dot8(qa, k0, k1) + ba[j]);
      mb = fmaxf(mb, dot8(qb, k0, k1) + bb[j]);
    }
    float la = 0.f, lb = 0.f;
    float aa[AT_HD], ab[AT_HD];
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) { aa[d] = 0.f; ab[d] = 0.f; }
    const uint32_t keya = drop_row_key(drop, ((uint64_t)b * H + h) * K + ia), keyb = drop_row_key(drop, ((uint64_t)b * H + h) * K + ib);
    for (int j = 0; j < K; ++j) {
      const float4* kp = reinterpret_cast<const float4*>(ks + (size_t)j * AT_D + h * AT_HD);
      const float4* vp = reinterpret_cast<const float4*>(vs + (size_t)j * AT_D + h * AT_HD);
      const float4 k0 = kp[0], k1 = kp[1], v0 = vp[0], v1 = vp[1];
      const float pa = expf(dot8(qa, k0, k1) + ba[j] - ma);
      const float pb = expf(dot8(qb, k0, k1) + bb[j] - mb);
      la += pa; lb += pb;
      const float da = pa * drop_rc(dctx, keya, (uint32_t)j), db = pb * drop_rc(dctx, keyb, (uint32_t)j);
      aa[0] = fmaf(da, v0.x, aa[0]); aa[1] = fmaf(da, v0.y, aa[1]); aa[2] = fmaf(da, v0.z, aa[2]); aa[3] = fmaf(da, v0.w, aa[3]);
      aa[4] = fmaf(da, v1.x, aa[4]); aa[5] = fmaf(da, v1.y, aa[5]); aa[6] = fmaf(da, v1.z, aa[6]); aa[7] = fmaf(da, v1.w, aa[7]);
      ab[0] = fmaf(db, v0.x, ab[0]); ab[1] = fmaf(db, v0.y, ab[1]); ab[2] = fmaf(db, v0.z, ab[2]); ab[3] = fmaf(db, v0.w, ab[3]);
      ab[4] = fmaf(db, v1.x, ab[4]); ab[5] = fmaf(db, v1.y, ab[5]); ab[6] = fmaf(db, v1.z, ab[6]); ab[7] = fmaf(db, v1.w, ab[7]);
    }
    {
      const float inv = 1.f / la;
      float4* op = reinterpret_cast<float4*>(o + (b * K + ia) * (int64_t)AT_D + h * AT_HD);
      op[0] = make_float4(aa[0] * inv, aa[1] * inv, aa[2] * inv, aa[3] * inv);
      op[1] = make_float4(aa[4] * inv, aa[5] * inv, aa[6] * inv, aa[7] * inv);
      lse[(b * H + h) * (int64_t)K + ia] = ma + logf(la);
    }
    if (has_b) {
      const float inv = 1.f / lb;
      float4* op = reinterpret_cast<float4*>(o + (b * K + ib) * (int64_t)AT_D + h * AT_HD);
      op[0] = make_float4(ab[0] * inv, ab[1] * inv, ab[2] * inv, ab[3] * inv);
      op[1] = make_float4(ab[4] * inv, ab[5] * inv, ab[6] * inv, ab[7] * inv);
      lse[(b * H + h) * (int64_t)K + ib] = mb + logf(lb);
    }
  }
}

__device__ __forceinline__ void axpy8(float* acc, float a, const float4& x0, const float4& x1) {
  acc[0] = fmaf(a, x0.x, acc[0]); acc[1] = fmaf(a, x0.y, acc[1]); acc[2] = fmaf(a, x0.z, acc[2]); acc[3] = fmaf(a, x0.w, acc[3]);
  acc[4] = fmaf(a, x1.x, acc[4]); acc[5] = fmaf(a, x1.y, acc[5]); acc[6] = fmaf(a, x1.z, acc[6]); acc[7] = fmaf(a, x1.w, acc[7]);
}

__global__ void __launch_bounds__(AT_THREADS)
attn_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ relmean, int max_len, const float* __restrict__ o,
                const float* __restrict__ lse, const float* __restrict__ d_o, int K, int H, float* __restrict__ dqkv,
                float* __restrict__ drelmean, b200ctr_dropout_t drop) {
  extern __shared__ __align__(16) float sm[];
  float* qs = sm;                            // K x 64 (pre-scaled by 1/sqrt(hd))
  float* ks = qs + (size_t)K * AT_D;
  float* vs = ks + (size_t)K * AT_D;
  float* dos = vs + (size_t)K * AT_D;
  float* lses = dos + (size_t)K * AT_D;      // K*H  [h*K+i]
  float* dsum = lses + (size_t)K * H;        // K*H
  float* bias = dsum + (size_t)K * H;        // 2K-1
  float* dbias = bias + (2 * K - 1);         // 2K-1
  const int64_t b = blockIdx.x;
  const float* base = qkv + b * (int64_t)K * 3 * AT_D;
  const float scale = rsqrtf((float)AT_HD);
  for (int i = threadIdx.x; i < K * (AT_D / 4); i += AT_THREADS) {
    int r = i / (AT_D / 4), c = i % (AT_D / 4);
    float4 qv = *reinterpret_cast<const float4*>(base + (int64_t)r * 3 * AT_D + c * 4);
    qv.x *= scale; qv.y *= scale; qv.z *= scale; qv.w *= scale;
    reinterpret_cast<float4*>(qs)[i] = qv;
    reinterpret_cast<float4*>(ks)[i] = *reinterpret_cast<const float4*>(base + (int64_t)r * 3 * AT_D + AT_D + c * 4);
    reinterpret_cast<float4*>(vs)[i] = *reinterpret_cast<const float4*>(base + (int64_t)r * 3 * AT_D + 2 * AT_D + c * 4);
    reinterpret_cast<float4*>(dos)[i] = *reinterpret_cast<const float4*>(d_o + (b * K + r) * (int64_t)AT_D + c * 4);
  }
  for (int t = threadIdx.x; t < 2 * K - 1; t += AT_THREADS) {
    int d = t - (K - 1);
    int di = min(max(d, -max_len), max_len) + max_len;
    bias[t] = relmean ? relmean[di] : 0.f;
    dbias[t] = 0.f;
  }
  __syncthreads();
  for (int item = threadIdx.x; item < K * H; item += AT_THREADS) {
    const int h = item / K, i = item % K;
    lses[item] = lse[(b * H + h) * (int64_t)K + i];
    const float4* op = reinterpret_cast<const float4*>(o + (b * K + i) * (int64_t)AT_D + h * AT_HD);
    const float4* dp = reinterpret_cast<const float4*>(dos + (size_t)i * AT_D + h * AT_HD);
    float4 o0 = op[0], o1 = op[1], d0 = dp[0], d1 = dp[1];
    dsum[item] = o0.x * d0.x + o0.y * d0.y + o0.z * d0.z + o0.w * d0.w + o1.x * d1.x + o1.y * d1.y + o1.z * d1.z + o1.w * d1.w;
  }
  __syncthreads();
  const int half = (K + 1) >> 1;
  const DropCtx dctx = drop_ctx(drop);
  // phase 1: thread per (head, query rows ia/ib) -> dq, dbias
  for (int item = threadIdx.x; item < half * H; item += AT_THREADS) {
    const int h = item / half, ia = item % half;
    const int ib_raw = ia + half;
    const bool has_b = ib_raw < K;
    const int ib = has_b ? ib_raw : ia;
    float qa[AT_HD], qb[AT_HD], da[AT_HD], db[AT_HD], ga[AT_HD], gb[AT_HD];
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) {
      qa[d] = qs[(size_t)ia * AT_D + h * AT_HD + d]; qb[d] = qs[(size_t)ib * AT_D + h * AT_HD + d];
      da[d] = dos[(size_t)ia * AT_D + h * AT_HD + d]; db[d] = dos[(size_t)ib * AT_D + h * AT_HD + d];
      ga[d] = 0.f; gb[d] = 0.f;
    }
    const float La = lses[h * K + ia], Lb = lses[h * K + ib], Da = dsum[h * K + ia], Db = dsum[h * K + ib];
    const float* ba = bias + (K - 1) - ia;
    const float* bb = bias + (K - 1) - ib;
    float* dba = dbias + (K - 1) - ia;
    float* dbb = dbias + (K - 1) - ib;
    const uint32_t keya = drop_row_key(drop, ((uint64_t)b * H + h) * K + ia), keyb = drop_row_key(drop, ((uint64_t)b * H + h) * K + ib);
    for (int j = 0; j < K; ++j) {
      const float4* kp = reinterpret_cast<const float4*>(ks + (size_t)j * AT_D + h * AT_HD);
      const float4* vp = reinterpret_cast<const float4*>(vs + (size_t)j * AT_D + h * AT_HD);
      const float4 k0 = kp[0], k1 = kp[1], v0 = vp[0], v1 = vp[1];
      const float pa = expf(dot8(qa, k0, k1) + ba[j] - La);
      const float pb = expf(dot8(qb, k0, k1) + bb[j] - Lb);
      const float dsa = pa * (dot8(da, v0, v1) * drop_rc(dctx, keya, (uint32_t)j) - Da);
      const float dsb = pb * (dot8(db, v0, v1) * drop_rc(dctx, keyb, (uint32_t)j) - Db);
      axpy8(ga, dsa, k0, k1);
      axpy8(gb, dsb, k0, k1);
      if (drelmean) {
        atomicAdd(dba + j, dsa);
        if (has_b) atomicAdd(dbb + j, dsb);
      }
    }
    float4* out = reinterpret_cast<float4*>(dqkv + (b * K + ia) * (int64_t)(3 * AT_D) + h * AT_HD);
    out[0] = make_float4(ga[0] * scale, ga[1] * scale, ga[2] * scale, ga[3] * scale);
    out[1] = make_float4(ga[4] * scale, ga[5] * scale, ga[6] * scale, ga[7] * scale);
    if (has_b) {
      out = reinterpret_cast<float4*>(dqkv + (b * K + ib) * (int64_t)(3 * AT_D) + h * AT_HD);
      out[0] = make_float4(gb[0] * scale, gb[1] * scale, gb[2] * scale, gb[3] * scale);
      out[1] = make_float4(gb[4] * scale, gb[5] * scale, gb[6] * scale, gb[7] * scale);
    }
  }
  // phase 2: thread per (head, key rows ja/jb) -> dk, dv
  for (int item = threadIdx.x; item < half * H; item += AT_THREADS) {
    const int h = item / half, ja = item % half;
    const int jb_raw = ja + half;
    const bool has_b = jb_raw < K;
    const int jb = has_b ? jb_raw : ja;
    float ka[AT_HD], kb[AT_HD], va[AT_HD], vb[AT_HD], dka[AT_HD], dkb[AT_HD], dva[AT_HD], dvb[AT_HD];
#pragma unroll
    for (int d = 0; d < AT_HD; ++d) {
      ka[d] = ks[(size_t)ja * AT_D + h * AT_HD + d]; kb[d] = ks[(size_t)jb * AT_D + h * AT_HD + d];
      va[d] = vs[(size_t)ja * AT_D + h * AT_HD + d]; vb[d] = vs[(size_t)jb * AT_D + h * AT_HD + d];
      dka[d] = 0.f; dkb[d] = 0.f; dva[d] = 0.f; dvb[d] = 0.f;
    }
    for (int i = 0; i < K; ++i) {
      const float4* qp = reinterpret_cast<const float4*>(qs + (size_t)i * AT_D + h * AT_HD);
      const float4* dp4 = reinterpret_cast<const float4*>(dos + (size_t)i * AT_D + h * AT_HD);
      const float4 q0 = qp[0], q1 = qp[1], d0 = dp4[0], d1 = dp4[1];
      const float L = lses[h * K + i], Di = dsum[h * K + i];
      const uint32_t keyi = dctx.on ? drop_row_key(drop, ((uint64_t)b * H + h) * K + i) : 0u;
      const float pa = expf(dot8(ka, q0, q1) + bias[(ja - i) + K - 1] - L);
      const float pb = expf(dot8(kb, q0, q1) + bias[(jb - i) + K - 1] - L);
      const float sa = drop_rc(dctx, keyi, (uint32_t)ja), sb = drop_rc(dctx, keyi, (uint32_t)jb);
      axpy8(dva, pa * sa, d0, d1);
      axpy8(dvb, pb * sb, d0, d1);
      const float dsa = pa * (dot8(va, d0, d1) * sa - Di);
      const float dsb = pb * (dot8(vb, d0, d1) * sb - Di);
      axpy8(dka, dsa, q0, q1);                       // qs carries the 1/sqrt(hd) scale
      axpy8(dkb, dsb, q0, q1);
    }
    float4* outk = reinterpret_cast<float4*>(dqkv + (b * K + ja) * (int64_t)(3 * AT_D) + AT_D + h * AT_HD);
    float4* outv = reinterpret_cast<float4*>(dqkv + (b * K + ja) * (int64_t)(3 * AT_D) + 2 * AT_D + h * AT_HD);
    outk[0] = make_float4(dka[0], dka[1], dka[2], dka[3]); outk[1] = make_float4(dka[4], dka[5], dka[6], dka[7]);
    outv[0] = make_float4(dva[0], dva[1], dva[2], dva[3]); outv[1] = make_float4(dva[4], dva[5], dva[6], dva[7]);
    if (has_b) {
      outk = reinterpret_cast<float4*>(dqkv + (b * K + jb) * (int64_t)(3 * AT_D) + AT_D + h * AT_HD);
      outv = reinterpret_cast<float4*>(dqkv + (b * K + jb) * (int64_t)(3 * AT_D) + 2 * AT_D + h * AT_HD);
      outk[0] = make_float4(dkb[0], dkb[1], dkb[2], dkb[3]); outk[1] = make_float4(dkb[4], dkb[5], dkb[6], dkb[7]);
      outv[0] = make_float4(dvb[0], dvb[1], dvb[2], dvb[3]); outv[1] = make_float4(dvb[4], dvb[5], dvb[6], dvb[7]);
    }
  }
  if (drelmean) {
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * K - 1; t += AT_THREADS) {
      int d = t - (K - 1);
      int di = min(max(d, -max_len), max_len) + max_len;
      // per-sample partial row (zero-initialised by the caller, |d| <= K-1 <= max_len-1 so no two t share a slot); the rows
      // are summed in sample order afterwards (b200ctr_colsum)
      drelmean[b * (int64_t)(2 * max_len + 1) + di] = dbias[t];
    }
  }
}

// warp per row, D == 64
__global__ void __launch_bounds__(256)
add_rmsnorm_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ w, float eps,
                       int64_t M, float* __restrict__ s_out, float* __restrict__ y_out, float* __restrict__ r_out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float2 wv = reinterpret_cast<const float2*>(w)[lane];
  for (int64_t m = warp; m < M; m += nwarps) {
    float2 s = reinterpret_cast<const float2*>(a + m * AT_D)[lane];
    if (b) { float2 t = reinterpret_cast<const float2*>(b + m * AT_D)[lane]; s.x += t.x; s.y += t.y; }
    float ss = warp_sum(s.x * s.x + s.y * s.y);
    float r = rsqrtf(ss * (1.f / AT_D) + eps);
    if (s_out) reinterpret_cast<float2*>(s_out + m * AT_D)[lane] = s;
    reinterpret_cast<float2*>(y_out + m * AT_D)[lane] = make_float2(wv.x * s.x * r, wv.y * s.y * r);
    if (lane == 0 && r_out) r_out[m] = r;
  }
}

__global__ void __launch_bounds__(256)
rmsnorm_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ s, const float* __restrict__ w,
                   const float* __restrict__ r, int64_t M, const float* __restrict__ dx_add, float* __restrict__ dx,
                   float* __restrict__ dw) {
  __shared__ float2 red[8][32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float2 wv = reinterpret_cast<const float2*>(w)[lane];
  float2 dwa = make_float2(0.f, 0.f);
  for (int64_t m = warp; m < M; m += nwarps) {
    float2 d = reinterpret_cast<const float2*>(dy + m * AT_D)[lane];
    float2 x = reinterpret_cast<const float2*>(s + m * AT_D)[lane];
    float rr = r[m];
    float2 g = make_float2(d.x * wv.x, d.y * wv.y);
    float dot = warp_sum(g.x * x.x + g.y * x.y) * (1.f / AT_D);
    float c = rr * rr * rr * dot;
    float2 out = make_float2(rr * g.x - x.x * c, rr * g.y - x.y * c);
    if (dx_add) { float2 t = reinterpret_cast<const float2*>(dx_add + m * AT_D)[lane]; out.x += t.x; out.y += t.y; }
    reinterpret_cast<float2*>(dx + m * AT_D)[lane] = out;
    dwa.x = fmaf(d.x * x.x, rr, dwa.x);
    dwa.y = fmaf(d.y * x.y, rr, dwa.y);
  }
  red[wid][lane] = dwa;
  __syncthreads();
  if (wid == 0) {
    float2 t = red[0][lane];
    for (int k = 1; k < 8; ++k) { t.x += red[k][lane].x; t.y += red[k][lane].y; }
    // one partial row per CTA, summed afterwards in row order (b200ctr_colsum): no float atomics, bit-identical runs
    reinterpret_cast<float2*>(dw + (int64_t)blockIdx.x * AT_D)[lane] = t;
  }
}

__global__ void relmean_fwd_kernel(const float* __restrict__ rel, int T, int H, float* __restrict__ relmean) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  float s = 0.f;
  for (int h = 0; h < H; ++h) s += rel[t * H + h];
  relmean[t] = s / (float)H;
}
__global__ void relmean_bwd_kernel(const float* __restrict__ drelmean, int T, int H, float* __restrict__ drel) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * H) return;
  drel[i] = drelmean[i / H] / (float)H;
}

static int set_attn_smem(const void* fn, size_t bytes) {
  if (bytes > 227 * 1024) {
    set_error("attention: K too large for one CTA's shared memory (%zu bytes)", bytes);
    return B200CTR_E_UNSUPPORTED_SHAPE;
  }
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attention smem)");
  }
  return B200CTR_OK;
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_attn_fwd(const float* qkv, const float* relmean, int max_len, int64_t B, int K, int H, int D, float* o,
                                float* lse, b200ctr_dropout_t drop, void* stream) {
  B200_REQUIRE(qkv && o && lse, B200CTR_E_INVALID_ARG, "attn_fwd: null pointer");
  B200_REQUIRE(D == AT_D && H * AT_HD == D, B200CTR_E_UNSUPPORTED_SHAPE,
               "attn_fwd: only emb_dim=64 with head_dim=8 is built (got D=%d, H=%d)", D, H);
  if (B == 0 || K == 0) return B200CTR_OK;
  size_t smem = ((size_t)2 * K * AT_D + 2 * K) * sizeof(float);
  int rc = set_attn_smem((const void*)attn_fwd_kernel, smem);
  if (rc) return rc;
  attn_fwd_kernel<<<(unsigned)B, AT_THREADS, smem, as_stream(stream)>>>(qkv, relmean, max_len, K, H, o, lse, drop);
  B200_LAUNCH_CHECK("attn_fwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_attn_bwd(const float* qkv, const float* relmean, int max_len, const float* o, const float* lse,
                                const float* d_o, int64_t B, int K, int H, int D, float* dqkv, float* drelmean,
                                b200ctr_dropout_t drop, void* stream) {
  B200_REQUIRE(qkv && o && lse && d_o && dqkv, B200CTR_E_INVALID_ARG, "attn_bwd: null pointer");
  B200_REQUIRE(D == AT_D && H * AT_HD == D, B200CTR_E_UNSUPPORTED_SHAPE,
               "attn_bwd: only emb_dim=64 with head_dim=8 is built (got D=%d, H=%d)", D, H);
  if (B == 0 || K == 0) return B200CTR_OK;
  size_t smem = ((size_t)4 * K * AT_D + 2 * (size_t)K * H + 4 * K) * sizeof(float);
  int rc = set_attn_smem((const void*)attn_bwd_kernel, smem);
  if (rc) return rc;
  attn_bwd_kernel<<<(unsigned)B, AT_THREADS, smem, as_stream(stream)>>>(qkv, relmean, max_len, o, lse, d_o, K, H, dqkv, drelmean, drop);
  B200_LAUNCH_CHECK("attn_bwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_add_rmsnorm_fwd(const float* a, const float* b, const float* w, float eps, int64_t M, int D,
                                       float* s_out, float* y_out, float* r_out, void* stream) {
  B200_REQUIRE(a && w && y_out, B200CTR_E_INVALID_ARG, "add_rmsnorm_fwd: null pointer");
  B200_REQUIRE(D == AT_D, B200CTR_E_UNSUPPORTED_SHAPE, "add_rmsnorm_fwd: D=%d unsupported (64 only)", D);
  if (M == 0) return B200CTR_OK;
  int grid = (int)std::min<int64_t>(ceil_div(M, 8), (int64_t)num_sms() * 16);
  add_rmsnorm_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, b, w, eps, M, s_out, y_out, r_out);
  B200_LAUNCH_CHECK("add_rmsnorm_fwd_kernel");
  return B200CTR_OK;
}

extern "C" int64_t b200ctr_rmsnorm_bwd_rows(int64_t M) {
  return std::max<int64_t>(1, std::min<int64_t>(ceil_div(M, (int64_t)8), (int64_t)num_sms() * 8));
}

extern "C" int b200ctr_rmsnorm_bwd(const float* dy, const float* s, const float* w, const float* r, int64_t M, int D,
                                   const float* dx_add, float* dx, float* dw, void* stream) {
  B200_REQUIRE(dy && s && w && r && dx && dw, B200CTR_E_INVALID_ARG, "rmsnorm_bwd: null pointer");
  B200_REQUIRE(D == AT_D, B200CTR_E_UNSUPPORTED_SHAPE, "rmsnorm_bwd: D=%d unsupported (64 only)", D);
  if (M == 0) return B200CTR_OK;
  int grid = (int)b200ctr_rmsnorm_bwd_rows(M);
  rmsnorm_bwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(dy, s, w, r, M, dx_add, dx, dw);
  B200_LAUNCH_CHECK("rmsnorm_bwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_relmean_fwd(const float* rel, int T, int H, float* relmean, void* stream) {
  B200_REQUIRE(rel && relmean, B200CTR_E_INVALID_ARG, "relmean_fwd: null pointer");
  relmean_fwd_kernel<<<(T + 127) / 128, 128, 0, as_stream(stream)>>>(rel, T, H, relmean);
  B200_LAUNCH_CHECK("relmean_fwd_kernel");
  return B200CTR_OK;
}
extern "C" int b200ctr_relmean_bwd(const float* drelmean, int T, int H, float* drel, void* stream) {
  B200_REQUIRE(drelmean && drel, B200CTR_E_INVALID_ARG, "relmean_bwd: null pointer");
  relmean_bwd_kernel<<<(T * H + 127) / 128, 128, 0, as_stream(stream)>>>(drelmean, T, H, drel);
  B200_LAUNCH_CHECK("relmean_bwd_kernel");
  return B200CTR_OK;
}
