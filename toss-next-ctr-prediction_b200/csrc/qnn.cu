// K3 pieces that are not plain contractions: row RMSNorm over F*D, the FM-trick reductions of the low-rank
// quadratic interaction, the batch-statistic SE gate and its application.
// Reference: src/models/qnn_alpha.py:6-12 (RMSNorm), 24-26 (SEBlock), 86-97 (_pair_interaction_all), 109-130.
#include "common.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace b200ctr {

constexpr int RR_THREADS = 256;
constexpr int RR_MAXV = 16;          // float4 slots per thread => N <= 16*4*256 = 16384

__global__ void __launch_bounds__(RR_THREADS)
rowrms_fwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ w, float eps, int64_t N,
                  float* __restrict__ z, int64_t ld_z, float* __restrict__ r, __nv_bfloat16* __restrict__ z16) {
  __shared__ float red[32];
  const int64_t b = blockIdx.x;
  const float4* xr = reinterpret_cast<const float4*>(x + b * ldx);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float4* zr = reinterpret_cast<float4*>(z ? z + b * ld_z : nullptr);
  const int n4 = (int)(N >> 2);
  float ss = 0.f;
  for (int i = threadIdx.x; i < n4; i += RR_THREADS) {
    float4 v = xr[i];
    ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  ss = block_sum(ss, red);
  const float rr = rsqrtf(ss / (float)N + eps);
  for (int i = threadIdx.x; i < n4; i += RR_THREADS) {
    float4 v = xr[i], ww = __ldg(w4 + i);
    const float4 o = make_float4(ww.x * v.x * rr, ww.y * v.y * rr, ww.z * v.z * rr, ww.w * v.w * rr);
    if (z) zr[i] = o;
    if (z16) {          // bf16 copy for the tensor-core contractions that consume z (row stride N)
      __nv_bfloat162 lo = __floats2bfloat162_rn(o.x, o.y), hi = __floats2bfloat162_rn(o.z, o.w);
      reinterpret_cast<uint2*>(z16 + b * N)[i] = make_uint2(*reinterpret_cast<uint32_t*>(&lo), *reinterpret_cast<uint32_t*>(&hi));
    }
  }
  if (threadIdx.x == 0) r[b] = rr;
}

// grid = n_partials CTAs; CTA p handles rows p, p+P, ...  The row's dz stays in registers between the dot-product pass and the
// output pass; x is read again in the output pass (an L2 hit: the CTA read the same 51 KB a few microseconds earlier) and the
// dw accumulators live in shared memory (each thread owns its elements: no conflicts, no atomics, same summation order as
// register accumulators).  That is 2 CTAs per SM instead of 1 (247 registers before): one CTA's load phase overlaps the other's
// reduction barrier and stores (ncu r02ai: 12.5 % of the warp slots active, 4.5 TB/s of DRAM traffic).
__global__ void __launch_bounds__(RR_THREADS, 2)
rowrms_bwd_kernel(const float* __restrict__ dz, int64_t ld_dz, const float* __restrict__ x, int64_t ldx,
                  const float* __restrict__ w, const float* __restrict__ r, int64_t B, int64_t N, float* __restrict__ dx,
                  int64_t ld_dx, float* __restrict__ dw_partial) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  float4* accs = reinterpret_cast<float4*>(rr_smem);       // [N / 4]
  __shared__ float red[32];
  const int n4 = (int)(N >> 2);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  for (int i = threadIdx.x; i < n4; i += RR_THREADS) accs[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    const float4* dzr = reinterpret_cast<const float4*>(dz + b * ld_dz);
    const float4* xr = reinterpret_cast<const float4*>(x + b * ldx);
    float4* dxr = reinterpret_cast<float4*>(dx + b * ld_dx);
    const float rr = r[b];
    float4 dreg[RR_MAXV];
    float dot = 0.f;
#pragma unroll
    for (int s = 0; s < RR_MAXV; ++s) {
      const int i = threadIdx.x + s * RR_THREADS;
      dreg[s] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n4) dreg[s] = dzr[i];
    }
#pragma unroll
    for (int s = 0; s < RR_MAXV; ++s) {
      const int i = threadIdx.x + s * RR_THREADS;
      if (i < n4) {
        const float4 ww = __ldg(w4 + i), xv = xr[i];
        dot += dreg[s].x * ww.x * xv.x + dreg[s].y * ww.y * xv.y + dreg[s].z * ww.z * xv.z + dreg[s].w * ww.w * xv.w;
      }
    }
    dot = block_sum(dot, red) / (float)N;
    const float c = rr * rr * rr * dot;
#pragma unroll
    for (int s = 0; s < RR_MAXV; ++s) {
      const int i = threadIdx.x + s * RR_THREADS;
      if (i < n4) {
        const float4 d = dreg[s], v = xr[i], ww = __ldg(w4 + i);   // w (51 KB, shared by every row) is an L1/L2 hit
        dxr[i] = make_float4(rr * d.x * ww.x - v.x * c, rr * d.y * ww.y - v.y * c, rr * d.z * ww.z - v.z * c, rr * d.w * ww.w - v.w * c);
        float4 a = accs[i];
        a.x = fmaf(d.x * v.x, rr, a.x); a.y = fmaf(d.y * v.y, rr, a.y);
        a.z = fmaf(d.z * v.z, rr, a.z); a.w = fmaf(d.w * v.w, rr, a.w);
        accs[i] = a;
      }
    }
  }
  float4* out = reinterpret_cast<float4*>(dw_partial + (int64_t)blockIdx.x * N);
  for (int i = threadIdx.x; i < n4; i += RR_THREADS) out[i] = accs[i];
}

// A (B,F,C): s[b,c] = sum_f A, quad = s^2 - sum_f A^2.  One CTA per sample.
__global__ void __launch_bounds__(256)
fm_reduce_kernel(const float* __restrict__ A, int F, int C, float* __restrict__ s, float* __restrict__ quad) {
  const int64_t b = blockIdx.x;
  const float* Ab = A + b * (int64_t)F * C;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float a = 0.f, q = 0.f;
    for (int f = 0; f < F; ++f) { float v = Ab[(int64_t)f * C + c]; a += v; q = fmaf(v, v, q); }
    s[b * C + c] = a;
    quad[b * C + c] = a * a - q;
  }
}

__global__ void fm_grad_kernel(float* __restrict__ A, const float* __restrict__ s, const float* __restrict__ dquad,
                               int64_t total, int F, int C) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  int64_t b = i / ((int64_t)F * C);
  int c = (int)(i % C);
  A[i] = dquad[b * C + c] * 2.f * (s[b * C + c] - A[i]);
}

// m = colsum/B ; h1 = relu(W1 m + b1) ; gate = sigmoid(W2 h1 + b2).  W1 (Cr,C), W2 (C,Cr).  Two grid-wide launches (the
// single-CTA version was 0.09 ms of pure latency): A = one warp per hidden unit j, B = one warp per output channel c.
__global__ void __launch_bounds__(256)
se_gate_fwd_a_kernel(const float* __restrict__ colsum, float inv_B, const float* __restrict__ w1, const float* __restrict__ b1,
                     int C, int Cr, float* __restrict__ mean, float* __restrict__ h1) {
  const int lane = threadIdx.x & 31;
  const int j = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (j >= Cr) return;
  float a = 0.f;
  for (int c = lane; c < C; c += 32) {
    const float m = colsum[c] * inv_B;
    if (j == 0) mean[c] = m;
    a = fmaf(w1[(int64_t)j * C + c], m, a);
  }
  a = warp_sum(a);
  if (lane == 0) h1[j] = fmaxf(a + b1[j], 0.f);
}
__global__ void __launch_bounds__(256)
se_gate_fwd_b_kernel(const float* __restrict__ h1, const float* __restrict__ w2, const float* __restrict__ b2, int C, int Cr,
                     float* __restrict__ gate) {
  const int lane = threadIdx.x & 31;
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (c >= C) return;
  float a = 0.f;
  for (int j = lane; j < Cr; j += 32) a = fmaf(w2[(int64_t)c * Cr + j], h1[j], a);
  a = warp_sum(a);
  if (lane == 0) gate[c] = sigmoidf_(a + b2[c]);
}

// SE gate backward as two grid-wide launches (the single-CTA version took 0.22 ms of pure latency per step):
//   A: CTA j  -> dh1[j] = relu'(h1[j]) * sum_c da2[c] w2[c][j]  (da2[c] = dgate[c] g (1-g)), db1[j], dw1[j][:] = dh1[j] * mean[:]
//   B: thread c -> da2[c], db2[c], dw2[c][:] = da2[c] * h1[:], dmean[c] = sum_j dh1[j] w1[j][c]      (reads dh1 == db1 from A)
__global__ void __launch_bounds__(256)
se_gate_bwd_a_kernel(const float* __restrict__ dgate, const float* __restrict__ mean, const float* __restrict__ h1,
                     const float* __restrict__ gate, const float* __restrict__ w2, int C, int Cr,
                     float* __restrict__ dw1, float* __restrict__ db1) {
  __shared__ float red[32];
  const int j = blockIdx.x;
  float a = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float g = gate[c];
    a = fmaf(dgate[c] * g * (1.f - g), w2[(int64_t)c * Cr + j], a);
  }
  a = block_sum(a, red);
  const float d = h1[j] > 0.f ? a : 0.f;
  if (threadIdx.x == 0) db1[j] = d;
  for (int c = threadIdx.x; c < C; c += blockDim.x) dw1[(int64_t)j * C + c] = d * mean[c];
}

__global__ void __launch_bounds__(128)
se_gate_bwd_b_kernel(const float* __restrict__ dgate, const float* __restrict__ h1, const float* __restrict__ gate,
                     const float* __restrict__ w1, const float* __restrict__ dh1, int C, int Cr, float* __restrict__ dw2,
                     float* __restrict__ db2, float* __restrict__ dmean) {
  extern __shared__ float sm[];
  float* hs = sm;            // Cr
  float* ds = hs + Cr;       // Cr
  for (int j = threadIdx.x; j < Cr; j += blockDim.x) { hs[j] = h1[j]; ds[j] = dh1[j]; }
  __syncthreads();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float g = gate[c];
  const float d = dgate[c] * g * (1.f - g);
  db2[c] = d;
  float a = 0.f;
  for (int j = 0; j < Cr; ++j) {
    dw2[(int64_t)c * Cr + j] = d * hs[j];
    a = fmaf(ds[j], w1[(int64_t)j * C + c], a);
  }
  dmean[c] = a;
}

__global__ void se_apply_fwd_kernel(const float* __restrict__ inter, const float* __restrict__ gate, int64_t B, int C,
                                    float* __restrict__ out, int64_t ld_out, b200ctr_dropout_t drop) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  int64_t b = i / C;
  int c = (int)(i % C);
  float g = gate ? gate[c] : 1.f;
  out[b * ld_out + c] = inter[i] * g * drop_scale(drop, (uint64_t)i);
}

__global__ void se_apply_bwd_kernel(const float* __restrict__ dout, int64_t ld_dout, const float* __restrict__ inter,
                                    const float* __restrict__ gate, const float* __restrict__ dmean, float inv_B, int64_t B,
                                    int C, float* __restrict__ dinter, float* __restrict__ gprod, b200ctr_dropout_t drop) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  int64_t b = i / C;
  int c = (int)(i % C);
  float d = dout[b * ld_dout + c] * drop_scale(drop, (uint64_t)i);
  if (gprod) gprod[i] = d * inter[i];
  if (dinter) dinter[i] = d * (gate ? gate[c] : 1.f) + (dmean ? dmean[c] * inv_B : 0.f);
}
// the same, one aligned element PAIR per thread (C even: a pair never straddles a row): one dropout hash word and one 64-bit
// division per pair instead of per element
__global__ void se_apply_bwd_pair_kernel(const float* __restrict__ dout, int64_t ld_dout, const float* __restrict__ inter,
                                         const float* __restrict__ gate, const float* __restrict__ dmean, float inv_B, int64_t B,
                                         int C, float* __restrict__ dinter, float* __restrict__ gprod, b200ctr_dropout_t drop) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2;
  if (i >= B * C) return;
  const int64_t b = i / C;
  const int c = (int)(i - b * C);
  float s0 = 1.f, s1 = 1.f;
  if (drop.p > 0.f) drop_scale2(drop, (uint64_t)i, s0, s1);
  const float d0 = dout[b * ld_dout + c] * s0, d1 = dout[b * ld_dout + c + 1] * s1;
  if (gprod) { gprod[i] = d0 * inter[i]; gprod[i + 1] = d1 * inter[i + 1]; }
  if (dinter) {
    dinter[i] = d0 * (gate ? gate[c] : 1.f) + (dmean ? dmean[c] * inv_B : 0.f);
    dinter[i + 1] = d1 * (gate ? gate[c + 1] : 1.f) + (dmean ? dmean[c + 1] * inv_B : 0.f);
  }
}


// ---- per-head value projection of the quadratic term (qnn_alpha.py:96: out_h = quad_h @ V_h) --------------------
// inter[b, h*P + p] = sum_r quad[b, h*R + r] * V[h][r][p].  Contraction depth R = 16: one launch for all heads on the
// CUDA cores (exact fp32) instead of H skinny GEMMs; V stays in shared memory.
constexpr int HP_THREADS = 384;
constexpr int HP_MAXC = 4;           // output columns per thread => H*P <= 1536

__global__ void __launch_bounds__(HP_THREADS)
headproj_fwd_kernel(const float* __restrict__ quad, const float* __restrict__ V, int64_t B, int H, int R, int P,
                    float* __restrict__ inter) {
  extern __shared__ float hp_sm[];
  float* Vs = hp_sm;                    // H*R*P
  float* qs = Vs + H * R * P;           // 4 rows x H*R
  const int C = H * R, HPd = H * P;
  for (int i = threadIdx.x; i < C * P; i += HP_THREADS) Vs[i] = V[i];
  for (int64_t b0 = (int64_t)blockIdx.x * 4; b0 < B; b0 += (int64_t)gridDim.x * 4) {
    __syncthreads();
    for (int i = threadIdx.x; i < 4 * C; i += HP_THREADS) {
      const int64_t b = b0 + i / C;
      qs[i] = b < B ? quad[b * C + i % C] : 0.f;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < HPd; c += HP_THREADS) {
      const int h = c / P, p = c % P;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      for (int r = 0; r < R; ++r) {
        const float v = Vs[(h * R + r) * P + p];
        a0 = fmaf(qs[h * R + r], v, a0); a1 = fmaf(qs[C + h * R + r], v, a1);
        a2 = fmaf(qs[2 * C + h * R + r], v, a2); a3 = fmaf(qs[3 * C + h * R + r], v, a3);
      }
      if (b0 < B) inter[b0 * HPd + c] = a0;
      if (b0 + 1 < B) inter[(b0 + 1) * HPd + c] = a1;
      if (b0 + 2 < B) inter[(b0 + 2) * HPd + c] = a2;
      if (b0 + 3 < B) inter[(b0 + 3) * HPd + c] = a3;
    }
  }
}

// dquad[b, h*R + r] = sum_p dinter[b, h*P + p] * V[h][r][p] ;  dV partial[cta][h][r][p] = sum_{b in cta} quad * dinter
// HP_RB batch rows per iteration share the two block barriers and the V reads.
constexpr int HP_RB = 4;
template <int RMAX>
__global__ void __launch_bounds__(HP_THREADS)
headproj_bwd_kernel(const float* __restrict__ quad, const float* __restrict__ dinter, const float* __restrict__ V, int64_t B,
                    int H, int R, int P, float* __restrict__ dquad, float* __restrict__ dv_partial) {
  extern __shared__ float hp_sm[];
  const int C = H * R, HPd = H * P, PS = P + 1;
  float* Vs = hp_sm;                    // C x (P+1)
  const int vs_floats = (C * PS + 3) & ~3;
  float* ds_all = Vs + vs_floats;       // [2] x HP_RB x H*P : dinter rows (double-buffered, filled by cp.async)
  float* qs_all = ds_all + 2 * HP_RB * HPd;   // [2] x HP_RB x C : quad rows
  for (int i = threadIdx.x; i < C * P; i += HP_THREADS) Vs[(i / P) * PS + i % P] = V[i];
  float acc[HP_MAXC][RMAX];
#pragma unroll
  for (int k = 0; k < HP_MAXC; ++k)
#pragma unroll
    for (int r = 0; r < RMAX; ++r) acc[k][r] = 0.f;
  int parts = 1;                        // threads cooperating on one dquad element: largest power of two <= min(32, threads/C)
  while (parts * 2 <= 32 && parts * 2 * C <= HP_THREADS) parts *= 2;
  // The rows of iteration i+1 are copied into the other buffer (16-byte cp.async, zero fill past the batch) while
  // iteration i computes: with one CTA per SM the exposed load latency of every 4-row iteration was most of the kernel.
  auto issue = [&](int64_t b0, int buf) {
    float* dsb = ds_all + buf * HP_RB * HPd;
    float* qsb = qs_all + buf * HP_RB * C;
    const int d4 = HPd >> 2, c4n = C >> 2;
    for (int i = threadIdx.x; i < HP_RB * d4; i += HP_THREADS) {
      const int rb = i / d4, c4 = i - rb * d4;
      const int64_t b = b0 + rb;
      const float* src = dinter + (b < B ? b : 0) * HPd + c4 * 4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dsb + rb * HPd + c4 * 4)),
                   "l"(src), "r"(b < B ? 16 : 0) : "memory");
    }
    for (int i = threadIdx.x; i < HP_RB * c4n; i += HP_THREADS) {
      const int rb = i / c4n, c4 = i - rb * c4n;
      const int64_t b = b0 + rb;
      const float* src = quad + (b < B ? b : 0) * C + c4 * 4;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(qsb + rb * C + c4 * 4)),
                   "l"(src), "r"(b < B ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  const int64_t bstep = (int64_t)gridDim.x * HP_RB;
  int buf = 0;
  if ((int64_t)blockIdx.x * HP_RB < B) issue((int64_t)blockIdx.x * HP_RB, 0);
  for (int64_t b0 = (int64_t)blockIdx.x * HP_RB; b0 < B; b0 += bstep, buf ^= 1) {
    if (b0 + bstep < B) {
      issue(b0 + bstep, buf ^ 1);                                  // the buffer the previous iteration finished reading
      asm volatile("cp.async.wait_group 1;" ::: "memory");
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    const float* ds = ds_all + buf * HP_RB * HPd;
    const float* qs = qs_all + buf * HP_RB * C;
#pragma unroll
    for (int k = 0; k < HP_MAXC; ++k) {
      const int c = threadIdx.x + k * HP_THREADS;
      if (c < HPd) {
        const int h = c / P;
#pragma unroll
        for (int rb = 0; rb < HP_RB; ++rb) {
          const float d = ds[rb * HPd + c];
#pragma unroll
          for (int r = 0; r < RMAX; ++r)
            if (r < R) acc[k][r] = fmaf(qs[rb * C + h * R + r], d, acc[k][r]);
        }
      }
    }
    if (threadIdx.x < C * parts) {
      const int cc = threadIdx.x / parts, part = threadIdx.x % parts, h = cc / R;
      const int span = (P + parts - 1) / parts, p0 = part * span, p1 = min(P, p0 + span);
      float a[HP_RB];
#pragma unroll
      for (int rb = 0; rb < HP_RB; ++rb) a[rb] = 0.f;
      for (int pp = p0; pp < p1; ++pp) {
        const float v = Vs[cc * PS + pp];
#pragma unroll
        for (int rb = 0; rb < HP_RB; ++rb) a[rb] = fmaf(ds[rb * HPd + h * P + pp], v, a[rb]);
      }
#pragma unroll
      for (int rb = 0; rb < HP_RB; ++rb) {
        float t = a[rb];
        for (int o = parts >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (part == 0 && b0 + rb < B) dquad[(b0 + rb) * C + cc] = t;
      }
    }
    __syncthreads();          // every thread is done with this buffer before the next iteration's copies overwrite it
  }
  float* out = dv_partial + (int64_t)blockIdx.x * C * P;
#pragma unroll
  for (int k = 0; k < HP_MAXC; ++k) {
    const int c = threadIdx.x + k * HP_THREADS;
    if (c < HPd) {
      const int h = c / P, p = c % P;
#pragma unroll
      for (int r = 0; r < RMAX; ++r)
        if (r < R) out[(h * R + r) * P + p] = acc[k][r];
    }
  }
}

__global__ void headproj_reduce_kernel(const float* __restrict__ partial, int n_part, int n, float* __restrict__ dV) {
  const int i = blockIdx.x * RED_TX + threadIdx.x;
  float a = 0.f;
  if (i < n)
    for (int c = threadIdx.y; c < n_part; c += RED_TY) a += partial[(int64_t)c * n + i];
  a = red_finish_32x8(a);
  if (threadIdx.y == 0 && i < n) dV[i] = a;
}

// ---- the MLP's final Linear(K -> 1) (qnn_alpha.py:85-93): a per-row dot product, not a GEMM ---------------------------
// forward: logit[b] = h[b, :] . w + bias ; one warp per row, float4 lanes
__global__ void __launch_bounds__(256)
lin1_fwd_kernel(const float* __restrict__ h, int64_t ldh, const float* __restrict__ w, const float* __restrict__ bias, int64_t B,
                int K, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t b = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= B) return;
  const float4* hr = reinterpret_cast<const float4*>(h + b * ldh);
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float acc = 0.f;
  for (int i = lane; i < (K >> 2); i += 32) {
    const float4 a = hr[i], ww = __ldg(w4 + i);
    acc += a.x * ww.x + a.y * ww.y + a.z * ww.z + a.w * ww.w;
  }
  acc = warp_sum(acc);
  if (lane == 0) out[b] = acc + (bias ? bias[0] : 0.f);
}
// backward: dh[b, n] = dl[b] * w[n] * (act[b, n] != 0 ? mask_scale : 0)   (ReLU + dropout mask of the layer's input)
//           dw[n]   = sum_b dl[b] * act[b, n] ;  db = sum_b dl[b]           (per-CTA partials [K + 1], reduced afterwards)
// CTA = 256 threads = 8 warps; a warp owns a row per iteration, lane owns K/128 float4 column groups
constexpr int L1_MAXV = 4;      // K <= 512
__global__ void __launch_bounds__(256)
lin1_bwd_kernel(const float* __restrict__ dl, const float* __restrict__ act, int64_t ldact, const float* __restrict__ w, int64_t B,
                int K, float mask_scale, float* __restrict__ dh, int64_t lddh, float* __restrict__ partial) {
  extern __shared__ float red[];              // [8][K + 1]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = K >> 7;                       // float4 groups per lane
  const float4* w4 = reinterpret_cast<const float4*>(w);
  float4 wv[L1_MAXV], acc[L1_MAXV];
#pragma unroll
  for (int v = 0; v < L1_MAXV; ++v) {
    acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    wv[v] = v < nv ? __ldg(w4 + v * 32 + lane) : acc[v];
  }
  float dsum = 0.f;
  for (int64_t b = (int64_t)blockIdx.x * 8 + warp; b < B; b += (int64_t)gridDim.x * 8) {
    const float d = dl[b];
    const float4* ar = reinterpret_cast<const float4*>(act + b * ldact);
    float4* dr = reinterpret_cast<float4*>(dh + b * lddh);
    dsum += d;
#pragma unroll
    for (int v = 0; v < L1_MAXV; ++v) {
      if (v < nv) {
        const float4 a = ar[v * 32 + lane];
        const float dm = d * mask_scale;
        dr[v * 32 + lane] = make_float4(a.x != 0.f ? dm * wv[v].x : 0.f, a.y != 0.f ? dm * wv[v].y : 0.f,
                                        a.z != 0.f ? dm * wv[v].z : 0.f, a.w != 0.f ? dm * wv[v].w : 0.f);
        acc[v].x = fmaf(d, a.x, acc[v].x); acc[v].y = fmaf(d, a.y, acc[v].y);
        acc[v].z = fmaf(d, a.z, acc[v].z); acc[v].w = fmaf(d, a.w, acc[v].w);
      }
    }
  }
#pragma unroll
  for (int v = 0; v < L1_MAXV; ++v)
    if (v < nv) *reinterpret_cast<float4*>(red + warp * (K + 4) + (v * 32 + lane) * 4) = acc[v];
  if (lane == 0) red[8 * (K + 4) + warp] = dsum;
  __syncthreads();
  for (int n = threadIdx.x; n <= K; n += 256) {
    float a = 0.f;
    for (int wv_ = 0; wv_ < 8; ++wv_) a += n < K ? red[wv_ * (K + 4) + n] : red[8 * (K + 4) + wv_];
    partial[(int64_t)blockIdx.x * (K + 1) + n] = a;
  }
}
__global__ void lin1_reduce_kernel(const float* __restrict__ partial, int n_part, int K, float* __restrict__ dw, float* __restrict__ db) {
  const int i = blockIdx.x * RED_TX + threadIdx.x;
  float a = 0.f;
  if (i <= K)
    for (int c = threadIdx.y; c < n_part; c += RED_TY) a += partial[(int64_t)c * (K + 1) + i];
  a = red_finish_32x8(a);
  if (threadIdx.y != 0 || i > K) return;
  if (i < K) dw[i] = a;
  else if (db) db[0] = a;
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_lin1_fwd(const float* h, int64_t ldh, const float* w, const float* bias, int64_t B, int K, float* out, void* stream) {
  B200_REQUIRE(h && w && out, B200CTR_E_INVALID_ARG, "lin1_fwd: null pointer");
  B200_REQUIRE(K % 4 == 0 && ldh % 4 == 0, B200CTR_E_UNSUPPORTED_SHAPE, "lin1_fwd: K and ldh must be multiples of 4");
  if (B == 0) return B200CTR_OK;
  lin1_fwd_kernel<<<(unsigned)ceil_div(B, (int64_t)8), 256, 0, as_stream(stream)>>>(h, ldh, w, bias, B, K, out);
  B200_LAUNCH_CHECK("lin1_fwd_kernel");
  return B200CTR_OK;
}
extern "C" int64_t b200ctr_lin1_bwd_workspace_bytes(int K) { return (int64_t)num_sms() * 2 * (K + 1) * 4; }
extern "C" int b200ctr_lin1_bwd(const float* dl, const float* act, int64_t ldact, const float* w, int64_t B, int K, float mask_scale,
                                float* dh, int64_t lddh, float* dw, float* db, void* workspace, void* stream) {
  B200_REQUIRE(dl && act && w && dh && dw && workspace, B200CTR_E_INVALID_ARG, "lin1_bwd: null pointer");
  B200_REQUIRE(K % 128 == 0 && K <= 128 * L1_MAXV && ldact % 4 == 0 && lddh % 4 == 0, B200CTR_E_UNSUPPORTED_SHAPE,
               "lin1_bwd: K=%d must be a multiple of 128 up to %d", K, 128 * L1_MAXV);
  B200_REQUIRE(B > 0, B200CTR_E_INVALID_ARG, "lin1_bwd: empty batch");
  const int grid = (int)std::min<int64_t>(ceil_div(B, (int64_t)8), (int64_t)num_sms() * 2);
  cudaStream_t st = as_stream(stream);
  lin1_bwd_kernel<<<grid, 256, (size_t)(8 * (K + 4) + 8) * 4, st>>>(dl, act, ldact, w, B, K, mask_scale, dh, lddh, (float*)workspace);
  B200_LAUNCH_CHECK("lin1_bwd_kernel");
  lin1_reduce_kernel<<<(K + RED_TX) / RED_TX, dim3(RED_TX, RED_TY), 0, st>>>((const float*)workspace, grid, K, dw, db);
  B200_LAUNCH_CHECK("lin1_reduce_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_rowrms_fwd_bf16(const float* x, int64_t ldx, const float* w, float eps, int64_t B, int64_t N, float* z,
                                       int64_t ld_z, float* r, void* z_bf16, void* stream) {
  B200_REQUIRE(x && w && r && (z || z_bf16), B200CTR_E_INVALID_ARG, "rowrms_fwd: null pointer");   // z may be NULL when only the bf16 copy is wanted
  B200_REQUIRE(N % 4 == 0 && ldx % 4 == 0 && ld_z % 4 == 0, B200CTR_E_UNSUPPORTED_SHAPE, "rowrms_fwd: N and leading dims must be multiples of 4");
  if (B == 0) return B200CTR_OK;
  rowrms_fwd_kernel<<<(unsigned)B, RR_THREADS, 0, as_stream(stream)>>>(x, ldx, w, eps, N, z, ld_z, r, (__nv_bfloat16*)z_bf16);
  B200_LAUNCH_CHECK("rowrms_fwd_kernel");
  return B200CTR_OK;
}
extern "C" int b200ctr_rowrms_fwd(const float* x, int64_t ldx, const float* w, float eps, int64_t B, int64_t N, float* z,
                                  int64_t ld_z, float* r, void* stream) {
  return b200ctr_rowrms_fwd_bf16(x, ldx, w, eps, B, N, z, ld_z, r, nullptr, stream);
}

extern "C" int b200ctr_rowrms_bwd(const float* dz, int64_t ld_dz, const float* x, int64_t ldx, const float* w, const float* r,
                                  int64_t B, int64_t N, float* dx, int64_t ld_dx, float* dw_partial, int n_partials, void* stream) {
  B200_REQUIRE(dz && x && w && r && dx && dw_partial, B200CTR_E_INVALID_ARG, "rowrms_bwd: null pointer");
  B200_REQUIRE(N % 4 == 0 && ldx % 4 == 0 && ld_dz % 4 == 0 && ld_dx % 4 == 0, B200CTR_E_UNSUPPORTED_SHAPE, "rowrms_bwd: alignment");
  B200_REQUIRE(N <= (int64_t)RR_MAXV * 4 * RR_THREADS, B200CTR_E_UNSUPPORTED_SHAPE, "rowrms_bwd: F*D=%lld exceeds %d", (long long)N, RR_MAXV * 4 * RR_THREADS);
  B200_REQUIRE(n_partials >= 1, B200CTR_E_INVALID_ARG, "rowrms_bwd: n_partials");
  const size_t smem = (size_t)N * 4;
  cudaError_t e = cudaFuncSetAttribute(rowrms_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(rowrms_bwd smem)");
  rowrms_bwd_kernel<<<n_partials, RR_THREADS, smem, as_stream(stream)>>>(dz, ld_dz, x, ldx, w, r, B, N, dx, ld_dx, dw_partial);
  B200_LAUNCH_CHECK("rowrms_bwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_fm_reduce(const float* A, int64_t B, int F, int C, float* s, float* quad, void* stream) {
  B200_REQUIRE(A && s && quad, B200CTR_E_INVALID_ARG, "fm_reduce: null pointer");
  if (B == 0) return B200CTR_OK;
  int threads = C <= 128 ? 128 : 256;
  fm_reduce_kernel<<<(unsigned)B, threads, 0, as_stream(stream)>>>(A, F, C, s, quad);
  B200_LAUNCH_CHECK("fm_reduce_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_fm_grad(float* A, const float* s, const float* dquad, int64_t B, int F, int C, void* stream) {
  B200_REQUIRE(A && s && dquad, B200CTR_E_INVALID_ARG, "fm_grad: null pointer");
  int64_t total = B * F * C;
  if (total == 0) return B200CTR_OK;
  fm_grad_kernel<<<(unsigned)ceil_div(total, 256), 256, 0, as_stream(stream)>>>(A, s, dquad, total, F, C);
  B200_LAUNCH_CHECK("fm_grad_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_se_gate_fwd(const float* colsum, float inv_B, const float* w1, const float* b1, const float* w2,
                                   const float* b2, int C, int Cr, float* mean, float* h1, float* gate, void* stream) {
  B200_REQUIRE(colsum && w1 && b1 && w2 && b2 && mean && h1 && gate, B200CTR_E_INVALID_ARG, "se_gate_fwd: null pointer");
  B200_REQUIRE(C > 0 && Cr > 0, B200CTR_E_INVALID_ARG, "se_gate_fwd: empty");
  se_gate_fwd_a_kernel<<<(Cr + 7) / 8, 256, 0, as_stream(stream)>>>(colsum, inv_B, w1, b1, C, Cr, mean, h1);
  B200_LAUNCH_CHECK("se_gate_fwd_a_kernel");
  se_gate_fwd_b_kernel<<<(C + 7) / 8, 256, 0, as_stream(stream)>>>(h1, w2, b2, C, Cr, gate);
  B200_LAUNCH_CHECK("se_gate_fwd_b_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_se_gate_bwd(const float* dgate, const float* mean, const float* h1, const float* gate, const float* w1,
                                   const float* w2, int C, int Cr, float* dw1, float* db1, float* dw2, float* db2, float* dmean,
                                   void* stream) {
  B200_REQUIRE(dgate && mean && h1 && gate && w1 && w2 && dw1 && db1 && dw2 && db2 && dmean, B200CTR_E_INVALID_ARG, "se_gate_bwd: null pointer");
  B200_REQUIRE(C > 0 && Cr > 0 && (size_t)2 * Cr * 4 <= 48 * 1024, B200CTR_E_UNSUPPORTED_SHAPE, "se_gate_bwd: C=%d Cr=%d unsupported", C, Cr);
  cudaStream_t st = as_stream(stream);
  se_gate_bwd_a_kernel<<<Cr, 256, 0, st>>>(dgate, mean, h1, gate, w2, C, Cr, dw1, db1);
  B200_LAUNCH_CHECK("se_gate_bwd_a_kernel");
  se_gate_bwd_b_kernel<<<(C + 127) / 128, 128, (size_t)2 * Cr * 4, st>>>(dgate, h1, gate, w1, db1, C, Cr, dw2, db2, dmean);
  B200_LAUNCH_CHECK("se_gate_bwd_b_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_se_apply_fwd(const float* inter, const float* gate, int64_t B, int C, float* out, int64_t ld_out,
                                    b200ctr_dropout_t drop, void* stream) {
  B200_REQUIRE(inter && out, B200CTR_E_INVALID_ARG, "se_apply_fwd: null pointer");
  if (B * C == 0) return B200CTR_OK;
  se_apply_fwd_kernel<<<(unsigned)ceil_div(B * C, 256), 256, 0, as_stream(stream)>>>(inter, gate, B, C, out, ld_out, drop);
  B200_LAUNCH_CHECK("se_apply_fwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_se_apply_bwd(const float* dout, int64_t ld_dout, const float* inter, const float* gate, const float* dmean,
                                    float inv_B, int64_t B, int C, float* dinter, float* gprod, b200ctr_dropout_t drop, void* stream) {
  B200_REQUIRE(dout && inter && (dinter || gprod), B200CTR_E_INVALID_ARG, "se_apply_bwd: null pointer");
  if (B * C == 0) return B200CTR_OK;
  if (C % 2 == 0)
    se_apply_bwd_pair_kernel<<<(unsigned)ceil_div(B * C / 2, 256), 256, 0, as_stream(stream)>>>(dout, ld_dout, inter, gate, dmean, inv_B, B, C, dinter, gprod, drop);
  else
    se_apply_bwd_kernel<<<(unsigned)ceil_div(B * C, 256), 256, 0, as_stream(stream)>>>(dout, ld_dout, inter, gate, dmean, inv_B, B, C, dinter, gprod, drop);
  B200_LAUNCH_CHECK("se_apply_bwd_kernel");
  return B200CTR_OK;
}

static bool headproj_ok(int H, int R, int P) {
  const int C = H * R;
  if (H <= 0 || R <= 0 || P <= 0 || R > 32 || H * P > HP_MAXC * HP_THREADS || C > HP_THREADS) return false;
  if ((H * P) % 4 != 0 || C % 4 != 0) return false;                       // 16-byte cp.async rows in the backward
  return ((size_t)C * (P + 1) + 4 + (size_t)2 * HP_RB * (H * P + C) + 4 * C) * 4 <= 200 * 1024;
}

extern "C" int b200ctr_headproj_fwd(const float* quad, const float* V, int64_t B, int H, int R, int P, float* inter, void* stream) {
  B200_REQUIRE(quad && V && inter, B200CTR_E_INVALID_ARG, "headproj_fwd: null pointer");
  B200_REQUIRE(headproj_ok(H, R, P), B200CTR_E_UNSUPPORTED_SHAPE, "headproj_fwd: heads=%d rank=%d proj=%d unsupported", H, R, P);
  if (B == 0) return B200CTR_OK;
  const size_t smem = ((size_t)H * R * P + 4 * H * R) * 4;
  cudaError_t e = cudaFuncSetAttribute(headproj_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(headproj_fwd smem)");
  const int grid = (int)std::min<int64_t>(ceil_div(B, 4), (int64_t)num_sms() * 2);
  headproj_fwd_kernel<<<grid, HP_THREADS, smem, as_stream(stream)>>>(quad, V, B, H, R, P, inter);
  B200_LAUNCH_CHECK("headproj_fwd_kernel");
  return B200CTR_OK;
}

extern "C" int64_t b200ctr_headproj_bwd_workspace_bytes(int H, int R, int P) { return (int64_t)num_sms() * H * R * P * 4; }

extern "C" int b200ctr_headproj_bwd(const float* quad, const float* dinter, const float* V, int64_t B, int H, int R, int P,
                                    float* dquad, float* dV, void* workspace, void* stream) {
  B200_REQUIRE(quad && dinter && V && dquad && dV && workspace, B200CTR_E_INVALID_ARG, "headproj_bwd: null pointer");
  B200_REQUIRE(headproj_ok(H, R, P), B200CTR_E_UNSUPPORTED_SHAPE, "headproj_bwd: heads=%d rank=%d proj=%d unsupported", H, R, P);
  B200_REQUIRE(B > 0, B200CTR_E_INVALID_ARG, "headproj_bwd: empty batch");
  const int C = H * R;
  const size_t smem = ((size_t)((C * (P + 1) + 3) & ~3) + (size_t)2 * HP_RB * (H * P + C)) * 4;
  const int grid = (int)std::min<int64_t>(ceil_div(B, (int64_t)HP_RB), (int64_t)num_sms());
  cudaStream_t st = as_stream(stream);
#define HP_LAUNCH(RM)                                                                                              \
  do {                                                                                                             \
    cudaError_t e = cudaFuncSetAttribute(headproj_bwd_kernel<RM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(headproj_bwd smem)");                          \
    headproj_bwd_kernel<RM><<<grid, HP_THREADS, smem, st>>>(quad, dinter, V, B, H, R, P, dquad, (float*)workspace); \
  } while (0)
  if (R <= 8) HP_LAUNCH(8); else if (R <= 16) HP_LAUNCH(16); else HP_LAUNCH(32);
#undef HP_LAUNCH
  B200_LAUNCH_CHECK("headproj_bwd_kernel");
  const int n = C * P;
  headproj_reduce_kernel<<<(n + RED_TX - 1) / RED_TX, dim3(RED_TX, RED_TY), 0, st>>>((const float*)workspace, grid, n, dV);
  B200_LAUNCH_CHECK("headproj_reduce_kernel");
  return B200CTR_OK;
}
