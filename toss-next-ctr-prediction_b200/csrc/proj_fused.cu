// Token-parallel projections of the DARE encoder layer's attention half on tcgen05 (the `amp: bf16` path of
// dare.py:64-66: MHA in_proj / out_proj around the attention core, residual, post-norm):
//
//   lin_fwd<192, false> : qkv = x . Win^T + b_in
//   lin_fwd<64,  true>  : s1 = x + o . Wout^T + b_out ; r1 = rsqrt(mean(s1^2)+eps) ; x1 = n1 * s1 * r1
//   lin_bwd<64,  true>  : d_s1 = RMSNorm'(dx1; s1, r1, n1) ; d_o = d_s1 . Wout ; dWout, db_out, dn1
//   lin_bwd<192, false> : dx = dqkv . Win + d_s1 ; dWin, db_in
//
// Same skeleton as ffn_fused.cu: persistent CTA per SM, the weight matrix resident in shared memory as a bf16 UMMA
// operand, 128-token tiles staged fp32 -> bf16 by loader warps (double buffered), one thread issuing tcgen05.mma,
// accumulators in TMEM, thread == token row in the epilogues, mbarrier hand-offs only.  The backward accumulates the
// weight gradient in TMEM across the persistent loop (M = 128 output features, K = 128 tokens, the K-major bf16 tile
// of dY re-read MN-major) and gets the BIAS gradient from the same MMAs: the B operand (the activation tile viewed
// MN-major, N = 64) is extended by 16 columns whose first one is all ones, so accumulator column 64 is colsum(dY).
#include "tc_common.cuh"
#include <algorithm>

namespace b200ctr {

constexpr int PJ_D = 64;
constexpr int PJ_TM = 128;
constexpr int PJ_NT = 13 * 32;
constexpr int PJ_X_BYTES = PJ_TM * PJ_D * 2;          // 16 KB bf16 activation tile
constexpr int PJ_ONES_BYTES = 2 * PJ_TM * 16;         // 4 KB: two extra 8-column groups (ones column + zeros)

// [128 tokens x 64] fp32 rows (stride ld floats) -> bf16 K-major core matrices (R = 128); NTHR cooperating threads
template <int NTHR>
__device__ __forceinline__ void pj_stage_x(unsigned char* dst, const float* __restrict__ x, int64_t ld, int64_t m0, int64_t T, int ptid) {
  constexpr int ITERS = 1024 / NTHR;
  float4 va[ITERS], vb[ITERS];
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int u = ptid + it * NTHR, r = u & 127, c = u >> 7;
    const int64_t m = m0 + r;
    va[it] = make_float4(0.f, 0.f, 0.f, 0.f); vb[it] = va[it];
    if (m < T) {
      const float4* p = reinterpret_cast<const float4*>(x + m * ld + c * 8);
      va[it] = __ldg(p); vb[it] = __ldg(p + 1);
    }
  }
#pragma unroll
  for (int it = 0; it < ITERS; ++it) {
    const int u = ptid + it * NTHR, r = u & 127, c = u >> 7;
    *reinterpret_cast<uint4*>(dst + (c * 16 + (r >> 3)) * 128 + (r & 7) * 16) =
        make_uint4(pack_bf16(va[it].x, va[it].y), pack_bf16(va[it].z, va[it].w), pack_bf16(vb[it].x, vb[it].y), pack_bf16(vb[it].z, vb[it].w));
  }
}

// =============================================================================================================
// forward
// =============================================================================================================
struct PjFwdBars {
  unsigned long long a_full[2], a_empty[2], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

template <int NOUT, bool POST_NORM>
__global__ void __launch_bounds__(PJ_NT, 1)
lin_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ resid, const float* __restrict__ nw, float eps, int64_t T,
               float* __restrict__ y, float* __restrict__ s_out, float* __restrict__ r_out) {
  constexpr int ACC_STRIDE = NOUT <= 64 ? 64 : 256;
  constexpr int TMEM_COLS = NOUT <= 64 ? 128 : 512;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* ws = smem;                                     // NOUT x 64 bf16, K-major B operand (R = NOUT)
  unsigned char* as = ws + NOUT * 128;                          // [2][16 KB]
  float* bs = reinterpret_cast<float*>(as + 2 * PJ_X_BYTES);    // NOUT
  float* ns = bs + NOUT;                                        // 64
  constexpr int SCR = 32 * ((POST_NORM ? 64 : 32) + 4);         // per-epilogue-warp transpose scratch (floats)
  float* scr_all = ns + PJ_D;                                   // 8 warps x SCR
  PjFwdBars* bars = reinterpret_cast<PjFwdBars*>(scr_all + 8 * SCR);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (T + PJ_TM - 1) / PJ_TM;
  const int n_it = (int)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (warp == 12) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->a_full[s]), 4); mbar_init(smem_u32(&bars->a_empty[s]), 1);
      mbar_init(smem_u32(&bars->acc_full[s]), 1); mbar_init(smem_u32(&bars->acc_empty[s]), 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int u = tid; u < NOUT * 8; u += PJ_NT) {
    const int n = u >> 3, kc = u & 7;
    const float4* p = reinterpret_cast<const float4*>(w + (size_t)n * PJ_D + kc * 8);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    *reinterpret_cast<uint4*>(ws + (kc * (NOUT / 8) + (n >> 3)) * 128 + (n & 7) * 16) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
  for (int i = tid; i < NOUT; i += PJ_NT) bs[i] = bias ? bias[i] : 0.f;
  if (tid < PJ_D) ns[tid] = (POST_NORM && nw) ? nw[tid] : 1.f;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 4) {
    // ---- loaders ----
    for (int it = 0; it < n_it; ++it) {
      const int ab = it & 1;
      if (it >= 2) mbar_wait(smem_u32(&bars->a_empty[ab]), (uint32_t)(((it >> 1) - 1) & 1));
      pj_stage_x<128>(as + ab * PJ_X_BYTES, x, PJ_D, ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * PJ_TM, T, tid);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->a_full[ab]));
    }
  } else if (warp < 12) {
    // ---- epilogue: group eg handles the tiles of parity eg; thread == token row ----
    const int ew = warp - 4, q = ew & 3, eg = ew >> 2;
    for (int it = eg; it < n_it; it += 2) {
      mbar_wait(smem_u32(&bars->acc_full[eg]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const int64_t m_w = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * PJ_TM + q * 32;      // first row of this warp
      const int64_t m = m_w + lane;
      const int rows_valid = (int)max((int64_t)0, min((int64_t)32, T - m_w));
      float* scr = scr_all + ew * SCR;
      const uint32_t taddr = tmem + eg * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
      if (POST_NORM) {
        uint32_t v[64];
        tmem_ld32(taddr, v);
        tmem_ld32(taddr + 32, v + 32);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[eg]));
        float ss = 0.f;
        if (m < T) {
          const float4* xr = reinterpret_cast<const float4*>(resid + m * PJ_D);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 xv = __ldg(xr + j);
            const float s0 = __uint_as_float(v[4 * j]) + bs[4 * j] + xv.x, s1 = __uint_as_float(v[4 * j + 1]) + bs[4 * j + 1] + xv.y;
            const float s2 = __uint_as_float(v[4 * j + 2]) + bs[4 * j + 2] + xv.z, s3 = __uint_as_float(v[4 * j + 3]) + bs[4 * j + 3] + xv.w;
            ss = fmaf(s0, s0, ss); ss = fmaf(s1, s1, ss); ss = fmaf(s2, s2, ss); ss = fmaf(s3, s3, ss);
            v[4 * j] = __float_as_uint(s0); v[4 * j + 1] = __float_as_uint(s1); v[4 * j + 2] = __float_as_uint(s2); v[4 * j + 3] = __float_as_uint(s3);
          }
        }
        const float r = rsqrtf(ss * (1.f / PJ_D) + eps);
        if (r_out && m < T) r_out[m] = r;
        if (s_out) warp_store_rows<64>(scr, v, s_out + m_w * PJ_D, PJ_D, rows_valid, lane);
#pragma unroll
        for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(ns[j] * __uint_as_float(v[j]) * r);
        warp_store_rows<64>(scr, v, y + m_w * PJ_D, PJ_D, rows_valid, lane);
      } else {
#pragma unroll 1
        for (int cc = 0; cc < NOUT / 32; ++cc) {
          uint32_t v[32];
          tmem_ld32(taddr + cc * 32, v);
          if (cc == NOUT / 32 - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[eg]));
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + bs[cc * 32 + j]);
          warp_store_rows<32>(scr, v, y + m_w * NOUT + cc * 32, NOUT, rows_valid, lane);
        }
      }
    }
  } else if (lane == 0) {
    // ---- MMA issuer ----
    const uint32_t idesc = umma_idesc(PJ_TM, NOUT, false, false);
    const uint32_t as_u = smem_u32(as), ws_u = smem_u32(ws);
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ab = it & 1;
      mbar_wait(smem_u32(&bars->a_full[ab]), (uint32_t)((it >> 1) & 1));
      if (it >= 2) mbar_wait(smem_u32(&bars->acc_empty[ab]), (uint32_t)(((it >> 1) - 1) & 1));
      tc_fence_after();
#pragma unroll 1
      for (int ks = 0; ks < PJ_D / 16; ++ks)
        umma_bf16(tmem + ab * ACC_STRIDE, make_desc(as_u + ab * PJ_X_BYTES + ks * 2 * (PJ_TM * 16), PJ_TM * 16, 128),
                  make_desc(ws_u + ks * 2 * (NOUT * 16), NOUT * 16, 128), idesc, ks > 0 ? 1u : 0u);
      umma_commit(smem_u32(&bars->acc_full[ab]));
      umma_commit(smem_u32(&bars->a_empty[ab]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    __syncwarp();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// =============================================================================================================
// backward
// =============================================================================================================
struct PjBwdBars {
  unsigned long long in_full[2], in_empty[2], dx_full[2], dx_empty[2], done;
  uint32_t tmem_base;
};

template <int NOUT, bool PRE_NORM>
__global__ void __launch_bounds__(PJ_NT, 1)
lin_bwd_kernel(const float* __restrict__ gin,      // PRE_NORM: dOut (T,64) = gradient of the norm's output ; else dY (T,NOUT)
               const float* __restrict__ s, const float* __restrict__ r, const float* __restrict__ nw,   // PRE_NORM only
               const float* __restrict__ xin,      // (T,64) forward input of the linear layer
               const float* __restrict__ w,        // (NOUT,64)
               const float* __restrict__ resid,    // optional (T,64), added to dX
               int64_t T, float* __restrict__ dy_out /* PRE_NORM: d_s (T,64) */, float* __restrict__ dx,
               float* __restrict__ wsp /* per-CTA partials */) {
  constexpr int NP = NOUT <= 128 ? 128 : 256;                   // dY columns padded to whole 128-row weight-gradient blocks
  constexpr int NBLK = NP / 128;
  constexpr int DY_BYTES = PJ_TM * NP * 2;
  constexpr int IN_BYTES = PJ_X_BYTES + PJ_ONES_BYTES + DY_BYTES;
  constexpr int TMEM_COLS = NBLK == 1 ? 256 : 512;
  constexpr int DW_COL = 128, DW_STRIDE = 128;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* wsm = smem;                                    // W as MN-major B operand (N = d_in, K = NOUT)
  unsigned char* ins = wsm + NOUT * 128;                        // [2][x 16 KB | ones 4 KB | dY]
  float* red = reinterpret_cast<float*>(ins + 2 * IN_BYTES);    // 8 warps x 64 (dn partials)
  // store-transpose scratch of the 4 epilogue warps (64-wide kernel only: at NOUT = 192 the extra 18 KB would push the
  // shared-memory carve-out to 228 KB and the 28 KB of L1 left no longer hold the loader's row lines -- measured slower)
  constexpr int SCR_FLOATS = NOUT <= 64 ? 4 * 32 * 36 : 0;
  float* scr_all = red + 8 * PJ_D;
  PjBwdBars* bars = reinterpret_cast<PjBwdBars*>(scr_all + SCR_FLOATS);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (T + PJ_TM - 1) / PJ_TM;
  const int n_it = (int)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const size_t ws_per_cta = (size_t)NP * 80 + PJ_D;             // dW|db accumulators [NP][80] + dn [64]
  float* wsc = wsp + (size_t)blockIdx.x * ws_per_cta;

  if (warp == 12) tmem_alloc(&bars->tmem_base, TMEM_COLS);
  if (tid == 0) {
    for (int sidx = 0; sidx < 2; ++sidx) {
      mbar_init(smem_u32(&bars->in_full[sidx]), 8); mbar_init(smem_u32(&bars->in_empty[sidx]), 1);
      mbar_init(smem_u32(&bars->dx_full[sidx]), 1); mbar_init(smem_u32(&bars->dx_empty[sidx]), 4);
    }
    mbar_init(smem_u32(&bars->done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // W [NOUT][64]: unit (row n, d-group ng) -> core (k-group n/8, n-group ng), row n%8, 8 consecutive d
  for (int u = tid; u < NOUT * 8; u += PJ_NT) {
    const int n = u >> 3, ng = u & 7;
    const float4* p = reinterpret_cast<const float4*>(w + (size_t)n * PJ_D + ng * 8);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    *reinterpret_cast<uint4*>(wsm + ((n >> 3) * 8 + ng) * 128 + (n & 7) * 16) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
  // static parts of both input buffers: the ones / zeros column groups after the x tile, the zero padding of dY
  for (int bsel = 0; bsel < 2; ++bsel) {
    unsigned char* ones = ins + bsel * IN_BYTES + PJ_X_BYTES;
    for (int i = tid; i < 2 * PJ_TM; i += PJ_NT)        // group 8: column 64 == 1.0 (bf16 0x3F80), group 9: zeros
      *reinterpret_cast<uint4*>(ones + i * 16) = make_uint4(i < PJ_TM ? 0x00003F80u : 0u, 0u, 0u, 0u);
    unsigned char* dyb = ones + PJ_ONES_BYTES;
    for (int i = tid; i < (NP - NOUT) / 8 * PJ_TM; i += PJ_NT)
      *reinterpret_cast<uint4*>(dyb + (NOUT / 8) * (PJ_TM * 16) + i * 16) = make_uint4(0u, 0u, 0u, 0u);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 8) {
    // ---- loaders: x tile, dY tile (PRE_NORM: dY = RMSNorm'(dOut) computed here, lane owns 2 columns) ----
    float2 dn_acc = make_float2(0.f, 0.f);
    float2 nwv = make_float2(0.f, 0.f);
    if (PRE_NORM) nwv = reinterpret_cast<const float2*>(nw)[lane];
    for (int it = 0; it < n_it; ++it) {
      const int ib = it & 1;
      if (it >= 2) mbar_wait(smem_u32(&bars->in_empty[ib]), (uint32_t)(((it >> 1) - 1) & 1));
      const int64_t m0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * PJ_TM;
      unsigned char* xs = ins + ib * IN_BYTES;
      unsigned char* dys = xs + PJ_X_BYTES + PJ_ONES_BYTES;
      pj_stage_x<256>(xs, xin, PJ_D, m0, T, tid);
      if (PRE_NORM) {
#pragma unroll 1
        for (int rr = 0; rr < 16; rr += 4) {
          float2 d[4], sv[4];
          float rv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t m = m0 + warp * 16 + rr + u;
            d[u] = make_float2(0.f, 0.f); sv[u] = d[u]; rv[u] = 0.f;
            if (m < T) {
              d[u] = __ldg(reinterpret_cast<const float2*>(gin + m * PJ_D) + lane);
              sv[u] = __ldg(reinterpret_cast<const float2*>(s + m * PJ_D) + lane);
              rv[u] = __ldg(r + m);
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int row = warp * 16 + rr + u;
            const int64_t m = m0 + row;
            const float2 g = make_float2(d[u].x * nwv.x, d[u].y * nwv.y);
            const float dot = warp_sum(g.x * sv[u].x + g.y * sv[u].y) * (1.f / PJ_D);
            const float c = rv[u] * rv[u] * rv[u] * dot;
            const float2 o = make_float2(rv[u] * g.x - sv[u].x * c, rv[u] * g.y - sv[u].y * c);
            dn_acc.x = fmaf(d[u].x * sv[u].x, rv[u], dn_acc.x);
            dn_acc.y = fmaf(d[u].y * sv[u].y, rv[u], dn_acc.y);
            if (m < T) reinterpret_cast<float2*>(dy_out + m * PJ_D)[lane] = o;
            *reinterpret_cast<uint32_t*>(dys + ((lane >> 2) * 16 + (row >> 3)) * 128 + (row & 7) * 16 + (lane & 3) * 4) = pack_bf16(o.x, o.y);
          }
        }
      } else {
        constexpr int UNITS = PJ_TM * (NOUT / 8), ITERS = UNITS / 256;
#pragma unroll 1
        for (int it0 = 0; it0 < ITERS; it0 += 4) {
          float4 va[4], vb[4];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int u = tid + (it0 + q4) * 256, rrow = u & 127, c = u >> 7;
            const int64_t m = m0 + rrow;
            va[q4] = make_float4(0.f, 0.f, 0.f, 0.f); vb[q4] = va[q4];
            if (m < T) {
              const float4* p = reinterpret_cast<const float4*>(gin + m * NOUT + c * 8);
              va[q4] = __ldg(p); vb[q4] = __ldg(p + 1);
            }
          }
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const int u = tid + (it0 + q4) * 256, rrow = u & 127, c = u >> 7;
            *reinterpret_cast<uint4*>(dys + (c * 16 + (rrow >> 3)) * 128 + (rrow & 7) * 16) =
                make_uint4(pack_bf16(va[q4].x, va[q4].y), pack_bf16(va[q4].z, va[q4].w), pack_bf16(vb[q4].x, vb[q4].y), pack_bf16(vb[q4].z, vb[q4].w));
          }
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->in_full[ib]));
    }
    if (PRE_NORM) reinterpret_cast<float2*>(red + warp * PJ_D)[lane] = dn_acc;
  } else if (warp < 12) {
    // ---- dX epilogue (thread == token row), then the weight-gradient accumulators ----
    const int q = warp & 3;
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ib = it & 1;
      mbar_wait(smem_u32(&bars->dx_full[ib]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      uint32_t v[64];
      const uint32_t taddr = tmem + ib * 64 + ((uint32_t)(q * 32) << 16);
      tmem_ld32(taddr, v);
      tmem_ld32(taddr + 32, v + 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->dx_empty[ib]));
      const int64_t m_w = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * PJ_TM + q * 32;
      const int64_t m = m_w + lane;
      const int rows_valid = (int)max((int64_t)0, min((int64_t)32, T - m_w));
      if (resid && m < T) {
        const float4* rp = reinterpret_cast<const float4*>(resid + m * PJ_D);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 a = __ldg(rp + j);
          v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + a.x); v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + a.y);
          v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + a.z); v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + a.w);
        }
      }
      if (NOUT <= 64) {          // measured: the transposed store pays off for the 64-wide kernel, not for the 192-wide one
        float* scr = scr_all + (warp - 8) * (32 * 36);
        warp_store_rows<32>(scr, v, dx + m_w * PJ_D, PJ_D, rows_valid, lane);
        warp_store_rows<32>(scr, v + 32, dx + m_w * PJ_D + 32, PJ_D, rows_valid, lane);
      } else if (m < T) {
        float4* op = reinterpret_cast<float4*>(dx + m * PJ_D);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          op[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
    }
    mbar_wait(smem_u32(&bars->done), 0);
    tc_fence_after();
    for (int b = 0; b < NBLK; ++b) {
      uint32_t v[80];
      const uint32_t taddr = tmem + DW_COL + b * DW_STRIDE + ((uint32_t)(q * 32) << 16);
      tmem_ld32(taddr, v);
      tmem_ld32(taddr + 32, v + 32);
      tmem_ld16(taddr + 64, v + 64);
      float4* dst = reinterpret_cast<float4*>(wsc + (size_t)(b * 128 + q * 32 + lane) * 80);
#pragma unroll
      for (int j = 0; j < 20; ++j)
        dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    }
  } else if (lane == 0) {
    // ---- MMA issuer ----
    const uint32_t idesc_dx = umma_idesc(PJ_TM, PJ_D, false, true);       // dY (K-major) . W (MN-major)
    const uint32_t idesc_dw = umma_idesc(128, 80, true, true);            // dY^T block (MN view) . [x | 1 | 0] (MN view)
    const uint32_t in_u = smem_u32(ins), w_u = smem_u32(wsm);
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ib = it & 1;
      mbar_wait(smem_u32(&bars->in_full[ib]), (uint32_t)((it >> 1) & 1));
      if (it >= 2) mbar_wait(smem_u32(&bars->dx_empty[ib]), (uint32_t)(((it >> 1) - 1) & 1));
      tc_fence_after();
      const uint32_t xs_u = in_u + ib * IN_BYTES, dy_u = xs_u + PJ_X_BYTES + PJ_ONES_BYTES;
#pragma unroll 1
      for (int ks = 0; ks < NOUT / 16; ++ks)
        umma_bf16(tmem + ib * 64, make_desc(dy_u + ks * 2 * (PJ_TM * 16), PJ_TM * 16, 128), make_desc(w_u + ks * 2 * 1024, 1024, 128),
                  idesc_dx, ks > 0 ? 1u : 0u);
      umma_commit(smem_u32(&bars->dx_full[ib]));
      for (int b = 0; b < NBLK; ++b) {
#pragma unroll 1
        for (int ks = 0; ks < 8; ++ks)      // K = 128 tokens; MN-major views: k-group (8 tokens) stride 128 B, mn-group stride 2048 B
          umma_bf16(tmem + DW_COL + b * DW_STRIDE, make_desc(dy_u + b * 16 * (PJ_TM * 16) + ks * 256, 128, 2048),
                    make_desc(xs_u + ks * 256, 128, 2048), idesc_dw, (it > 0 || ks > 0) ? 1u : 0u);
      }
      umma_commit(smem_u32(&bars->in_empty[ib]));
    }
    umma_commit(smem_u32(&bars->done));
  }
  tc_fence_before();
  __syncthreads();
  if (PRE_NORM && tid < PJ_D) {
    float a = 0.f;
    for (int wv = 0; wv < 8; ++wv) a += red[wv * PJ_D + tid];
    wsc[(size_t)NP * 80 + tid] = a;
  }
  if (warp == 12) {
    __syncwarp();
    tmem_dealloc(tmem, TMEM_COLS);
  }
}

// deterministic cross-CTA reduction: dW (NOUT,64), db (NOUT), dn (64)
__global__ void lin_bwd_reduce_kernel(const float* __restrict__ wsp, int n_cta, int NP, int NOUT, float* __restrict__ dw,
                                      float* __restrict__ db, float* __restrict__ dn) {
  const size_t per = (size_t)NP * 80 + PJ_D;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < NOUT * 65) {
    const int n = i / 65, c = i % 65;
    float acc = 0.f;
    for (int k = 0; k < n_cta; ++k) acc += wsp[(size_t)k * per + (size_t)n * 80 + c];
    if (c < 64) dw[n * PJ_D + c] = acc;
    else if (db) db[n] = acc;
  } else if (i < NOUT * 65 + PJ_D && dn) {
    const int d = i - NOUT * 65;
    float acc = 0.f;
    for (int k = 0; k < n_cta; ++k) acc += wsp[(size_t)k * per + (size_t)NP * 80 + d];
    dn[d] = acc;
  }
}

template <int NOUT, bool POST_NORM>
static int launch_lin_fwd(const float* x, const float* w, const float* bias, const float* resid, const float* nw, float eps,
                          int64_t T, float* y, float* s_out, float* r_out, cudaStream_t st) {
  const size_t smem = (size_t)NOUT * 128 + 2 * PJ_X_BYTES + ((size_t)NOUT + PJ_D) * 4 +
                      (size_t)8 * 32 * ((POST_NORM ? 64 : 32) + 4) * 4 + sizeof(PjFwdBars) + 16;
  auto fn = lin_fwd_kernel<NOUT, POST_NORM>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(lin_fwd smem)");
  const int grid = (int)std::min<int64_t>(ceil_div(T, (int64_t)PJ_TM), (int64_t)num_sms());
  fn<<<grid, PJ_NT, smem, st>>>(x, w, bias, resid, nw, eps, T, y, s_out, r_out);
  return B200CTR_OK;
}

template <int NOUT, bool PRE_NORM>
static int launch_lin_bwd(const float* gin, const float* s, const float* r, const float* nw, const float* xin, const float* w,
                          const float* resid, int64_t T, float* dy_out, float* dx, float* dw, float* db, float* dn, float* wsp,
                          cudaStream_t st) {
  constexpr int NP = NOUT <= 128 ? 128 : 256;
  const size_t smem = (size_t)NOUT * 128 + 2 * (size_t)(PJ_X_BYTES + PJ_ONES_BYTES + PJ_TM * NP * 2) + 8 * PJ_D * 4 + (size_t)(NOUT <= 64 ? 4 * 32 * 36 * 4 : 0) +
                      sizeof(PjBwdBars) + 16;
  auto fn = lin_bwd_kernel<NOUT, PRE_NORM>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(lin_bwd smem)");
  const int grid = (int)std::min<int64_t>(ceil_div(T, (int64_t)PJ_TM), (int64_t)num_sms());
  fn<<<grid, PJ_NT, smem, st>>>(gin, s, r, nw, xin, w, resid, T, dy_out, dx, wsp);
  cudaError_t le = cudaPeekAtLastError();
  if (le != cudaSuccess) return cuda_fail(le, "lin_bwd_kernel");
  const int n_out = NOUT * 65 + PJ_D;
  lin_bwd_reduce_kernel<<<(n_out + 255) / 256, 256, 0, st>>>(wsp, grid, NP, NOUT, dw, db, PRE_NORM ? dn : nullptr);
  return B200CTR_OK;
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_inproj_fwd(const float* x, const float* w_in, const float* b_in, int64_t T, int D, float* qkv, void* stream) {
  B200_REQUIRE(x && w_in && qkv, B200CTR_E_INVALID_ARG, "inproj_fwd: null pointer");
  B200_REQUIRE(D == PJ_D, B200CTR_E_UNSUPPORTED_SHAPE, "inproj_fwd: D=%d unsupported (64 only)", D);
  if (T == 0) return B200CTR_OK;
  int rc = launch_lin_fwd<192, false>(x, w_in, b_in, nullptr, nullptr, 0.f, T, qkv, nullptr, nullptr, as_stream(stream));
  if (rc) return rc;
  B200_LAUNCH_CHECK("lin_fwd_kernel<192>");
  return B200CTR_OK;
}

extern "C" int b200ctr_outproj_norm_fwd(const float* o, const float* w_out, const float* b_out, const float* x_resid,
                                        const float* n1w, float eps, int64_t T, int D, float* s1_out, float* r1_out,
                                        float* x1_out, void* stream) {
  B200_REQUIRE(o && w_out && x_resid && n1w && x1_out, B200CTR_E_INVALID_ARG, "outproj_norm_fwd: null pointer");
  B200_REQUIRE(D == PJ_D, B200CTR_E_UNSUPPORTED_SHAPE, "outproj_norm_fwd: D=%d unsupported (64 only)", D);
  if (T == 0) return B200CTR_OK;
  int rc = launch_lin_fwd<64, true>(o, w_out, b_out, x_resid, n1w, eps, T, x1_out, s1_out, r1_out, as_stream(stream));
  if (rc) return rc;
  B200_LAUNCH_CHECK("lin_fwd_kernel<64,norm>");
  return B200CTR_OK;
}

extern "C" int64_t b200ctr_proj_bwd_workspace_bytes(void) { return (int64_t)num_sms() * ((int64_t)256 * 80 + PJ_D) * 4; }

extern "C" int b200ctr_outproj_norm_bwd(const float* dx1, const float* s1, const float* r1, const float* n1w, const float* o,
                                        const float* w_out, int64_t T, int D, float* ds1_out, float* d_o, float* dw_out,
                                        float* db_out, float* dn1, void* workspace, void* stream) {
  B200_REQUIRE(dx1 && s1 && r1 && n1w && o && w_out && ds1_out && d_o && dw_out && db_out && dn1 && workspace, B200CTR_E_INVALID_ARG,
               "outproj_norm_bwd: null pointer");
  B200_REQUIRE(D == PJ_D, B200CTR_E_UNSUPPORTED_SHAPE, "outproj_norm_bwd: D=%d unsupported (64 only)", D);
  B200_REQUIRE(T > 0, B200CTR_E_INVALID_ARG, "outproj_norm_bwd: empty input");
  int rc = launch_lin_bwd<64, true>(dx1, s1, r1, n1w, o, w_out, nullptr, T, ds1_out, d_o, dw_out, db_out, dn1, (float*)workspace,
                                    as_stream(stream));
  if (rc) return rc;
  B200_LAUNCH_CHECK("lin_bwd_kernel<64,norm>");
  return B200CTR_OK;
}

extern "C" int b200ctr_inproj_bwd(const float* dqkv, const float* x, const float* w_in, const float* d_resid, int64_t T, int D,
                                  float* dx, float* dw_in, float* db_in, void* workspace, void* stream) {
  B200_REQUIRE(dqkv && x && w_in && dx && dw_in && db_in && workspace, B200CTR_E_INVALID_ARG, "inproj_bwd: null pointer");
  B200_REQUIRE(D == PJ_D, B200CTR_E_UNSUPPORTED_SHAPE, "inproj_bwd: D=%d unsupported (64 only)", D);
  B200_REQUIRE(T > 0, B200CTR_E_INVALID_ARG, "inproj_bwd: empty input");
  int rc = launch_lin_bwd<192, false>(dqkv, nullptr, nullptr, nullptr, x, w_in, d_resid, T, nullptr, dx, dw_in, db_in, nullptr,
                                      (float*)workspace, as_stream(stream));
  if (rc) return rc;
  B200_LAUNCH_CHECK("lin_bwd_kernel<192>");
  return B200CTR_OK;
}
