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
#include "h2_common.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace b200ctr {

constexpr int PJ_D = 64;
constexpr int PJ_TM = 128;
constexpr int PJ_NT = 13 * 32;
// K-direction stride of the 8-column core-matrix groups of a staged tile: 128 rows x 16 bytes + 16 bytes of padding, so
// that the column groups of one token row fall into different shared-memory banks (the coalesced loaders below write
// consecutive column groups of the same row from neighbouring lanes)
constexpr int PJ_LBO = PJ_TM * 16 + 16;
constexpr int PJ_X_BYTES = 8 * PJ_LBO;                // bf16 activation tile (64 columns)
constexpr int PJ_ONES_BYTES = 2 * PJ_LBO;             // two extra 8-column groups (ones column + zeros)

// Coalesced staging of a [128 tokens x W] fp32 tile as a bf16 K-major UMMA operand.  The tile is a flat list of 16-byte
// chunks; one pass = NTHR consecutive chunks, i.e. every LDG.128 of a warp reads 512 contiguous bytes (4 L1 wavefronts;
// the row-per-lane pattern this replaces cost 32, and the kernels were bound by exactly that).  Two passes (A, B) are
// converted together: 4 fp32 -> 8 bytes of bf16, neighbouring lanes hold the two halves of one 16-byte core-matrix row,
// so even lanes collect the A halves and odd lanes the B halves with one exchange and every lane stores 16 bytes.
template <int W, int NTHR>
__device__ __forceinline__ void pj_load_pairs(float4 (&va)[4], float4 (&vb)[4], const float* __restrict__ src, int64_t m0,
                                              int64_t T, int pair0, int tid) {
  constexpr int C4 = W / 4;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int qa = 2 * (pair0 + k) * NTHR + tid, qb = qa + NTHR;
    const int ra = qa / C4, rb = qb / C4;
    va[k] = make_float4(0.f, 0.f, 0.f, 0.f); vb[k] = va[k];
    if (m0 + ra < T) va[k] = __ldg(reinterpret_cast<const float4*>(src + (m0 + ra) * W) + (qa - ra * C4));
    if (m0 + rb < T) vb[k] = __ldg(reinterpret_cast<const float4*>(src + (m0 + rb) * W) + (qb - rb * C4));
  }
}
template <int W, int NTHR>
__device__ __forceinline__ void pj_store_pairs(unsigned char* dst, const float4 (&va)[4], const float4 (&vb)[4], int pair0, int tid) {
  constexpr int C4 = W / 4;
  const bool odd = tid & 1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const uint32_t a0 = pack_bf16(va[k].x, va[k].y), a1 = pack_bf16(va[k].z, va[k].w);
    const uint32_t b0 = pack_bf16(vb[k].x, vb[k].y), b1 = pack_bf16(vb[k].z, vb[k].w);
    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, odd ? a0 : b0, 1), r1 = __shfl_xor_sync(0xffffffffu, odd ? a1 : b1, 1);
    const int q = (2 * (pair0 + k) + (odd ? 1 : 0)) * NTHR + tid;          // the chunk whose 16-byte group this lane stores
    const int row = q / C4, c8 = (q - row * C4) >> 1;
    *reinterpret_cast<uint4*>(dst + c8 * PJ_LBO + (row >> 3) * 128 + (row & 7) * 16) =
        odd ? make_uint4(r0, r1, b0, b1) : make_uint4(a0, a1, r0, r1);
  }
}

// row-per-lane variant (each thread reads 32 contiguous bytes of its own row; 32 L1 wavefronts per LDG.128)
template <int NTHR>
__device__ __forceinline__ void pj_stage_x_rows(unsigned char* dst, const float* __restrict__ x, int64_t ld, int64_t m0, int64_t T, int ptid) {
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
    *reinterpret_cast<uint4*>(dst + c * PJ_LBO + (r >> 3) * 128 + (r & 7) * 16) =
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

template <int NOUT, bool POST_NORM, bool COAL>
__global__ void __launch_bounds__(PJ_NT, 1)
lin_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
               const float* __restrict__ resid, const float* __restrict__ nw, float eps, int64_t T,
               float* __restrict__ y, float* __restrict__ s_out, float* __restrict__ r_out, uint32_t* __restrict__ y16, float qscale,
               int x16 /* x is bf16 (T, 64): a 16-byte chunk already is one core-matrix row of the A operand */) {
  // y16 (NOUT = 192 only): the output goes out as fp16 pairs, (T, NOUT / 2) words, with its first 64 columns (q) times qscale --
  // the operand format of the packed-half attention kernels (attn_h16.cu), which otherwise convert the fp32 tensor themselves
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
    // ---- loaders: two batches of 4 pass pairs per tile; the next batch's loads are in flight while one is stored ----
    auto tile_m0 = [&](int it) { return ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * PJ_TM; };
    float4 ra[2][4], rb[2][4];
    if (COAL && !x16 && n_it > 0) pj_load_pairs<PJ_D, 128>(ra[0], rb[0], x, tile_m0(0), T, 0, tid);
    for (int it = 0; it < n_it; ++it) {
      const int ab = it & 1;
      if (it >= 2) mbar_wait(smem_u32(&bars->a_empty[ab]), (uint32_t)(((it >> 1) - 1) & 1));
      unsigned char* dst = as + ab * PJ_X_BYTES;
      if (x16) {
        const int64_t m0 = tile_m0(it);
        uint4 c[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int q = k * 128 + tid, row = q >> 3;
          c[k] = make_uint4(0u, 0u, 0u, 0u);
          if (m0 + row < T) c[k] = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint16_t*>(x) + (m0 + row) * PJ_D) + (q & 7));
        }
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int q = k * 128 + tid, row = q >> 3;
          *reinterpret_cast<uint4*>(dst + (q & 7) * PJ_LBO + (row >> 3) * 128 + (row & 7) * 16) = c[k];
        }
      } else if (COAL) {
        pj_load_pairs<PJ_D, 128>(ra[1], rb[1], x, tile_m0(it), T, 4, tid);
        pj_store_pairs<PJ_D, 128>(dst, ra[0], rb[0], 0, tid);
        if (it + 1 < n_it) pj_load_pairs<PJ_D, 128>(ra[0], rb[0], x, tile_m0(it + 1), T, 0, tid);
        pj_store_pairs<PJ_D, 128>(dst, ra[1], rb[1], 4, tid);
      } else {
        pj_stage_x_rows<128>(dst, x, PJ_D, tile_m0(it), T, tid);
      }
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
        // bias in the row-per-lane domain, then through the scratch: 16 lanes x 16 bytes cover one row, so the residual
        // load and both stores are whole-row accesses and the row's sum of squares is a 16-lane shuffle reduction
#pragma unroll
        for (int j = 0; j < 16; ++j)
          *reinterpret_cast<float4*>(scr + lane * 68 + 4 * j) =
              make_float4(__uint_as_float(v[4 * j]) + bs[4 * j], __uint_as_float(v[4 * j + 1]) + bs[4 * j + 1],
                          __uint_as_float(v[4 * j + 2]) + bs[4 * j + 2], __uint_as_float(v[4 * j + 3]) + bs[4 * j + 3]);
        __syncwarp();
        const int rsub = lane >> 4, c = lane & 15;
        const float4 nv = *reinterpret_cast<const float4*>(ns + c * 4);
#pragma unroll 1
        for (int r0 = 0; r0 < 32; r0 += 8) {
          float4 xv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int row = r0 + 2 * u + rsub;
            xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < rows_valid) xv[u] = __ldg(reinterpret_cast<const float4*>(resid + (m_w + row) * PJ_D) + c);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int row = r0 + 2 * u + rsub;
            const float4 a = *reinterpret_cast<const float4*>(scr + row * 68 + c * 4);
            const float4 sv = make_float4(a.x + xv[u].x, a.y + xv[u].y, a.z + xv[u].z, a.w + xv[u].w);
            float ss = sv.x * sv.x + sv.y * sv.y + sv.z * sv.z + sv.w * sv.w;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
            const float r = rsqrtf(ss * (1.f / PJ_D) + eps);
            if (row < rows_valid) {
              const int64_t mr = m_w + row;
              if (s_out) *reinterpret_cast<float4*>(s_out + mr * PJ_D + c * 4) = sv;
              *reinterpret_cast<float4*>(y + mr * PJ_D + c * 4) = make_float4(nv.x * sv.x * r, nv.y * sv.y * r, nv.z * sv.z * r, nv.w * sv.w * r);
              if (r_out && c == 0) r_out[mr] = r;
            }
          }
        }
        __syncwarp();
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
          if (y16) {
            const float sc = cc < 2 ? qscale : 1.f;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              v[j] = h2_pack((__uint_as_float(v[2 * j]) + bs[cc * 32 + 2 * j]) * sc, (__uint_as_float(v[2 * j + 1]) + bs[cc * 32 + 2 * j + 1]) * sc);
            warp_store_rows<16>(scr, v, reinterpret_cast<float*>(y16) + m_w * (NOUT / 2) + cc * 16, NOUT / 2, rows_valid, lane);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + bs[cc * 32 + j]);
            warp_store_rows<32>(scr, v, y + m_w * NOUT + cc * 32, NOUT, rows_valid, lane);
          }
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
        umma_bf16(tmem + ab * ACC_STRIDE, make_desc(as_u + ab * PJ_X_BYTES + ks * 2 * PJ_LBO, PJ_LBO, 128),
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
__global__ void __launch_bounds__(PJ_NT, 1)      // 13 warps: 4 on one SM sub-partition (16K registers) -> 128 registers per thread at most
lin_bwd_kernel(const float* __restrict__ gin,      // PRE_NORM: dOut (T,64) = gradient of the norm's output ; else dY (T,NOUT)
               const float* __restrict__ s, const float* __restrict__ r, const float* __restrict__ nw,   // PRE_NORM only
               const float* __restrict__ xin,      // (T,64) forward input of the linear layer
               const float* __restrict__ w,        // (NOUT,64)
               const float* __restrict__ resid,    // optional (T,64), added to dX
               int64_t T, float* __restrict__ dy_out /* PRE_NORM: d_s (T,64) */, float* __restrict__ dx,
               float* __restrict__ wsp /* per-CTA partials */, int gin16 /* !PRE_NORM: dY is bf16 (T,NOUT) */,
               int xin16 /* xin is bf16 (T,64) */, int dx16 /* dX is written as bf16 (T,64); no residual */) {
  constexpr int NP = NOUT <= 128 ? 128 : 256;                   // dY columns padded to whole 128-row weight-gradient blocks
  constexpr int NBLK = NP / 128;
  constexpr int DY_BYTES = (NP / 8) * PJ_LBO;
  constexpr int IN_BYTES = PJ_X_BYTES + PJ_ONES_BYTES + DY_BYTES;
  constexpr int TMEM_COLS = NBLK == 1 ? 256 : 512;
  constexpr int DW_COL = 128, DW_STRIDE = 128;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* wsm = smem;                                    // W as MN-major B operand (N = d_in, K = NOUT)
  unsigned char* ins = wsm + NOUT * 128;                        // [2][x 16 KB | ones 4 KB | dY]
  float* red = reinterpret_cast<float*>(ins + 2 * IN_BYTES);    // 8 warps x 64 (dn partials)
  // store-transpose scratch of the 4 epilogue warps (the loaders read whole lines once, so the small L1 left by the
  // 228 KB carve-out at NOUT = 192 no longer matters)
  constexpr int SCR_FLOATS = 4 * 32 * 36;
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
      *reinterpret_cast<uint4*>(ones + (i >> 7) * PJ_LBO + (i & 127) * 16) = make_uint4(i < PJ_TM ? 0x00003F80u : 0u, 0u, 0u, 0u);
    unsigned char* dyb = ones + PJ_ONES_BYTES;
    for (int i = tid; i < (NP - NOUT) / 8 * PJ_TM; i += PJ_NT)
      *reinterpret_cast<uint4*>(dyb + (NOUT / 8 + (i >> 7)) * PJ_LBO + (i & 127) * 16) = make_uint4(0u, 0u, 0u, 0u);
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
    // Batch 0 = the x tile (4 pass pairs of 256 chunks), batches 1.. = the dY tile.  The loads of batch b+1 (and of the
    // next tile's first batch) are issued before batch b is converted and stored: 32 KB per SM stay in flight.
    constexpr int NBATCH = 1 + (PRE_NORM ? 0 : PJ_TM * (NOUT / 4) / 256 / 8);
    static_assert(NBATCH % 2 == 0 || NBATCH == 1, "register sets alternate with a static parity");
    auto tile_m0 = [&](int it) { return ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * PJ_TM; };
    // bf16 dY: a 16-byte chunk already is one core-matrix row (8 bf16), 128 * NOUT / 8 chunks = 4 passes per batch
    constexpr int C8 = NOUT / 8;
    auto load_batch = [&](int it, int bi, float4 (&va)[4], float4 (&vb)[4]) {
      if (bi == 0 && xin16) {       // bf16 x tile: 128 rows x 8 chunks of 16 bytes, 4 per thread
        const int64_t m0 = tile_m0(it);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int q = k * 256 + tid, row = q >> 3;
          va[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m0 + row < T) va[k] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const uint16_t*>(xin) + (m0 + row) * PJ_D) + (q & 7));
        }
      } else if (bi == 0) pj_load_pairs<PJ_D, 256>(va, vb, xin, tile_m0(it), T, 0, tid);
      else if (!gin16) pj_load_pairs<NOUT, 256>(va, vb, gin, tile_m0(it), T, (bi - 1) * 4, tid);
      else {
        const int64_t m0 = tile_m0(it);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int q = ((bi - 1) * 4 + k) * 256 + tid, row = q / C8;
          va[k] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m0 + row < T)
            va[k] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const __nv_bfloat16*>(gin) + (m0 + row) * NOUT) + (q - row * C8));
        }
      }
    };
    auto store_batch = [&](unsigned char* xs, unsigned char* dys, int bi, const float4 (&va)[4], const float4 (&vb)[4]) {
      if (bi == 0 && xin16) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int q = k * 256 + tid, row = q >> 3;
          *reinterpret_cast<float4*>(xs + (q & 7) * PJ_LBO + (row >> 3) * 128 + (row & 7) * 16) = va[k];
        }
      } else if (bi == 0) pj_store_pairs<PJ_D, 256>(xs, va, vb, 0, tid);
      else if (!gin16) pj_store_pairs<NOUT, 256>(dys, va, vb, (bi - 1) * 4, tid);
      else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int q = ((bi - 1) * 4 + k) * 256 + tid, row = q / C8, c8 = q - row * C8;
          *reinterpret_cast<float4*>(dys + c8 * PJ_LBO + (row >> 3) * 128 + (row & 7) * 16) = va[k];
        }
      }
    };
    float4 ra[2][4], rb[2][4];
    if (n_it > 0) load_batch(0, 0, ra[0], rb[0]);
    for (int it = 0; it < n_it; ++it) {
      const int ib = it & 1;
      if (it >= 2) mbar_wait(smem_u32(&bars->in_empty[ib]), (uint32_t)(((it >> 1) - 1) & 1));
      const int64_t m0 = tile_m0(it);
      unsigned char* xs = ins + ib * IN_BYTES;
      unsigned char* dys = xs + PJ_X_BYTES + PJ_ONES_BYTES;
      if (PRE_NORM) {
        // rows of this warp, 8 at a time: gradient of the norm's output, saved sum and 1/rms -> dY = RMSNorm'(dOut)
        float2 d[8], sv[8];
        float rv[8];
        auto load_rows = [&](int rr) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int64_t m = m0 + warp * 16 + rr + u;
            d[u] = make_float2(0.f, 0.f); sv[u] = d[u]; rv[u] = 0.f;
            if (m < T) {
              d[u] = __ldg(reinterpret_cast<const float2*>(gin + m * PJ_D) + lane);
              sv[u] = __ldg(reinterpret_cast<const float2*>(s + m * PJ_D) + lane);
              rv[u] = __ldg(r + m);
            }
          }
        };
        auto norm_rows = [&](int rr) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const int row = warp * 16 + rr + u;
            const int64_t m = m0 + row;
            const float2 g = make_float2(d[u].x * nwv.x, d[u].y * nwv.y);
            const float dot = warp_sum(g.x * sv[u].x + g.y * sv[u].y) * (1.f / PJ_D);
            const float c = rv[u] * rv[u] * rv[u] * dot;
            const float2 o = make_float2(rv[u] * g.x - sv[u].x * c, rv[u] * g.y - sv[u].y * c);
            dn_acc.x = fmaf(d[u].x * sv[u].x, rv[u], dn_acc.x);
            dn_acc.y = fmaf(d[u].y * sv[u].y, rv[u], dn_acc.y);
            if (m < T) reinterpret_cast<float2*>(dy_out + m * PJ_D)[lane] = o;
            *reinterpret_cast<uint32_t*>(dys + (lane >> 2) * PJ_LBO + (row >> 3) * 128 + (row & 7) * 16 + (lane & 3) * 4) = pack_bf16(o.x, o.y);
          }
        };
        load_rows(0);                                  // in flight together with the x tile (batch 0)
        store_batch(xs, dys, 0, ra[0], rb[0]);
        norm_rows(0);
        load_rows(8);
        if (it + 1 < n_it) load_batch(it + 1, 0, ra[0], rb[0]);
        norm_rows(8);
      } else {
#pragma unroll
        for (int bi = 0; bi < NBATCH; ++bi) {
          if (bi + 1 < NBATCH) load_batch(it, bi + 1, ra[(bi + 1) & 1], rb[(bi + 1) & 1]);
          else if (it + 1 < n_it) load_batch(it + 1, 0, ra[0], rb[0]);
          store_batch(xs, dys, bi, ra[bi & 1], rb[bi & 1]);
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
      float* scr = scr_all + (warp - 8) * (32 * 36);
      if (dx16) {               // bf16 pairs: 32 words per row
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
        warp_store_rows<32>(scr, v, dx + m_w * (PJ_D / 2), PJ_D / 2, rows_valid, lane);
      } else {
        warp_store_rows_add<32>(scr, v, dx + m_w * PJ_D, PJ_D, resid ? resid + m_w * PJ_D : nullptr, PJ_D, rows_valid, lane);
        warp_store_rows_add<32>(scr, v + 32, dx + m_w * PJ_D + 32, PJ_D, resid ? resid + m_w * PJ_D + 32 : nullptr, PJ_D, rows_valid, lane);
      }
      (void)m;
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
        umma_bf16(tmem + ib * 64, make_desc(dy_u + ks * 2 * PJ_LBO, PJ_LBO, 128), make_desc(w_u + ks * 2 * 1024, 1024, 128),
                  idesc_dx, ks > 0 ? 1u : 0u);
      umma_commit(smem_u32(&bars->dx_full[ib]));
      for (int b = 0; b < NBLK; ++b) {
#pragma unroll 1
        for (int ks = 0; ks < 8; ++ks)      // K = 128 tokens; MN-major views: k-group (8 tokens) stride 128 B, mn-group stride PJ_LBO
          umma_bf16(tmem + DW_COL + b * DW_STRIDE, make_desc(dy_u + b * 16 * PJ_LBO + ks * 256, 128, PJ_LBO),
                    make_desc(xs_u + ks * 256, 128, PJ_LBO), idesc_dw, (it > 0 || ks > 0) ? 1u : 0u);
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
  const int i = blockIdx.x * RED_TX + threadIdx.x, ty = threadIdx.y;
  float acc = 0.f;
  size_t off = 0;
  bool valid = false;
  if (i < NOUT * 65) { off = (size_t)(i / 65) * 80 + (i % 65); valid = true; }
  else if (i < NOUT * 65 + PJ_D && dn) { off = (size_t)NP * 80 + (i - NOUT * 65); valid = true; }
  if (valid)
    for (int k = ty; k < n_cta; k += RED_TY) acc += wsp[(size_t)k * per + off];
  acc = red_finish_32x8(acc);
  if (ty != 0 || !valid) return;
  if (i < NOUT * 65) {
    const int n = i / 65, c = i % 65;
    if (c < 64) dw[n * PJ_D + c] = acc;
    else if (db) db[n] = acc;
  } else dn[i - NOUT * 65] = acc;
}

template <int NOUT, bool POST_NORM>
static int launch_lin_fwd(const float* x, const float* w, const float* bias, const float* resid, const float* nw, float eps,
                          int64_t T, float* y, float* s_out, float* r_out, cudaStream_t st, uint32_t* y16 = nullptr, float qscale = 1.f,
                          int x16 = 0) {
  const size_t smem = (size_t)NOUT * 128 + 2 * PJ_X_BYTES + ((size_t)NOUT + PJ_D) * 4 +
                      (size_t)8 * 32 * ((POST_NORM ? 64 : 32) + 4) * 4 + sizeof(PjFwdBars) + 16;
  // measured on B200: the coalesced loader wins for the 64-wide kernel (0.103 vs 0.117 ms per launch), the row-per-lane
  // loader for the 192-wide one (0.102 vs 0.112 ms), whose epilogue already keeps the L1 data pipe busy with 96 KB of stores per tile
  auto fn = lin_fwd_kernel<NOUT, POST_NORM, POST_NORM>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(lin_fwd smem)");
  const int grid = (int)std::min<int64_t>(ceil_div(T, (int64_t)PJ_TM), (int64_t)num_sms());
  fn<<<grid, PJ_NT, smem, st>>>(x, w, bias, resid, nw, eps, T, y, s_out, r_out, y16, qscale, x16);
  return B200CTR_OK;
}

template <int NOUT, bool PRE_NORM>
static int launch_lin_bwd(const float* gin, const float* s, const float* r, const float* nw, const float* xin, const float* w,
                          const float* resid, int64_t T, float* dy_out, float* dx, float* dw, float* db, float* dn, float* wsp,
                          int gin16, cudaStream_t st, int xin16 = 0, int dx16 = 0) {
  constexpr int NP = NOUT <= 128 ? 128 : 256;
  const size_t smem = (size_t)NOUT * 128 + 2 * (size_t)(PJ_X_BYTES + PJ_ONES_BYTES + (NP / 8) * PJ_LBO) + 8 * PJ_D * 4 + (size_t)(4 * 32 * 36 * 4) +
                      sizeof(PjBwdBars) + 16;
  auto fn = lin_bwd_kernel<NOUT, PRE_NORM>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(lin_bwd smem)");
  const int grid = (int)std::min<int64_t>(ceil_div(T, (int64_t)PJ_TM), (int64_t)num_sms());
  fn<<<grid, PJ_NT, smem, st>>>(gin, s, r, nw, xin, w, resid, T, dy_out, dx, wsp, gin16, xin16, dx16);
  cudaError_t le = cudaPeekAtLastError();
  if (le != cudaSuccess) return cuda_fail(le, "lin_bwd_kernel");
  const int n_out = NOUT * 65 + PJ_D;
  lin_bwd_reduce_kernel<<<(n_out + RED_TX - 1) / RED_TX, dim3(RED_TX, RED_TY), 0, st>>>(wsp, grid, NP, NOUT, dw, db, PRE_NORM ? dn : nullptr);
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

extern "C" int b200ctr_inproj_fwd_h16(const float* x, const float* w_in, const float* b_in, int64_t T, int D, void* qkv16, void* stream) {
  B200_REQUIRE(x && w_in && qkv16, B200CTR_E_INVALID_ARG, "inproj_fwd_h16: null pointer");
  B200_REQUIRE(D == PJ_D, B200CTR_E_UNSUPPORTED_SHAPE, "inproj_fwd_h16: D=%d unsupported (64 only)", D);
  if (T == 0) return B200CTR_OK;
  // q leaves pre-multiplied by log2(e) / sqrt(head_dim): exactly the value attn_h16 would form from the fp32 tensor
  int rc = launch_lin_fwd<192, false>(x, w_in, b_in, nullptr, nullptr, 0.f, T, nullptr, nullptr, nullptr, as_stream(stream),
                                      (uint32_t*)qkv16, H2_QSCALE);
  if (rc) return rc;
  B200_LAUNCH_CHECK("lin_fwd_kernel<192,h16>");
  return B200CTR_OK;
}

extern "C" int b200ctr_outproj_norm_fwd(const void* o, int o_bf16, const float* w_out, const float* b_out, const float* x_resid,
                                        const float* n1w, float eps, int64_t T, int D, float* s1_out, float* r1_out,
                                        float* x1_out, void* stream) {
  B200_REQUIRE(o && w_out && x_resid && n1w && x1_out, B200CTR_E_INVALID_ARG, "outproj_norm_fwd: null pointer");
  B200_REQUIRE(D == PJ_D, B200CTR_E_UNSUPPORTED_SHAPE, "outproj_norm_fwd: D=%d unsupported (64 only)", D);
  if (T == 0) return B200CTR_OK;
  int rc = launch_lin_fwd<64, true>((const float*)o, w_out, b_out, x_resid, n1w, eps, T, x1_out, s1_out, r1_out, as_stream(stream), nullptr,
                                    1.f, o_bf16);
  if (rc) return rc;
  B200_LAUNCH_CHECK("lin_fwd_kernel<64,norm>");
  return B200CTR_OK;
}

extern "C" int64_t b200ctr_proj_bwd_workspace_bytes(void) { return (int64_t)num_sms() * ((int64_t)256 * 80 + PJ_D) * 4; }

extern "C" int b200ctr_outproj_norm_bwd(const float* dx1, const float* s1, const float* r1, const float* n1w, const void* o, int o_bf16,
                                        const float* w_out, int64_t T, int D, float* ds1_out, void* d_o, int do_bf16, float* dw_out,
                                        float* db_out, float* dn1, void* workspace, void* stream) {
  B200_REQUIRE(dx1 && s1 && r1 && n1w && o && w_out && ds1_out && d_o && dw_out && db_out && dn1 && workspace, B200CTR_E_INVALID_ARG,
               "outproj_norm_bwd: null pointer");
  B200_REQUIRE(D == PJ_D, B200CTR_E_UNSUPPORTED_SHAPE, "outproj_norm_bwd: D=%d unsupported (64 only)", D);
  B200_REQUIRE(T > 0, B200CTR_E_INVALID_ARG, "outproj_norm_bwd: empty input");
  int rc = launch_lin_bwd<64, true>(dx1, s1, r1, n1w, (const float*)o, w_out, nullptr, T, ds1_out, (float*)d_o, dw_out, db_out, dn1,
                                    (float*)workspace, 0, as_stream(stream), o_bf16, do_bf16);
  if (rc) return rc;
  B200_LAUNCH_CHECK("lin_bwd_kernel<64,norm>");
  return B200CTR_OK;
}

extern "C" int b200ctr_inproj_bwd(const void* dqkv, int dqkv_bf16, const float* x, const float* w_in, const float* d_resid, int64_t T,
                                  int D, float* dx, float* dw_in, float* db_in, void* workspace, void* stream) {
  B200_REQUIRE(dqkv && x && w_in && dx && dw_in && db_in && workspace, B200CTR_E_INVALID_ARG, "inproj_bwd: null pointer");
  B200_REQUIRE(D == PJ_D, B200CTR_E_UNSUPPORTED_SHAPE, "inproj_bwd: D=%d unsupported (64 only)", D);
  B200_REQUIRE(T > 0, B200CTR_E_INVALID_ARG, "inproj_bwd: empty input");
  int rc = launch_lin_bwd<192, false>((const float*)dqkv, nullptr, nullptr, nullptr, x, w_in, d_resid, T, nullptr, dx, dw_in, db_in,
                                      nullptr, (float*)workspace, dqkv_bf16, as_stream(stream));
  if (rc) return rc;
  B200_LAUNCH_CHECK("lin_bwd_kernel<192>");
  return B200CTR_OK;
}
