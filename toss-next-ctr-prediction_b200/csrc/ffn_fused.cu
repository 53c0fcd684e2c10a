// Fused encoder FFN block on tcgen05 (the `amp: bf16` path of DAREEncoderLayer's second half, dare.py:67-69):
//
//     x2 = RMSNorm_n2( x1 + W2 . dropout(GELU(W1 . x1 + b1)) + b2 )
//
// The unfused composition streams the (tokens x ffn_hidden) pre-activation and activation through HBM
// (2 x 755 MB per layer at B=8192, K=60) and is bound by that traffic at 27 FLOP/B; here a 128-token tile never
// leaves the SM: both weight matrices stay resident in shared memory as bf16 UMMA operands, the hidden dimension
// is walked in 64-wide chunks whose pre-activation lives in TMEM, the GELU/dropout epilogue writes the bf16
// activation straight into the A-operand layout of the second contraction, and the residual + RMSNorm run on the
// TMEM accumulator of that contraction with one thread per token row.
//
//   persistent CTA (one per SM), 13 warps:
//     warps 0-3   stream x1 tiles (fp32 -> bf16 K-major core matrices, double buffered), then run the final
//                 residual + RMSNorm epilogue of the previous tile (thread == token row == TMEM lane)
//     warps 4-19  GELU epilogue: TMEM chunk -> +b1 -> erf-GELU -> dropout -> bf16 -> smem A operand of GEMM2
//                 (16 warps = 4 TMEM lane groups x 4 column quarters: the epilogue is the bottleneck, see profiles/)
//     warp  20    one thread issues every tcgen05.mma; tcgen05.commit -> mbarriers hand buffers back
//
// Shared-memory operand layout is the canonical SWIZZLE_NONE one of gemm_tc.cu: 8 x 16-byte core matrices,
// k-group stride R*16 bytes ("LBO"), mn-group stride 128 bytes ("SBO").
#include "tc_common.cuh"
#include "h2_common.cuh"
#include <algorithm>

namespace b200ctr {

constexpr int FF_D = 64;          // emb_dim
constexpr int FF_TM = 128;        // tokens per tile (UMMA M)
constexpr int FF_HC = 64;         // hidden chunk (UMMA N of GEMM1, K of GEMM2)
constexpr int FFW_EPI_WARPS = 16;                 // forward kernel: 4 loader/RMSNorm + 16 GELU-epilogue + 1 MMA warp
constexpr int FFW_MMA_WARP = 4 + FFW_EPI_WARPS;
constexpr int FFW_NT = (FFW_MMA_WARP + 1) * 32;
// forward kernel: 4 tile-producer + 4 final-epilogue + 16 GELU-epilogue + 1 MMA warp
constexpr int FFF_MMA_WARP = 8 + FFW_EPI_WARPS;
constexpr int FFF_NT = (FFF_MMA_WARP + 1) * 32;
constexpr int FF_A_BYTES = FF_TM * FF_D * 2;        // 16 KB: x1 tile, bf16
constexpr int FF_WCHUNK = FF_HC * FF_D * 2;         // 8 KB of W1 or W2 per hidden chunk

struct FfnBars {
  unsigned long long a_full[2], a_empty[2], p_full[2], p_empty[2], fd_full[2], fd_empty[2], h2_full[2], h2_empty[2];
  uint32_t tmem_base;
};

// W1 [hidden][64] fp32 -> per 64-row chunk a K-major B operand (N = hidden rows, K = d)
__device__ __forceinline__ void ffn_stage_w1(unsigned char* w1s, const float* __restrict__ w1, int hidden, int tid, int nt) {
  for (int u = tid; u < hidden * 8; u += nt) {
    const int h = u >> 3, kc = u & 7;
    const float4* p = reinterpret_cast<const float4*>(w1 + (size_t)h * FF_D + kc * 8);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    const int c = h >> 6, r = h & 63;
    *reinterpret_cast<uint4*>(w1s + c * FF_WCHUNK + (kc * 8 + (r >> 3)) * 128 + (r & 7) * 16) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
}
// W2 [64][hidden] fp32 -> per 64-column chunk a K-major B operand (N = d rows, K = hidden within the chunk);
// one 128-token x 64 fp32 tile -> bf16 K-major A operand (R = 128); 128 threads, rows >= T are zero-filled
__device__ __forceinline__ void ffn_stage_tile(unsigned char* dst, const float* __restrict__ x, int64_t m0, int64_t T, int ptid) {
  float4 va[8], vb[8];
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int u = ptid + it * 128;
    const int r = u & 127, c = u >> 7;
    const int64_t m = m0 + r;
    va[it] = make_float4(0.f, 0.f, 0.f, 0.f); vb[it] = va[it];
    if (m < T) {
      const float4* p = reinterpret_cast<const float4*>(x + m * FF_D + c * 8);
      va[it] = __ldg(p); vb[it] = __ldg(p + 1);
    }
  }
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int u = ptid + it * 128;
    const int r = u & 127, c = u >> 7;
    *reinterpret_cast<uint4*>(dst + (c * 16 + (r >> 3)) * 128 + (r & 7) * 16) =
        make_uint4(pack_bf16(va[it].x, va[it].y), pack_bf16(va[it].z, va[it].w), pack_bf16(vb[it].x, vb[it].y), pack_bf16(vb[it].z, vb[it].w));
  }
}

// Normal CDF for the bf16 path: Phi(x) ~= sigmoid(x * (a + b u + c u^2)), u = min(x^2, 64)  (minimax fit of the three
// coefficients: |error| < 2.1e-5 in Phi and < 4.2e-5 in x * Phi(x), two orders below bf16 resolution; 9 instructions
// with 2 MUFU instead of the 18 of the erf formula in common.cuh, which stays the fp32 path's GELU).  The
// coefficients carry the -log2(e) of the exponential.  x -> -inf gives rcp(inf) = 0, x -> +inf gives rcp(1) = 1.
__device__ __forceinline__ float ffn_cdf(float x, float& u) {
  u = fminf(x * x, 64.f);
  float q = fmaf(u, 9.4134924e-4f, -0.10628247f);           // -log2e * (c u + b)
  q = fmaf(u, q, -2.3016721f);                               // -log2e * a
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * q));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}
__device__ __forceinline__ float ffn_gelu(float x) {
  float u;
  return x * ffn_cdf(x, u);
}
// cdf = Phi(x); returns GELU'(x) = Phi(x) + x * phi(x), phi = exp(-x^2/2) / sqrt(2 pi)
__device__ __forceinline__ float ffn_gelu_grad(float x, float& cdf) {
  float u, e;
  cdf = ffn_cdf(x, u);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fmaf(u, -0.72134752f, -1.3257480f)));   // phi(x), log2(1/sqrt(2 pi)) folded
  return fmaf(x, e, cdf);
}

// D[tmem] (+)= A[smem, K-major, R_A rows] . B[smem, K-major, R_B rows]^T over `ksteps` K=16 steps
__device__ __forceinline__ void ffn_mma_kk(uint32_t tmem_d, uint32_t a_base, uint32_t lbo_a, uint32_t b_base, uint32_t lbo_b,
                                           uint32_t idesc, int ksteps, bool accumulate) {
#pragma unroll 1
  for (int ks = 0; ks < ksteps; ++ks) {
    const uint64_t da = make_desc(a_base + ks * 2 * lbo_a, lbo_a, 128);
    const uint64_t db = make_desc(b_base + ks * 2 * lbo_b, lbo_b, 128);
    umma_bf16(tmem_d, da, db, idesc, (accumulate || ks > 0) ? 1u : 0u);
  }
}

// HC = hidden chunk width.  Every chunk costs one trip around the barrier ring (epilogue reads P -> issuer -> GEMM1 -> epilogue;
// epilogue writes fd -> issuer -> GEMM2 -> fd buffer free), ~1.3 us whatever the chunk holds (measured by ablation: with the
// GELU removed the kernel went 0.274 -> 0.262 ms, with the output stores removed too -> 0.21 ms).  128-wide chunks halve the
// number of trips; HC = 64 remains for ffn_hidden not divisible by 128.
template <int HC>
__global__ void __launch_bounds__(FFF_NT, 1)
ffn_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ w1, const float* __restrict__ b1,
               const float* __restrict__ w2, const float* __restrict__ b2, const float* __restrict__ n2w, float eps,
               int64_t T, int hidden, float* __restrict__ s2_out, float* __restrict__ r2_out, float* __restrict__ x2_out,
               uint16_t* __restrict__ keep, b200ctr_dropout_t drop) {
  // keep (optional): one 16-bit word of dropout keep flags per (tile, 64-wide chunk, 16-column quarter, token row) --
  // so that the backward tests bits instead of re-hashing
  constexpr int WCH = HC * FF_D * 2;            // bytes of W1 or W2 per chunk
  constexpr int FDB = FF_TM * HC * 2;           // bytes of one activation chunk buffer
  constexpr int CW = HC / 4;                    // columns per epilogue warp and chunk (16 or 32)
  extern __shared__ __align__(128) unsigned char smem[];
  const int NC = hidden / HC;
  unsigned char* w1s = smem;
  unsigned char* w2s = w1s + (size_t)NC * WCH;
  unsigned char* as = w2s + (size_t)NC * WCH;                  // [2][16 KB]
  unsigned char* fds = as + 2 * FF_A_BYTES;                    // [2][FDB]
  float* b1s = reinterpret_cast<float*>(fds + 2 * FDB);
  float* b2s = b1s + hidden;
  float* n2s = b2s + FF_D;
  float* scr_all = n2s + FF_D;                                 // 4 final-epilogue warps x 32 x 36 floats (store transpose, 32 columns at a time)
  FfnBars* bars = reinterpret_cast<FfnBars*>(scr_all + 4 * 32 * 36);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (T + FF_TM - 1) / FF_TM;
  const int n_it = (int)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);      // tiles of this CTA

  if (warp == FFF_MMA_WARP) tmem_alloc(&bars->tmem_base, 256);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(smem_u32(&bars->a_full[s]), 4); mbar_init(smem_u32(&bars->a_empty[s]), 1);
      mbar_init(smem_u32(&bars->fd_full[s]), FFW_EPI_WARPS); mbar_init(smem_u32(&bars->fd_empty[s]), 1);
      mbar_init(smem_u32(&bars->h2_full[s]), 1); mbar_init(smem_u32(&bars->h2_empty[s]), 4);
    }
    mbar_init(smem_u32(&bars->p_full[0]), 1); mbar_init(smem_u32(&bars->p_empty[0]), FFW_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // W1 [hidden][64] -> per chunk a K-major B operand (N = HC hidden rows, K = d): core (k-group kc, n-group r/8)
  for (int u = tid; u < hidden * 8; u += FFF_NT) {
    const int h = u >> 3, kc = u & 7;
    const float4* p = reinterpret_cast<const float4*>(w1 + (size_t)h * FF_D + kc * 8);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    const int c = h / HC, r = h % HC;
    *reinterpret_cast<uint4*>(w1s + c * WCH + (kc * (HC / 8) + (r >> 3)) * 128 + (r & 7) * 16) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
  // W2 [64][hidden] -> per chunk a K-major B operand (N = d rows, K = HC hidden within the chunk), fp16, times 1/(1-p)
  {
    const float scale = drop.p > 0.f ? 1.f / (1.f - drop.p) : 1.f;
    const int hg_per_row = hidden >> 3;
    for (int u = tid; u < FF_D * hg_per_row; u += FFF_NT) {
      const int d = u / hg_per_row, hg = u % hg_per_row;
      const float4* p = reinterpret_cast<const float4*>(w2 + (size_t)d * hidden + hg * 8);
      const float4 a = __ldg(p), b = __ldg(p + 1);
      const int c = hg / (HC / 8), kc = hg % (HC / 8);
      *reinterpret_cast<uint4*>(w2s + c * WCH + (kc * 8 + (d >> 3)) * 128 + (d & 7) * 16) =
          make_uint4(h2_pack(a.x * scale, a.y * scale), h2_pack(a.z * scale, a.w * scale), h2_pack(b.x * scale, b.y * scale),
                     h2_pack(b.z * scale, b.w * scale));
    }
  }
  uint32_t* b1h = reinterpret_cast<uint32_t*>(b1s);                // bias as half2 pairs
  for (int i = tid; i < hidden / 2; i += FFF_NT) b1h[i] = h2_pack(b1[2 * i], b1[2 * i + 1]);
  if (tid < FF_D) { b2s[tid] = b2[tid]; n2s[tid] = n2w[tid]; }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  // TMEM: the pre-activation chunk at columns 0 .. HC-1 (one buffer: two measured slower, profiles/README.md); the x1 . W1^T
  // product of chunk g + 1 is issued the moment the epilogue has READ chunk g; h2 buffers at columns 128 and 192
  const uint32_t P_COL = 0, H2_COL = 128;

  if (warp < 4) {
    // ---------------------------------------- tile producer --------------------------------------------------------
    // (its own four warps: with one set of warps doing "stage tile it + 1, then the final epilogue of tile it" the two
    // phases' global-load latencies added up on the kernel's critical path -- removing the output stores alone moved the
    // kernel by 15 %)
    for (int it = 0; it < n_it; ++it) {
      const int nb = it & 1;
      if (it >= 2) mbar_wait(smem_u32(&bars->a_empty[nb]), (uint32_t)(((it >> 1) - 1) & 1));
      ffn_stage_tile(as + nb * FF_A_BYTES, x1, ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * FF_TM, T, tid);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->a_full[nb]));
    }
  } else if (warp < 8) {
    // ---------------------------------------- final epilogue: residual + RMSNorm on the h2 accumulator --------------
    const int fw = warp - 4;                                      // TMEM lane group = warp % 4
    for (int it = 0; it < n_it; ++it) {
      const int hb = it & 1;
      mbar_wait(smem_u32(&bars->h2_full[hb]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      uint32_t v[64];
      const uint32_t taddr = tmem + H2_COL + hb * 64 + ((uint32_t)(fw * 32) << 16);
      tmem_ld32(taddr, v);
      tmem_ld32(taddr + 32, v + 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->h2_empty[hb]));
      const int64_t m_w = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * FF_TM + fw * 32;
      const int64_t m = m_w + lane;
      const int rows_valid = (int)max((int64_t)0, min((int64_t)32, T - m_w));
      float ss = 0.f;
      if (m < T) {
        const float4* xr = reinterpret_cast<const float4*>(x1 + m * FF_D);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 xv = __ldg(xr + j);
          float s0 = __uint_as_float(v[4 * j]) + b2s[4 * j] + xv.x;
          float s1 = __uint_as_float(v[4 * j + 1]) + b2s[4 * j + 1] + xv.y;
          float s2 = __uint_as_float(v[4 * j + 2]) + b2s[4 * j + 2] + xv.z;
          float s3 = __uint_as_float(v[4 * j + 3]) + b2s[4 * j + 3] + xv.w;
          ss = fmaf(s0, s0, ss); ss = fmaf(s1, s1, ss); ss = fmaf(s2, s2, ss); ss = fmaf(s3, s3, ss);
          v[4 * j] = __float_as_uint(s0); v[4 * j + 1] = __float_as_uint(s1);
          v[4 * j + 2] = __float_as_uint(s2); v[4 * j + 3] = __float_as_uint(s3);
        }
      }
      const float r = rsqrtf(ss * (1.f / FF_D) + eps);
      if (r2_out && m < T) r2_out[m] = r;
      float* scr = scr_all + fw * (32 * 36);
      if (s2_out) {
        warp_store_rows<32>(scr, v, s2_out + m_w * FF_D, FF_D, rows_valid, lane);
        warp_store_rows<32>(scr, v + 32, s2_out + m_w * FF_D + 32, FF_D, rows_valid, lane);
      }
#pragma unroll
      for (int j = 0; j < 64; ++j) v[j] = __float_as_uint(n2s[j] * __uint_as_float(v[j]) * r);
      warp_store_rows<32>(scr, v, x2_out + m_w * FF_D, FF_D, rows_valid, lane);
      warp_store_rows<32>(scr, v + 32, x2_out + m_w * FF_D + 32, FF_D, rows_valid, lane);
    }
  } else if (warp < FFF_MMA_WARP) {
    // ---------------------------------------- GELU epilogue -----------------------------------------------------
    const int ew = warp - 8, q = ew & 3, part = ew >> 2;          // 16 warps: TMEM lane group x CW-column quarter of the chunk
    const int r = q * 32 + lane;
    uint32_t g = 0;
    const DropCtx dctx = drop_ctx(drop);
    const uint32_t thr2 = dctx.thr | (dctx.thr << 16);
    for (int it = 0; it < n_it; ++it) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      const int64_t m = tile * FF_TM + r;
      const uint32_t rkey = dctx.on ? drop_row_key(drop, (uint64_t)m) : 0u;
      for (int c = 0; c < NC; ++c, ++g) {
        mbar_wait(smem_u32(&bars->p_full[0]), g & 1);
        tc_fence_after();
        uint32_t v[CW];
        const uint32_t taddr = tmem + P_COL + part * CW + ((uint32_t)(q * 32) << 16);
        if (CW == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->p_empty[0]));
        // per-element work on half2 pairs: pack, + b1, GELU (h2_common.cuh), keep mask = one AND + one HSET2 on the hash
        // word; the activation goes to GEMM2 as fp16 (11 bits, against the 8 of bf16), masked but not rescaled (W2 carries
        // the 1/(1-p)).  10 instructions per element instead of 27.
        uint32_t pk[CW / 2];
#pragma unroll
        for (int sq = 0; sq < CW / 16; ++sq) {
          const int n0 = c * HC + part * CW + 16 * sq;      // first hidden column of this 16-column quarter
          const uint4 bq0 = *reinterpret_cast<const uint4*>(b1h + (n0 >> 1)), bq1 = *reinterpret_cast<const uint4*>(b1h + (n0 >> 1) + 4);
          const uint32_t bh[8] = {bq0.x, bq0.y, bq0.z, bq0.w, bq1.x, bq1.y, bq1.z, bq1.w};
          uint32_t kb = 0u;
#pragma unroll
          for (int jp = 0; jp < 8; ++jp) {
            uint32_t f = h2_gelu(h2_add(h2_pack(__uint_as_float(v[16 * sq + 2 * jp]), __uint_as_float(v[16 * sq + 2 * jp + 1])), bh[jp]));
            if (dctx.on) {
              const uint32_t mk = drop_keep_mask2(drop_pair_hash(rkey, (uint32_t)(n0 >> 1) + jp), thr2);
              f &= mk;
              kb |= mk & (0x00010001u << (2 * jp));      // bit 2jp <- even column, bit 16 + 2jp <- odd column
            }
            pk[8 * sq + jp] = f;
          }
          kb = (kb & 0x5555u) | ((kb >> 15) & 0xAAAAu);      // stored word: bit j = keep flag of column j
          if (keep && dctx.on) keep[((tile * (hidden / 64) + (n0 >> 6)) * 4 + ((n0 & 63) >> 4)) * FF_TM + r] = (uint16_t)kb;
        }
        const int fb = g & 1;
        if (g >= 2) mbar_wait(smem_u32(&bars->fd_empty[fb]), ((g >> 1) - 1) & 1);
        unsigned char* dst = fds + fb * FDB;
#pragma unroll
        for (int j4 = 0; j4 < CW / 8; ++j4) {
          const int kg = part * (CW / 8) + j4;
          *reinterpret_cast<uint4*>(dst + (kg * 16 + (r >> 3)) * 128 + (r & 7) * 16) =
              make_uint4(pk[4 * j4], pk[4 * j4 + 1], pk[4 * j4 + 2], pk[4 * j4 + 3]);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->fd_full[fb]));
      }
    }
  } else if (lane == 0) {
    // ---------------------------------------- MMA issuer ---------------------------------------------------------
    const uint32_t idesc = umma_idesc(FF_TM, HC, false, false);          // M=128, N=HC, both operands K-major (bf16: x1 . W1^T)
    const uint32_t idesc2 = umma_idesc_f16(FF_TM, FF_D, false, false);   // fp16: activation . W2^T, N = 64
    const uint32_t as_u = smem_u32(as), fd_u = smem_u32(fds), w1_u = smem_u32(w1s), w2_u = smem_u32(w2s);
    const uint32_t total = (uint32_t)n_it * (uint32_t)NC;
    // descriptors are built once and advanced by an add; chunk / tile counters are carried (no divisions)
    const uint64_t dA[2] = {make_desc(as_u, FF_TM * 16, 128), make_desc(as_u + FF_A_BYTES, FF_TM * 16, 128)};
    const uint64_t dF[2] = {make_desc(fd_u, FF_TM * 16, 128), make_desc(fd_u + FDB, FF_TM * 16, 128)};
    const uint64_t dW1 = make_desc(w1_u, HC * 16, 128), dW2 = make_desc(w2_u, FF_D * 16, 128);
    uint32_t g1 = 0, it1 = 0, c1 = 0, g2 = 0, it2 = 0, c2 = 0;
    auto gemm1 = [&]() {                             // next chunk -> the P buffer (free once the epilogue has read the previous chunk)
      const uint32_t ab = it1 & 1;
      if (c1 == 0) mbar_wait(smem_u32(&bars->a_full[ab]), (it1 >> 1) & 1);
      if (g1 >= 1) mbar_wait(smem_u32(&bars->p_empty[0]), (g1 - 1) & 1);
      tc_fence_after();
      const uint64_t da = dA[ab], db = desc_advance(dW1, c1 * WCH);
      const uint32_t td = tmem + P_COL;
      umma_bf16_c<false>(td, da, db, idesc);
#pragma unroll
      for (int ks = 1; ks < FF_D / 16; ++ks)
        umma_bf16_c<true>(td, desc_advance(da, ks * 2 * FF_TM * 16), desc_advance(db, ks * 2 * HC * 16), idesc);
      umma_commit(smem_u32(&bars->p_full[0]));
      ++g1;
      if (++c1 == (uint32_t)NC) { umma_commit(smem_u32(&bars->a_empty[ab])); c1 = 0; ++it1; }
    };
    auto gemm2 = [&]() {
      const uint32_t fb = g2 & 1, hb = it2 & 1;
      mbar_wait(smem_u32(&bars->fd_full[fb]), (g2 >> 1) & 1);
      if (c2 == 0 && it2 >= 2) mbar_wait(smem_u32(&bars->h2_empty[hb]), ((it2 >> 1) - 1) & 1);
      tc_fence_after();
      const uint64_t da = dF[fb], db = desc_advance(dW2, c2 * WCH);
      const uint32_t td = tmem + H2_COL + hb * 64;
      if (c2 == 0) umma_bf16_c<false>(td, da, db, idesc2); else umma_bf16_c<true>(td, da, db, idesc2);
#pragma unroll
      for (int ks = 1; ks < HC / 16; ++ks)
        umma_bf16_c<true>(td, desc_advance(da, ks * 2 * FF_TM * 16), desc_advance(db, ks * 2 * FF_D * 16), idesc2);
      umma_commit(smem_u32(&bars->fd_empty[fb]));
      ++g2;
      if (++c2 == (uint32_t)NC) { umma_commit(smem_u32(&bars->h2_full[hb])); c2 = 0; ++it2; }
    };
    if (g1 < total) gemm1();
    for (uint32_t g = 0; g < total; ++g) {
      if (g1 < total) gemm1();
      gemm2();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FFF_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem, 256);
  }
}


// =============================================================================================================
// Backward, kernel 1 of 2 (per 128-token tile; weight gradients accumulate in TMEM across the persistent loop):
//   d_s2  = RMSNorm'(dy; s2, r2, n2)                                  (loader warps, fp32, also -> HBM for kernel 2)
//   P     = x1 . W1_c^T (+b1)   Q = d_s2 . W2[:, c]                   (two 128x64x64 MMAs per 64-wide hidden chunk)
//   fd    = dropout(GELU(P)) ;  dpre = Q * mask * GELU'(P)            (epilogue warps; dropout mask regenerated)
//   dW2^T[blk] += fd^T . d_s2 ;  dW1[blk] += dpre^T . x1              (M = 128 hidden rows, K = 128 tokens; the bf16
//        activation / gradient buffers written K-major by the epilogue are re-read MN-major: same bytes,
//        LBO/SBO exchanged in the descriptor, no transpose)
//   dpre  -> HBM as the bf16 shared-memory image (kernel 2 = dx1 = dpre . W1 + d_s2 bulk-copies it back)
//   db1 (transpose-reduce over the 32 rows of a warp), db2, dn2 (column-owner lanes of the loader warps)
// TMEM: P 64 | Q 64 | dW1 hidden/2 | dW2^T hidden/2 columns = 512 at ffn_hidden = 384.
// =============================================================================================================
struct FfnBwdBars {
  unsigned long long in_full[2], in_empty[2], pq_full, pq_empty, blk_full, blk_empty, done;
  uint32_t tmem_base;
};
constexpr int FF_IN_BYTES = 2 * FF_A_BYTES;            // x1 tile + d_s2 tile (bf16)
constexpr int FF_BLK_BYTES = FF_TM * 128 * 2;          // 32 KB: 128 tokens x 128 hidden, bf16

// W2 [64][hidden] fp32 -> per 64-column chunk an MN-major B operand (N = hidden within the chunk, K = d)
__device__ __forceinline__ void ffn_stage_w2_mn(unsigned char* w2s, const float* __restrict__ w2, int hidden, float scale, int tid, int nt) {
  const int hg_per_row = hidden >> 3;
  for (int u = tid; u < FF_D * hg_per_row; u += nt) {
    const int d = u / hg_per_row, hg = u % hg_per_row;
    const float4* p = reinterpret_cast<const float4*>(w2 + (size_t)d * hidden + hg * 8);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    const int c = hg >> 3, ng = hg & 7;
    *reinterpret_cast<uint4*>(w2s + c * FF_WCHUNK + ((d >> 3) * 8 + ng) * 128 + (d & 7) * 16) =
        make_uint4(pack_bf16(a.x * scale, a.y * scale), pack_bf16(a.z * scale, a.w * scale), pack_bf16(b.x * scale, b.y * scale),
                   pack_bf16(b.z * scale, b.w * scale));
  }
}

// lane j <- sum over the warp's 32 lanes of v[j]  (31 shuffles; v is destroyed)
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

// 16 columns: fold the two half-warps, then the same halving network; lanes j and j + 16 <- column j
__device__ __forceinline__ float warp_colsum16(float (&v)[16], int lane) {
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], 16);
#pragma unroll
  for (int off = 8; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

template <bool KEEP>
__global__ void __launch_bounds__(FFW_NT, 1)
ffn_bwd1_kernel(const float* __restrict__ dy, const float* __restrict__ s2, const float* __restrict__ r2,
                const float* __restrict__ n2w, const float* __restrict__ x1, const float* __restrict__ w1,
                const float* __restrict__ b1, const float* __restrict__ w2, int64_t T, int hidden,
                float* __restrict__ ds2_out, unsigned char* __restrict__ dpre_spill, float* __restrict__ ws,
                const uint16_t* __restrict__ keep, b200ctr_dropout_t drop) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int NC = hidden / FF_HC, NB = hidden / 128;
  unsigned char* w1s = smem;
  unsigned char* w2s = w1s + (size_t)NC * FF_WCHUNK;
  unsigned char* ins = w2s + (size_t)NC * FF_WCHUNK;           // [2][x1 16 KB | d_s2 16 KB]
  unsigned char* fdb = ins + 2 * FF_IN_BYTES;                  // 32 KB activation block
  unsigned char* dpb = fdb + FF_BLK_BYTES;                     // 32 KB dpre block
  float* b1s = reinterpret_cast<float*>(dpb + FF_BLK_BYTES);
  FfnBwdBars* bars = reinterpret_cast<FfnBwdBars*>(b1s + hidden);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (T + FF_TM - 1) / FF_TM;
  const int n_it = (int)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
  // per-CTA partial sums: dW1 [hidden][64] | dW2^T [hidden][64] | db1 [4][hidden] | db2 [4][64] | dn2 [4][64]
  const size_t ws_per_cta = (size_t)2 * hidden * FF_D + (size_t)4 * hidden + 8 * FF_D;
  float* wsc = ws + (size_t)blockIdx.x * ws_per_cta;
  float* ws_db1 = wsc + (size_t)2 * hidden * FF_D;
  float* ws_db2 = ws_db1 + (size_t)4 * hidden;
  float* ws_dn2 = ws_db2 + 4 * FF_D;

  if (warp == FFW_MMA_WARP) tmem_alloc(&bars->tmem_base, 512);
  if (tid == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bars->in_full[s]), 4); mbar_init(smem_u32(&bars->in_empty[s]), 1); }
    mbar_init(smem_u32(&bars->pq_full), 1); mbar_init(smem_u32(&bars->pq_empty), FFW_EPI_WARPS);
    mbar_init(smem_u32(&bars->blk_full), FFW_EPI_WARPS); mbar_init(smem_u32(&bars->blk_empty), 1);
    mbar_init(smem_u32(&bars->done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  ffn_stage_w1(w1s, w1, hidden, tid, FFW_NT);
  ffn_stage_w2_mn(w2s, w2, hidden, drop.p > 0.f ? 1.f / (1.f - drop.p) : 1.f, tid, FFW_NT);      // Q arrives times 1/(1-p)
  uint32_t* b1h = reinterpret_cast<uint32_t*>(b1s);                // bias as half2 pairs
  for (int i = tid; i < hidden / 2; i += FFW_NT) b1h[i] = h2_pack(b1[2 * i], b1[2 * i + 1]);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  const uint32_t P_COL = 0, Q_COL = 64, DW1_COL = 128, DW2_COL = 128 + (uint32_t)NB * 64;

  if (warp < 4) {
    // ------------------------------- loader: x1 tile + d_s2 = RMSNorm'(dy) -----------------------------------
    const float2 nw = reinterpret_cast<const float2*>(n2w)[lane];
    float2 dn_acc = make_float2(0.f, 0.f), db_acc = make_float2(0.f, 0.f);
    for (int it = 0; it < n_it; ++it) {
      const int ib = it & 1;
      if (it >= 2) mbar_wait(smem_u32(&bars->in_empty[ib]), (uint32_t)(((it >> 1) - 1) & 1));
      const int64_t m0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * FF_TM;
      unsigned char* xs = ins + ib * FF_IN_BYTES;
      unsigned char* dss = xs + FF_A_BYTES;
      ffn_stage_tile(xs, x1, m0, T, tid);
#pragma unroll 1
      for (int rr = 0; rr < 32; rr += 8) {      // 8 rows (24 loads) in flight per lane: the loader is latency-bound
        float2 d[8], s[8];
        float r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int64_t m = m0 + warp * 32 + rr + u;
          d[u] = make_float2(0.f, 0.f); s[u] = d[u]; r[u] = 0.f;
          if (m < T) {
            d[u] = __ldg(reinterpret_cast<const float2*>(dy + m * FF_D) + lane);
            s[u] = __ldg(reinterpret_cast<const float2*>(s2 + m * FF_D) + lane);
            r[u] = __ldg(r2 + m);
          }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int row = warp * 32 + rr + u;
          const int64_t m = m0 + row;
          const float2 g = make_float2(d[u].x * nw.x, d[u].y * nw.y);
          const float dot = warp_sum(g.x * s[u].x + g.y * s[u].y) * (1.f / FF_D);
          const float c = r[u] * r[u] * r[u] * dot;
          const float2 o = make_float2(r[u] * g.x - s[u].x * c, r[u] * g.y - s[u].y * c);
          dn_acc.x = fmaf(d[u].x * s[u].x, r[u], dn_acc.x);
          dn_acc.y = fmaf(d[u].y * s[u].y, r[u], dn_acc.y);
          db_acc.x += o.x; db_acc.y += o.y;
          if (m < T) reinterpret_cast<float2*>(ds2_out + m * FF_D)[lane] = o;
          *reinterpret_cast<uint32_t*>(dss + ((lane >> 2) * 16 + (row >> 3)) * 128 + (row & 7) * 16 + (lane & 3) * 4) = pack_bf16(o.x, o.y);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->in_full[ib]));
    }
    reinterpret_cast<float2*>(ws_db2 + warp * FF_D)[lane] = db_acc;
    reinterpret_cast<float2*>(ws_dn2 + warp * FF_D)[lane] = dn_acc;
  } else if (warp < FFW_MMA_WARP) {
    // ------------------------------- epilogue: fd, dpre, db1 ---------------------------------------------------
    const int ew = warp - 4, q = ew & 3, part = ew >> 2;          // 16 warps: TMEM lane group x 16-column quarter of the chunk
    const int r = q * 32 + lane;
    float db1_acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    uint32_t g = 0;
    const DropCtx dctx = drop_ctx(drop);
    const uint32_t thr2 = dctx.thr | (dctx.thr << 16);
    for (int it = 0; it < n_it; ++it) {
      const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      const int64_t m = tile * FF_TM + r;
      const uint32_t rkey = (!KEEP && dctx.on) ? drop_row_key(drop, (uint64_t)m) : 0u;
      uint32_t kb_next = 0u;
      if (KEEP) kb_next = keep[((tile * NC) * 4 + part) * FF_TM + r];
      for (int c = 0; c < NC; ++c, ++g) {
        // keep flags of this chunk were requested one chunk ago (the load right before the wait was NOT hidden: 7 % of the
        // kernel's stall samples sat on its first use)
        const uint32_t kb16 = kb_next;
        if (KEEP && c + 1 < NC) kb_next = keep[((tile * NC + c + 1) * 4 + part) * FF_TM + r];
        mbar_wait(smem_u32(&bars->pq_full), g & 1);
        tc_fence_after();
        uint32_t pv[16], qv[16];
        const uint32_t lane_bits = (uint32_t)(q * 32) << 16;
        tmem_ld16_pair(tmem + P_COL + part * 16 + lane_bits, pv, tmem + Q_COL + part * 16 + lane_bits, qv);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&bars->pq_empty));
        const int n0 = c * FF_HC + part * 16;
        // GELU, GELU' and the mask on half2 pairs (h2_common.cuh; the forward's arithmetic), the products with Q -- gradient
        // sized, outside fp16's range -- in fp32:  f = x Phi(x) & keep ;  dpre = Q (Phi(x) + x phi(x)) & keep
        // (Q carries the 1/(1-p) through W2; f gets it below).  21 instructions per element instead of 36.
        const uint4 bq0 = *reinterpret_cast<const uint4*>(b1h + (n0 >> 1)), bq1 = *reinterpret_cast<const uint4*>(b1h + (n0 >> 1) + 4);
        const uint32_t bh[8] = {bq0.x, bq0.y, bq0.z, bq0.w, bq1.x, bq1.y, bq1.z, bq1.w};
        const uint32_t kb32 = (kb16 & 0x5555u) | ((kb16 & 0xAAAAu) << 15);      // pair jp: bits 2jp and 16 + 2jp
        const uint32_t inv2 = h2_const_rt(dctx.inv);
        uint32_t fpk[8], dpk[8];
        float dp[16];
#pragma unroll
        for (int jp = 0; jp < 8; ++jp) {
          const uint32_t x = h2_add(h2_pack(__uint_as_float(pv[2 * jp]), __uint_as_float(pv[2 * jp + 1])), bh[jp]);
          const uint32_t u = h2_min(h2_mul(x, x), h2_const(64.f));
          uint32_t qq = h2_fma(u, h2_const(-3.2625e-4f), h2_const(0.036835f));
          qq = h2_fma(u, qq, h2_const(0.797687f));
          const uint32_t cdf = h2_fma(h2_tanh(h2_mul(x, qq)), h2_const(0.5f), h2_const(0.5f));
          const uint32_t phi = h2_ex2(h2_fma(u, h2_const(-0.72134752f), h2_const(-1.3257480f)));      // exp(-x^2/2) / sqrt(2 pi)
          uint32_t gg = h2_fma(x, phi, cdf), f = h2_mul(x, cdf);
          if (KEEP) {
            const uint32_t mk = h2_mask_from_signs(kb32 << (15 - 2 * jp));
            gg &= mk;
            f = h2_mul(f & mk, inv2);
          } else if (dctx.on) {
            const uint32_t mk = drop_keep_mask2(drop_pair_hash(rkey, (uint32_t)(n0 >> 1) + jp), thr2);
            gg &= mk;
            f = h2_mul(f & mk, inv2);
          }
          dp[2 * jp] = __uint_as_float(qv[2 * jp]) * h2_lo(gg);
          dp[2 * jp + 1] = __uint_as_float(qv[2 * jp + 1]) * h2_hi(gg);
          fpk[jp] = pack_bf16(h2_lo(f), h2_hi(f));
          dpk[jp] = pack_bf16(dp[2 * jp], dp[2 * jp + 1]);
        }
        const uint32_t bk = (uint32_t)it * NB + (uint32_t)(c >> 1);
        if ((c & 1) == 0 && bk >= 1) mbar_wait(smem_u32(&bars->blk_empty), (bk - 1) & 1);
        unsigned char* spill = dpre_spill + ((size_t)tile * NB + (c >> 1)) * FF_BLK_BYTES;
#pragma unroll
        for (int j4 = 0; j4 < 2; ++j4) {
          const int kg = (c & 1) * 8 + part * 2 + j4;
          const uint32_t off = (uint32_t)((kg * 16 + (r >> 3)) * 128 + (r & 7) * 16);
          const uint4 dv = make_uint4(dpk[4 * j4], dpk[4 * j4 + 1], dpk[4 * j4 + 2], dpk[4 * j4 + 3]);
          *reinterpret_cast<uint4*>(fdb + off) = make_uint4(fpk[4 * j4], fpk[4 * j4 + 1], fpk[4 * j4 + 2], fpk[4 * j4 + 3]);
          *reinterpret_cast<uint4*>(dpb + off) = dv;
          *reinterpret_cast<uint4*>(spill + off) = dv;
        }
        if (c & 1) {
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->blk_full));
        }
        const float cs = warp_colsum16(dp, lane);          // lane j (and j + 16) <- column j summed over the warp's 32 rows
#pragma unroll
        for (int k = 0; k < 6; ++k) db1_acc[k] += (k == c) ? cs : 0.f;
      }
    }
    // weight-gradient accumulators: TMEM -> this CTA's partial-sum slab
    mbar_wait(smem_u32(&bars->done), 0);
    tc_fence_after();
    for (int which = 0; which < 2; ++which) {
      for (int b = 0; b < NB; ++b) {
        uint32_t v[16];
        tmem_ld16(tmem + (which ? DW2_COL : DW1_COL) + b * 64 + part * 16 + ((uint32_t)(q * 32) << 16), v);
        float4* dst = reinterpret_cast<float4*>(wsc + ((size_t)which * hidden + b * 128 + r) * FF_D + part * 16);
#pragma unroll
        for (int j = 0; j < 4; ++j)
          dst[j] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
      }
    }
    if (lane < 16) {
      for (int c = 0; c < NC; ++c) {
        float vsel = 0.f;
#pragma unroll
        for (int k = 0; k < 6; ++k) vsel = (k == c) ? db1_acc[k] : vsel;
        ws_db1[(size_t)q * hidden + c * FF_HC + part * 16 + lane] = vsel;
      }
    }
  } else if (lane == 0) {
    // ------------------------------- MMA issuer ---------------------------------------------------------------------
    const uint32_t idesc_p = umma_idesc(FF_TM, FF_HC, false, false);     // x1 (K-major) . W1 chunk (K-major)
    const uint32_t idesc_q = umma_idesc(FF_TM, FF_HC, false, true);      // d_s2 (K-major) . W2 chunk (MN-major)
    const uint32_t idesc_w = umma_idesc(128, FF_D, true, true);          // block^T (MN view) . tile (MN view)
    const uint32_t in_u = smem_u32(ins), w1_u = smem_u32(w1s), w2_u = smem_u32(w2s), fd_u = smem_u32(fdb), dp_u = smem_u32(dpb);
    const uint32_t total = (uint32_t)n_it * (uint32_t)NC;
    auto wgrad = [&](uint32_t bk) {
      const uint32_t it2 = bk / NB, b2 = bk % NB, ib2 = it2 & 1;
      mbar_wait(smem_u32(&bars->blk_full), bk & 1);
      tc_fence_after();
      const uint32_t xs_u = in_u + ib2 * FF_IN_BYTES, ds_u = xs_u + FF_A_BYTES;
#pragma unroll 1
      for (int ks = 0; ks < 8; ++ks) {        // K = 128 tokens; MN-major views: k-group stride 128 B, mn-group stride 2048 B
        const uint32_t acc = (it2 > 0 || ks > 0) ? 1u : 0u;
        umma_bf16(tmem + DW2_COL + b2 * 64, make_desc(fd_u + ks * 256, 128, 2048), make_desc(ds_u + ks * 256, 128, 2048), idesc_w, acc);
        umma_bf16(tmem + DW1_COL + b2 * 64, make_desc(dp_u + ks * 256, 128, 2048), make_desc(xs_u + ks * 256, 128, 2048), idesc_w, acc);
      }
      umma_commit(smem_u32(&bars->blk_empty));
      if (b2 == (uint32_t)NB - 1) umma_commit(smem_u32(&bars->in_empty[ib2]));
    };
    for (uint32_t g = 0; g < total; ++g) {
      const uint32_t it = g / NC, c = g % NC, ib = it & 1;
      if (c == 0) mbar_wait(smem_u32(&bars->in_full[ib]), (it >> 1) & 1);
      if (g > 0) mbar_wait(smem_u32(&bars->pq_empty), (g - 1) & 1);
      tc_fence_after();
      const uint32_t xs_u = in_u + ib * FF_IN_BYTES, ds_u = xs_u + FF_A_BYTES;
      ffn_mma_kk(tmem + P_COL, xs_u, FF_TM * 16, w1_u + c * FF_WCHUNK, FF_HC * 16, idesc_p, FF_D / 16, false);
      ffn_mma_kk(tmem + Q_COL, ds_u, FF_TM * 16, w2_u + c * FF_WCHUNK, FF_HC * 16, idesc_q, FF_D / 16, false);
      umma_commit(smem_u32(&bars->pq_full));
      if (g > 0 && ((g - 1) % NC) % 2 == 1) wgrad((g - 1) / 2);
    }
    if (total > 0) wgrad(total / 2 - 1);
    umma_commit(smem_u32(&bars->done));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FFW_MMA_WARP) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

// deterministic cross-CTA reduction of kernel 1's partial sums (and the dW2^T -> dW2 transpose).  The two weight-gradient
// slabs (2 * hidden * 64 floats per CTA) are summed four outputs per thread with 16-byte loads: a warp reads 512 contiguous
// bytes of one CTA's slab per load instead of 128 (the slabs are 200 KB apart, every load opens another DRAM page; ncu r02an:
// 0.033 ms for 29 MB).  Same summation order per output as the scalar form: partials ty, ty + 8, .. then the 8 slices.
__global__ void ffn_bwd_reduce_kernel(const float* __restrict__ ws, int n_cta, int hidden, float* __restrict__ dw1,
                                      float* __restrict__ db1, float* __restrict__ dw2, float* __restrict__ db2,
                                      float* __restrict__ dn2) {
  const size_t per = (size_t)2 * hidden * FF_D + (size_t)4 * hidden + 8 * FF_D;
  const int HD = hidden * FF_D;
  const int n4 = 2 * HD / 4, n4blocks = (n4 + RED_TX - 1) / RED_TX;
  const int ty = threadIdx.y;
  if ((int)blockIdx.x < n4blocks) {
    __shared__ float4 red4[RED_TY][RED_TX];
    const int i4 = blockIdx.x * RED_TX + threadIdx.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i4 < n4)
      for (int c = ty; c < n_cta; c += RED_TY) {
        const float4 v = *reinterpret_cast<const float4*>(ws + (size_t)c * per + (size_t)4 * i4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    red4[ty][threadIdx.x] = acc;
    __syncthreads();
    if (ty != 0 || i4 >= n4) return;
#pragma unroll
    for (int y = 1; y < RED_TY; ++y) {
      const float4 v = red4[y][threadIdx.x];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    const int i = 4 * i4;
    if (i < HD) {
      dw1[i] = acc.x; dw1[i + 1] = acc.y; dw1[i + 2] = acc.z; dw1[i + 3] = acc.w;   // a view into the flat gradient buffer: 4-byte aligned only
    } else {
      const int j = i - HD, h = j / FF_D, d = j % FF_D;          // the four outputs share h (FF_D is a multiple of 4)
      dw2[(size_t)d * hidden + h] = acc.x; dw2[(size_t)(d + 1) * hidden + h] = acc.y;
      dw2[(size_t)(d + 2) * hidden + h] = acc.z; dw2[(size_t)(d + 3) * hidden + h] = acc.w;
    }
    return;
  }
  const int i = 2 * HD + ((int)blockIdx.x - n4blocks) * RED_TX + threadIdx.x;
  float acc = 0.f;
  if (i < 2 * HD + hidden) {
    const int h = i - 2 * HD;
    for (int p = ty; p < 4 * n_cta; p += RED_TY) acc += ws[(size_t)(p >> 2) * per + 2 * HD + (size_t)(p & 3) * hidden + h];
  } else if (i < 2 * HD + hidden + 2 * FF_D) {
    const int j = i - 2 * HD - hidden, which = j / FF_D, d = j % FF_D;
    for (int p = ty; p < 4 * n_cta; p += RED_TY)
      acc += ws[(size_t)(p >> 2) * per + 2 * HD + (size_t)4 * hidden + which * 4 * FF_D + (p & 3) * FF_D + d];
  }
  acc = red_finish_32x8(acc);
  if (ty != 0) return;
  if (i < 2 * HD + hidden) db1[i - 2 * HD] = acc;
  else if (i < 2 * HD + hidden + 2 * FF_D) { const int j = i - 2 * HD - hidden; (j / FF_D ? dn2 : db2)[j % FF_D] = acc; }
}

// =============================================================================================================
// Backward, kernel 2 of 2:  dx1 = dpre . W1 + d_s2.   dpre tiles come back from HBM with one bulk (TMA) copy per
// 32 KB block, already in UMMA operand order.  6 warps: 0 = bulk-copy producer, 1 = MMA, 2-5 = epilogue.
// =============================================================================================================
constexpr int FF2_STAGES = 4;
constexpr int FF2_NT = 6 * 32;
struct Ffn2Bars {
  unsigned long long full[FF2_STAGES], empty[FF2_STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(FF2_NT, 1)
ffn_bwd2_kernel(const unsigned char* __restrict__ dpre_spill, const float* __restrict__ w1, const float* __restrict__ ds2,
                int64_t T, int hidden, float* __restrict__ dx1) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int NB = hidden / 128;
  unsigned char* w1s = smem;                                   // NB x 16 KB: MN-major B operand (N = d, K = hidden)
  unsigned char* ring = w1s + (size_t)NB * 16384;              // FF2_STAGES x 32 KB
  float* scr_all = reinterpret_cast<float*>(ring + (size_t)FF2_STAGES * FF_BLK_BYTES);       // 4 epilogue warps x 32 x 68 floats
  Ffn2Bars* bars = reinterpret_cast<Ffn2Bars*>(scr_all + 4 * 32 * 68);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t tiles = (T + FF_TM - 1) / FF_TM;
  const int n_it = (int)((tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (warp == 1) tmem_alloc(&bars->tmem_base, 128);
  if (tid == 0) {
    for (int s = 0; s < FF2_STAGES; ++s) { mbar_init(smem_u32(&bars->full[s]), 1); mbar_init(smem_u32(&bars->empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bars->acc_full[s]), 1); mbar_init(smem_u32(&bars->acc_empty[s]), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // W1 [hidden][64]: unit (hidden row h, d-group ng) -> core (k-group h/8, n-group ng), row h%8, 8 consecutive d
  for (int u = tid; u < hidden * 8; u += FF2_NT) {
    const int h = u >> 3, ng = u & 7;
    const float4* p = reinterpret_cast<const float4*>(w1 + (size_t)h * FF_D + ng * 8);
    const float4 a = __ldg(p), b = __ldg(p + 1);
    const int blk = h >> 7, hh = h & 127;
    *reinterpret_cast<uint4*>(w1s + blk * 16384 + ((hh >> 3) * 8 + ng) * 128 + (hh & 7) * 16) =
        make_uint4(pack_bf16(a.x, a.y), pack_bf16(a.z, a.w), pack_bf16(b.x, b.y), pack_bf16(b.z, b.w));
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int it = 0; it < n_it; ++it) {
        const int64_t tile = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
        for (int b = 0; b < NB; ++b, ++g) {
          const uint32_t s = g % FF2_STAGES, use = g / FF2_STAGES;
          if (use > 0) mbar_wait(smem_u32(&bars->empty[s]), (use - 1) & 1);
          const uint32_t bar = smem_u32(&bars->full[s]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)FF_BLK_BYTES) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(ring + (size_t)s * FF_BLK_BYTES)), "l"(dpre_spill + ((size_t)tile * NB + b) * FF_BLK_BYTES),
                         "r"((uint32_t)FF_BLK_BYTES), "r"(bar) : "memory");
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(FF_TM, FF_D, false, true);
      const uint32_t ring_u = smem_u32(ring), w1_u = smem_u32(w1s);
      uint32_t g = 0;
      for (int it = 0; it < n_it; ++it) {
        const uint32_t ab = it & 1;
        if (it >= 2) mbar_wait(smem_u32(&bars->acc_empty[ab]), ((it >> 1) - 1) & 1);
        for (int b = 0; b < NB; ++b, ++g) {
          const uint32_t s = g % FF2_STAGES;
          mbar_wait(smem_u32(&bars->full[s]), (g / FF2_STAGES) & 1);
          tc_fence_after();
#pragma unroll 1
          for (int ks = 0; ks < 8; ++ks)
            umma_bf16(tmem + ab * 64, make_desc(ring_u + s * FF_BLK_BYTES + ks * 2 * 2048, 2048, 128),
                      make_desc(w1_u + b * 16384 + ks * 2 * 1024, 1024, 128), idesc, (b > 0 || ks > 0) ? 1u : 0u);
          umma_commit(smem_u32(&bars->empty[s]));
        }
        umma_commit(smem_u32(&bars->acc_full[ab]));
      }
    }
  } else {
    const int q = warp & 3;
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ab = it & 1;
      mbar_wait(smem_u32(&bars->acc_full[ab]), (it >> 1) & 1);
      tc_fence_after();
      uint32_t v[64];
      const uint32_t taddr = tmem + ab * 64 + ((uint32_t)(q * 32) << 16);
      tmem_ld32(taddr, v);
      tmem_ld32(taddr + 32, v + 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[ab]));
      const int64_t m_w = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * FF_TM + q * 32;
      const int64_t m = m_w + lane;
      const int rows_valid = (int)max((int64_t)0, min((int64_t)32, T - m_w));
      if (m < T) {
        const float4* rp = reinterpret_cast<const float4*>(ds2 + m * FF_D);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 a = __ldg(rp + j);
          v[4 * j] = __float_as_uint(__uint_as_float(v[4 * j]) + a.x); v[4 * j + 1] = __float_as_uint(__uint_as_float(v[4 * j + 1]) + a.y);
          v[4 * j + 2] = __float_as_uint(__uint_as_float(v[4 * j + 2]) + a.z); v[4 * j + 3] = __float_as_uint(__uint_as_float(v[4 * j + 3]) + a.w);
        }
      }
      warp_store_rows<64>(scr_all + (warp - 2) * (32 * 68), v, dx1 + m_w * FF_D, FF_D, rows_valid, lane);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem, 128);
  }
}

static size_t ffn_fwd_smem(int hidden, int hc) {
  return (size_t)2 * hidden * FF_D * 2 + 2 * FF_A_BYTES + (size_t)2 * FF_TM * hc * 2 + ((size_t)hidden + 2 * FF_D) * 4 +
         (size_t)4 * 32 * 36 * 4 + sizeof(FfnBars) + 16;
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_ffn_fwd(const float* x1, const float* w1, const float* b1, const float* w2, const float* b2,
                               const float* n2w, float eps, int64_t T, int D, int hidden, float* s2_out, float* r2_out,
                               float* x2_out, void* keep_bits, b200ctr_dropout_t drop, void* stream) {
  B200_REQUIRE(x1 && w1 && b1 && w2 && b2 && n2w && x2_out, B200CTR_E_INVALID_ARG, "ffn_fwd: null pointer");
  B200_REQUIRE(D == FF_D, B200CTR_E_UNSUPPORTED_SHAPE, "ffn_fwd: D=%d unsupported (64 only)", D);
  B200_REQUIRE(hidden > 0 && hidden % FF_HC == 0 && hidden <= 512, B200CTR_E_UNSUPPORTED_SHAPE,
               "ffn_fwd: ffn_hidden=%d unsupported (multiple of 64, <= 512)", hidden);
  if (T == 0) return B200CTR_OK;
  const int hc = (hidden % 128 == 0 && hidden <= 384) ? 128 : 64;      // 128-wide chunks need 213 KB of shared memory at 384
  const size_t smem = ffn_fwd_smem(hidden, hc);
  auto fn = hc == 128 ? ffn_fwd_kernel<128> : ffn_fwd_kernel<64>;
  cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(ffn_fwd smem)");
  const int64_t tiles = ceil_div(T, (int64_t)FF_TM);
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)num_sms());
  fn<<<grid, FFF_NT, smem, as_stream(stream)>>>(x1, w1, b1, w2, b2, n2w, eps, T, hidden, s2_out, r2_out, x2_out,
                                                (uint16_t*)keep_bits, drop);
  B200_LAUNCH_CHECK("ffn_fwd_kernel");
  return B200CTR_OK;
}

extern "C" int64_t b200ctr_ffn_keep_bytes(int64_t T, int hidden) {
  if (T <= 0 || hidden <= 0) return 0;
  return ceil_div(T, (int64_t)FF_TM) * FF_TM * (hidden / 16) * 2;
}

extern "C" int64_t b200ctr_ffn_bwd_workspace_bytes(int64_t T, int hidden) {
  if (T <= 0 || hidden <= 0) return 0;
  const int64_t tiles = ceil_div(T, (int64_t)FF_TM);
  const int64_t spill = tiles * (hidden / 128) * (int64_t)FF_BLK_BYTES;
  const int64_t per_cta = ((int64_t)2 * hidden * FF_D + (int64_t)4 * hidden + 8 * FF_D) * 4;
  return spill + per_cta * num_sms() + T * FF_D * 4;
}

extern "C" int b200ctr_ffn_bwd(const float* dy, const float* s2, const float* r2, const float* n2w, const float* x1,
                               const float* w1, const float* b1, const float* w2, int64_t T, int D, int hidden,
                               float* dx1, float* dw1, float* db1, float* dw2, float* db2, float* dn2,
                               void* workspace, const void* keep_bits, b200ctr_dropout_t drop, void* stream) {
  B200_REQUIRE(dy && s2 && r2 && n2w && x1 && w1 && b1 && w2 && dx1 && dw1 && db1 && dw2 && db2 && dn2 && workspace,
               B200CTR_E_INVALID_ARG, "ffn_bwd: null pointer");
  B200_REQUIRE(D == FF_D, B200CTR_E_UNSUPPORTED_SHAPE, "ffn_bwd: D=%d unsupported (64 only)", D);
  B200_REQUIRE(hidden > 0 && hidden % 128 == 0 && hidden <= 384, B200CTR_E_UNSUPPORTED_SHAPE,
               "ffn_bwd: ffn_hidden=%d unsupported (multiple of 128, <= 384)", hidden);
  B200_REQUIRE(T > 0, B200CTR_E_INVALID_ARG, "ffn_bwd: empty input");
  cudaStream_t st = as_stream(stream);
  const int64_t tiles = ceil_div(T, (int64_t)FF_TM);
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)num_sms());
  const int NB = hidden / 128;
  unsigned char* spill = (unsigned char*)workspace;
  float* ws = (float*)(spill + (size_t)tiles * NB * FF_BLK_BYTES);
  const size_t per_cta = (size_t)2 * hidden * FF_D + (size_t)4 * hidden + 8 * FF_D;
  float* ds2 = ws + per_cta * (size_t)num_sms();
  const size_t smem1 = (size_t)2 * (hidden / FF_HC) * FF_WCHUNK + 2 * FF_IN_BYTES + 2 * FF_BLK_BYTES + (size_t)hidden * 4 + sizeof(FfnBwdBars) + 16;
  const bool use_keep = keep_bits != nullptr && drop.p > 0.f;
  auto fn1 = use_keep ? ffn_bwd1_kernel<true> : ffn_bwd1_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(fn1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(ffn_bwd1 smem)");
  fn1<<<grid, FFW_NT, smem1, st>>>(dy, s2, r2, n2w, x1, w1, b1, w2, T, hidden, ds2, spill, ws,
                                   use_keep ? (const uint16_t*)keep_bits : nullptr, drop);
  B200_LAUNCH_CHECK("ffn_bwd1_kernel");
  const int red_blocks = (2 * hidden * FF_D / 4 + RED_TX - 1) / RED_TX + (hidden + 2 * FF_D + RED_TX - 1) / RED_TX;
  ffn_bwd_reduce_kernel<<<red_blocks, dim3(RED_TX, RED_TY), 0, st>>>(ws, grid, hidden, dw1, db1, dw2, db2, dn2);
  B200_LAUNCH_CHECK("ffn_bwd_reduce_kernel");
  const size_t smem2 = (size_t)NB * 16384 + (size_t)FF2_STAGES * FF_BLK_BYTES + (size_t)4 * 32 * 68 * 4 + sizeof(Ffn2Bars) + 16;
  e = cudaFuncSetAttribute(ffn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(ffn_bwd2 smem)");
  ffn_bwd2_kernel<<<grid, FF2_NT, smem2, st>>>(spill, w1, ds2, T, hidden, dx1);
  B200_LAUNCH_CHECK("ffn_bwd2_kernel");
  return B200CTR_OK;
}
