// Fused low-rank quadratic interaction of QNN-alpha on tcgen05 (qnn_alpha.py:86-97, `amp: bf16` path):
//
//   A = z_b . Ucat            (F x C per sample; F = 200 features, C = heads*rank = 96)
//   s = sum_f A ,  quad = s^2 - sum_f A^2                                   (forward; A is NEVER written to HBM)
//   dA = dquad * 2 (s - A) ,  dz = dA . Ucat^T (+ the MLP branch's dz) ,  dUcat += z^T . dA            (backward)
//
// The unfused composition materialises A (B*F x C fp32 = 629 MB at B = 8192) and streams it three more times
// (fm_reduce, fm_grad, the dz / dUcat contractions).  Here a persistent CTA handles one sample per iteration: the
// sample's z rows (bf16, emitted by the row-RMSNorm kernel) are copied into UMMA operand layout, A lives in TMEM,
// the column statistics are reduced across the 128 lanes with a shuffle transpose-reduce, and the backward recomputes A,
// writes dA as a bf16 operand tile, and accumulates dUcat^T in TMEM across the persistent loop (the K-major dA / z
// tiles are re-read MN-major for that product, as in ffn_fused.cu).
//   13 warps: 0-3 loaders, 4-11 epilogue (warp = (128-row block, TMEM lane group)), 12 = MMA issuer.
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace b200ctr {

constexpr int FM_D = 64;
constexpr int FM_NT = 13 * 32;
// K-direction stride of the 8-column core-matrix groups of a z block: 128 rows x 16 bytes + 16 bytes of padding, so that
// the 8 chunks of one row (written by 8 neighbouring lanes of the coalesced loader) fall into different banks
constexpr int FM_ZLBO = 128 * 16 + 16;
constexpr int FM_ZBLK = 8 * FM_ZLBO;           // one 128-row block of z, bf16 K-major
constexpr int FM_ZTILE = 2 * FM_ZBLK;          // up to 256 feature rows per sample
constexpr int FM_CP = 128;                     // column stride of the A accumulators / dA tile

// lane j <- sum over the warp's 32 lanes of v[j]  (31 shuffles; v is destroyed)
__device__ __forceinline__ float fm_colsum32(float (&v)[32], int lane) {
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

// z16 rows of sample b (F x 64 bf16, contiguous) -> K-major core matrices, 128-row blocks; rows >= F stay zero (static).
// The sample is a flat list of F * 8 16-byte chunks read in passes of 128 consecutive chunks (512 contiguous bytes per
// warp instruction = 4 L1 wavefronts; one row per lane cost 32).  All loads of a thread are issued before the first store.
__device__ __forceinline__ void fm_stage_z(unsigned char* dst, const __nv_bfloat16* __restrict__ z16, int64_t b, int F, int ptid) {
  const uint4* src = reinterpret_cast<const uint4*>(z16 + b * (int64_t)F * FM_D);
  const int nq = F * 8;
  uint4 v[16];
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const int q = p * 128 + ptid;
    v[p] = q < nq ? __ldg(src + q) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const int q = p * 128 + ptid, row = q >> 3, c = q & 7;
    if (q < nq)
      *reinterpret_cast<uint4*>(dst + (row >> 7) * FM_ZBLK + c * FM_ZLBO + ((row & 127) >> 3) * 128 + (row & 7) * 16) = v[p];
  }
}

// Ucat (64 x C fp32, row-major [d][c]) -> MN-major B operand (N = c, K = d): core (k-group d/8, n-group c/8), row d%8
__device__ __forceinline__ void fm_stage_u_mn(unsigned char* dst, const float* __restrict__ U, int C, int tid, int nt) {
  const int ng_n = C / 8;
  for (int u = tid; u < FM_D * ng_n; u += nt) {
    const int d = u / ng_n, ng = u % ng_n;
    const float* p = U + (size_t)d * C + ng * 8;
    *reinterpret_cast<uint4*>(dst + ((d >> 3) * ng_n + ng) * 128 + (d & 7) * 16) =
        make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
  }
}
// Ucat -> K-major B operand of dz = dA . Ucat^T (N = d, 64 rows; K = c): core (n-group d/8, k-group c/8) at (kc*8 + g)*128
__device__ __forceinline__ void fm_stage_u_k(unsigned char* dst, const float* __restrict__ U, int C, int tid, int nt) {
  const int kc_n = C / 8;
  for (int u = tid; u < FM_D * kc_n; u += nt) {
    const int d = u / kc_n, kc = u % kc_n;
    const float* p = U + (size_t)d * C + kc * 8;
    *reinterpret_cast<uint4*>(dst + (kc * 8 + (d >> 3)) * 128 + (d & 7) * 16) =
        make_uint4(pack_bf16(p[0], p[1]), pack_bf16(p[2], p[3]), pack_bf16(p[4], p[5]), pack_bf16(p[6], p[7]));
  }
}

struct FmFwdBars {
  unsigned long long z_full[2], z_empty[2], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(FM_NT, 1)
fm_fwd_kernel(const __nv_bfloat16* __restrict__ z16, const float* __restrict__ U, int64_t B, int F, int C,
              float* __restrict__ s_out, float* __restrict__ quad_out) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* us = smem;                                      // MN-major Ucat: 64 x C bf16
  unsigned char* zs = us + FM_D * FM_CP * 2;                     // [2][32 KB]
  float* part = reinterpret_cast<float*>(zs + 2 * FM_ZTILE);     // [2 parity][2 stats][8 warps][C]
  FmFwdBars* bars = reinterpret_cast<FmFwdBars*>(part + 2 * 2 * 8 * FM_CP);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = (F + 127) >> 7;
  const int n_it = (int)((B - blockIdx.x + gridDim.x - 1) / gridDim.x);

  if (warp == 12) tmem_alloc(&bars->tmem_base, 512);
  if (tid == 0) {
    for (int k = 0; k < 2; ++k) {
      mbar_init(smem_u32(&bars->z_full[k]), 4); mbar_init(smem_u32(&bars->z_empty[k]), 1);
      mbar_init(smem_u32(&bars->acc_full[k]), 1); mbar_init(smem_u32(&bars->acc_empty[k]), 8);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fm_stage_u_mn(us, U, C, tid, FM_NT);
  for (int i = tid; i < 2 * FM_ZTILE / 16; i += FM_NT) reinterpret_cast<uint4*>(zs)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 4) {
    for (int it = 0; it < n_it; ++it) {
      const int ib = it & 1;
      if (it >= 2) mbar_wait(smem_u32(&bars->z_empty[ib]), (uint32_t)(((it >> 1) - 1) & 1));
      fm_stage_z(zs + ib * FM_ZTILE, z16, (int64_t)blockIdx.x + (int64_t)it * gridDim.x, F, tid);
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->z_full[ib]));
    }
  } else if (warp < 12) {
    const int ew = warp - 4, blk = ew >> 2, q = ew & 3, et = tid - 128;          // et: 0..255 within the epilogue group
    for (int it = 0; it < n_it; ++it) {
      const int ib = it & 1;
      const int64_t b = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      float* ps = part + (size_t)ib * 2 * 8 * FM_CP;
      mbar_wait(smem_u32(&bars->acc_full[ib]), (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      if (blk < nblk) {
        for (int cc = 0; cc * 32 < C; ++cc) {
          uint32_t v[32];
          tmem_ld32(tmem + ib * 256 + blk * FM_CP + cc * 32 + ((uint32_t)(q * 32) << 16), v);
          float a[32], a2[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) { a[j] = __uint_as_float(v[j]); a2[j] = a[j] * a[j]; }
          const float cs = fm_colsum32(a, lane), cq = fm_colsum32(a2, lane);
          ps[ew * FM_CP + cc * 32 + lane] = cs;
          ps[8 * FM_CP + ew * FM_CP + cc * 32 + lane] = cq;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[ib]));
      asm volatile("bar.sync 1, 256;" ::: "memory");                 // the 8 epilogue warps
      if (et < C) {
        float sv = 0.f, qv = 0.f;
        for (int w8 = 0; w8 < 8; ++w8)
          if ((w8 >> 2) < nblk) { sv += ps[w8 * FM_CP + et]; qv += ps[8 * FM_CP + w8 * FM_CP + et]; }
        s_out[b * C + et] = sv;
        quad_out[b * C + et] = sv * sv - qv;
      }
      // the partial-sum arrays alternate with the sample parity, so the next sample's writes cannot race these reads
    }
  } else if (lane == 0) {
    const uint32_t idesc = umma_idesc(128, C, false, true);
    const uint32_t zs_u = smem_u32(zs), us_u = smem_u32(us);
    const uint32_t lbo_u = (uint32_t)(C / 8) * 128;
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ib = it & 1;
      mbar_wait(smem_u32(&bars->z_full[ib]), (uint32_t)((it >> 1) & 1));
      if (it >= 2) mbar_wait(smem_u32(&bars->acc_empty[ib]), (uint32_t)(((it >> 1) - 1) & 1));
      tc_fence_after();
      for (int blk = 0; blk < nblk; ++blk)
#pragma unroll 1
        for (int ks = 0; ks < FM_D / 16; ++ks)
          umma_bf16(tmem + ib * 256 + blk * FM_CP, make_desc(zs_u + ib * FM_ZTILE + blk * FM_ZBLK + ks * 2 * FM_ZLBO, FM_ZLBO, 128),
                    make_desc(us_u + ks * 2 * lbo_u, lbo_u, 128), idesc, ks > 0 ? 1u : 0u);
      umma_commit(smem_u32(&bars->acc_full[ib]));
      umma_commit(smem_u32(&bars->z_empty[ib]));
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

struct FmBwdBars {
  unsigned long long in_full[2], in_empty[2], a_full, a_empty, da_full, da_empty, dz_full, dz_empty, done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(FM_NT, 1)
fm_bwd_kernel(const __nv_bfloat16* __restrict__ z16, const float* __restrict__ U, const float* __restrict__ s_in,
              const float* __restrict__ dquad, const float* __restrict__ dz_resid, int64_t B, int F, int C,
              float* __restrict__ dz, float* __restrict__ wsp) {
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* us_mn = smem;                                   // Ucat MN-major (A = z . Ucat)
  unsigned char* us_k = us_mn + FM_D * FM_CP * 2;                // Ucat K-major  (dz = dA . Ucat^T)
  unsigned char* zs = us_k + FM_D * FM_CP * 2;                   // [2][32 KB]
  unsigned char* das = zs + 2 * FM_ZTILE;                        // dA tile: 2 blocks x [128 rows][128 cols] bf16 = 64 KB
  float* vecs = reinterpret_cast<float*>(das + 2 * 128 * FM_CP * 2);     // [2 parity][2][C]: s, 2*dquad
  float* scr_all = vecs + 2 * 2 * FM_CP;                                  // 8 epilogue warps x 32 x 36 floats (store transpose)
  FmBwdBars* bars = reinterpret_cast<FmBwdBars*>(scr_all + 8 * 32 * 36);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nblk = (F + 127) >> 7;
  const int n_it = (int)((B - blockIdx.x + gridDim.x - 1) / gridDim.x);
  const uint32_t A_COL = 0, DZ_COL = 256, DU_COL = 384;

  if (warp == 12) tmem_alloc(&bars->tmem_base, 512);
  if (tid == 0) {
    for (int k = 0; k < 2; ++k) { mbar_init(smem_u32(&bars->in_full[k]), 4); mbar_init(smem_u32(&bars->in_empty[k]), 1); }
    mbar_init(smem_u32(&bars->a_full), 1); mbar_init(smem_u32(&bars->a_empty), 8);
    mbar_init(smem_u32(&bars->da_full), 8); mbar_init(smem_u32(&bars->da_empty), 1);
    mbar_init(smem_u32(&bars->dz_full), 1); mbar_init(smem_u32(&bars->dz_empty), 8);
    mbar_init(smem_u32(&bars->done), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fm_stage_u_mn(us_mn, U, C, tid, FM_NT);
  fm_stage_u_k(us_k, U, C, tid, FM_NT);
  for (int i = tid; i < (2 * FM_ZTILE + 2 * 128 * FM_CP * 2) / 16; i += FM_NT) reinterpret_cast<uint4*>(zs)[i] = make_uint4(0u, 0u, 0u, 0u);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp < 4) {
    for (int it = 0; it < n_it; ++it) {
      const int ib = it & 1;
      const int64_t b = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      if (it >= 2) mbar_wait(smem_u32(&bars->in_empty[ib]), (uint32_t)(((it >> 1) - 1) & 1));
      fm_stage_z(zs + ib * FM_ZTILE, z16, b, F, tid);
      if (tid < C) {
        vecs[ib * 2 * FM_CP + tid] = s_in[b * C + tid];
        vecs[ib * 2 * FM_CP + FM_CP + tid] = 2.f * dquad[b * C + tid];
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->in_full[ib]));
    }
  } else if (warp < 12) {
    const int ew = warp - 4, blk = ew >> 2, q = ew & 3;
    const int r = q * 32 + lane;                                  // row within the block
    for (int it = 0; it < n_it; ++it) {
      const int ib = it & 1;
      const int64_t b = (int64_t)blockIdx.x + (int64_t)it * gridDim.x;
      const float* sv = vecs + ib * 2 * FM_CP;
      const float* dq2 = sv + FM_CP;
      const bool row_ok = blk * 128 + r < F;
      // ---- phase 1: dA = 2 dquad (s - A) -> bf16 operand tile ----
      mbar_wait(smem_u32(&bars->a_full), (uint32_t)(it & 1));
      tc_fence_after();
      if (it >= 1) mbar_wait(smem_u32(&bars->da_empty), (uint32_t)((it - 1) & 1));
      for (int cc = 0; cc * 32 < C; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem + A_COL + blk * FM_CP + cc * 32 + ((uint32_t)(q * 32) << 16), v);
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 32; j += 2) {
          const int c0 = cc * 32 + j;
          const float d0 = row_ok ? dq2[c0] * (sv[c0] - __uint_as_float(v[j])) : 0.f;
          const float d1 = row_ok ? dq2[c0 + 1] * (sv[c0 + 1] - __uint_as_float(v[j + 1])) : 0.f;
          pk[j >> 1] = pack_bf16(d0, d1);
        }
        if (blk < nblk) {
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const int kg = cc * 4 + j4;
            *reinterpret_cast<uint4*>(das + blk * (128 * FM_CP * 2) + (kg * 16 + (r >> 3)) * 128 + (r & 7) * 16) =
                make_uint4(pk[4 * j4], pk[4 * j4 + 1], pk[4 * j4 + 2], pk[4 * j4 + 3]);
          }
        }
      }
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) { mbar_arrive(smem_u32(&bars->a_empty)); mbar_arrive(smem_u32(&bars->da_full)); }
      // ---- phase 2: dz rows = dA . Ucat^T (+ the other branch's gradient) ----
      mbar_wait(smem_u32(&bars->dz_full), (uint32_t)(it & 1));
      tc_fence_after();
      uint32_t w[64];
      const uint32_t taddr = tmem + DZ_COL + blk * 64 + ((uint32_t)(q * 32) << 16);
      tmem_ld32(taddr, w);
      tmem_ld32(taddr + 32, w + 32);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&bars->dz_empty));
      const int64_t row_w = b * F + blk * 128 + q * 32;                 // first row of this warp
      const int rows_valid = max(0, min(32, F - (blk * 128 + q * 32)));
      // the other branch's gradient is added in the coalesced domain of the store transpose (whole-row loads; a row-per-lane
      // 16-byte load costs one L1 wavefront per lane).  In-place use (dz == dz_resid) is safe: a warp reads and writes only
      // its own 32 rows, and each half's residual is loaded before that half is stored.
      float* scr = scr_all + ew * (32 * 36);
      const float* rres = dz_resid ? dz_resid + row_w * FM_D : nullptr;
      warp_store_rows_add<32>(scr, w, dz + row_w * FM_D, FM_D, rres, FM_D, rows_valid, lane);
      warp_store_rows_add<32>(scr, w + 32, dz + row_w * FM_D + 32, FM_D, rres ? rres + 32 : nullptr, FM_D, rows_valid, lane);
      (void)row_ok;
    }
    // ---- dUcat^T accumulator (rows = c, 64 columns = d) -> this CTA's partial slab ----
    mbar_wait(smem_u32(&bars->done), 0);
    tc_fence_after();
    if (blk == 0) {
      uint32_t w[64];
      const uint32_t taddr = tmem + DU_COL + ((uint32_t)(q * 32) << 16);
      tmem_ld32(taddr, w);
      tmem_ld32(taddr + 32, w + 32);
      if (r < C) {
        float4* dst = reinterpret_cast<float4*>(wsp + ((size_t)blockIdx.x * FM_CP + r) * FM_D);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          dst[j] = make_float4(__uint_as_float(w[4 * j]), __uint_as_float(w[4 * j + 1]), __uint_as_float(w[4 * j + 2]), __uint_as_float(w[4 * j + 3]));
      }
    }
  } else if (lane == 0) {
    const uint32_t idesc_a = umma_idesc(128, C, false, true);        // z (K-major) . Ucat (MN-major)
    const uint32_t idesc_dz = umma_idesc(128, FM_D, false, false);   // dA (K-major) . Ucat^T (K-major B)
    const uint32_t idesc_du = umma_idesc(128, FM_D, true, true);     // dA^T (MN view) . z (MN view)
    const uint32_t zs_u = smem_u32(zs), umn_u = smem_u32(us_mn), uk_u = smem_u32(us_k), da_u = smem_u32(das);
    const uint32_t lbo_u = (uint32_t)(C / 8) * 128;
    auto gemm_a = [&](int it) {
      const uint32_t ib = it & 1;
      mbar_wait(smem_u32(&bars->in_full[ib]), (uint32_t)((it >> 1) & 1));
      if (it >= 1) mbar_wait(smem_u32(&bars->a_empty), (uint32_t)((it - 1) & 1));
      tc_fence_after();
      for (int blk = 0; blk < nblk; ++blk)
#pragma unroll 1
        for (int ks = 0; ks < FM_D / 16; ++ks)
          umma_bf16(tmem + A_COL + blk * FM_CP, make_desc(zs_u + ib * FM_ZTILE + blk * FM_ZBLK + ks * 2 * FM_ZLBO, FM_ZLBO, 128),
                    make_desc(umn_u + ks * 2 * lbo_u, lbo_u, 128), idesc_a, ks > 0 ? 1u : 0u);
      umma_commit(smem_u32(&bars->a_full));
    };
    if (n_it > 0) gemm_a(0);
    for (int it = 0; it < n_it; ++it) {
      const uint32_t ib = it & 1;
      mbar_wait(smem_u32(&bars->da_full), (uint32_t)(it & 1));
      if (it >= 1) mbar_wait(smem_u32(&bars->dz_empty), (uint32_t)((it - 1) & 1));
      tc_fence_after();
      for (int blk = 0; blk < nblk; ++blk)
#pragma unroll 1
        for (int ks = 0; ks < C / 16; ++ks)
          umma_bf16(tmem + DZ_COL + blk * 64, make_desc(da_u + blk * (128 * FM_CP * 2) + ks * 2 * 2048, 2048, 128),
                    make_desc(uk_u + ks * 2 * 1024, 1024, 128), idesc_dz, ks > 0 ? 1u : 0u);
      umma_commit(smem_u32(&bars->dz_full));
      for (int blk = 0; blk < nblk; ++blk)
#pragma unroll 1
        for (int ks = 0; ks < 8; ++ks)            // K = 128 feature rows per block; MN views: k-group stride 128 B, mn-group 2048 B
          umma_bf16(tmem + DU_COL, make_desc(da_u + blk * (128 * FM_CP * 2) + ks * 256, 128, 2048),
                    make_desc(zs_u + ib * FM_ZTILE + blk * FM_ZBLK + ks * 256, 128, FM_ZLBO), idesc_du, (it > 0 || blk > 0 || ks > 0) ? 1u : 0u);
      umma_commit(smem_u32(&bars->da_empty));
      umma_commit(smem_u32(&bars->in_empty[ib]));
      if (it + 1 < n_it) gemm_a(it + 1);
    }
    umma_commit(smem_u32(&bars->done));
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

// dUcat[d][c] = sum over CTAs of partial[cta][c][d]
__global__ void fm_bwd_reduce_kernel(const float* __restrict__ wsp, int n_cta, int C, float* __restrict__ dU) {
  const int i = blockIdx.x * RED_TX + threadIdx.x, ty = threadIdx.y;       // i = c * 64 + d: coalesced partial reads
  const bool valid = i < FM_D * C;
  const int c = i / FM_D, d = i % FM_D;
  float acc = 0.f;
  if (valid)
    for (int k = ty; k < n_cta; k += RED_TY) acc += wsp[((size_t)k * FM_CP + c) * FM_D + d];
  acc = red_finish_32x8(acc);
  if (ty == 0 && valid) dU[d * C + c] = acc;
}

static bool fm_shape_ok(int F, int C, int D) { return D == FM_D && F >= 1 && F <= 256 && C >= 16 && C <= 128 && C % 32 == 0; }

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_fm_fused_fwd(const void* z_bf16, const float* Ucat, int64_t B, int F, int C, int D, float* s_out,
                                    float* quad_out, void* stream) {
  B200_REQUIRE(z_bf16 && Ucat && s_out && quad_out, B200CTR_E_INVALID_ARG, "fm_fused_fwd: null pointer");
  B200_REQUIRE(fm_shape_ok(F, C, D), B200CTR_E_UNSUPPORTED_SHAPE, "fm_fused_fwd: F=%d C=%d D=%d unsupported", F, C, D);
  if (B == 0) return B200CTR_OK;
  const size_t smem = (size_t)FM_D * FM_CP * 2 + 2 * FM_ZTILE + (size_t)2 * 2 * 8 * FM_CP * 4 + sizeof(FmFwdBars) + 16;
  cudaError_t e = cudaFuncSetAttribute(fm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fm_fwd smem)");
  const int grid = (int)std::min<int64_t>(B, (int64_t)num_sms());
  fm_fwd_kernel<<<grid, FM_NT, smem, as_stream(stream)>>>((const __nv_bfloat16*)z_bf16, Ucat, B, F, C, s_out, quad_out);
  B200_LAUNCH_CHECK("fm_fwd_kernel");
  return B200CTR_OK;
}

extern "C" int64_t b200ctr_fm_fused_bwd_workspace_bytes(void) { return (int64_t)num_sms() * FM_CP * FM_D * 4; }

extern "C" int b200ctr_fm_fused_bwd(const void* z_bf16, const float* Ucat, const float* s, const float* dquad,
                                    const float* dz_resid, int64_t B, int F, int C, int D, float* dz, float* dUcat,
                                    void* workspace, void* stream) {
  B200_REQUIRE(z_bf16 && Ucat && s && dquad && dz && dUcat && workspace, B200CTR_E_INVALID_ARG, "fm_fused_bwd: null pointer");
  B200_REQUIRE(fm_shape_ok(F, C, D), B200CTR_E_UNSUPPORTED_SHAPE, "fm_fused_bwd: F=%d C=%d D=%d unsupported", F, C, D);
  B200_REQUIRE(B > 0, B200CTR_E_INVALID_ARG, "fm_fused_bwd: empty batch");
  cudaStream_t st = as_stream(stream);
  const size_t smem = (size_t)2 * FM_D * FM_CP * 2 + 2 * FM_ZTILE + (size_t)2 * 128 * FM_CP * 2 + (size_t)2 * 2 * FM_CP * 4 +
                      (size_t)8 * 32 * 36 * 4 + sizeof(FmBwdBars) + 16;
  cudaError_t e = cudaFuncSetAttribute(fm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fm_bwd smem)");
  const int grid = (int)std::min<int64_t>(B, (int64_t)num_sms());
  fm_bwd_kernel<<<grid, FM_NT, smem, st>>>((const __nv_bfloat16*)z_bf16, Ucat, s, dquad, dz_resid, B, F, C, dz, (float*)workspace);
  B200_LAUNCH_CHECK("fm_bwd_kernel");
  fm_bwd_reduce_kernel<<<(FM_D * C + RED_TX - 1) / RED_TX, dim3(RED_TX, RED_TY), 0, st>>>((const float*)workspace, grid, C, dUcat);
  B200_LAUNCH_CHECK("fm_bwd_reduce_kernel");
  return B200CTR_OK;
}
