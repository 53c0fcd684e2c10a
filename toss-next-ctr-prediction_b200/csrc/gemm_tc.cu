// Dense contraction on the 5th-gen tensor cores (tcgen05 + TMEM), bf16 operands / fp32 accumulation.
//
// This is the `amp: bf16` path of the reference (autocast casts matmul operands to bf16): activations and
// weights stay fp32 in HBM (so every other kernel is shared with the fp32 parity path) and are converted to bf16
// on their way into shared memory.  Because the operands pass through registers for the conversion anyway, the
// kernel writes them straight into the canonical UMMA shared-memory layout (SWIZZLE_NONE "interleaved" core
// matrices: 8 rows x 16 bytes), choosing K-major or MN-major per operand from its global contiguity, so all of
// NN / NT / TN contractions run without a transpose pass.
//
//   CTA = 256 threads, tile 128 x BN (BN in {64,96,128,192,256}) x 64, 2 smem stages.
//   loop: all warps stage tile kt (LDG.128 -> cvt.bf16x2 -> STS.128) | fence.proxy.async | bar |
//         one thread issues 4 x tcgen05.mma.cta_group::1.kind::f16 (K=16 each), tcgen05.commit -> mbarrier[stage];
//         the MMAs of tile kt run asynchronously under the staging of tile kt+1.
//   epilogue: tcgen05.ld 32x32b.x32 (TMEM -> registers), transposed through shared memory so that the fused
//         epilogue (bias / act / dropout / masks / residual) reads and writes global memory coalesced.
//   split-K: grid.z slices of K write raw partials; the shared reduce kernel applies the epilogue.
#include "gemm_common.cuh"
#include "tc_common.cuh"
#include <cuda_bf16.h>
#include <type_traits>
#include <algorithm>
#include <stdlib.h>

namespace b200ctr {

constexpr int TC_BM = 128;
constexpr int TC_BK = 64;
constexpr int TC_NT = 256;
constexpr int TC_STAGES = 2;

// Stage one operand tile (R mn-rows x 64 k) into the canonical interleaved layout:
//   core matrix (mn-group g, k-group c) at ((c * R/8) + g) * 128 bytes; SBO = 128, LBO = R*16.
//   KMAJOR  (global [mn][k], k contiguous): unit = (row r, k-group c), 16-B smem row = 8 consecutive k of row r.
//   MNMAJOR (global [k][mn], mn contiguous): unit = (k row, mn-group g), 16-B smem row = 8 consecutive mn at row k.
// Consecutive threads take consecutive smem rows (16-B stride) => conflict-free STS.128, 32-B global sectors.
// SPLIT > 1 (split-precision mode, see gemm_tc_kernel): the tile is written SPLIT times -- bf16(v) at `off`, then the bf16
// of what is still missing at off + part * R*TC_BK*2 -- so that the parts together carry 16 (SPLIT = 2) or 24 (SPLIT = 3)
// mantissa bits of every fp32 operand.
__device__ __forceinline__ void resid8(float* v, const uint4& p) {       // v -= the bf16 values just packed (exact in fp32)
  v[0] -= __uint_as_float(p.x << 16); v[1] -= __uint_as_float(p.x & 0xffff0000u);
  v[2] -= __uint_as_float(p.y << 16); v[3] -= __uint_as_float(p.y & 0xffff0000u);
  v[4] -= __uint_as_float(p.z << 16); v[5] -= __uint_as_float(p.z & 0xffff0000u);
  v[6] -= __uint_as_float(p.w << 16); v[7] -= __uint_as_float(p.w & 0xffff0000u);
}

template <int R, bool KMAJOR, int SPLIT = 1>
__device__ __forceinline__ void stage_operand(unsigned char* smem, const float* __restrict__ g, int64_t ld, int64_t mn0,
                                              int64_t mn_end, int64_t k0, int64_t k_end, bool vec_ok, int tid) {
  constexpr int UNITS = R * 8;
  static_assert(UNITS % TC_NT == 0, "tile/threads mismatch");
#pragma unroll 4
  for (int it = 0; it < UNITS / TC_NT; ++it) {
    const int u = tid + it * TC_NT;
    float v[8];
    uint32_t off;
    if (KMAJOR) {
      const int r = u % R, c = u / R;
      off = (uint32_t)((c * (R / 8) + (r >> 3)) * 128 + (r & 7) * 16);
      const int64_t m = mn0 + r, k = k0 + c * 8;
      if (m < mn_end && vec_ok && k + 7 < k_end) {
        const float4* p = reinterpret_cast<const float4*>(g + m * ld + k);
        float4 a = __ldg(p), b = __ldg(p + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (m < mn_end && k + j < k_end) ? g[m * ld + k + j] : 0.f;
      }
    } else {
      const int kk = u % TC_BK, gidx = u / TC_BK;
      off = (uint32_t)(((kk >> 3) * (R / 8) + gidx) * 128 + (kk & 7) * 16);
      const int64_t k = k0 + kk, m = mn0 + gidx * 8;
      if (k < k_end && vec_ok && m + 7 < mn_end) {
        const float4* p = reinterpret_cast<const float4*>(g + k * ld + m);
        float4 a = __ldg(p), b = __ldg(p + 1);
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (k < k_end && m + j < mn_end) ? g[k * ld + m + j] : 0.f;
      }
    }
    uint4 packed = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    *reinterpret_cast<uint4*>(smem + off) = packed;
#pragma unroll
    for (int part = 1; part < SPLIT; ++part) {          // residual tiles: bf16(v - hi), bf16(v - hi - mid)
      resid8(v, packed);
      packed = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
      *reinterpret_cast<uint4*>(smem + off + part * (R * TC_BK * 2)) = packed;
    }
  }
}

// SPLIT = split-precision fp32 emulation for the `amp: none` mode (SURVEY hard part 5): every fp32 operand is split into
// bf16 parts a = a0 + a1 (+ a2) (each the bf16 rounding of what the previous parts miss) and the product is accumulated as
// the sum of a_i.b_j over i + j < SPLIT in the same fp32 TMEM accumulator: SPLIT = 2 -> 3 MMAs per k-step, error ~2^-17 per
// product term; SPLIT = 3 -> 6 MMAs, dropped terms ~2^-24 = fp32 rounding itself.  Operands are fp32 in HBM either way, so
// the kernel stays bound by the operand loads / the epilogue stores and the extra MMAs are nearly free.
// NST = ring depth actually needed (1 when the whole contraction depth fits one k-tile: K <= 64, the encoder projections).
template <int BN, bool TA, bool TB, int SPLIT = 1, int NST = TC_STAGES>
__global__ void __launch_bounds__(TC_NT, 2)
gemm_tc_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B,
               int64_t ldb, float* __restrict__ C, int64_t ldc, EpiDev epi, int split_k, int64_t k_chunk,
               float* __restrict__ partial, int vec_epi) {
  constexpr int A_HALF = TC_BM * TC_BK * 2, B_HALF = BN * TC_BK * 2;      // one bf16 tile; X3 stores the residual tile behind it
  constexpr int A_BYTES = A_HALF * SPLIT;
  constexpr int B_BYTES = B_HALF * SPLIT;
  constexpr int TMEM_COLS = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* sA = smem;                              // [NST][A_BYTES]
  unsigned char* sB = smem + NST * A_BYTES;              // [NST][B_BYTES]
  __shared__ __align__(8) unsigned long long mbar[NST];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t m0 = (int64_t)blockIdx.y * TC_BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_chunk;
  const int64_t kend = min(K, kbeg + k_chunk);
  const int KT = (int)((kend - kbeg + TC_BK - 1) / TC_BK);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < NST; ++s) mbar_init(smem_u32(&mbar[s]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = tmem_base_s;

  const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool b_vec = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
  // instruction descriptor: D=f32, A=B=bf16, majors, N>>3, M>>4
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((TA ? 1u : 0u) << 15) | ((TB ? 0u : 1u) << 16) |
                         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  constexpr uint32_t LBO_A = TC_BM * 16, LBO_B = BN * 16, SBO = 128;

  for (int kt = 0; kt < KT; ++kt) {
    const int s = kt % NST;
    if (kt >= NST) mbar_wait(smem_u32(&mbar[s]), (uint32_t)(((kt / NST) - 1) & 1));   // stage free again
    const int64_t k0 = kbeg + (int64_t)kt * TC_BK;
    // A operand: opA(A)[m,k].  !TA: stored [M,K] -> K-major.  TA: stored [K,M] -> MN-major.
    stage_operand<TC_BM, !TA, SPLIT>(sA + s * A_BYTES, A, lda, m0, M, k0, kend, a_vec, tid);
    // B operand: opB(B)[k,n] as the UMMA "B" (N x K).  TB: stored [N,K] -> K-major.  !TB: stored [K,N] -> MN-major.
    stage_operand<BN, TB, SPLIT>(sB + s * B_BYTES, B, ldb, n0, N, k0, kend, b_vec, tid);
    fence_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_base = smem_u32(sA + s * A_BYTES), b_base = smem_u32(sB + s * B_BYTES);
#pragma unroll
      for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
        const uint64_t da = make_desc(a_base + k4 * 2 * LBO_A, LBO_A, SBO);
        const uint64_t db = make_desc(b_base + k4 * 2 * LBO_B, LBO_B, SBO);
        umma_bf16(tmem_d, da, db, idesc, (kt > 0 || k4 > 0) ? 1u : 0u);
#pragma unroll
        for (int i = 0; i < SPLIT; ++i)
#pragma unroll
          for (int j = 0; j < SPLIT - i; ++j)
            if (i + j > 0)                                   // the cross terms a_i . b_j, i + j < SPLIT
              umma_bf16(tmem_d, make_desc(a_base + i * A_HALF + k4 * 2 * LBO_A, LBO_A, SBO),
                        make_desc(b_base + j * B_HALF + k4 * 2 * LBO_B, LBO_B, SBO), idesc, 1u);
      }
      umma_commit(smem_u32(&mbar[s]));
    }
  }
  if (KT > 0) mbar_wait(smem_u32(&mbar[(KT - 1) % NST]), (uint32_t)(((KT - 1) / NST) & 1));
  tc_fence_after();
  __syncthreads();     // every warp is past its last smem/MMA dependency: the stage buffers become epilogue scratch

  // ---- epilogue: TMEM -> registers -> smem transpose -> coalesced, 128-bit global accesses ---------------------
  // per warp: a 32 x 32 fp32 chunk goes through a 32 x 36 scratch (16-B aligned rows, conflict-free both ways);
  // then 8 lanes cover one row (float4 each), 4 rows per instruction, auxiliary operands loaded 4 rows ahead.
  constexpr int SROW = 36;
  float* scratch = reinterpret_cast<float*>(smem) + warp * (32 * SROW);
  const int lane_grp = warp & 3;                 // TMEM lanes [32*lane_grp, +32) are visible to this warp
  const int col_half = warp >> 2;                // warps 0-3: even column chunks, 4-7: odd column chunks
  constexpr int NCHUNK = BN / 32;
  const int rg = lane >> 3, c4 = lane & 7;
  for (int cc = col_half; cc < NCHUNK; cc += 2) {
    uint32_t r[32];
    if (KT > 0) tmem_ld32(tmem_d + ((uint32_t)(lane_grp * 32) << 16) + (uint32_t)(cc * 32), r);
    else {
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = 0u;
    }
#pragma unroll
    for (int q = 0; q < 8; ++q)
      *reinterpret_cast<uint4*>(scratch + lane * SROW + q * 4) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
    __syncwarp();
    if (vec_epi) {
      const int64_t n = n0 + cc * 32 + c4 * 4;
#pragma unroll
      for (int pg = 0; pg < 2; ++pg) {
        float4 acc[4];
        EpiAux4 aux[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int row = (pg * 4 + u) * 4 + rg;
          const int64_t m = m0 + lane_grp * 32 + row;
          acc[u] = *reinterpret_cast<const float4*>(scratch + row * SROW + c4 * 4);
          if (split_k == 1 && m < M && n < N) epi_load4(aux[u], m, n, epi);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int row = (pg * 4 + u) * 4 + rg;
          const int64_t m = m0 + lane_grp * 32 + row;
          if (m < M && n < N) {
            if (split_k > 1) *reinterpret_cast<float4*>(partial + ((int64_t)blockIdx.z * M + m) * N + n) = acc[u];
            else *reinterpret_cast<float4*>(C + m * ldc + n) = apply_epi4(acc[u], aux[u], m, n, N, epi);
          }
        }
      }
    } else {
      const int64_t n = n0 + cc * 32 + lane;
#pragma unroll 4
      for (int rr = 0; rr < 32; ++rr) {
        const int64_t m = m0 + lane_grp * 32 + rr;
        if (m < M && n < N) {
          float acc = scratch[rr * SROW + lane];
          if (split_k > 1) partial[((int64_t)blockIdx.z * M + m) * N + n] = acc;
          else C[m * ldc + n] = apply_epi(acc, m, n, N, epi);
        }
      }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(TMEM_COLS) : "memory");
  }
}


// =============================================================================================================
// v2: persistent, warp-specialised pipeline.
//   warps 0-7  producers : stream operand tiles (LDG.128 -> bf16 -> STS.128) into a WS_STAGES-deep ring
//   warps 8-11 epilogue  : TMEM -> registers -> smem transpose -> fused epilogue -> coalesced 128-bit stores
//   warp  12   MMA       : one thread issues tcgen05.mma, tcgen05.commit releases ring slots / publishes accumulators
// Two TMEM accumulator buffers: the epilogue of tile i overlaps the loads and MMAs of tile i+1 (for the K=64
// encoder shapes a tile is a single ring slot, so the ring runs WS_STAGES tiles ahead of the epilogue).
// Each CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...; tile -> (m-tile, n-tile, k-split).
// =============================================================================================================
// role split (producer warps, epilogue warps): (4, 8) for the output-heavy small-K shapes, (8, 4) for deep-K shapes
constexpr int WS_NT = 13 * 32;     // 416 threads in both splits


template <int R, bool KMAJOR, int WS_PROD_THREADS, int SPLIT = 1>
__device__ __forceinline__ void stage_operand_ws(unsigned char* smem, const float* __restrict__ g, int64_t ld, int64_t mn0,
                                                 int64_t mn_end, int64_t k0, int64_t k_end, bool vec_ok, int ptid) {
  constexpr int UNITS = R * 8;
  static_assert(UNITS % WS_PROD_THREADS == 0, "tile/threads mismatch");
  constexpr int ITERS = UNITS / WS_PROD_THREADS;
  constexpr int UNR = (ITERS % 4 == 0) ? 4 : ((ITERS % 3 == 0) ? 3 : ((ITERS % 2 == 0) ? 2 : 1));
#pragma unroll 1
  for (int it0 = 0; it0 < ITERS; it0 += UNR) {
    float4 va[UNR], vb[UNR];
    uint32_t off[UNR];
#pragma unroll
    for (int q = 0; q < UNR; ++q) {
      const int u = ptid + (it0 + q) * WS_PROD_THREADS;
      va[q] = make_float4(0.f, 0.f, 0.f, 0.f); vb[q] = va[q];
      if (KMAJOR) {
        const int r = u % R, c = u / R;
        off[q] = (uint32_t)((c * (R / 8) + (r >> 3)) * 128 + (r & 7) * 16);
        const int64_t m = mn0 + r, k = k0 + c * 8;
        if (m < mn_end) {
          const float* p = g + m * ld + k;
          if (vec_ok && k + 7 < k_end) { va[q] = __ldg(reinterpret_cast<const float4*>(p)); vb[q] = __ldg(reinterpret_cast<const float4*>(p) + 1); }
          else {
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = (k + j < k_end) ? p[j] : 0.f;
            va[q] = make_float4(t[0], t[1], t[2], t[3]); vb[q] = make_float4(t[4], t[5], t[6], t[7]);
          }
        }
      } else {
        const int kk = u % TC_BK, gidx = u / TC_BK;
        off[q] = (uint32_t)(((kk >> 3) * (R / 8) + gidx) * 128 + (kk & 7) * 16);
        const int64_t k = k0 + kk, m = mn0 + gidx * 8;
        if (k < k_end) {
          const float* p = g + k * ld + m;
          if (vec_ok && m + 7 < mn_end) { va[q] = __ldg(reinterpret_cast<const float4*>(p)); vb[q] = __ldg(reinterpret_cast<const float4*>(p) + 1); }
          else {
            float t[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) t[j] = (m + j < mn_end) ? p[j] : 0.f;
            va[q] = make_float4(t[0], t[1], t[2], t[3]); vb[q] = make_float4(t[4], t[5], t[6], t[7]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < UNR; ++q) {
      uint4 packed = make_uint4(pack_bf16(va[q].x, va[q].y), pack_bf16(va[q].z, va[q].w), pack_bf16(vb[q].x, vb[q].y), pack_bf16(vb[q].z, vb[q].w));
      *reinterpret_cast<uint4*>(smem + off[q]) = packed;
      if (SPLIT > 1) {
        float v[8] = {va[q].x, va[q].y, va[q].z, va[q].w, vb[q].x, vb[q].y, vb[q].z, vb[q].w};
#pragma unroll
        for (int part = 1; part < SPLIT; ++part) {
          resid8(v, packed);
          packed = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          *reinterpret_cast<uint4*>(smem + off[q] + part * (R * TC_BK * 2)) = packed;
        }
      }
    }
  }
}

// Same layout, operands already bf16 in global memory: one 16-byte cp.async (LDGSTS) == one core-matrix row, no
// registers in between, so a producer thread can keep several ring stages of loads in flight (the operand tiles mostly
// come from L2, ~1 us away; with register staging only ~32 KB per SM were in flight and the tensor pipe starved).
// Partially valid rows use the src-size form (zero fill); rows out of range are zeroed with a plain store.
template <int R, bool KMAJOR, int WS_PROD_THREADS, int SPLIT = 1>
__device__ __forceinline__ void stage_operand_ws(unsigned char* smem, const __nv_bfloat16* __restrict__ g, int64_t ld, int64_t mn0,
                                                 int64_t mn_end, int64_t k0, int64_t k_end, bool vec_ok, int ptid) {
  constexpr int UNITS = R * 8;
  static_assert(UNITS % WS_PROD_THREADS == 0, "tile/threads mismatch");
  constexpr int ITERS = UNITS / WS_PROD_THREADS;
#pragma unroll 4
  for (int it = 0; it < ITERS; ++it) {
    const int u = ptid + it * WS_PROD_THREADS;
    uint32_t off;
    int64_t row, col, row_end, col_end;          // 8 consecutive elements [row][col .. col+7] of the global matrix
    // MN-major operands: a warp copies 8 k-rows x 4 adjacent 16-byte chunks, so every 32-byte sector it touches in global
    // memory is fully used (64 contiguous bytes per row) and every quarter-warp writes one whole 128-byte core matrix
    const int r8 = u & 7, c4 = (u >> 3) & 3, rest = u >> 5;
    if (KMAJOR) {        // measured: for K-major operands consecutive threads on consecutive rows is the faster mapping
      const int r = u % R, c = u / R;
      off = (uint32_t)((c * (R / 8) + (r >> 3)) * 128 + (r & 7) * 16);
      row = mn0 + r; col = k0 + c * 8; row_end = mn_end; col_end = k_end;
    } else {
      const int gq = rest % (R / 32), kg = rest / (R / 32);
      const int gidx = gq * 4 + c4, kk = kg * 8 + r8;
      off = (uint32_t)((kg * (R / 8) + gidx) * 128 + r8 * 16);
      row = k0 + kk; col = mn0 + gidx * 8; row_end = k_end; col_end = mn_end;
    }
    const __nv_bfloat16* p = g + row * ld + col;
    if (row < row_end && col < col_end && vec_ok) {
      const int valid = (int)min((int64_t)8, col_end - col) * 2;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem + off)), "l"(p), "r"(valid) : "memory");
    } else {
      unsigned short t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = (row < row_end && col + j < col_end) ? __bfloat16_as_ushort(p[j]) : (unsigned short)0;
      *reinterpret_cast<uint4*>(smem + off) = make_uint4((uint32_t)t[0] | ((uint32_t)t[1] << 16), (uint32_t)t[2] | ((uint32_t)t[3] << 16),
                                                         (uint32_t)t[4] | ((uint32_t)t[5] << 16), (uint32_t)t[6] | ((uint32_t)t[7] << 16));
    }
  }
}

template <int BN, bool TA, bool TB, int STAGES, int WS_PROD_WARPS, int WS_EPI_WARPS, typename TIn, int SPLIT = 1>
__global__ void __launch_bounds__(WS_NT, 1)
gemm_tc_ws_kernel(int64_t M, int64_t N, int64_t K, const TIn* __restrict__ A, int64_t lda, const TIn* __restrict__ B,
                  int64_t ldb, float* __restrict__ C, int64_t ldc, EpiDev epi, int split_k, int64_t k_chunk,
                  float* __restrict__ partial, int vec_epi, int tiles_m, int tiles_n) {
  static_assert(SPLIT == 1 || std::is_same<TIn, float>::value, "split precision applies to fp32 operands");
  constexpr int A_HALF = TC_BM * TC_BK * 2, B_HALF = BN * TC_BK * 2;
  constexpr int A_BYTES = A_HALF * SPLIT;
  constexpr int B_BYTES = B_HALF * SPLIT;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int ACC_COLS = BN <= 64 ? 64 : (BN <= 128 ? 128 : 256);     // columns per accumulator buffer
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  constexpr int SROW = 36;
  extern __shared__ __align__(128) unsigned char smem[];
  float* epi_scratch = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);       // WS_EPI_WARPS x 32 x SROW floats
  __shared__ __align__(8) unsigned long long full_bar[STAGES], empty_bar[STAGES], accfull_bar[2], accempty_bar[2];
  __shared__ uint32_t tmem_base_s;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total_tiles = tiles_m * tiles_n * split_k;

  if (warp == WS_PROD_WARPS + WS_EPI_WARPS) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) { mbar_init(smem_u32(&full_bar[s]), WS_PROD_WARPS); mbar_init(smem_u32(&empty_bar[s]), 1); }
    mbar_init(smem_u32(&accfull_bar[0]), 1); mbar_init(smem_u32(&accfull_bar[1]), 1);
    mbar_init(smem_u32(&accempty_bar[0]), WS_EPI_WARPS); mbar_init(smem_u32(&accempty_bar[1]), WS_EPI_WARPS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;

  if (warp < WS_PROD_WARPS) {
    // ------------------------------------------------ producers ------------------------------------------------
    constexpr int64_t VEC_LD = std::is_same<TIn, float>::value ? 3 : 7;         // 16-byte rows: ld % 4 (fp32) / % 8 (bf16)
    const bool a_vec = ((lda & VEC_LD) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    const bool b_vec = ((ldb & VEC_LD) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);
    constexpr bool ASYNC_STAGING = !std::is_same<TIn, float>::value;           // bf16 operands: cp.async straight into the ring
    constexpr uint32_t LAG = STAGES - 1;
    uint32_t g = 0, done = 0;                             // ring position across all tiles of this CTA; published stages
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
      const int nt = t % tiles_n, mt = (t / tiles_n) % tiles_m, z = t / (tiles_n * tiles_m);
      const int64_t m0 = (int64_t)mt * TC_BM, n0 = (int64_t)nt * BN;
      const int64_t kbeg = (int64_t)z * k_chunk, kend = min(K, kbeg + k_chunk);
      const int KT = (int)((kend - kbeg + TC_BK - 1) / TC_BK);
      for (int kt = 0; kt < KT; ++kt, ++g) {
        const int s = g % STAGES;
        const uint32_t use = g / STAGES;
        if (use > 0) mbar_wait(smem_u32(&empty_bar[s]), (use - 1) & 1);
        unsigned char* sa = smem + s * STAGE_BYTES;
        unsigned char* sb = sa + A_BYTES;
        const int64_t k0 = kbeg + (int64_t)kt * TC_BK;
        stage_operand_ws<TC_BM, !TA, WS_PROD_WARPS * 32, SPLIT>(sa, A, lda, m0, M, k0, kend, a_vec, tid);
        stage_operand_ws<BN, TB, WS_PROD_WARPS * 32, SPLIT>(sb, B, ldb, n0, N, k0, kend, b_vec, tid);
        if (ASYNC_STAGING) {
          // stage g is in flight; publish stage g - LAG once this thread's (and, after the warp sync, this warp's) copies of
          // it have landed.  LAG = STAGES - 1 keeps the whole ring in flight without ever waiting on a slot the MMA warp
          // cannot free yet.
          asm volatile("cp.async.commit_group;" ::: "memory");
          if (g + 1 - done > LAG) {
            asm volatile("cp.async.wait_group %0;" ::"n"(LAG) : "memory");
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&full_bar[done % STAGES]));
            ++done;
          }
        } else {
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&full_bar[s]));
        }
      }
    }
    if (ASYNC_STAGING) {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      fence_async_smem();
      __syncwarp();
      for (; done < g; ++done)
        if (lane == 0) mbar_arrive(smem_u32(&full_bar[done % STAGES]));
    }
  } else if (warp == WS_PROD_WARPS + WS_EPI_WARPS) {
    // ------------------------------------------------ MMA issuer -----------------------------------------------
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((TA ? 1u : 0u) << 15) | ((TB ? 0u : 1u) << 16) |
                             ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
      constexpr uint32_t LBO_A = TC_BM * 16, LBO_B = BN * 16, SBO = 128;
      uint32_t g = 0, a = 0;
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++a) {
        const int z = t / (tiles_n * tiles_m);
        const int64_t kbeg = (int64_t)z * k_chunk, kend = min(K, kbeg + k_chunk);
        const int KT = (int)((kend - kbeg + TC_BK - 1) / TC_BK);
        const uint32_t buf = a & 1, use = a >> 1;
        if (use > 0) mbar_wait(smem_u32(&accempty_bar[buf]), (use - 1) & 1);      // epilogue drained this buffer
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + buf * ACC_COLS;
        for (int kt = 0; kt < KT; ++kt, ++g) {
          const int s = g % STAGES;
          mbar_wait(smem_u32(&full_bar[s]), (g / STAGES) & 1);
          tc_fence_after();
          const uint32_t a_base = smem_u32(smem + s * STAGE_BYTES), b_base = a_base + A_BYTES;
#pragma unroll
          for (int k4 = 0; k4 < TC_BK / 16; ++k4) {
            const uint64_t da = make_desc(a_base + k4 * 2 * LBO_A, LBO_A, SBO);
            const uint64_t db = make_desc(b_base + k4 * 2 * LBO_B, LBO_B, SBO);
            umma_bf16(tmem_d, da, db, idesc, (kt > 0 || k4 > 0) ? 1u : 0u);
#pragma unroll
            for (int i = 0; i < SPLIT; ++i)
#pragma unroll
              for (int j = 0; j < SPLIT - i; ++j)
                if (i + j > 0)
                  umma_bf16(tmem_d, make_desc(a_base + i * A_HALF + k4 * 2 * LBO_A, LBO_A, SBO),
                            make_desc(b_base + j * B_HALF + k4 * 2 * LBO_B, LBO_B, SBO), idesc, 1u);
          }
          umma_commit(smem_u32(&empty_bar[s]));            // ring slot reusable once these MMAs have read it
        }
        umma_commit(smem_u32(&accfull_bar[buf]));          // accumulator complete (also covers KT == 0)
      }
    }
  } else {
    // ------------------------------------------------ epilogue -------------------------------------------------
    static_assert(WS_PROD_WARPS % 4 == 0 && WS_EPI_WARPS % 4 == 0 && WS_PROD_WARPS + WS_EPI_WARPS == 12, "role split");
    const int ewarp = warp - WS_PROD_WARPS;                // 0 .. WS_EPI_WARPS-1
    const int ew = ewarp & 3;                              // TMEM lane group == warp % 4
    const int cgrp = ewarp >> 2;                           // column group of this warp
    constexpr int NCG = WS_EPI_WARPS / 4;
    float* scratch = epi_scratch + ewarp * (32 * SROW);
    constexpr int NCHUNK = BN / 32;
    const int rg = lane >> 3, c4 = lane & 7;
    uint32_t a = 0;
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++a) {
      const int nt = t % tiles_n, mt = (t / tiles_n) % tiles_m, z = t / (tiles_n * tiles_m);
      const int64_t m0 = (int64_t)mt * TC_BM, n0 = (int64_t)nt * BN;
      const int64_t kbeg = (int64_t)z * k_chunk, kend = min(K, kbeg + k_chunk);
      const bool has_k = kend > kbeg;
      const uint32_t buf = a & 1, use = a >> 1;
      mbar_wait(smem_u32(&accfull_bar[buf]), use & 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + buf * ACC_COLS;
      if (cgrp >= NCHUNK) {                                   // (BN = 32*k with fewer chunks than column groups)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&accempty_bar[buf]));
      }
      for (int cc = cgrp; cc < NCHUNK; cc += NCG) {
        uint32_t r[32];
        if (has_k) tmem_ld32(tmem_d + ((uint32_t)(ew * 32) << 16) + (uint32_t)(cc * 32), r);
        else {
#pragma unroll
          for (int j = 0; j < 32; ++j) r[j] = 0u;
        }
        if (cc + NCG >= NCHUNK) {        // this warp's last TMEM read of the buffer: hand it back to the MMA warp early
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&accempty_bar[buf]));
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
          *reinterpret_cast<uint4*>(scratch + lane * SROW + q * 4) = make_uint4(r[4 * q], r[4 * q + 1], r[4 * q + 2], r[4 * q + 3]);
        __syncwarp();
        if (vec_epi) {
          const int64_t n = n0 + cc * 32 + c4 * 4;
#pragma unroll
          for (int pg = 0; pg < 2; ++pg) {
            float4 acc[4];
            EpiAux4 aux[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int row = (pg * 4 + u) * 4 + rg;
              const int64_t m = m0 + ew * 32 + row;
              acc[u] = *reinterpret_cast<const float4*>(scratch + row * SROW + c4 * 4);
              if (split_k == 1 && m < M && n < N) epi_load4(aux[u], m, n, epi);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int row = (pg * 4 + u) * 4 + rg;
              const int64_t m = m0 + ew * 32 + row;
              if (m < M && n < N) {
                if (split_k > 1) *reinterpret_cast<float4*>(partial + ((int64_t)z * M + m) * N + n) = acc[u];
                else *reinterpret_cast<float4*>(C + m * ldc + n) = apply_epi4(acc[u], aux[u], m, n, N, epi);
              }
            }
          }
        } else {
          const int64_t n = n0 + cc * 32 + lane;
#pragma unroll 4
          for (int rr = 0; rr < 32; ++rr) {
            const int64_t m = m0 + ew * 32 + rr;
            if (m < M && n < N) {
              float acc = scratch[rr * SROW + lane];
              if (split_k > 1) partial[((int64_t)z * M + m) * N + n] = acc;
              else C[m * ldc + n] = apply_epi(acc, m, n, N, epi);
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WS_PROD_WARPS + WS_EPI_WARPS) {
    __syncwarp();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
  }
}

template <int BN, typename TIn = float, int SPLIT = 1>
static int launch_tc_ws(int tA, int tB, int64_t M, int64_t N, int64_t K, const TIn* A, int64_t lda, const TIn* B,
                        int64_t ldb, float* C, int64_t ldc, const EpiDev& e, int split_k, float* ws, cudaStream_t st) {
  int64_t k_chunk = ceil_div(ceil_div(K, (int64_t)split_k), (int64_t)TC_BK) * TC_BK;
  if (k_chunk <= 0) k_chunk = TC_BK;
  const int tiles_m = (int)ceil_div(M, (int64_t)TC_BM), tiles_n = (int)ceil_div(N, (int64_t)BN);
  const int64_t total = (int64_t)tiles_m * tiles_n * split_k;
  constexpr bool X3 = SPLIT > 1;
  static_assert(!X3 || BN <= 128, "split precision: tiles up to 128 columns (SPLIT bf16 tiles per operand per stage)");
  // ring depth: what fits next to the epilogue scratch in 227 KB
  constexpr int STAGES = SPLIT == 3 ? 2 : (SPLIT == 2 ? (BN <= 64 ? 4 : 3) : (BN <= 128 ? 5 : (BN <= 192 ? 4 : 3)));
  // deep-K: loads dominate -> 8 producer warps; else 8 epilogue warps (output-heavy shapes: the epilogue is the long pole)
  const bool deep_k = k_chunk > 512;
  size_t smem = (size_t)STAGES * (TC_BM * TC_BK * 2 + BN * TC_BK * 2) * SPLIT + (size_t)(deep_k ? 4 : 8) * 32 * 36 * 4;
  if (smem > 232448) {
    set_error("gemm_tc_ws: tile %d x %d parts needs %zu bytes of shared memory", BN, SPLIT, smem);
    return B200CTR_E_UNSUPPORTED_SHAPE;
  }
  const int vec = epi_vec_ok(e, C, ldc, N) ? 1 : 0;
  const int grid = (int)std::min<int64_t>(total, (int64_t)num_sms());
#define WS_LAUNCH1(TA_, TB_, PW_, EW_)                                                                            \
  do {                                                                                                            \
    auto fn = gemm_tc_ws_kernel<BN, TA_, TB_, STAGES, PW_, EW_, TIn, SPLIT>;                                         \
    cudaError_t ee = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
    if (ee != cudaSuccess) return cuda_fail(ee, "cudaFuncSetAttribute(gemm_tc_ws smem)");                         \
    fn<<<grid, WS_NT, smem, st>>>(M, N, K, A, lda, B, ldb, C, ldc, e, split_k, k_chunk, ws, vec, tiles_m, tiles_n); \
  } while (0)
#define WS_LAUNCH(TA_, TB_)                                                                                       \
  do {                                                                                                            \
    if (deep_k) WS_LAUNCH1(TA_, TB_, 8, 4); else WS_LAUNCH1(TA_, TB_, 4, 8);                                      \
  } while (0)
  if (!tA && !tB) WS_LAUNCH(false, false);
  else if (!tA && tB) WS_LAUNCH(false, true);
  else if (tA && !tB) WS_LAUNCH(true, false);
  else {
    set_error("gemm_bf16: transA && transB is not built");
    return B200CTR_E_UNSUPPORTED_SHAPE;
  }
#undef WS_LAUNCH
#undef WS_LAUNCH1
  return B200CTR_OK;
}

template <int BN, int SPLIT = 1, int NST = TC_STAGES>
static int launch_tc(int tA, int tB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                     int64_t ldb, float* C, int64_t ldc, const EpiDev& e, int split_k, float* ws, cudaStream_t st) {
  int64_t k_chunk = ceil_div(ceil_div(K, (int64_t)split_k), (int64_t)TC_BK) * TC_BK;
  if (k_chunk <= 0) k_chunk = TC_BK;
  dim3 grid((unsigned)ceil_div(N, (int64_t)BN), (unsigned)ceil_div(M, (int64_t)TC_BM), (unsigned)split_k);
  size_t smem = (size_t)NST * (TC_BM * TC_BK * 2 + BN * TC_BK * 2) * SPLIT;
  smem = std::max(smem, (size_t)8 * 32 * 36 * 4);
  const int vec = epi_vec_ok(e, C, ldc, N) ? 1 : 0;
#define TC_LAUNCH(TA_, TB_)                                                                                       \
  do {                                                                                                            \
    auto fn = gemm_tc_kernel<BN, TA_, TB_, SPLIT, NST>;                                                                  \
    if (smem + 1024 > 48 * 1024) {                                                                                \
      cudaError_t ee = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);         \
      if (ee != cudaSuccess) return cuda_fail(ee, "cudaFuncSetAttribute(gemm_tc smem)");                          \
    }                                                                                                             \
    fn<<<grid, TC_NT, smem, st>>>(M, N, K, A, lda, B, ldb, C, ldc, e, split_k, k_chunk, ws, vec);                 \
  } while (0)
  if (!tA && !tB) TC_LAUNCH(false, false);
  else if (!tA && tB) TC_LAUNCH(false, true);
  else if (tA && !tB) TC_LAUNCH(true, false);
  else {
    set_error("gemm_bf16: transA && transB is not built");
    return B200CTR_E_UNSUPPORTED_SHAPE;
  }
#undef TC_LAUNCH
  return B200CTR_OK;
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_gemm_bf16(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                                 const float* B, int64_t ldb, float* C, int64_t ldc, const b200ctr_epilogue_t* eh,
                                 int split_k, void* workspace, void* stream) {
  B200_REQUIRE(M >= 0 && N >= 0 && K >= 0, B200CTR_E_INVALID_ARG, "gemm_bf16: negative dimension");
  if (M == 0 || N == 0) return B200CTR_OK;
  B200_REQUIRE(A && B && C, B200CTR_E_INVALID_ARG, "gemm_bf16: null operand");
  B200_REQUIRE(split_k >= 1 && split_k <= 1024, B200CTR_E_INVALID_ARG, "gemm_bf16: split_k out of range");
  B200_REQUIRE(split_k == 1 || workspace, B200CTR_E_INVALID_ARG, "gemm_bf16: split_k > 1 needs a workspace");
  B200_REQUIRE(ceil_div(M, (int64_t)TC_BM) <= 65535, B200CTR_E_UNSUPPORTED_SHAPE, "gemm_bf16: M too large for one launch");
  EpiDev e{};
  e.alpha = 1.f;
  if (eh) {
    e.alpha = eh->alpha; e.bias = eh->bias; e.act = eh->act; e.pre_out = eh->pre_out; e.ld_pre = eh->ld_pre;
    e.drop = eh->drop; e.mask_mode = eh->mask_mode; e.mask_src = eh->mask_src; e.ld_mask = eh->ld_mask;
    e.mask_scale = eh->mask_scale; e.residual = eh->residual; e.ld_res = eh->ld_res;
    e.accum = eh->accum; e.ld_accum = eh->ld_accum;
    B200_REQUIRE(e.mask_mode == B200CTR_MASK_NONE || e.mask_src, B200CTR_E_INVALID_ARG, "gemm_bf16: mask_mode without mask_src");
  }
  cudaStream_t st = as_stream(stream);
  int rc;
  float* ws = (float*)workspace;
  static const bool use_v1 = getenv("B200CTR_TC_V1") != nullptr;      // A/B switch: the non-persistent first version
#define TC_DISPATCH(FN)                                                                                          \
  if (N <= 64) rc = FN<64>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, split_k, ws, st);                 \
  else if (N <= 96) rc = FN<96>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, split_k, ws, st);            \
  else if (N <= 128) rc = FN<128>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, split_k, ws, st);          \
  else if (N % 192 == 0 && N <= 768) rc = FN<192>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, split_k, ws, st); \
  else if (N % 256 == 0) rc = FN<256>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, split_k, ws, st);      \
  else rc = FN<128>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, split_k, ws, st);
  // measured on B200 (profiles/): the persistent pipeline wins when the K loop is deep (MLP / weight-gradient shapes);
  // for the single-k-tile, output-heavy encoder shapes two resident non-persistent CTAs (16 epilogue warps) are faster.
  static const bool force_ws = getenv("B200CTR_TC_WS") != nullptr;
  const int64_t k_per_split = ceil_div(K, (int64_t)split_k);
  if (!force_ws && (use_v1 || k_per_split <= 512)) { TC_DISPATCH(launch_tc) } else { TC_DISPATCH(launch_tc_ws) }
#undef TC_DISPATCH
  if (rc) return rc;
  B200_LAUNCH_CHECK("gemm_tc_kernel");
  if (split_k > 1) {
    rc = launch_splitk_reduce((const float*)workspace, split_k, M, N, C, ldc, e, st);
    if (rc) return rc;
  }
  return B200CTR_OK;
}

// ---- split-precision contraction of fp32 operands: the tensor-core path of the `amp: none` mode ----
namespace b200ctr {
template <int SPLIT>
static int gemm_split(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                      int64_t ldb, float* C, int64_t ldc, const EpiDev& e, int split_k, float* ws, cudaStream_t st) {
  const int64_t k_per_split = ceil_div(K, (int64_t)split_k);
#define SP_ARGS transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, split_k, ws, st
  // A/B switch (measured on B200, profiles/): the persistent warp-specialised pipeline also for the shallow-K shapes --
  // 4 producer + 8 epilogue warps, double-buffered TMEM accumulators, tiles of at most 96 columns (shared-memory budget)
  // Measured on B200 (profiles/README.md, r02g): the persistent pipeline wins for the narrow-output / several-k-tile shapes
  // (491,520 x 64 x 384: 0.57 vs 0.76 ms) and loses for the wide-output single-k-tile ones (x 384 x 64: 1.6 vs 1.2 ms).
  static const bool ws_all = getenv("B200CTR_SPLIT_WS") != nullptr, ws_none = getenv("B200CTR_SPLIT_V1") != nullptr;
  if (!ws_none && k_per_split <= 512 && M >= 4096 && (ws_all || (N <= 64 && k_per_split > TC_BK))) {
    if (N <= 64) return launch_tc_ws<64, float, SPLIT>(SP_ARGS);
    return launch_tc_ws<96, float, SPLIT>(SP_ARGS);
  }
  if (k_per_split <= TC_BK) {
    // one k-tile: a single ring slot, so 2-3 CTAs stay resident per SM and one CTA's epilogue stores overlap the others' loads
    if (N <= 64) return launch_tc<64, SPLIT, 1>(SP_ARGS);
    if (N <= 96) return launch_tc<96, SPLIT, 1>(SP_ARGS);
    return launch_tc<128, SPLIT, 1>(SP_ARGS);
  }
  if (k_per_split <= 512) {
    if (N <= 64) return launch_tc<64, SPLIT, 2>(SP_ARGS);
    if (N <= 96) return launch_tc<96, SPLIT, 2>(SP_ARGS);
    return launch_tc<128, SPLIT, 2>(SP_ARGS);
  }
  if (N <= 64) return launch_tc_ws<64, float, SPLIT>(SP_ARGS);
  if (N <= 96) return launch_tc_ws<96, float, SPLIT>(SP_ARGS);
  return launch_tc_ws<128, float, SPLIT>(SP_ARGS);
#undef SP_ARGS
}
}  // namespace b200ctr

extern "C" int b200ctr_gemm_split(int parts, int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                                  const float* B, int64_t ldb, float* C, int64_t ldc, const b200ctr_epilogue_t* eh,
                                  int split_k, void* workspace, void* stream) {
  B200_REQUIRE(parts == 2 || parts == 3, B200CTR_E_INVALID_ARG, "gemm_split: parts must be 2 (3 MMAs per product) or 3 (6 MMAs)");
  B200_REQUIRE(M >= 0 && N >= 0 && K >= 0, B200CTR_E_INVALID_ARG, "gemm_split: negative dimension");
  if (M == 0 || N == 0) return B200CTR_OK;
  B200_REQUIRE(A && B && C, B200CTR_E_INVALID_ARG, "gemm_split: null operand");
  B200_REQUIRE(split_k >= 1 && split_k <= 1024, B200CTR_E_INVALID_ARG, "gemm_split: split_k out of range");
  B200_REQUIRE(split_k == 1 || workspace, B200CTR_E_INVALID_ARG, "gemm_split: split_k > 1 needs a workspace");
  B200_REQUIRE(ceil_div(M, (int64_t)TC_BM) <= 65535, B200CTR_E_UNSUPPORTED_SHAPE, "gemm_split: M too large for one launch");
  B200_REQUIRE(!(transA && transB), B200CTR_E_UNSUPPORTED_SHAPE, "gemm_split: transA && transB is not built");
  EpiDev e{};
  e.alpha = 1.f;
  if (eh) {
    e.alpha = eh->alpha; e.bias = eh->bias; e.act = eh->act; e.pre_out = eh->pre_out; e.ld_pre = eh->ld_pre;
    e.drop = eh->drop; e.mask_mode = eh->mask_mode; e.mask_src = eh->mask_src; e.ld_mask = eh->ld_mask;
    e.mask_scale = eh->mask_scale; e.residual = eh->residual; e.ld_res = eh->ld_res;
    e.accum = eh->accum; e.ld_accum = eh->ld_accum;
    B200_REQUIRE(e.mask_mode == B200CTR_MASK_NONE || e.mask_src, B200CTR_E_INVALID_ARG, "gemm_split: mask_mode without mask_src");
  }
  cudaStream_t st = as_stream(stream);
  float* ws = (float*)workspace;
  int rc = parts == 2 ? b200ctr::gemm_split<2>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, split_k, ws, st)
                      : b200ctr::gemm_split<3>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, e, split_k, ws, st);
  if (rc) return rc;
  B200_LAUNCH_CHECK("gemm_tc_kernel(split precision)");
  if (split_k > 1) {
    rc = launch_splitk_reduce((const float*)workspace, split_k, M, N, C, ldc, e, st);
    if (rc) return rc;
  }
  return B200CTR_OK;
}

// ---- operands already in bf16 (activations the producer kernel emitted in bf16, weights converted once per step) ----
__global__ void to_bf16_kernel(const float* __restrict__ x, int64_t n8, int64_t n, __nv_bfloat16* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n8) {
    const float4 a = reinterpret_cast<const float4*>(x)[2 * i], b = reinterpret_cast<const float4*>(x)[2 * i + 1];
    reinterpret_cast<uint4*>(out)[i] = make_uint4(b200ctr::pack_bf16(a.x, a.y), b200ctr::pack_bf16(a.z, a.w),
                                                  b200ctr::pack_bf16(b.x, b.y), b200ctr::pack_bf16(b.z, b.w));
  } else if (i == n8) {
    for (int64_t j = n8 * 8; j < n; ++j) out[j] = __float2bfloat16(x[j]);
  }
}

extern "C" int b200ctr_to_bf16(const float* x, int64_t n, void* out_bf16, void* stream) {
  B200_REQUIRE(x && out_bf16, B200CTR_E_INVALID_ARG, "to_bf16: null pointer");
  B200_REQUIRE((((uintptr_t)x | (uintptr_t)out_bf16) & 15) == 0, B200CTR_E_INVALID_ARG, "to_bf16: pointers must be 16-byte aligned");
  if (n == 0) return B200CTR_OK;
  const int64_t n8 = n / 8;
  to_bf16_kernel<<<(unsigned)ceil_div(n8 + 1, 256), 256, 0, as_stream(stream)>>>(x, n8, n, (__nv_bfloat16*)out_bf16);
  B200_LAUNCH_CHECK("to_bf16_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_gemm_bf16in(int transA, int transB, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                                   const void* B, int64_t ldb, float* C, int64_t ldc, const b200ctr_epilogue_t* eh,
                                   int split_k, void* workspace, void* stream) {
  B200_REQUIRE(M >= 0 && N >= 0 && K >= 0, B200CTR_E_INVALID_ARG, "gemm_bf16in: negative dimension");
  if (M == 0 || N == 0) return B200CTR_OK;
  B200_REQUIRE(A && B && C, B200CTR_E_INVALID_ARG, "gemm_bf16in: null operand");
  B200_REQUIRE(split_k >= 1 && split_k <= 1024, B200CTR_E_INVALID_ARG, "gemm_bf16in: split_k out of range");
  B200_REQUIRE(split_k == 1 || workspace, B200CTR_E_INVALID_ARG, "gemm_bf16in: split_k > 1 needs a workspace");
  EpiDev e{};
  e.alpha = 1.f;
  if (eh) {
    e.alpha = eh->alpha; e.bias = eh->bias; e.act = eh->act; e.pre_out = eh->pre_out; e.ld_pre = eh->ld_pre;
    e.drop = eh->drop; e.mask_mode = eh->mask_mode; e.mask_src = eh->mask_src; e.ld_mask = eh->ld_mask;
    e.mask_scale = eh->mask_scale; e.residual = eh->residual; e.ld_res = eh->ld_res;
    e.accum = eh->accum; e.ld_accum = eh->ld_accum;
    B200_REQUIRE(e.mask_mode == B200CTR_MASK_NONE || e.mask_src, B200CTR_E_INVALID_ARG, "gemm_bf16in: mask_mode without mask_src");
  }
  cudaStream_t st = as_stream(stream);
  const __nv_bfloat16* a = (const __nv_bfloat16*)A;
  const __nv_bfloat16* b = (const __nv_bfloat16*)B;
  float* ws = (float*)workspace;
  int rc;
  if (N <= 64) rc = launch_tc_ws<64, __nv_bfloat16>(transA, transB, M, N, K, a, lda, b, ldb, C, ldc, e, split_k, ws, st);
  else if (N <= 128) rc = launch_tc_ws<128, __nv_bfloat16>(transA, transB, M, N, K, a, lda, b, ldb, C, ldc, e, split_k, ws, st);
  else rc = launch_tc_ws<256, __nv_bfloat16>(transA, transB, M, N, K, a, lda, b, ldb, C, ldc, e, split_k, ws, st);
  if (rc) return rc;
  B200_LAUNCH_CHECK("gemm_tc_ws_kernel(bf16 operands)");
  if (split_k > 1) {
    rc = launch_splitk_reduce((const float*)workspace, split_k, M, N, C, ldc, e, st);
    if (rc) return rc;
  }
  return B200CTR_OK;
}
