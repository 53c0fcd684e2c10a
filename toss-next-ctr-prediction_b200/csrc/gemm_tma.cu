// Deep-K bf16 contraction fed by TMA:  C[M][N] (fp32) = A[M][K] . B[N][K]^T, both operands bf16 with K contiguous.
//
// This is the first QNN MLP layer of the `amp: bf16` path (qnn_alpha.py:123-129): h = z . W0z^T (8192 x 512 x 12800),
// dW0z = dh^T . z and dz = dh . W0z (the same 105 GFLOP each, expressed as K-major products over bf16 copies /
// transposes that cost < 0.1 ms to make).  gemm_tc.cu stages operands through the SMs' load/store units into the
// SWIZZLE_NONE core-matrix layout, which tops out near 240 TFLOP/s; here the operand tiles are brought in by the TMA
// unit (cp.async.bulk.tensor.2d, 128-byte swizzle, completion on an mbarrier) so no thread touches operand bytes:
//   warp 0 : one thread issues the TMA loads of a 4-stage ring (A 128 x 64, B 256 x 64 per stage = 48 KB)
//   warp 1 : one thread issues tcgen05.mma (M 128, N 256, K 16; SWIZZLE_128B K-major descriptors, +32 B per k step)
//   warps 2-5 : epilogue, TMEM -> registers -> warp-transposed coalesced fp32 stores; two 256-column accumulators
//               alternate so the epilogue of tile i overlaps the loads and MMAs of tile i+1
// Persistent: CTA c walks tiles c, c + grid, ...
#include "tc_common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>

namespace b200ctr {

constexpr int TG_BM = 128, TG_BN = 256, TG_BK = 64, TG_STAGES = 4;
constexpr int TG_A_BYTES = TG_BM * TG_BK * 2, TG_B_BYTES = TG_BN * TG_BK * 2, TG_STAGE_BYTES = TG_A_BYTES + TG_B_BYTES;
constexpr int TG_NT = 6 * 32;

struct TgBars {
  unsigned long long full[TG_STAGES], empty[TG_STAGES], acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

// SWIZZLE_128B, K-major: 8-row groups of 128-byte rows (1024 B apart); LBO unused
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
}

__global__ void __launch_bounds__(TG_NT, 1)
gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* __restrict__ C, int64_t ldc,
                int64_t M, int64_t N, int64_t K, int tiles_m, int tiles_n) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* scr_all = reinterpret_cast<float*>(smem + TG_STAGES * TG_STAGE_BYTES);            // 4 warps x 32 x 36 floats
  TgBars* bars = reinterpret_cast<TgBars*>(scr_all + 4 * 32 * 36);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total = tiles_m * tiles_n;
  const int KT = (int)((K + TG_BK - 1) / TG_BK);

  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  if (tid == 0) {
    for (int s = 0; s < TG_STAGES; ++s) { mbar_init(smem_u32(&bars->full[s]), 1); mbar_init(smem_u32(&bars->empty[s]), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(smem_u32(&bars->acc_full[s]), 1); mbar_init(smem_u32(&bars->acc_empty[s]), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t g = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x) {
        const int mt = t / tiles_n, nt = t % tiles_n;          // consecutive CTAs share the A row block (L2 reuse)
        for (int kt = 0; kt < KT; ++kt, ++g) {
          const uint32_t s = g % TG_STAGES, use = g / TG_STAGES;
          if (use > 0) mbar_wait(smem_u32(&bars->empty[s]), (use - 1) & 1);
          const uint32_t bar = smem_u32(&bars->full[s]);
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)TG_STAGE_BYTES) : "memory");
          const uint32_t sa = smem_u32(smem + s * TG_STAGE_BYTES);
          tma_load_2d(sa, &tmA, bar, kt * TG_BK, mt * TG_BM);
          tma_load_2d(sa + TG_A_BYTES, &tmB, bar, kt * TG_BK, nt * TG_BN);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc(TG_BM, TG_BN, false, false);
      uint32_t g = 0, a = 0;
      for (int t = blockIdx.x; t < total; t += gridDim.x, ++a) {
        const uint32_t buf = a & 1, use = a >> 1;
        if (use > 0) mbar_wait(smem_u32(&bars->acc_empty[buf]), (use - 1) & 1);
        tc_fence_after();
        for (int kt = 0; kt < KT; ++kt, ++g) {
          const uint32_t s = g % TG_STAGES;
          mbar_wait(smem_u32(&bars->full[s]), (g / TG_STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * TG_STAGE_BYTES);
          const uint64_t da = make_desc_sw128(sa), db = make_desc_sw128(sa + TG_A_BYTES);
#pragma unroll
          for (int k = 0; k < TG_BK / 16; ++k)
            umma_bf16(tmem + buf * 256, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, (kt > 0 || k > 0) ? 1u : 0u);
          umma_commit(smem_u32(&bars->empty[s]));
        }
        umma_commit(smem_u32(&bars->acc_full[buf]));
      }
    }
  } else {
    const int q = warp & 3;
    float* scr = scr_all + (warp - 2) * (32 * 36);
    uint32_t a = 0;
    for (int t = blockIdx.x; t < total; t += gridDim.x, ++a) {
      const int mt = t / tiles_n, nt = t % tiles_n;
      const uint32_t buf = a & 1, use = a >> 1;
      mbar_wait(smem_u32(&bars->acc_full[buf]), use & 1);
      tc_fence_after();
      const int64_t m_w = (int64_t)mt * TG_BM + q * 32;
      const int rows_valid = (int)max((int64_t)0, min((int64_t)32, M - m_w));
      const int64_t n0 = (int64_t)nt * TG_BN;
#pragma unroll 1
      for (int cc = 0; cc < TG_BN / 32; ++cc) {
        uint32_t v[32];
        tmem_ld32(tmem + buf * 256 + cc * 32 + ((uint32_t)(q * 32) << 16), v);
        if (cc == TG_BN / 32 - 1) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&bars->acc_empty[buf]));
        }
        if (n0 + cc * 32 < N)            // N % 32 == 0 (checked by the host)
          warp_store_rows<32>(scr, v, C + m_w * ldc + n0 + cc * 32, ldc, rows_valid, lane);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

// fp32 [R][C] (row stride ld_in) -> bf16 [C][R] (contiguous): 32 x 32 tiles through shared memory
__global__ void to_bf16_transposed_kernel(const float* __restrict__ in, int64_t R, int64_t Cc, int64_t ld_in, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int64_t r0 = (int64_t)blockIdx.y * 32, c0 = (int64_t)blockIdx.x * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? in[r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int64_t c = c0 + i, r = r0 + threadIdx.x;
    if (c < Cc && r < R) out[c * R + r] = __float2bfloat16(tile[threadIdx.x][i]);
  }
}
// bf16 [R][C] (contiguous) -> bf16 [C][R]: 64 x 64 tiles, 4-byte accesses on both sides
__global__ void transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int64_t R, int64_t Cc, __nv_bfloat16* __restrict__ out) {
  __shared__ __nv_bfloat16 tile[64][66];
  const int64_t r0 = (int64_t)blockIdx.y * 64, c0 = (int64_t)blockIdx.x * 64;
  const int tx = threadIdx.x, ty = threadIdx.y;            // 32 x 8
  for (int i = ty; i < 64; i += 8) {
    const int64_t r = r0 + i, c = c0 + 2 * tx;
    __nv_bfloat162 v = __floats2bfloat162_rn(0.f, 0.f);
    if (r < R && c + 1 < Cc) v = *reinterpret_cast<const __nv_bfloat162*>(in + r * Cc + c);
    else if (r < R && c < Cc) v.x = in[r * Cc + c];
    tile[i][2 * tx] = v.x; tile[i][2 * tx + 1] = v.y;
  }
  __syncthreads();
  for (int i = ty; i < 64; i += 8) {
    const int64_t c = c0 + i, r = r0 + 2 * tx;
    if (c < Cc && r + 1 < R) {
      __nv_bfloat162 v;
      v.x = tile[2 * tx][i]; v.y = tile[2 * tx + 1][i];
      *reinterpret_cast<__nv_bfloat162*>(out + c * R + r) = v;
    } else if (c < Cc && r < R) {
      out[c * R + r] = tile[2 * tx][i];
    }
  }
}

// The same transposition with 16-byte accesses on both sides (R, Cc multiples of 8, 16-byte aligned bases).  A 64 x 64 tile is
// held as 32-bit words W[row][column pair] (row stride 33: the four scalar stores of a loaded quad and the transposed reads below are
// bank-conflict free); output word (c, rp) = {in[2 rp][c], in[2 rp + 1][c]} is one byte permute of W[2 rp][c / 2] and
// W[2 rp + 1][c / 2].  A warp writes 16 output rows x 32 bytes per store (the 2-byte version: 3.5 M shared-memory bank conflicts
// per launch, 2.7 TB/s -- ncu r02ai).
__global__ void __launch_bounds__(256)
transpose_bf16_v8_kernel(const uint4* __restrict__ in, int64_t R, int64_t Cc, uint4* __restrict__ out) {
  __shared__ uint32_t W[64][33];
  const int64_t r0 = (int64_t)blockIdx.y * 64, c0 = (int64_t)blockIdx.x * 64;
  const int tid = threadIdx.x;
  const int64_t in_ld = Cc >> 3, out_ld = R >> 3;          // row strides in 16-byte units
  uint4 v[2];
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = tid + 256 * k, row = idx >> 3, seg = idx & 7;
    v[k] = make_uint4(0u, 0u, 0u, 0u);
    if (r0 + row < R && c0 + seg * 8 < Cc) v[k] = in[(r0 + row) * in_ld + (c0 >> 3) + seg];
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int idx = tid + 256 * k, row = idx >> 3, seg = idx & 7;
    W[row][seg * 4 + 0] = v[k].x; W[row][seg * 4 + 1] = v[k].y; W[row][seg * 4 + 2] = v[k].z; W[row][seg * 4 + 3] = v[k].w;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int o = tid + 256 * k;
    const int c = (o & 127) >> 1, seg = ((o >> 7) << 1) | (o & 1);     // a warp: 16 output rows x 2 adjacent 16-byte segments
    const uint32_t sel = (c & 1) ? 0x7632u : 0x5410u;                  // high / low halves of the two source words
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int rp = seg * 4 + q;
      w[q] = __byte_perm(W[2 * rp][c >> 1], W[2 * rp + 1][c >> 1], sel);
    }
    if (c0 + c < Cc && r0 + seg * 8 < R) out[(c0 + c) * out_ld + (r0 >> 3) + seg] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

typedef CUresult (*tg_encode_fn_t)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int tg_make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
  static tg_encode_fn_t fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || !p) return cuda_fail(e, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
    fn = (tg_encode_fn_t)p;
  }
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TG_BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with driver error %d (rows=%lld K=%lld ld=%lld)", (int)r, (long long)rows, (long long)K, (long long)ld);
    return B200CTR_E_CUDA;
  }
  return B200CTR_OK;
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_gemm_bf16_kk(const void* A, int64_t lda, const void* B, int64_t ldb, float* C, int64_t ldc, int64_t M,
                                    int64_t N, int64_t K, void* stream) {
  B200_REQUIRE(A && B && C, B200CTR_E_INVALID_ARG, "gemm_bf16_kk: null pointer");
  B200_REQUIRE(M > 0 && N > 0 && K > 0, B200CTR_E_INVALID_ARG, "gemm_bf16_kk: empty problem");
  B200_REQUIRE(N % 32 == 0 && K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0 && ldc % 4 == 0 && (((uintptr_t)A | (uintptr_t)B | (uintptr_t)C) & 15) == 0,
               B200CTR_E_UNSUPPORTED_SHAPE, "gemm_bf16_kk: needs N %% 32 == 0, K, lda, ldb %% 8 == 0, ldc %% 4 == 0 and 16-byte aligned bases");
  alignas(64) CUtensorMap ma, mb;
  int rc = tg_make_map(&ma, A, M, K, lda, TG_BM);
  if (rc) return rc;
  rc = tg_make_map(&mb, B, N, K, ldb, TG_BN);
  if (rc) return rc;
  const int tiles_m = (int)ceil_div(M, (int64_t)TG_BM), tiles_n = (int)ceil_div(N, (int64_t)TG_BN);
  const size_t smem = (size_t)TG_STAGES * TG_STAGE_BYTES + (size_t)4 * 32 * 36 * 4 + sizeof(TgBars) + 1024 + 16;
  cudaError_t e = cudaFuncSetAttribute(gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(gemm_tma smem)");
  const int grid = std::min(tiles_m * tiles_n, num_sms());
  gemm_tma_kernel<<<grid, TG_NT, smem, as_stream(stream)>>>(ma, mb, C, ldc, M, N, K, tiles_m, tiles_n);
  B200_LAUNCH_CHECK("gemm_tma_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_to_bf16_transposed(const float* in, int64_t R, int64_t C, int64_t ld_in, void* out_bf16, void* stream) {
  B200_REQUIRE(in && out_bf16, B200CTR_E_INVALID_ARG, "to_bf16_transposed: null pointer");
  if (R == 0 || C == 0) return B200CTR_OK;
  dim3 grid((unsigned)ceil_div(C, 32), (unsigned)ceil_div(R, 32));
  to_bf16_transposed_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(in, R, C, ld_in, (__nv_bfloat16*)out_bf16);
  B200_LAUNCH_CHECK("to_bf16_transposed_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_transpose_bf16(const void* in_bf16, int64_t R, int64_t C, void* out_bf16, void* stream) {
  B200_REQUIRE(in_bf16 && out_bf16, B200CTR_E_INVALID_ARG, "transpose_bf16: null pointer");
  B200_REQUIRE(C % 2 == 0 && R % 2 == 0, B200CTR_E_UNSUPPORTED_SHAPE, "transpose_bf16: even dimensions only");
  if (R == 0 || C == 0) return B200CTR_OK;
  dim3 grid((unsigned)ceil_div(C, 64), (unsigned)ceil_div(R, 64));
  if (R % 8 == 0 && C % 8 == 0 && (((uintptr_t)in_bf16 | (uintptr_t)out_bf16) & 15) == 0)
    transpose_bf16_v8_kernel<<<grid, 256, 0, as_stream(stream)>>>((const uint4*)in_bf16, R, C, (uint4*)out_bf16);
  else
    transpose_bf16_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>((const __nv_bfloat16*)in_bf16, R, C, (__nv_bfloat16*)out_bf16);
  B200_LAUNCH_CHECK("transpose_bf16_kernel");
  return B200CTR_OK;
}
