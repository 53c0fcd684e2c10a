// Dense contraction kernels (fp32 CUDA-core path) + column sums.
//
// C[M,N] = epilogue(alpha * opA(A) . opB(B)).  This is the exact-fp32 path used for parity (the reference runs
// true fp32: TF32 is off by default) and for every shape the tcgen05 path does not cover.  Tiles: 128 x {64,128} x 16,
// 256 threads, 8x4 / 8x8 register micro-tiles arranged as 4x4 blocks so shared-memory float4 reads are
// conflict-free; register-prefetched double buffering; optional deterministic split-K (partials + reduce).
#include "gemm_common.cuh"
#include <algorithm>

namespace b200ctr {

constexpr int BK = 16;
constexpr int NT = 256;

template <int BM, int BN, bool TA, bool TB>
__global__ void __launch_bounds__(NT)
gemm_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t lda,
            const float* __restrict__ B, int64_t ldb, float* __restrict__ C, int64_t ldc,
            EpiDev epi, int split_k, int64_t k_chunk, float* __restrict__ partial) {
  constexpr int TM = BM / 16, TN = BN / 16;        // per-thread micro tile
  constexpr int MB = TM / 4, NB = TN / 4;          // 4x4 blocks
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  constexpr int A4 = BM * BK / 4 / NT, B4 = BK * BN / 4 / NT;
  __shared__ __align__(16) float As[2][BK][LDA_S];
  __shared__ __align__(16) float Bs[2][BK][LDB_S];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM, n0 = (int64_t)blockIdx.x * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * k_chunk;
  const int64_t kend = min(K, kbeg + k_chunk);
  const bool a_vec = ((lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
  const bool b_vec = ((ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(B) & 15) == 0);

  float4 ra[A4], rb[B4];
  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  auto load_tile = [&](int64_t k0) {
#pragma unroll
    for (int s = 0; s < A4; ++s) {
      int slot = tid + s * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!TA) {  // A[M,K]: float4 along k
        int row = slot >> 2, kq = slot & 3;
        int64_t m = m0 + row, k = k0 + kq * 4;
        if (m < M) {
          const float* p = A + m * lda + k;
          if (a_vec && k + 3 < kend) v = *reinterpret_cast<const float4*>(p);
          else {
            if (k < kend) v.x = p[0];
            if (k + 1 < kend) v.y = p[1];
            if (k + 2 < kend) v.z = p[2];
            if (k + 3 < kend) v.w = p[3];
          }
        }
      } else {    // A[K,M]: float4 along m
        int krow = slot / (BM / 4), mq = slot % (BM / 4);
        int64_t k = k0 + krow, m = m0 + mq * 4;
        if (k < kend) {
          const float* p = A + k * lda + m;
          if (a_vec && m + 3 < M) v = *reinterpret_cast<const float4*>(p);
          else {
            if (m < M) v.x = p[0];
            if (m + 1 < M) v.y = p[1];
            if (m + 2 < M) v.z = p[2];
            if (m + 3 < M) v.w = p[3];
          }
        }
      }
      ra[s] = v;
    }
#pragma unroll
    for (int s = 0; s < B4; ++s) {
      int slot = tid + s * NT;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!TB) {  // B[K,N]: float4 along n
        int krow = slot / (BN / 4), nq = slot % (BN / 4);
        int64_t k = k0 + krow, n = n0 + nq * 4;
        if (k < kend) {
          const float* p = B + k * ldb + n;
          if (b_vec && n + 3 < N) v = *reinterpret_cast<const float4*>(p);
          else {
            if (n < N) v.x = p[0];
            if (n + 1 < N) v.y = p[1];
            if (n + 2 < N) v.z = p[2];
            if (n + 3 < N) v.w = p[3];
          }
        }
      } else {    // B[N,K]: float4 along k
        int nrow = slot >> 2, kq = slot & 3;
        int64_t n = n0 + nrow, k = k0 + kq * 4;
        if (n < N) {
          const float* p = B + n * ldb + k;
          if (b_vec && k + 3 < kend) v = *reinterpret_cast<const float4*>(p);
          else {
            if (k < kend) v.x = p[0];
            if (k + 1 < kend) v.y = p[1];
            if (k + 2 < kend) v.z = p[2];
            if (k + 3 < kend) v.w = p[3];
          }
        }
      }
      rb[s] = v;
    }
  };
  auto store_tile = [&](int buf) {
#pragma unroll
    for (int s = 0; s < A4; ++s) {
      int slot = tid + s * NT;
      if (!TA) {
        int row = slot >> 2, kq = slot & 3;
        As[buf][kq * 4 + 0][row] = ra[s].x; As[buf][kq * 4 + 1][row] = ra[s].y;
        As[buf][kq * 4 + 2][row] = ra[s].z; As[buf][kq * 4 + 3][row] = ra[s].w;
      } else {
        int krow = slot / (BM / 4), mq = slot % (BM / 4);
        *reinterpret_cast<float4*>(&As[buf][krow][mq * 4]) = ra[s];
      }
    }
#pragma unroll
    for (int s = 0; s < B4; ++s) {
      int slot = tid + s * NT;
      if (!TB) {
        int krow = slot / (BN / 4), nq = slot % (BN / 4);
        *reinterpret_cast<float4*>(&Bs[buf][krow][nq * 4]) = rb[s];
      } else {
        int nrow = slot >> 2, kq = slot & 3;
        Bs[buf][kq * 4 + 0][nrow] = rb[s].x; Bs[buf][kq * 4 + 1][nrow] = rb[s].y;
        Bs[buf][kq * 4 + 2][nrow] = rb[s].z; Bs[buf][kq * 4 + 3][nrow] = rb[s].w;
      }
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    load_tile(kbeg);
    store_tile(0);
  }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    const bool has_next = (k0 + BK) < kend;
    if (has_next) load_tile(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int ib = 0; ib < MB; ++ib) {
        float4 v = *reinterpret_cast<const float4*>(&As[buf][kk][ib * (BM / MB) + ty * 4]);
        a[ib * 4 + 0] = v.x; a[ib * 4 + 1] = v.y; a[ib * 4 + 2] = v.z; a[ib * 4 + 3] = v.w;
      }
#pragma unroll
      for (int jb = 0; jb < NB; ++jb) {
        float4 v = *reinterpret_cast<const float4*>(&Bs[buf][kk][jb * (BN / NB) + tx * 4]);
        b[jb * 4 + 0] = v.x; b[jb * 4 + 1] = v.y; b[jb * 4 + 2] = v.z; b[jb * 4 + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) store_tile(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  // epilogue / partial store
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + (i / 4) * (BM / MB) + ty * 4 + (i & 3);
    if (m >= M) continue;
#pragma unroll
    for (int jb = 0; jb < NB; ++jb) {
      int64_t n = n0 + jb * (BN / NB) + tx * 4;
      if (split_k > 1) {
        float* dst = partial + ((int64_t)blockIdx.z * M + m) * N + n;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) dst[j] = acc[i][jb * 4 + j];
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) C[m * ldc + n + j] = apply_epi(acc[i][jb * 4 + j], m, n + j, N, epi);
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ partial, int split_k, int64_t M, int64_t N,
                                     float* __restrict__ C, int64_t ldc, EpiDev epi) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  int64_t m = i / N, n = i % N;
  float s = 0.f;
  for (int z = 0; z < split_k; ++z) s += partial[(int64_t)z * M * N + i];   // fixed order: deterministic
  C[m * ldc + n] = apply_epi(s, m, n, N, epi);
}

int launch_splitk_reduce(const float* partial, int split_k, int64_t M, int64_t N, float* C, int64_t ldc, const EpiDev& epi,
                         cudaStream_t st) {
  int64_t tot = M * N;
  if (tot == 0) return B200CTR_OK;
  splitk_reduce_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>(partial, split_k, M, N, C, ldc, epi);
  B200_LAUNCH_CHECK("splitk_reduce_kernel");
  return B200CTR_OK;
}

// column sums, two deterministic passes: [M,N] -> partial[P,N] -> out[N].
// pass 1 (N % 4 == 0, N <= 1024): a block tiles rows x float4-columns (N/4 lanes per row, 256/(N/4) rows per step),
// strides over the rows with 4 independent row-steps in flight, then reduces its row-lanes through shared memory.
constexpr int CS_PARTS = 256;
constexpr int CS_MAXP = 1184;
__global__ void __launch_bounds__(256)
colsum_pass1_v4(const float* __restrict__ X, int64_t ldx, int64_t M, int N, float* __restrict__ part) {
  __shared__ float4 red[256];
  const int n4 = N >> 2;
  const int lanes_per_row = n4;                       // <= 256
  const int rows_per_step = 256 / lanes_per_row;      // >= 1
  const int c = threadIdx.x % lanes_per_row, rl = threadIdx.x / lanes_per_row;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < rows_per_step) {
    const int64_t stride = (int64_t)gridDim.x * rows_per_step;
    int64_t r = (int64_t)blockIdx.x * rows_per_step + rl;
    for (; r + 3 * stride < M; r += 4 * stride) {
      float4 a0 = *reinterpret_cast<const float4*>(X + r * ldx + c * 4);
      float4 a1 = *reinterpret_cast<const float4*>(X + (r + stride) * ldx + c * 4);
      float4 a2 = *reinterpret_cast<const float4*>(X + (r + 2 * stride) * ldx + c * 4);
      float4 a3 = *reinterpret_cast<const float4*>(X + (r + 3 * stride) * ldx + c * 4);
      acc.x += (a0.x + a1.x) + (a2.x + a3.x); acc.y += (a0.y + a1.y) + (a2.y + a3.y);
      acc.z += (a0.z + a1.z) + (a2.z + a3.z); acc.w += (a0.w + a1.w) + (a2.w + a3.w);
    }
    for (; r < M; r += stride) {
      float4 a0 = *reinterpret_cast<const float4*>(X + r * ldx + c * 4);
      acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (rl == 0) {
    for (int k = 1; k < rows_per_step; ++k) {
      float4 t = red[k * lanes_per_row + c];
      acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
    }
    *reinterpret_cast<float4*>(part + (int64_t)blockIdx.x * N + c * 4) = acc;
  }
}
__global__ void colsum_pass1(const float* __restrict__ X, int64_t ldx, int64_t M, int64_t N, float* __restrict__ part) {
  int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int p = blockIdx.y;
  int64_t rows = ceil_div(M, (int64_t)CS_PARTS);
  int64_t r0 = p * rows, r1 = min(M, r0 + rows);
  float s = 0.f;
  for (int64_t r = r0; r < r1; ++r) s += X[r * ldx + n];
  part[(int64_t)p * N + n] = s;
}
__global__ void colsum_pass2(const float* __restrict__ part, int P, int64_t N, float* __restrict__ out) {
  const int64_t n = (int64_t)blockIdx.x * RED_TX + threadIdx.x;
  float s = 0.f;
  if (n < N)
    for (int p = threadIdx.y; p < P; p += RED_TY) s += part[(int64_t)p * N + n];
  s = red_finish_32x8(s);
  if (threadIdx.y == 0 && n < N) out[n] = s;
}

template <int BM, int BN>
static int launch_gemm(int tA, int tB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                       int64_t ldb, float* C, int64_t ldc, const EpiDev& e, int split_k, float* ws, cudaStream_t st) {
  int64_t k_chunk = ceil_div(ceil_div(K, (int64_t)split_k), (int64_t)BK) * BK;
  if (k_chunk <= 0) k_chunk = BK;
  dim3 grid((unsigned)ceil_div(N, (int64_t)BN), (unsigned)ceil_div(M, (int64_t)BM), (unsigned)split_k);
  if (grid.y > 65535u) {  // huge M: fold into x by swapping roles is not possible; use 2D->(x=N tiles, y<=65535) guard
    set_error("gemm: M=%lld too large for one launch (tile rows > 65535)", (long long)M);
    return B200CTR_E_UNSUPPORTED_SHAPE;
  }
#define LAUNCH(TA_, TB_) \
  gemm_kernel<BM, BN, TA_, TB_><<<grid, NT, 0, st>>>(M, N, K, A, lda, B, ldb, C, ldc, e, split_k, k_chunk, ws)
  if (!tA && !tB) LAUNCH(false, false);
  else if (!tA && tB) LAUNCH(false, true);
  else if (tA && !tB) LAUNCH(true, false);
  else LAUNCH(true, true);
#undef LAUNCH
  return B200CTR_OK;
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int64_t b200ctr_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int split_k) {
  (void)K;
  return split_k > 1 ? (int64_t)split_k * M * N * (int64_t)sizeof(float) : 0;
}

extern "C" int b200ctr_gemm(int transA, int transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda,
                            const float* B, int64_t ldb, float* C, int64_t ldc, const b200ctr_epilogue_t* eh,
                            int split_k, void* workspace, void* stream) {
  B200_REQUIRE(M >= 0 && N >= 0 && K >= 0, B200CTR_E_INVALID_ARG, "gemm: negative dimension");
  if (M == 0 || N == 0) return B200CTR_OK;
  B200_REQUIRE(A && B && C, B200CTR_E_INVALID_ARG, "gemm: null operand");
  B200_REQUIRE(split_k >= 1 && split_k <= 1024, B200CTR_E_INVALID_ARG, "gemm: split_k out of range");
  B200_REQUIRE(split_k == 1 || workspace, B200CTR_E_INVALID_ARG, "gemm: split_k > 1 needs a workspace");
  EpiDev e{};
  e.alpha = 1.f;
  if (eh) {
    e.alpha = eh->alpha; e.bias = eh->bias; e.act = eh->act; e.pre_out = eh->pre_out; e.ld_pre = eh->ld_pre;
    e.drop = eh->drop; e.mask_mode = eh->mask_mode; e.mask_src = eh->mask_src; e.ld_mask = eh->ld_mask;
    e.mask_scale = eh->mask_scale; e.residual = eh->residual; e.ld_res = eh->ld_res;
    e.accum = eh->accum; e.ld_accum = eh->ld_accum;
    B200_REQUIRE(e.mask_mode == B200CTR_MASK_NONE || e.mask_src, B200CTR_E_INVALID_ARG, "gemm: mask_mode without mask_src");
  }
  cudaStream_t st = as_stream(stream);
  // 16M-row tiles: grid.y limit 65535 * 128 rows = 8.3M rows; larger M is chunked here.
  const int64_t M_CHUNK = 65535LL * 128;
  for (int64_t mo = 0; mo < M; mo += M_CHUNK) {
    int64_t Mc = M - mo < M_CHUNK ? M - mo : M_CHUNK;
    B200_REQUIRE(split_k == 1 || M <= M_CHUNK, B200CTR_E_UNSUPPORTED_SHAPE, "gemm: split_k with chunked M");
    const float* Ac = transA ? A + mo : A + mo * lda;
    EpiDev ec = e;
    if (ec.pre_out) ec.pre_out += mo * ec.ld_pre;
    if (ec.mask_src) ec.mask_src += mo * ec.ld_mask;
    if (ec.residual) ec.residual += mo * ec.ld_res;
    if (ec.accum) ec.accum += mo * ec.ld_accum;
    // NOTE: dropout element ids are relative to the chunk when M is chunked (only for M > 8.3M rows).
    int rc;
    if (N > 64) rc = launch_gemm<128, 128>(transA, transB, Mc, N, K, Ac, lda, B, ldb, C + mo * ldc, ldc, ec, split_k, (float*)workspace, st);
    else rc = launch_gemm<128, 64>(transA, transB, Mc, N, K, Ac, lda, B, ldb, C + mo * ldc, ldc, ec, split_k, (float*)workspace, st);
    if (rc) return rc;
    B200_LAUNCH_CHECK("gemm_kernel");
    if (split_k > 1) {
      int64_t tot = Mc * N;
      splitk_reduce_kernel<<<(unsigned)ceil_div(tot, 256), 256, 0, st>>>((const float*)workspace, split_k, Mc, N, C + mo * ldc, ldc, ec);
      B200_LAUNCH_CHECK("splitk_reduce_kernel");
    }
  }
  return B200CTR_OK;
}

extern "C" int b200ctr_colsum(const float* X, int64_t ldx, int64_t M, int64_t N, float* out, void* workspace, void* stream) {
  B200_REQUIRE(X && out && workspace, B200CTR_E_INVALID_ARG, "colsum: null pointer");
  if (N == 0) return B200CTR_OK;
  cudaStream_t st = as_stream(stream);
  const bool v4 = (N % 4 == 0) && N <= 1024 && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  if (v4) {
    int rows_per_step = 256 / (int)(N / 4);
    int P = (int)std::min<int64_t>(std::max<int64_t>(1, ceil_div(M, (int64_t)rows_per_step * 8)), std::min<int64_t>(CS_MAXP, (int64_t)num_sms() * 8));
    colsum_pass1_v4<<<P, 256, 0, st>>>(X, ldx, M, (int)N, (float*)workspace);
    B200_LAUNCH_CHECK("colsum_pass1_v4");
    colsum_pass2<<<(unsigned)ceil_div(N, (int64_t)RED_TX), dim3(RED_TX, RED_TY), 0, st>>>((const float*)workspace, P, N, out);
  } else {
    dim3 g1((unsigned)ceil_div(N, 128), CS_PARTS);
    colsum_pass1<<<g1, 128, 0, st>>>(X, ldx, M, N, (float*)workspace);
    B200_LAUNCH_CHECK("colsum_pass1");
    colsum_pass2<<<(unsigned)ceil_div(N, (int64_t)RED_TX), dim3(RED_TX, RED_TY), 0, st>>>((const float*)workspace, CS_PARTS, N, out);
  }
  B200_LAUNCH_CHECK("colsum_pass2");
  return B200CTR_OK;
}
