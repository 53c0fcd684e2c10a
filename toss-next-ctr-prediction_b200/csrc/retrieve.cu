// K2: DARE target-aware retrieval (reference src/models/dare.py:116-138).
//
// forward, one CTA per sample:
//   1. the (L,) token row is staged in shared memory (coalesced);
//   2. every non-pad position's emb_att row (D=64 fp32 = 256 B) is fetched by one half-warp as 16 x 128-bit
//      no-allocate loads, 4 rows in flight per half-warp (64 rows / 16 KB in flight per CTA); pad positions
//      are never fetched (their score is the constant -1e9);
//   3. score = dot(row, q) in a FIXED order (per-lane 4-term fma chain, then xor-tree 8,4,2,1) + rec[p];
//   4. scores and positions are packed into 64-bit keys (~orderable(score) << 32 | position) and sorted by a
//      CTA bitonic network => sorted top-K directly in the canonical order (score desc, position asc);
//   5. only the K selected emb_rep rows are gathered and written with vals / idx / tok.
// Algorithmic bytes per sample: 8L (ids) + 4D*l (att rows) + 4D*Kv (rep rows) + 4D (q) + 4D*K + 12K (outputs).
//
// backward, one CTA per sample: dq = sum_k dvals_k * emb_att[tok_k]; table gradients are row-sparse
// (K rows per table per sample) and accumulated with 128-bit vector atomics.
#include "common.cuh"

namespace b200ctr {

constexpr int RT_D = 64;
constexpr int RT_THREADS = 256;
constexpr float RT_PAD_SCORE = -1e9f;

__device__ __forceinline__ uint32_t orderable(float s) {
  uint32_t u = __float_as_uint(s);
  return (u >> 31) ? ~u : (u ^ 0x80000000u);
}

template <typename IT>
__global__ void __launch_bounds__(RT_THREADS)
retrieve_fwd_kernel(const IT* __restrict__ seq, const float* __restrict__ q, const float* __restrict__ emb_att,
                    const float* __restrict__ emb_rep, const float* __restrict__ rec, int L, int LP, int K, int pad_id,
                    int64_t vocab, float* __restrict__ sel_rep, float* __restrict__ vals, int32_t* __restrict__ idx,
                    int32_t* __restrict__ tok_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);      // LP
  float* scores = reinterpret_cast<float*>(keys + LP);                              // L
  int32_t* toks = reinterpret_cast<int32_t*>(scores + L);                           // L

  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const int hl = tid & 15;             // lane within the half-warp
  const int hw = tid >> 4;             // half-warp id, 0..15
  constexpr int NHW = RT_THREADS / 16;

  for (int p = tid; p < L; p += RT_THREADS) {
    int64_t t = (int64_t)seq[b * L + p];
    t = t < 0 ? 0 : (t >= vocab ? vocab - 1 : t);     // memory-safe clamp (reference would raise on OOB)
    toks[p] = (int32_t)t;
  }
  const float4 qv = *reinterpret_cast<const float4*>(q + b * RT_D + hl * 4);
  __syncthreads();

  const float4* att4 = reinterpret_cast<const float4*>(emb_att);
  // loop bounds are warp-uniform (both half-warps of a warp run the same trip count) so the full-mask
  // shuffles below are always executed by all 32 lanes
  for (int base = 0; base < L; base += NHW * 4) {
    float4 r[4];
    int t[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int p = base + hw + u * NHW;
      t[u] = (p < L) ? toks[p] : pad_id;
      r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t[u] != pad_id) r[u] = ldg_nc_f4(att4 + (int64_t)t[u] * (RT_D / 4) + hl);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int p = base + hw + u * NHW;
      float d = r[u].x * qv.x;
      d = fmaf(r[u].y, qv.y, d);
      d = fmaf(r[u].z, qv.z, d);
      d = fmaf(r[u].w, qv.w, d);
      d += __shfl_xor_sync(0xffffffffu, d, 8);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      if (hl == 0 && p < L) {
        float s = (t[u] != pad_id) ? d + rec[p] : RT_PAD_SCORE;
        scores[p] = s;
        keys[p] = ((unsigned long long)(~orderable(s)) << 32) | (unsigned)p;
      }
    }
  }
  for (int p = L + tid; p < LP; p += RT_THREADS) keys[p] = ~0ULL;
  __syncthreads();

  // bitonic sort, ascending in the packed key == (score desc, position asc)
  for (int k = 2; k <= LP; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = tid; i < LP; i += RT_THREADS) {
        int ixj = i ^ j;
        if (ixj > i) {
          unsigned long long a = keys[i], c = keys[ixj];
          bool up = ((i & k) == 0);
          if ((a > c) == up) { keys[i] = c; keys[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }

  for (int k = tid; k < K; k += RT_THREADS) {
    int p = (int)(keys[k] & 0xffffffffu);
    vals[b * K + k] = scores[p];
    idx[b * K + k] = p;
    tok_out[b * K + k] = toks[p];
  }
  const float4* rep4 = reinterpret_cast<const float4*>(emb_rep);
  float4* out4 = reinterpret_cast<float4*>(sel_rep + b * (int64_t)K * RT_D);
  for (int k0 = hw; k0 < K; k0 += NHW * 4) {
    float4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int k = k0 + u * NHW;
      r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) {
        int t = toks[(int)(keys[k] & 0xffffffffu)];
        if (t != pad_id) r[u] = ldg_nc_f4(rep4 + (int64_t)t * (RT_D / 4) + hl);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int k = k0 + u * NHW;
      if (k < K) out4[k * (RT_D / 4) + hl] = r[u];
    }
  }
}

__global__ void __launch_bounds__(128)
retrieve_bwd_kernel(const int32_t* __restrict__ tok, const float* __restrict__ q, const float* __restrict__ emb_att,
                    const float* __restrict__ dsel_rep, const float* __restrict__ dvals, int K, int pad_id,
                    float* __restrict__ dq, float* __restrict__ datt, float* __restrict__ drep) {
  __shared__ float4 red[128];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, hl = tid & 15, hw = tid >> 4;
  constexpr int NHW = 8;
  const float4 qv = *reinterpret_cast<const float4*>(q + b * RT_D + hl * 4);
  const float4* att4 = reinterpret_cast<const float4*>(emb_att);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = hw; k < K; k += NHW) {
    int t = tok[b * K + k];
    if (t == pad_id) continue;                      // masked_fill blocks the gradient; padding_idx row gets none
    float dv = dvals[b * K + k];
    float4 a = ldg_nc_f4(att4 + (int64_t)t * (RT_D / 4) + hl);
    acc.x = fmaf(dv, a.x, acc.x); acc.y = fmaf(dv, a.y, acc.y); acc.z = fmaf(dv, a.z, acc.z); acc.w = fmaf(dv, a.w, acc.w);
    if (datt) {
      float4 g = make_float4(dv * qv.x, dv * qv.y, dv * qv.z, dv * qv.w);
      atomicAdd(reinterpret_cast<float4*>(datt + (int64_t)t * RT_D) + hl, g);
    }
    if (drep) {
      float4 g = *reinterpret_cast<const float4*>(dsel_rep + (b * K + k) * (int64_t)RT_D + hl * 4);
      atomicAdd(reinterpret_cast<float4*>(drep + (int64_t)t * RT_D) + hl, g);
    }
  }
  red[tid] = acc;
  __syncthreads();
  if (hw == 0) {
    float4 s = red[hl];
#pragma unroll
    for (int w = 1; w < NHW; ++w) {
      float4 v = red[w * 16 + hl];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<float4*>(dq + b * RT_D + hl * 4) = s;
  }
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_retrieve_fwd(const void* seq, int idx_is_64, const float* q, const float* emb_att,
                                    const float* emb_rep, const float* rec, int64_t B, int L, int K, int D, int pad_id,
                                    int64_t vocab, float* sel_rep, float* vals, int32_t* idx, int32_t* tok, void* stream) {
  B200_REQUIRE(seq && q && emb_att && emb_rep && rec && sel_rep && vals && idx && tok, B200CTR_E_INVALID_ARG,
               "retrieve_fwd: null pointer");
  B200_REQUIRE(D == RT_D, B200CTR_E_UNSUPPORTED_SHAPE, "retrieve_fwd: emb_dim=%d unsupported (kernel is specialised for 64)", D);
  B200_REQUIRE(L >= 1 && L <= 4096, B200CTR_E_UNSUPPORTED_SHAPE, "retrieve_fwd: max_len=%d unsupported (1..4096)", L);
  B200_REQUIRE(K >= 1 && K <= L, B200CTR_E_INVALID_ARG, "retrieve_fwd: need 1 <= K <= L (K=%d, L=%d)", K, L);
  if (B == 0) return B200CTR_OK;
  int LP = 2;
  while (LP < L) LP <<= 1;
  size_t smem = (size_t)LP * 8 + (size_t)L * 8;
  cudaStream_t st = as_stream(stream);
  if (idx_is_64)
    retrieve_fwd_kernel<int64_t><<<(unsigned)B, RT_THREADS, smem, st>>>((const int64_t*)seq, q, emb_att, emb_rep, rec, L, LP, K,
                                                                          pad_id, vocab, sel_rep, vals, idx, tok);
  else
    retrieve_fwd_kernel<int32_t><<<(unsigned)B, RT_THREADS, smem, st>>>((const int32_t*)seq, q, emb_att, emb_rep, rec, L, LP, K,
                                                                          pad_id, vocab, sel_rep, vals, idx, tok);
  B200_LAUNCH_CHECK("retrieve_fwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_retrieve_bwd(const int32_t* tok, const float* q, const float* emb_att, const float* dsel_rep,
                                    const float* dvals, int64_t B, int K, int D, int pad_id, float* dq, float* datt_dense,
                                    float* drep_dense, void* stream) {
  B200_REQUIRE(tok && q && emb_att && dvals && dq, B200CTR_E_INVALID_ARG, "retrieve_bwd: null pointer");
  B200_REQUIRE(!drep_dense || dsel_rep, B200CTR_E_INVALID_ARG, "retrieve_bwd: drep without dsel_rep");
  B200_REQUIRE(D == RT_D, B200CTR_E_UNSUPPORTED_SHAPE, "retrieve_bwd: emb_dim=%d unsupported", D);
  if (B == 0) return B200CTR_OK;
  retrieve_bwd_kernel<<<(unsigned)B, 128, 0, as_stream(stream)>>>(tok, q, emb_att, dsel_rep, dvals, K, pad_id, dq, datt_dense,
                                                                   drep_dense);
  B200_LAUNCH_CHECK("retrieve_bwd_kernel");
  return B200CTR_OK;
}
