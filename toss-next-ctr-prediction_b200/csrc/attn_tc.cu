// Tiny-head attention (head_dim 8) of the DARE encoder layer on the warp-level tensor-core path
// (mma.sync m16n8k8, TF32 operands, fp32 accumulation) -- the `amp: bf16` counterpart of encoder.cu's exact-fp32
// CUDA-core kernels.  Reference: torch MultiheadAttention math inside src/models/dare.py:53-70 (q scaled by
// 1/sqrt(hd), additive Toeplitz bias from PositionalBias dare.py:29-37, dropout on the probabilities).
//
// head_dim = 8 is exactly the k extent of m16n8k8, so Q.K^T needs one MMA per 16 x 8 score tile and P.V one per
// 8 keys.  (tcgen05.mma kind::tf32 also has k = 8, so the contraction would map; what rules it out is that the MMAs are
// 2 % of this kernel's issue slots -- the exp / mask / dropout work per score is the bound, on either tensor path.)
// SUPERSEDED as the default of the `amp: bf16` path by the packed-half kernels of attn_h16.cu (1.6-1.9x faster: the per-score
// work runs on half2 pairs there); kept selectable with B200CTR_ATTN=tf32 for A/B runs and as a second implementation the
// tests hold to the same fp32 reference.
//
// One CTA per sample; q (pre-scaled), k, v [, dO] of all heads staged once in shared memory as TF32 with a row
// stride of 68 words, which makes every fragment load bank-conflict free.  A warp owns (head, 16-query tile):
//   forward : S = QK^T for all keys stays in registers -> exact two-pass softmax -> dropout -> O = P.V
//             (the score tile's accumulator layout is fed back as the A operand by permuting the key index:
//              k-slot t <-> key 2t, k-slot t+4 <-> key 2t+1, with V rows read in the same order; no shuffles)
//   backward: warps = (head, key part); one pass over the scores: recompute P, dP = dO.V^T, dS -> dQ += dS.K, dbias;
//             P and dS tiles are transposed in registers (movmatrix on bf16 pairs) for dV += P^T.dO, dK += dS^T.Q;
//             q / dO fragments come from TF32 copies in shared memory (K <= 80), dQ is summed over the key parts
//             through a shared-memory hand-off, dbias in per-warp copies (no shared-memory atomics: they are CAS loops)
// Probabilities are never stored (only the per-row log-sum-exp, and -- with dropout -- one keep bit per probability
// so that the backward does not re-hash the mask); dqkv can be emitted as bf16 for the fused in-projection backward.
#include "common.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace b200ctr {

constexpr int TA_D = 64;
constexpr int TA_HD = 8;
constexpr int TA_H = 8;
constexpr int TA_RS = 68;          // shared-memory row stride in 32-bit words
constexpr int TA_THREADS = 256;

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return u;
}
// Hot-loop variant for FINITE values (probabilities, dS): cvt.rna.tf32 compiles to a compare against +inf plus a predicated
// add of 0x1000 per value (2 instructions + a constant move per tile, ~8 % of the forward's issue slots); the plain integer
// add is the same round-to-nearest on the magnitude for every finite input.
__device__ __forceinline__ uint32_t to_tf32_finite(float x) { return __float_as_uint(x) + 0x1000u; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 2^x for x <= 0: the scores are kept in log2 units (q and the bias carry a log2(e) factor), so a softmax numerator is
// one MUFU with no multiply in front of it
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float TA_LOG2E = 1.4426950408889634f, TA_LN2 = 0.6931471805599453f;
// keep-bit words of one (sample, head, 16-query tile): word w, lane l holds the 4 keep flags (rows g / g+8, key pair t)
// of the 8-key tiles 8w .. 8w+7 -- written by the forward pass, read by the backward pass instead of re-hashing
__device__ __forceinline__ int64_t ta_keep_index(int64_t b, int h, int mt, int i, int nw, int w, int lane) {
  return ((((b * TA_H + h) * mt + i) * nw + w) << 5) + lane;
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// stage one [K x 64] fp32 slice (row stride ld) as TF32 into smem [Kp][68]; rows >= K are zero
__device__ __forceinline__ void ta_stage(uint32_t* dst, const float* __restrict__ src, int64_t ld, int K, int Kp, float scale) {
  for (int i = threadIdx.x; i < Kp * 16; i += TA_THREADS) {
    const int r = i >> 4, c = i & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < K) v = *reinterpret_cast<const float4*>(src + (int64_t)r * ld + c * 4);
    *reinterpret_cast<uint4*>(dst + r * TA_RS + c * 4) =
        make_uint4(to_tf32(v.x * scale), to_tf32(v.y * scale), to_tf32(v.z * scale), to_tf32(v.w * scale));
  }
}

// A fragment (16 rows x 8 head dims) straight from global memory: each fragment is used by exactly one warp
__device__ __forceinline__ void ta_load_afrag(uint32_t (&a)[4], const float* __restrict__ base, int64_t ld, int ra, int rb, int K,
                                              int col, float scale) {
  const float x0 = ra < K ? __ldg(base + (int64_t)ra * ld + col) : 0.f, x1 = rb < K ? __ldg(base + (int64_t)rb * ld + col) : 0.f;
  const float x2 = ra < K ? __ldg(base + (int64_t)ra * ld + col + 4) : 0.f, x3 = rb < K ? __ldg(base + (int64_t)rb * ld + col + 4) : 0.f;
  a[0] = to_tf32(x0 * scale); a[1] = to_tf32(x1 * scale); a[2] = to_tf32(x2 * scale); a[3] = to_tf32(x3 * scale);
}

// NT = number of 8-key tiles (keys padded to a multiple of 16); Kp = 8 * NT
template <int NT>
__global__ void __launch_bounds__(TA_THREADS, (NT <= 8 ? 4 : 2))
attn_tc_fwd_kernel(const float* __restrict__ qkv, const float* __restrict__ relmean, int max_len, int K,
                   float* __restrict__ o, float* __restrict__ lse, uint32_t* __restrict__ keep, b200ctr_dropout_t drop) {
  constexpr int Kp = NT * 8, MT = NT / 2, NW = (NT + 7) / 8;
  extern __shared__ __align__(16) uint32_t sm_u[];
  uint32_t* ks = sm_u;
  uint32_t* vs = ks + Kp * TA_RS;
  float* bias = reinterpret_cast<float*>(vs + Kp * TA_RS);         // 2K-1 : bias[(j-i)+K-1]
  const int64_t b = blockIdx.x;
  const float* base = qkv + b * (int64_t)K * 3 * TA_D;
  ta_stage(ks, base + TA_D, 3 * TA_D, K, Kp, 1.f);
  ta_stage(vs, base + 2 * TA_D, 3 * TA_D, K, Kp, 1.f);
  for (int t = threadIdx.x; t < 2 * K - 1; t += TA_THREADS) {
    const int d = t - (K - 1);
    bias[t] = relmean ? relmean[min(max(d, -max_len), max_len) + max_len] * TA_LOG2E : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const DropCtx dctx = drop_ctx(drop);
  for (int item = warp; item < TA_H * MT; item += TA_THREADS / 32) {
    const int h = item / MT, i = item % MT;
    const int qa = 16 * i + g, qb = qa + 8;
    uint32_t a[4];
    ta_load_afrag(a, base, 3 * TA_D, qa, qb, K, h * 8 + t, rsqrtf((float)TA_HD) * TA_LOG2E);   // scores in log2 units
    float s[NT][4];
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
      mma_tf32(s[j], a, ks[(8 * j + g) * TA_RS + h * 8 + t], ks[(8 * j + g) * TA_RS + h * 8 + t + 4]);
    }
    const float* ba = bias + (K - 1) - min(qa, K - 1);
    const float* bb = bias + (K - 1) - min(qb, K - 1);
    float ma = -INFINITY, mb = -INFINITY;
    float pb0 = 0.f, pb1 = 0.f;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      const int kc = 8 * j + 2 * t;
      // rows g+8 see the bias of rows g shifted by one 8-key tile: carry the previous tile's pair instead of reloading
      const float a0 = ba[kc], a1 = ba[kc + 1];
      if (j == 0) { pb0 = bb[kc]; pb1 = bb[kc + 1]; }
      if (NT <= 10 && 8 * j + 8 <= K) {          // warp-uniform: whole tile valid, no per-score mask (small kernels only: code size)
        s[j][0] += a0; s[j][1] += a1; s[j][2] += pb0; s[j][3] += pb1;
      } else {
        s[j][0] = kc < K ? s[j][0] + a0 : -INFINITY;
        s[j][1] = kc + 1 < K ? s[j][1] + a1 : -INFINITY;
        s[j][2] = kc < K ? s[j][2] + pb0 : -INFINITY;
        s[j][3] = kc + 1 < K ? s[j][3] + pb1 : -INFINITY;
      }
      pb0 = a0; pb1 = a1;
      ma = fmaxf(ma, fmaxf(s[j][0], s[j][1]));
      mb = fmaxf(mb, fmaxf(s[j][2], s[j][3]));
    }
    ma = quad_max(ma); mb = quad_max(mb);
    float la = 0.f, lb = 0.f;
    float oc[4] = {0.f, 0.f, 0.f, 0.f};
    // dropout: row id = (b, h, query), column = key
    const uint32_t keya = dctx.on ? drop_row_key(drop, ((uint64_t)b * TA_H + h) * K + qa) : 0u;
    const uint32_t keyb = dctx.on ? drop_row_key(drop, ((uint64_t)b * TA_H + h) * K + qb) : 0u;
    uint32_t kw[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) kw[w] = 0u;
#pragma unroll
    for (int j = 0; j < NT; ++j) {
      float p0 = fast_exp2(s[j][0] - ma), p1 = fast_exp2(s[j][1] - ma), p2 = fast_exp2(s[j][2] - mb), p3 = fast_exp2(s[j][3] - mb);
      la += p0 + p1; lb += p2 + p3;
      if (dctx.on) {
        const uint32_t ha = drop_pair_hash(keya, (uint32_t)(4 * j + t)), hb = drop_pair_hash(keyb, (uint32_t)(4 * j + t));
        const bool k0 = drop_keep_lo(ha, dctx.thr), k1 = drop_keep_hi(ha, dctx.thr);
        const bool k2 = drop_keep_lo(hb, dctx.thr), k3 = drop_keep_hi(hb, dctx.thr);
        p0 = k0 ? p0 : 0.f; p1 = k1 ? p1 : 0.f;           // the 1/(1-p) scale is applied once, to the output row
        p2 = k2 ? p2 : 0.f; p3 = k3 ? p3 : 0.f;
        const uint32_t sh = 4 * (j & 7);
        kw[j >> 3] |= (k0 ? 1u << sh : 0u) | (k1 ? 2u << sh : 0u) | (k2 ? 4u << sh : 0u) | (k3 ? 8u << sh : 0u);
      }
      const uint32_t pa[4] = {to_tf32_finite(p0), to_tf32_finite(p2), to_tf32_finite(p1), to_tf32_finite(p3)};
      mma_tf32(oc, pa, vs[(8 * j + 2 * t) * TA_RS + h * 8 + g], vs[(8 * j + 2 * t + 1) * TA_RS + h * 8 + g]);
    }
    la = quad_sum(la); lb = quad_sum(lb);
    if (keep && dctx.on) {
#pragma unroll
      for (int w = 0; w < NW; ++w) keep[ta_keep_index(b, h, MT, i, NW, w, lane)] = kw[w];
    }
    if (qa < K) {
      const float inv = dctx.inv / la;
      *reinterpret_cast<float2*>(o + (b * K + qa) * (int64_t)TA_D + h * 8 + 2 * t) = make_float2(oc[0] * inv, oc[1] * inv);
      if (t == 0) lse[(b * TA_H + h) * (int64_t)K + qa] = (ma + log2f(la)) * TA_LN2;       // natural-log units
    }
    if (qb < K) {
      const float inv = dctx.inv / lb;
      *reinterpret_cast<float2*>(o + (b * K + qb) * (int64_t)TA_D + h * 8 + 2 * t) = make_float2(oc[2] * inv, oc[3] * inv);
      if (t == 0) lse[(b * TA_H + h) * (int64_t)K + qb] = (mb + log2f(lb)) * TA_LN2;
    }
  }
}

// ---- backward: ONE pass over the scores, one warp per head ----------------------------------------------------
// For every (16-query tile i, 16-key block jb) the warp recomputes P and dS once (TF32 MMAs for S = QK^T and
// dP = dO.V^T), accumulates dQ += dS.K (TF32), and then needs the TRANSPOSED tiles for dV += P^T.dO and
// dK += dS^T.Q: the accumulator fragments are packed to bf16 pairs and transposed in registers with
// movmatrix (8x8 b16), which yields exactly the A fragment of mma.m16n8k16 (16 keys x 16 queries); its B fragment
// (queries x head dims) is read from bf16 copies of Q^T and dO^T in shared memory.  dK / dV for all keys of the
// head live in the warp's registers; no atomics anywhere in the kernel except the final dbias -> drelmean reduction.
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t movmatrix_t(uint32_t x) {
  uint32_t r;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// two adjacent gradient elements at element offset `off` of the (M, 192) tensor: fp32 pair or one packed bf16 pair
__device__ __forceinline__ void ta_store2(float* base, int as_bf16, int64_t off, float a, float b) {
  if (as_bf16) reinterpret_cast<uint32_t*>(base)[off >> 1] = pack_bf16x2(a, b);
  else *reinterpret_cast<float2*>(base + off) = make_float2(a, b);
}

// KS = key split: the warps (head h, part) with part in [0, KS) own disjoint 16-key blocks of head h, so dK / dV / dbias
// accumulators stay small and an SM holds KS x more warps than with one warp per head (KS = 2 everywhere: 512 threads at
// up to 128 registers; 4 parts = 1024 threads capped the kernel at 64 registers and spilled); only dQ
// needs a cross-warp sum.
// KEEP: the dropout keep flags come from the words the forward pass stored (2 instructions per score instead of a hash)
template <int NT, int KS, bool KEEP>
// One CTA per SM: ~120 registers, no spills (two CTAs at 64 registers spilled the dK / dV / dbias accumulators: 2.28 vs
// 1.83 ms per step).  dQ is summed over the key parts through a small shared-memory hand-off (one named barrier per head).
__global__ void __launch_bounds__(TA_THREADS * KS, 1)
attn_tc_bwd_kernel(const float* __restrict__ qkv, const float* __restrict__ relmean, int max_len, const float* __restrict__ o,
                   const float* __restrict__ lse, const float* __restrict__ d_o, int K, float* __restrict__ dqkv, int dq16,
                   float* __restrict__ drelmean, const uint32_t* __restrict__ keep, b200ctr_dropout_t drop) {
  // dq16: dqkv is written as bf16 (same (M, 192) layout) -- its only consumer, the fused in-projection backward, rounds
  // it to bf16 for its tensor-core products anyway, so the results are identical and the tensor costs half the HBM traffic
  constexpr int Kp = NT * 8, MT = NT / 2, NB = NT / 2, NW = (NT + 7) / 8;
  constexpr bool SQD = NT <= 10;      // q / dO also staged as TF32 (A fragments from shared memory); does not fit for long sequences
  constexpr int NTH = TA_THREADS * KS;
  constexpr int NBW = (NB + KS - 1) / KS;                     // 16-key blocks per warp
  constexpr int NTW = 2 * NBW;                                 // 8-key tiles per warp
  constexpr int KPA = 16 * NBW * KS;                           // padded key extent covered by the dbias slots (>= Kp)
  constexpr int TS = Kp + 8;                                   // bf16 row stride of the transposed copies
  extern __shared__ __align__(16) uint32_t sm_u[];
  uint32_t* ks = sm_u;
  uint32_t* vs = ks + Kp * TA_RS;
  uint32_t* qs = vs + Kp * TA_RS;                               // q * scale (TF32): A fragments of S = Q.K^T   (SQD only)
  uint32_t* dos = qs + (SQD ? Kp * TA_RS : 0);                  // dO (TF32): A fragments of dP = dO.V^T         (SQD only)
  __nv_bfloat16* qT = reinterpret_cast<__nv_bfloat16*>(dos + (SQD ? Kp * TA_RS : 0));     // [64][TS]  (q * scale)^T
  __nv_bfloat16* doT = qT + 64 * TS;                                          // [64][TS]  dO^T
  float* lses = reinterpret_cast<float*>(doT + 64 * TS);        // [H][Kp]  (+inf on padded rows -> P = 0), log2 units
  float* dsum = lses + TA_H * Kp;                               // [H][Kp]
  float* bias = dsum + TA_H * Kp;                               // 2K-1 (log2 units)
  // dbias[(key - query) + KPA], one private copy per warp: shared-memory float AND 64-bit integer atomic adds are
  // CAS loops (ATOMS.CAST.SPIN), so each warp sums the lanes that share a diagonal with two shuffles and does a plain
  // read-modify-write on its own copy; the copies are added in warp order at the end (deterministic)
  constexpr int DBS = 2 * KPA + 2;
  float* dbias = bias + 2 * K;                                  // [8 * KS][DBS]
  // dQ hand-off between the key parts of a head: parts 1.. write their 16 x 8 tile, part 0 adds and stores the rows
  float* dqx = dbias + 8 * KS * DBS;                            // [H][2 (query-tile parity)][KS - 1][128]
  const int64_t b = blockIdx.x;
  const float* base = qkv + b * (int64_t)K * 3 * TA_D;
  const float* dob = d_o + b * (int64_t)K * TA_D;
  const float scale = rsqrtf((float)TA_HD);
  // Everything the main loop reads comes from shared memory: the loop was bound by the L1 data pipe, a third of it
  // fragment loads from global memory (8 rows = 8 tag lookups per LDG) and another tenth the dQ reductions.
  for (int i = threadIdx.x; i < Kp * 16; i += NTH) {
    const int r = i >> 4, c = i & 15;
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), kv = qv, vv = qv, dv = qv;
    if (r < K) {
      const float4* row = reinterpret_cast<const float4*>(base + (int64_t)r * 3 * TA_D);
      kv = row[16 + c]; vv = row[32 + c];
      if (SQD) { qv = row[c]; dv = *reinterpret_cast<const float4*>(dob + (int64_t)r * TA_D + c * 4); }
    }
    *reinterpret_cast<uint4*>(ks + r * TA_RS + c * 4) = make_uint4(to_tf32(kv.x), to_tf32(kv.y), to_tf32(kv.z), to_tf32(kv.w));
    *reinterpret_cast<uint4*>(vs + r * TA_RS + c * 4) = make_uint4(to_tf32(vv.x), to_tf32(vv.y), to_tf32(vv.z), to_tf32(vv.w));
    if (SQD) {
      *reinterpret_cast<uint4*>(qs + r * TA_RS + c * 4) = make_uint4(to_tf32(qv.x * scale), to_tf32(qv.y * scale), to_tf32(qv.z * scale), to_tf32(qv.w * scale));
      *reinterpret_cast<uint4*>(dos + r * TA_RS + c * 4) = make_uint4(to_tf32(dv.x), to_tf32(dv.y), to_tf32(dv.z), to_tf32(dv.w));
    }
  }
  if (!SQD) {
    constexpr int TOT = Kp * 64, PER = (TOT + NTH - 1) / NTH;   // coalesced global reads (all in flight), transposed smem writes
    float qv[PER], dv[PER];
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int i = threadIdx.x + u * NTH, r = i >> 6, c = i & 63;
      qv[u] = 0.f; dv[u] = 0.f;
      if (i < TOT && r < K) { qv[u] = base[(int64_t)r * 3 * TA_D + c]; dv[u] = dob[(int64_t)r * TA_D + c]; }
    }
#pragma unroll
    for (int u = 0; u < PER; ++u) {
      const int i = threadIdx.x + u * NTH, r = i >> 6, c = i & 63;
      if (i < TOT) { qT[c * TS + r] = __float2bfloat16(qv[u] * scale); doT[c * TS + r] = __float2bfloat16(dv[u]); }
    }
  }
  for (int tt = threadIdx.x; tt < 2 * K - 1; tt += NTH) {
    const int d = tt - (K - 1);
    bias[tt] = relmean ? relmean[min(max(d, -max_len), max_len) + max_len] * TA_LOG2E : 0.f;
  }
  if (drelmean)
    for (int tt = threadIdx.x; tt < 8 * KS * DBS; tt += NTH) dbias[tt] = 0.f;
  for (int item = threadIdx.x; item < Kp * TA_H; item += NTH) {
    const int h = item / Kp, i = item % Kp;
    float L = INFINITY, dsv = 0.f;
    if (i < K) {
      L = lse[(b * TA_H + h) * (int64_t)K + i] * TA_LOG2E;
      const float4* op = reinterpret_cast<const float4*>(o + (b * K + i) * (int64_t)TA_D + h * TA_HD);
      const float4* dp = reinterpret_cast<const float4*>(dob + (int64_t)i * TA_D + h * TA_HD);
      const float4 o0 = op[0], o1 = op[1], d0 = dp[0], d1 = dp[1];
      dsv = o0.x * d0.x + o0.y * d0.y + o0.z * d0.z + o0.w * d0.w + o1.x * d1.x + o1.y * d1.y + o1.z * d1.z + o1.w * d1.w;
    }
    lses[item] = L; dsum[item] = dsv;
  }
  __syncthreads();
  // bf16 transposes of the staged q / dO tiles (B operands of dK += dS^T.Q, dV += P^T.dO): a thread packs rows 2rp, 2rp+1
  // of column c; lanes = 8 columns x 4 row pairs, which is bank-conflict free on both the reads and the writes
  for (int u = threadIdx.x; SQD && u < (Kp / 2) * 64; u += NTH) {
    const int lane_ = u & 31, blk = u >> 5;                     // blk -> (column group of 8, row-pair group of 4)
    const int c = (blk & 7) * 8 + (lane_ & 7), rp = (blk >> 3) * 4 + (lane_ >> 3);
    const uint32_t q0 = qs[(2 * rp) * TA_RS + c], q1 = qs[(2 * rp + 1) * TA_RS + c];
    const uint32_t d0 = dos[(2 * rp) * TA_RS + c], d1 = dos[(2 * rp + 1) * TA_RS + c];
    reinterpret_cast<uint32_t*>(qT + c * TS)[rp] = pack_bf16x2(__uint_as_float(q0), __uint_as_float(q1));
    reinterpret_cast<uint32_t*>(doT + c * TS)[rp] = pack_bf16x2(__uint_as_float(d0), __uint_as_float(d1));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int h = warp & 7, part = warp >> 3;                     // (head, key part)
  const int jb0 = part * NBW;                                   // first 16-key block of this warp
  const DropCtx dctx = drop_ctx(drop);
  float dk[NBW][4], dv[NBW][4];
#pragma unroll
  for (int jl = 0; jl < NBW; ++jl) { dk[jl][0] = dk[jl][1] = dk[jl][2] = dk[jl][3] = 0.f; dv[jl][0] = dv[jl][1] = dv[jl][2] = dv[jl][3] = 0.f; }
  // dbias: a thread's four dS values of tile (i, j) lie on the diagonals 8(j-2i) + (2t-g) {+0, +1} (rows g) and the same
  // shifted by -8 (rows g+8), so they accumulate in registers in slots indexed by the local tile (+1) that rotate by two
  // per query tile; only the two slots that leave the window are flushed (shared-memory atomics) per query tile.
  float A0[NTW + 1], A1[NTW + 1];
#pragma unroll
  for (int sl = 0; sl <= NTW; ++sl) { A0[sl] = 0.f; A1[sl] = 0.f; }
  float* dbt = dbias + warp * DBS + KPA + 2 * t - g + 8 * (2 * jb0);    // + 8 * (local slot - 1 - 2i)
  // lanes (g, t), (g+2, t+1), (g+4, t+2), (g+6, t+3) = lane, lane+9, lane+18, lane+27 lie on the same diagonal 2t-g:
  // two shuffle steps leave the chain sum in its first lane (t == 0 or g < 2); those 14 lanes hit 14 distinct words
  const bool ch1 = t < 3 && g < 6, ch2 = t < 2 && g < 4, chead = t == 0 || g < 2;
  auto dbias_add = [&](float* p, float v) {
    float x = __shfl_down_sync(0xffffffffu, v, 9);
    if (ch1) v += x;
    x = __shfl_down_sync(0xffffffffu, v, 18);
    if (ch2) v += x;
    if (chead) *p += v;
    __syncwarp();
  };
  const uint32_t* qT32 = reinterpret_cast<const uint32_t*>(qT + (h * 8 + g) * TS);
  const uint32_t* doT32 = reinterpret_cast<const uint32_t*>(doT + (h * 8 + g) * TS);
#pragma unroll 1
  for (int i = 0; i < MT; ++i) {
    const int qa = 16 * i + g, qb = qa + 8;
    uint32_t aq[4], ad[4];
    if (SQD) {
      const int oa = qa * TA_RS + h * 8 + t, ob = qb * TA_RS + h * 8 + t;      // rows >= K are zero in shared memory
      aq[0] = qs[oa]; aq[1] = qs[ob]; aq[2] = qs[oa + 4]; aq[3] = qs[ob + 4];
      ad[0] = dos[oa]; ad[1] = dos[ob]; ad[2] = dos[oa + 4]; ad[3] = dos[ob + 4];
    } else {
      ta_load_afrag(aq, base, 3 * TA_D, qa, qb, K, h * 8 + t, scale);
      ta_load_afrag(ad, dob, TA_D, qa, qb, K, h * 8 + t, 1.f);
    }
    uint32_t kw[NW];
    if (KEEP) {
#pragma unroll
      for (int w = 0; w < NW; ++w) kw[w] = __ldg(keep + ta_keep_index(b, h, MT, i, NW, w, lane));
    }
    // B fragments (k = 16 queries of this tile, n = head dim g) of the bf16 m16n8k16 MMAs
    const uint32_t bq0 = qT32[8 * i + t], bq1 = qT32[8 * i + t + 4], bd0 = doT32[8 * i + t], bd1 = doT32[8 * i + t + 4];
    const float La = lses[h * Kp + qa], Lb = lses[h * Kp + qb], Da = dsum[h * Kp + qa], Db = dsum[h * Kp + qb];
    const int qac = min(qa, K - 1), qbc = min(qb, K - 1);
    const float* ba = bias + (K - 1) - qac;
    const float* bb = bias + (K - 1) - qbc;
    const uint32_t keya = (!KEEP && dctx.on) ? drop_row_key(drop, ((uint64_t)b * TA_H + h) * K + qa) : 0u;
    const uint32_t keyb = (!KEEP && dctx.on) ? drop_row_key(drop, ((uint64_t)b * TA_H + h) * K + qb) : 0u;
    float dq[4] = {0.f, 0.f, 0.f, 0.f};
    // rows g+8 see the bias of rows g shifted by one 8-key tile: carry the previous tile's pair instead of reloading
    float pb0 = bb[16 * jb0 + 2 * t], pb1 = bb[16 * jb0 + 2 * t + 1];
#pragma unroll
    for (int jl = 0; jl < NBW; ++jl) {
      if (jb0 + jl < NB) {                                      // warp-uniform
        uint32_t pd_lo[2], pd_hi[2], ds_lo[2], ds_hi[2];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int jt = 2 * jl + hf;                           // local tile
          const int j = 2 * jb0 + jt;                           // global tile
          const int kc = 8 * j + 2 * t;
          const uint32_t kb0 = ks[(8 * j + g) * TA_RS + h * 8 + t], kb1 = ks[(8 * j + g) * TA_RS + h * 8 + t + 4];
          const uint32_t vb0 = vs[(8 * j + g) * TA_RS + h * 8 + t], vb1 = vs[(8 * j + g) * TA_RS + h * 8 + t + 4];
          float s[4] = {0.f, 0.f, 0.f, 0.f}, dp[4] = {0.f, 0.f, 0.f, 0.f};
          mma_tf32(s, aq, kb0, kb1);
          mma_tf32(dp, ad, vb0, vb1);
          const float a0 = ba[kc], a1 = ba[kc + 1];
          float p0, p1, p2, p3;
          if (NT <= 10 && 8 * j + 8 <= K) {       // warp-uniform: whole tile valid (small kernels only: code size)
            p0 = fast_exp2(fmaf(s[0], TA_LOG2E, a0 - La)); p1 = fast_exp2(fmaf(s[1], TA_LOG2E, a1 - La));
            p2 = fast_exp2(fmaf(s[2], TA_LOG2E, pb0 - Lb)); p3 = fast_exp2(fmaf(s[3], TA_LOG2E, pb1 - Lb));
          } else {
            const bool v0 = kc < K, v1 = kc + 1 < K;
            p0 = v0 ? fast_exp2(fmaf(s[0], TA_LOG2E, a0 - La)) : 0.f; p1 = v1 ? fast_exp2(fmaf(s[1], TA_LOG2E, a1 - La)) : 0.f;
            p2 = v0 ? fast_exp2(fmaf(s[2], TA_LOG2E, pb0 - Lb)) : 0.f; p3 = v1 ? fast_exp2(fmaf(s[3], TA_LOG2E, pb1 - Lb)) : 0.f;
          }
          pb0 = a0; pb1 = a1;
          float c0 = 1.f, c1 = 1.f, c2 = 1.f, c3 = 1.f;
          if (KEEP) {
            const int w = j >> 3;
            const uint32_t word = NW == 1 ? kw[0] : (w == 0 ? kw[0] : (w == 1 ? kw[NW > 1 ? 1 : 0] : kw[NW > 2 ? 2 : 0]));
            const uint32_t nib = word >> (4 * (j & 7));
            c0 = (nib & 1u) ? dctx.inv : 0.f; c1 = (nib & 2u) ? dctx.inv : 0.f;
            c2 = (nib & 4u) ? dctx.inv : 0.f; c3 = (nib & 8u) ? dctx.inv : 0.f;
          } else if (dctx.on) {
            drop_pair(dctx, keya, (uint32_t)(4 * j + t), c0, c1);
            drop_pair(dctx, keyb, (uint32_t)(4 * j + t), c2, c3);
          }
          const float ds0 = p0 * (dp[0] * c0 - Da), ds1 = p1 * (dp[1] * c1 - Da);
          const float ds2 = p2 * (dp[2] * c2 - Db), ds3 = p3 * (dp[3] * c3 - Db);
          A0[jt + 1] += ds0; A1[jt + 1] += ds1; A0[jt] += ds2; A1[jt] += ds3;    // padded rows / keys contribute exact zeros
          const uint32_t pa[4] = {to_tf32_finite(ds0), to_tf32_finite(ds2), to_tf32_finite(ds1), to_tf32_finite(ds3)};
          mma_tf32(dq, pa, ks[(8 * j + 2 * t) * TA_RS + h * 8 + g], ks[(8 * j + 2 * t + 1) * TA_RS + h * 8 + g]);
          pd_lo[hf] = pack_bf16x2(p0 * c0, p1 * c1); pd_hi[hf] = pack_bf16x2(p2 * c2, p3 * c3);
          ds_lo[hf] = pack_bf16x2(ds0, ds1); ds_hi[hf] = pack_bf16x2(ds2, ds3);
        }
        // A = tile^T: rows = 16 keys of the block (even tile -> rows 0-7, odd tile -> rows 8-15), k = 16 queries
        const uint32_t apd[4] = {movmatrix_t(pd_lo[0]), movmatrix_t(pd_lo[1]), movmatrix_t(pd_hi[0]), movmatrix_t(pd_hi[1])};
        const uint32_t ads[4] = {movmatrix_t(ds_lo[0]), movmatrix_t(ds_lo[1]), movmatrix_t(ds_hi[0]), movmatrix_t(ds_hi[1])};
        mma_bf16_16816(dv[jl], apd, bd0, bd1);
        mma_bf16_16816(dk[jl], ads, bq0, bq1);
      }
    }
    {
      float2* xq = reinterpret_cast<float2*>(dqx + (size_t)((h * 2 + (i & 1)) * (KS - 1)) * 128);
      if (part != 0) {
        xq[(part - 1) * 64 + lane] = make_float2(dq[0], dq[1]);            // row g, dims 2t, 2t+1
        xq[(part - 1) * 64 + 32 + lane] = make_float2(dq[2], dq[3]);       // row g + 8
      }
      asm volatile("bar.sync %0, %1;" :: "r"(h + 1), "r"(32 * KS) : "memory");   // the KS warps of head h
      if (part == 0) {
#pragma unroll
        for (int pp = 0; pp < KS - 1; ++pp) {
          const float2 x0 = xq[pp * 64 + lane], x1 = xq[pp * 64 + 32 + lane];
          dq[0] += x0.x; dq[1] += x0.y; dq[2] += x1.x; dq[3] += x1.y;
        }
        if (qa < K) ta_store2(dqkv, dq16, (b * K + qa) * (int64_t)(3 * TA_D) + h * 8 + 2 * t, dq[0] * scale, dq[1] * scale);
        if (qb < K) ta_store2(dqkv, dq16, (b * K + qb) * (int64_t)(3 * TA_D) + h * 8 + 2 * t, dq[2] * scale, dq[3] * scale);
      }
    }
    if (drelmean) {
      // local slot sl holds the diagonal block u = 2*jb0 + sl - 1 - 2i
      if (i + 1 < MT) {
        dbias_add(dbt + 8 * (NTW - 1 - 2 * i), A0[NTW]); dbias_add(dbt + 8 * (NTW - 1 - 2 * i) + 1, A1[NTW]);
        dbias_add(dbt + 8 * (NTW - 2 - 2 * i), A0[NTW - 1]); dbias_add(dbt + 8 * (NTW - 2 - 2 * i) + 1, A1[NTW - 1]);
#pragma unroll
        for (int sl = NTW; sl >= 2; --sl) { A0[sl] = A0[sl - 2]; A1[sl] = A1[sl - 2]; }
        A0[0] = A0[1] = A1[0] = A1[1] = 0.f;
      } else {
#pragma unroll
        for (int sl = 0; sl <= NTW; ++sl) {
          dbias_add(dbt + 8 * (sl - 1 - 2 * i), A0[sl]); dbias_add(dbt + 8 * (sl - 1 - 2 * i) + 1, A1[sl]);
        }
      }
    }
  }
#pragma unroll
  for (int jl = 0; jl < NBW; ++jl) {
    const int ka = 16 * (jb0 + jl) + g, kb = ka + 8;
    if (ka < K) {
      ta_store2(dqkv, dq16, (b * K + ka) * (int64_t)(3 * TA_D) + TA_D + h * 8 + 2 * t, dk[jl][0], dk[jl][1]);
      ta_store2(dqkv, dq16, (b * K + ka) * (int64_t)(3 * TA_D) + 2 * TA_D + h * 8 + 2 * t, dv[jl][0], dv[jl][1]);
    }
    if (kb < K) {
      ta_store2(dqkv, dq16, (b * K + kb) * (int64_t)(3 * TA_D) + TA_D + h * 8 + 2 * t, dk[jl][2], dk[jl][3]);
      ta_store2(dqkv, dq16, (b * K + kb) * (int64_t)(3 * TA_D) + 2 * TA_D + h * 8 + 2 * t, dv[jl][2], dv[jl][3]);
    }
  }
  if (drelmean) {
    __syncthreads();
    for (int tt = threadIdx.x; tt < 2 * K - 1; tt += NTH) {
      const int d = tt - (K - 1);
      float acc = 0.f;
      for (int w = 0; w < 8 * KS; ++w) acc += dbias[w * DBS + d + KPA];
      // per-sample partial row (zero-initialised by the caller; |d| <= K - 1 < max_len: one writer per slot), summed in sample
      // order afterwards (b200ctr_colsum): no float atomics
      drelmean[b * (int64_t)(2 * max_len + 1) + min(max(d, -max_len), max_len) + max_len] = acc;
    }
  }
}

static int ta_set_smem(const void* fn, size_t bytes) {
  if (bytes > 227 * 1024) {
    set_error("attention (tensor-core path): K too large for one CTA's shared memory (%zu bytes)", bytes);
    return B200CTR_E_UNSUPPORTED_SHAPE;
  }
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_tc smem)");
  }
  return B200CTR_OK;
}

template <int NT>
static int ta_launch_fwd(const float* qkv, const float* relmean, int max_len, int64_t B, int K, float* o, float* lse,
                         uint32_t* keep, b200ctr_dropout_t drop, cudaStream_t st) {
  const size_t smem = ((size_t)2 * NT * 8 * TA_RS + K + NT * 8 + 8) * 4;   // bias: masked lanes read (unused) entries up to K - 1 + Kp
  int rc = ta_set_smem((const void*)attn_tc_fwd_kernel<NT>, smem);
  if (rc) return rc;
  attn_tc_fwd_kernel<NT><<<(unsigned)B, TA_THREADS, smem, st>>>(qkv, relmean, max_len, K, o, lse, keep, drop);
  return B200CTR_OK;
}
template <int NT>
static int ta_launch_bwd(const float* qkv, const float* relmean, int max_len, const float* o, const float* lse, const float* d_o,
                         int64_t B, int K, float* dqkv, int dq16, float* drelmean, const uint32_t* keep, b200ctr_dropout_t drop,
                         cudaStream_t st) {
  constexpr int KS = 2;      // 2 key parts: 512 threads at up to 128 registers (4 parts = 1024 threads at 64 registers spilled)
  constexpr int NBW = (NT / 2 + KS - 1) / KS, KPA = 16 * NBW * KS, Kp = NT * 8;
  const size_t smem = ((size_t)(NT <= 10 ? 4 : 2) * Kp * TA_RS + 2 * TA_H * Kp + 2 * K) * 4 + (size_t)(8 * KS) * (2 * KPA + 2) * 4 +
                      (size_t)TA_H * 2 * (KS - 1) * 128 * 4 + (size_t)2 * 64 * (Kp + 8) * 2;
  const bool use_keep = keep != nullptr && drop.p > 0.f;
  int rc = use_keep ? ta_set_smem((const void*)attn_tc_bwd_kernel<NT, KS, true>, smem)
                    : ta_set_smem((const void*)attn_tc_bwd_kernel<NT, KS, false>, smem);
  if (rc) return rc;
  if (use_keep)
    attn_tc_bwd_kernel<NT, KS, true><<<(unsigned)B, TA_THREADS * KS, smem, st>>>(qkv, relmean, max_len, o, lse, d_o, K, dqkv, dq16,
                                                                                 drelmean, keep, drop);
  else
    attn_tc_bwd_kernel<NT, KS, false><<<(unsigned)B, TA_THREADS * KS, smem, st>>>(qkv, relmean, max_len, o, lse, d_o, K, dqkv, dq16,
                                                                                  drelmean, nullptr, drop);
  return B200CTR_OK;
}

}  // namespace b200ctr

using namespace b200ctr;

#define TA_DISPATCH(CALL)                                                            \
  switch ((K + 15) / 16) {                                                           \
    case 1: rc = CALL(2); break;   case 2: rc = CALL(4); break;   case 3: rc = CALL(6); break;   \
    case 4: rc = CALL(8); break;   case 5: rc = CALL(10); break;  case 6: rc = CALL(12); break;  \
    case 7: rc = CALL(14); break;  case 8: rc = CALL(16); break;  case 9: rc = CALL(18); break;  \
    case 10: rc = CALL(20); break;                                                   \
    default:                                                                         \
      set_error("attention (tensor-core path): top_k=%d > 160 is not built", K);     \
      return B200CTR_E_UNSUPPORTED_SHAPE;                                            \
  }

extern "C" int64_t b200ctr_attn_tc_keep_words(int64_t B, int K) {
  const int64_t nt = 2 * (((int64_t)K + 15) / 16);
  return B * TA_H * (nt / 2) * ((nt + 7) / 8) * 32;
}

extern "C" int b200ctr_attn_tc_fwd(const float* qkv, const float* relmean, int max_len, int64_t B, int K, int H, int D, float* o,
                                   float* lse, uint32_t* keep, b200ctr_dropout_t drop, void* stream) {
  B200_REQUIRE(qkv && o && lse, B200CTR_E_INVALID_ARG, "attn_tc_fwd: null pointer");
  B200_REQUIRE(D == TA_D && H == TA_H, B200CTR_E_UNSUPPORTED_SHAPE,
               "attn_tc_fwd: only emb_dim=64 with 8 heads of head_dim=8 is built (got D=%d, H=%d)", D, H);
  if (B == 0 || K == 0) return B200CTR_OK;
  int rc;
#define CALL(NT_) ta_launch_fwd<NT_>(qkv, relmean, max_len, B, K, o, lse, keep, drop, as_stream(stream))
  TA_DISPATCH(CALL)
#undef CALL
  if (rc) return rc;
  B200_LAUNCH_CHECK("attn_tc_fwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_attn_tc_bwd(const float* qkv, const float* relmean, int max_len, const float* o, const float* lse,
                                   const float* d_o, int64_t B, int K, int H, int D, void* dqkv, int dqkv_bf16, float* drelmean,
                                   const uint32_t* keep, b200ctr_dropout_t drop, void* stream) {
  B200_REQUIRE(qkv && o && lse && d_o && dqkv, B200CTR_E_INVALID_ARG, "attn_tc_bwd: null pointer");
  B200_REQUIRE(D == TA_D && H == TA_H, B200CTR_E_UNSUPPORTED_SHAPE,
               "attn_tc_bwd: only emb_dim=64 with 8 heads of head_dim=8 is built (got D=%d, H=%d)", D, H);
  if (B == 0 || K == 0) return B200CTR_OK;
  int rc;
#define CALL(NT_) ta_launch_bwd<NT_>(qkv, relmean, max_len, o, lse, d_o, B, K, (float*)dqkv, dqkv_bf16, drelmean, keep, drop, as_stream(stream))
  TA_DISPATCH(CALL)
#undef CALL
  if (rc) return rc;
  B200_LAUNCH_CHECK("attn_tc_bwd_kernel");
  return B200CTR_OK;
}
