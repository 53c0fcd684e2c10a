// Tiny-head attention (head_dim 8) of the DARE encoder layer in PACKED HALF arithmetic -- the `amp: bf16` path.
// Reference: torch MultiheadAttention math inside src/models/dare.py:53-70 (q scaled by 1/sqrt(hd), additive Toeplitz bias
// from PositionalBias dare.py:29-37, dropout on the probabilities).
//
// Why this kernel exists: the TF32 kernels of attn_tc.cu spend 2 % of their issue slots on MMAs and the rest on fp32
// per-score work (bias add, max, exp, dropout compare / select / bit packing, TF32 rounding, bf16 packing): 23 (forward) and
// 31 (backward) thread instructions per score.  Under torch autocast the reference itself rounds the scores AND the
// probabilities to bf16 (8 bits); fp16 keeps 11, so everything between the two contractions can run on half2 registers --
// two scores per instruction -- without leaving the reference's own error band:
//   * S = Q.K^T + bias is ONE mma.m16n8k8 with an fp16 accumulator whose C operand is the packed bias pair (no add);
//     its D fragment (row g: keys 2t, 2t+1 | row g+8: same) IS the A fragment of the m16n8k16 products that follow, so
//     probabilities never get repacked, rounded or permuted;
//   * max / subtract / exp2 / mask are HMNMX2 / HADD2 / MUFU.EX2.F16 / LOP3 on pairs; the dropout decision of a pair is one
//     AND + one HSET2 on the hash word (common.cuh: drop_keep_mask2), the keep-bit word one LOP3;
//   * the softmax denominators come out of the tensor core (P . ones, fp32 accumulate);
//   * every operand fragment is fetched with ldmatrix from four TRANSPOSED fp16 images X^T[dim][position] (+ .trans where the
//     operand wants the other orientation): 2-3 shared-memory instructions per 16 x 16 score block instead of 12.
// Backward: P is recomputed as 2^(s - m') * f with m' = the row's log-sum-exp rounded to fp16 and the fp32 correction
// f = 2^(m' - lse) folded into the dO row; dO is also multiplied by 1/(1-p) and by a per-sample power of two that puts
// 8 max|dO| max|V| below 2^9 (gradients of a mean-reduced loss are ~1e-6: below fp16's range), undone on the way out.
// dS = e (dP' & keep - D'), dQ += dS.K straight from the registers, dK / dV through movmatrix transposes; all fp32 accumulate.
// Range: inputs are converted with saturation, but the score accumulator itself is fp16: |q.k / sqrt(hd) * log2 e + bias| must stay
// below 65504 (a softmax logit of 45,000 -- the layer's inputs are RMS-normalised, trained logits are O(10)); the TF32 kernels
// (B200CTR_ATTN=tf32) have fp32 range.
// No atomics; dbias as in attn_tc.cu (register slots per diagonal block, per-warp shared copies, per-sample partial row).
#include "h2_common.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace b200ctr {

constexpr int HA_D = 64;
constexpr int HA_H = 8;
constexpr int HA_HPC = 4;          // heads per CTA of the backward kernel (two CTAs per sample)
constexpr float HA_LOG2E = 1.4426950408889634f, HA_LN2 = 0.6931471805599453f;
constexpr uint32_t HA_ONES = 0x3c003c00u;       // half2(1, 1)

// D(16 x 8, fp16 pairs) = A(16 x 8) . B(8 x 8) + C
__device__ __forceinline__ void mma_1688_h(uint32_t (&d)[2], const uint32_t (&a)[2], uint32_t b, uint32_t c0, uint32_t c1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f16.f16.f16.f16 {%0,%1}, {%2,%3}, {%4}, {%5,%6};"
      : "=r"(d[0]), "=r"(d[1]) : "r"(a[0]), "r"(a[1]), "r"(b), "r"(c0), "r"(c1));
}
// C(16 x 8, fp32) += A(16 x 16, fp16) . B(16 x 8, fp16)
__device__ __forceinline__ void mma_16816_h(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
  uint32_t r;
  asm("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
// ldmatrix: lane l supplies the address of row (l & 7) of matrix (l >> 3); plain: lane (g, t) gets M[g][2t], M[g][2t+1];
// .trans: M[2t][g], M[2t+1][g]
__device__ __forceinline__ void ldsm_x4(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ void ldsm_x2_t(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Staging: the fp32 [K x 192] qkv block of the sample -> three TRANSPOSED fp16 images X^T[dim][position] (row stride TSW
// 32-bit words = 2 TSW positions; TSW = 4 mod 8).  A unit = (row pair rp, 4 dims c): two float4 loads per image, four packed
// stores; the lanes of a warp are 16 row pairs x 2 dim quads, which makes the stores bank-conflict free (16 * odd * cl + rp)
// and the loads whole 32-byte sectors.  Rows >= K are zero.
// QD = dim quads the CTA stages (16 = all 8 heads, 8 = the 4 heads of a head-split CTA)
struct HaUnit { int c, rp; bool act; };
template <int KP, int QD>
__device__ __forceinline__ HaUnit ha_unit(int u) {
  constexpr int RG = (KP / 2 + 15) / 16;
  HaUnit r;
  const int cl = u & 1, rpl = (u >> 1) & 15, rest = u >> 5;
  r.c = (rest % (QD / 2)) * 2 + cl;
  r.rp = (rest / (QD / 2)) * 16 + rpl;
  r.act = u < RG * 16 * QD && r.rp < KP / 2;
  return r;
}
__device__ __forceinline__ float ha_absmax4(float m, const float4& x) {
  return fmaxf(m, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
}
__device__ __forceinline__ void ha_store_T(uint32_t* dst, int TSW, int c, int rp, const float4& x0, const float4& x1, float s0, float s1) {
  uint32_t* d = dst + (4 * c) * TSW + rp;
  d[0] = h2_pack(x0.x * s0, x1.x * s1);
  d[TSW] = h2_pack(x0.y * s0, x1.y * s1);
  d[2 * TSW] = h2_pack(x0.z * s0, x1.z * s1);
  d[3 * TSW] = h2_pack(x0.w * s0, x1.w * s1);
}
// unscaled (k, v): no multiplies in the staging loop
__device__ __forceinline__ void ha_store_T1(uint32_t* dst, int TSW, int c, int rp, const float4& x0, const float4& x1) {
  uint32_t* d = dst + (4 * c) * TSW + rp;
  d[0] = h2_pack(x0.x, x1.x);
  d[TSW] = h2_pack(x0.y, x1.y);
  d[2 * TSW] = h2_pack(x0.z, x1.z);
  d[3 * TSW] = h2_pack(x0.w, x1.w);
}
// q (scaled), k, v of the sample; the six 128-bit loads of a unit (twelve with two units per thread) are all in flight
// before the first conversion: the staging phase is pure load latency.  Returns the thread's max |v|.
template <int NTH, int KP, int QD>
__device__ __forceinline__ float ha_stage_qkv(uint32_t* qT, uint32_t* kT, uint32_t* vT, const float* __restrict__ base, int K, float qscale) {
  constexpr int RG = (KP / 2 + 15) / 16, UPT = (RG * 16 * QD + NTH - 1) / NTH, TSW = (KP + 8) / 2;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  float vmax = 0.f;
#pragma unroll
  for (int it0 = 0; it0 < UPT; it0 += 2) {
    float4 x[2][3][2];
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      const HaUnit un = ha_unit<KP, QD>(threadIdx.x + (it0 + ii) * NTH);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int row = 2 * un.rp + r;
        const bool ok = it0 + ii < UPT && un.act && row < K;
        const float4* src = reinterpret_cast<const float4*>(base + (int64_t)row * (3 * HA_D)) + un.c;
        x[ii][0][r] = ok ? src[0] : z; x[ii][1][r] = ok ? src[16] : z; x[ii][2][r] = ok ? src[32] : z;
      }
    }
#pragma unroll
    for (int ii = 0; ii < 2; ++ii) {
      const HaUnit un = ha_unit<KP, QD>(threadIdx.x + (it0 + ii) * NTH);
      if (it0 + ii < UPT && un.act) {
        ha_store_T(qT, TSW, un.c, un.rp, x[ii][0][0], x[ii][0][1], qscale, qscale);
        ha_store_T1(kT, TSW, un.c, un.rp, x[ii][1][0], x[ii][1][1]);
        ha_store_T1(vT, TSW, un.c, un.rp, x[ii][2][0], x[ii][2][1]);
        vmax = ha_absmax4(ha_absmax4(vmax, x[ii][2][0]), x[ii][2][1]);
      }
    }
  }
  return vmax;
}

// The same three images from an fp16 qkv tensor ((T, 96) words per row: q~ | k | v with q ALREADY times H2_QSCALE, written by
// b200ctr_inproj_fwd_h16): a unit = (row pair, 8 dims), two 16-byte loads per image, eight byte-permutes and stores -- no
// conversions, half the bytes.  The lanes of a warp are 32 row pairs of one dim octet: stores hit 32 consecutive words.
template <int NTH, int KP, int QD>
__device__ __forceinline__ float ha_stage_qkv16(uint32_t* qT, uint32_t* kT, uint32_t* vT, const uint32_t* __restrict__ base16, int K) {
  constexpr int TSW = (KP + 8) / 2, NOCT = QD / 2, RPG = (KP / 2 + 31) / 32, UNITS = RPG * 32 * NOCT, UPT = (UNITS + NTH - 1) / NTH;
  uint32_t vm2 = 0u;
#pragma unroll
  for (int it = 0; it < UPT; ++it) {
    const int u = threadIdx.x + it * NTH;
    const int rpl = u & 31, rest = u >> 5, oc = rest % NOCT, rp = (rest / NOCT) * 32 + rpl;
    if (u >= UNITS || rp >= KP / 2) continue;
    uint4 x[3][2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = 2 * rp + r;
      const uint4* src = reinterpret_cast<const uint4*>(base16 + (int64_t)row * (3 * HA_D / 2)) + oc;
      const uint4 z = make_uint4(0u, 0u, 0u, 0u);
      x[0][r] = row < K ? src[0] : z; x[1][r] = row < K ? src[8] : z; x[2][r] = row < K ? src[16] : z;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      uint32_t* img = (a == 0 ? qT : (a == 1 ? kT : vT)) + (8 * oc) * TSW + rp;
      const uint32_t w0[4] = {x[a][0].x, x[a][0].y, x[a][0].z, x[a][0].w}, w1[4] = {x[a][1].x, x[a][1].y, x[a][1].z, x[a][1].w};
#pragma unroll
      for (int d = 0; d < 8; ++d) img[d * TSW] = __byte_perm(w0[d >> 1], w1[d >> 1], (d & 1) ? 0x7632 : 0x5410);
      if (a == 2) {
#pragma unroll
        for (int w = 0; w < 4; ++w) vm2 = __vmaxu2(vm2, __vmaxu2(w0[w] & 0x7fff7fffu, w1[w] & 0x7fff7fffu));     // |v| as ordered bit patterns
      }
    }
  }
  return fmaxf(h2_lo(vm2), h2_hi(vm2));
}

// bias pairs in log2 units: bp[8 + d] = half2(bias[d], bias[d + 1]), d = (key - query) + K - 1 in [0, 2K - 2]; zero elsewhere
// (the 8 leading entries serve the rows g + 8 of the first key tile; the tail serves padded keys / clamped padded queries)
template <int NTH>
__device__ __forceinline__ void ha_fill_bias(uint32_t* bp, int n, const float* __restrict__ relmean, int max_len, int K) {
  for (int i = threadIdx.x; i < n; i += NTH) {
    const int d = i - 8;
    float b0 = 0.f, b1 = 0.f;
    if (relmean && d >= 0 && d <= 2 * K - 2) {
      b0 = relmean[min(max(d - (K - 1), -max_len), max_len) + max_len] * HA_LOG2E;
      if (d + 1 <= 2 * K - 2) b1 = relmean[min(max(d + 1 - (K - 1), -max_len), max_len) + max_len] * HA_LOG2E;
    }
    bp[i] = h2_pack(b0, b1);
  }
}

// keep-flag words: same size and addressing as attn_tc.cu (b200ctr_attn_tc_keep_words); word w of a (sample, head, 16-query
// tile) holds, for lane (g, t), the key tiles 8w .. 8w+7: bit q / bit 16 + q = key 2t / 2t + 1 of slot q = 2 * (tile & 7) + row half
__device__ __forceinline__ int64_t ha_keep_index(int64_t b, int h, int mt, int i, int nw, int w, int lane) {
  return ((((b * HA_H + h) * mt + i) * nw + w) << 5) + lane;
}

// and / or words that turn the scores of padded keys (>= K) into -inf (forward) or the probabilities into 0 (backward)
__device__ __forceinline__ void ha_key_masks(int kc, int K, uint32_t& andm, uint32_t& orm) {
  const bool v0 = kc < K, v1 = kc + 1 < K;
  andm = (v0 ? 0xffffu : 0u) | (v1 ? 0xffff0000u : 0u);
  orm = (v0 ? 0u : 0xfc00u) | (v1 ? 0u : 0xfc000000u);
}

// two adjacent elements at element offset `off` of an output tensor: fp32 pair or one packed bf16 pair
__device__ __forceinline__ void ha_store2(float* base, int as_bf16, int64_t off, float a, float b) {
  if (as_bf16) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    reinterpret_cast<uint32_t*>(base)[off >> 1] = r;
  } else {
    *reinterpret_cast<float2*>(base + off) = make_float2(a, b);
  }
}

// four consecutive elements of an fp32 or bf16 tensor at element offset `off` (a multiple of 4)
__device__ __forceinline__ float4 ha_load4(const float* base, int is_bf16, int64_t off) {
  if (!is_bf16) return *reinterpret_cast<const float4*>(base + off);
  const uint2 w = *reinterpret_cast<const uint2*>(reinterpret_cast<const uint16_t*>(base) + off);
  return make_float4(__uint_as_float(w.x << 16), __uint_as_float(w.x & 0xffff0000u), __uint_as_float(w.y << 16), __uint_as_float(w.y & 0xffff0000u));
}

// NT = number of 8-key tiles (keys padded to a multiple of 16).  Like the backward, two CTAs per sample with four heads each
// (one warp per head): half the staging per CTA, twice as many CTAs to hide it behind.
template <int NT, bool DROP>
__global__ void __launch_bounds__(32 * HA_HPC, (NT <= 14 ? 32 : 24) / HA_HPC)
attn_h16_fwd_kernel(const float* __restrict__ qkv, int in16, const float* __restrict__ relmean, int max_len, int K,
                    float* __restrict__ o, int o16, float* __restrict__ lse, uint32_t* __restrict__ keep, b200ctr_dropout_t drop) {
  // o16: o is written as bf16 (same (M, 64) layout) -- both of its consumers' tensor-core products (the out-projection and its
  // weight gradient) round it to bf16 anyway, so only the backward's D = dO . o sees the rounding
  constexpr int Kp = NT * 8, MT = NT / 2, NB = NT / 2, NW = (NT + 7) / 8, TS = Kp + 8, TSW = TS / 2, NBP = 2 * Kp + 16;
  constexpr bool KEEP_S = NT <= 8;
  extern __shared__ __align__(16) uint32_t sm_u[];
  constexpr int HPC = HA_HPC, NTH = 32 * HA_HPC;
  uint32_t* qT = sm_u;                       // [8 HPC][TSW]  (q / sqrt(hd) * log2 e)^T of this CTA's heads
  uint32_t* kT = qT + 8 * HPC * TSW;
  uint32_t* vT = kT + 8 * HPC * TSW;
  uint32_t* bp = vT + 8 * HPC * TSW;         // [NBP]
  uint32_t* keys = bp + NBP;                 // [HPC][Kp] dropout row keys (DROP only)
  const int64_t b = blockIdx.x / (HA_H / HPC);
  const int hbase = (int)(blockIdx.x % (HA_H / HPC)) * HPC;
  if (in16)      // (T, 96) words of fp16 pairs, q pre-scaled
    ha_stage_qkv16<NTH, Kp, 2 * HPC>(qT, kT, vT, reinterpret_cast<const uint32_t*>(qkv) + b * (int64_t)K * (3 * HA_D / 2) + hbase * 4, K);
  else
    ha_stage_qkv<NTH, Kp, 2 * HPC>(qT, kT, vT, qkv + b * (int64_t)K * 3 * HA_D + hbase * 8, K, H2_QSCALE);
  ha_fill_bias<NTH>(bp, NBP, relmean, max_len, K);
  if (DROP)
    for (int i = threadIdx.x; i < HPC * Kp; i += NTH) {
      const int hl = i / Kp, r = i % Kp;
      keys[i] = r < K ? drop_row_key(drop, ((uint64_t)b * HA_H + hbase + hl) * K + r) : 0u;
    }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int hl = warp, h = hbase + hl;
  const DropCtx dctx = drop_ctx(drop);
  const uint32_t thr2 = dctx.thr | (dctx.thr << 16);
  // ldmatrix lane addresses (bytes): row (h * 8 + (lane & 7)) of the image, position offset 8 * ((lane >> 3) & 1)
  const uint32_t lrow = (uint32_t)(((hl * 8 + (lane & 7)) * TS + 8 * ((lane >> 3) & 1)) * 2);
  const uint32_t q_addr = smem_u32(qT) + lrow, k_addr = smem_u32(kT) + lrow, v_addr = smem_u32(vT) + lrow;
  uint32_t andm[2], orm[2];
  ha_key_masks(8 * (NT - 2) + 2 * t, K, andm[0], orm[0]);
  ha_key_masks(8 * (NT - 1) + 2 * t, K, andm[1], orm[1]);
  const uint32_t tphi = (uint32_t)t * 0x9E3779B9U;
#pragma unroll 1
  for (int i = 0; i < MT; ++i) {
    const int qa = 16 * i + g, qb = qa + 8;
    uint32_t a[2];
    ldsm_x2_t(a[0], a[1], q_addr + 32 * i);
    const uint32_t* bpa = bp + 8 + 2 * t + (K - 1) - min(qa, K - 1);     // + 8 * tile
    // score tile j of this query tile: S = Q.K^T with the packed bias pair as the accumulator input; rows g + 8 see the bias
    // of rows g shifted by one key tile (cprev).  Up to 64 keys the tiles stay in registers between the max pass and the
    // exp pass; above, the second pass recomputes them (one more HMMA + ldmatrix per tile instead of 2 NT live registers,
    // which is what decides how many CTAs fit on an SM)
    auto score_tile = [&](uint32_t (&st)[2], uint32_t kfrag, int j, uint32_t& cprev) {
      const uint32_t ca = bpa[8 * j];
      mma_1688_h(st, a, kfrag, ca, cprev);
      cprev = ca;
      if (j >= NT - 2) {
        st[0] = (st[0] & andm[j - (NT - 2)]) | orm[j - (NT - 2)];
        st[1] = (st[1] & andm[j - (NT - 2)]) | orm[j - (NT - 2)];
      }
    };
    uint32_t s[KEEP_S ? NT : 1][2];
    uint32_t mxa = 0xfc00fc00u, mxb = 0xfc00fc00u;      // -inf
    {
      uint32_t cprev = bpa[-8];
#pragma unroll
      for (int jb = 0; jb < NB; ++jb) {
        uint32_t kb[2];
        ldsm_x2_t(kb[0], kb[1], k_addr + 32 * jb);
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int j = 2 * jb + tt;
          uint32_t st[2];
          score_tile(st, kb[tt], j, cprev);
          mxa = h2_max(mxa, st[0]); mxb = h2_max(mxb, st[1]);
          if (KEEP_S) { s[j][0] = st[0]; s[j][1] = st[1]; }
        }
      }
    }
    mxa = h2_max(mxa, h2_swap(mxa)); mxb = h2_max(mxb, h2_swap(mxb));
    uint32_t pk = __byte_perm(mxa, mxb, 0x5410);                            // (max row g, max row g + 8)
    pk = h2_max(pk, __shfl_xor_sync(0xffffffffu, pk, 1));
    pk = h2_max(pk, __shfl_xor_sync(0xffffffffu, pk, 2));
    const uint32_t nma = h2_dup_lo(pk) ^ 0x80008000u, nmb = h2_dup_hi(pk) ^ 0x80008000u;      // -max, both halves
    float oc[4] = {0.f, 0.f, 0.f, 0.f}, ls[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t keya = 0u, keyb = 0u;
    if (DROP) { keya = keys[hl * Kp + qa]; keyb = keys[hl * Kp + qb]; }
    uint32_t kw[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) kw[w] = 0u;
    uint32_t cprev2 = bpa[-8];
#pragma unroll
    for (int jb = 0; jb < NB; ++jb) {
      uint32_t vb0, vb1, kb[2];
      ldsm_x2(vb0, vb1, v_addr + 32 * jb);
      if (!KEEP_S) ldsm_x2_t(kb[0], kb[1], k_addr + 32 * jb);
      uint32_t e[4];                                                          // A fragment: (t0 row g, t0 row g+8, t1 row g, t1 row g+8)
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) {
        const int j = 2 * jb + tt;
        uint32_t st[2];
        if (KEEP_S) { st[0] = s[j][0]; st[1] = s[j][1]; }
        else score_tile(st, kb[tt], j, cprev2);
        e[tt * 2] = h2_ex2(h2_add(st[0], nma));
        e[tt * 2 + 1] = h2_ex2(h2_add(st[1], nmb));
      }
      mma_16816_h(ls, e[0], e[1], e[2], e[3], HA_ONES, HA_ONES);          // softmax denominators (before dropout)
      if (DROP) {
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int j = 2 * jb + tt;
          const uint32_t cj = (uint32_t)(4 * j) * 0x9E3779B9U + tphi;      // (4j + t) * golden
          const uint32_t ma = drop_keep_mask2(drop_pair_hash_pre(keya, cj), thr2), mb = drop_keep_mask2(drop_pair_hash_pre(keyb, cj), thr2);
          e[tt * 2] &= ma; e[tt * 2 + 1] &= mb;
          kw[j >> 3] |= (ma & (0x00010001u << (2 * (j & 7)))) | (mb & (0x00010001u << (2 * (j & 7) + 1)));
        }
      }
      mma_16816_h(oc, e[0], e[1], e[2], e[3], vb0, vb1);
    }
    if (DROP && keep) {
#pragma unroll
      for (int w = 0; w < NW; ++w) keep[ha_keep_index(b, h, MT, i, NW, w, lane)] = kw[w];
    }
    if (qa < K) {
      const float inv = dctx.inv / ls[0];
      ha_store2(o, o16, (b * K + qa) * (int64_t)HA_D + h * 8 + 2 * t, oc[0] * inv, oc[1] * inv);
      if (t == 0) lse[(b * HA_H + h) * (int64_t)K + qa] = (h2_lo(pk) + log2f(ls[0])) * HA_LN2;       // natural-log units
    }
    if (qb < K) {
      const float inv = dctx.inv / ls[2];
      ha_store2(o, o16, (b * K + qb) * (int64_t)HA_D + h * 8 + 2 * t, oc[2] * inv, oc[3] * inv);
      if (t == 0) lse[(b * HA_H + h) * (int64_t)K + qb] = (h2_hi(pk) + log2f(ls[2])) * HA_LN2;
    }
  }
}

// ---- backward ---------------------------------------------------------------------------------------------------------
// TWO CTAs per sample, four heads each (HA_HPC): a CTA stages only its heads' 32 columns of q / k / v / dO / o, so the per-sample
// staging phase (pure global-load latency) is half as long and twice as many, smaller CTAs interleave it with the others'
// compute.  warps = (head, key part): the KS parts of a head own disjoint runs of 16-key blocks (dK / dV accumulators in
// registers), dQ is summed over the parts through a shared-memory hand-off.  KS = 1 up to 64 keys, 2 above.
// drelmean: one partial row per CTA = 2 B rows (b200ctr_attn_h16_drel_rows).
// DROP: dropout on; KEEP: the keep flags come from the words the forward stored (otherwise re-hashed: same mask).
template <int NT, int KS, bool DROP, bool KEEP>
__global__ void __launch_bounds__(32 * HA_HPC * KS, (KS == 1 ? 24 : 8) / HA_HPC)
attn_h16_bwd_kernel(const float* __restrict__ qkv, int in16, const float* __restrict__ relmean, int max_len, const float* __restrict__ o,
                    const float* __restrict__ lse, const float* __restrict__ d_o, int io16 /* bit 0: o is bf16, bit 1: d_o is bf16 */,
                    int K, float* __restrict__ dqkv, int dq16, float* __restrict__ drelmean, const uint32_t* __restrict__ keep,
                    b200ctr_dropout_t drop) {
  constexpr int Kp = NT * 8, MT = NT / 2, NB = NT / 2, NW = (NT + 7) / 8, TS = Kp + 8, TSW = TS / 2, NBP = 2 * Kp + 16;
  constexpr int NTH = 32 * HA_HPC * KS, HPC = HA_HPC, QD = 2 * HA_HPC;
  constexpr int NBW = (NB + KS - 1) / KS;                     // 16-key blocks per warp
  constexpr int NTW = 2 * NBW;                                 // 8-key tiles per warp
  constexpr int KPA = 16 * NBW * KS;                           // padded key extent covered by the dbias slots (>= Kp)
  constexpr int DBS = 2 * KPA + 2;
  extern __shared__ __align__(16) uint32_t sm_u[];
  uint32_t* qT = sm_u;                       // [8 HPC][TSW]  (q / sqrt(hd) * log2 e)^T of this CTA's heads
  uint32_t* kT = qT + 8 * HPC * TSW;
  uint32_t* vT = kT + 8 * HPC * TSW;
  uint32_t* doT = vT + 8 * HPC * TSW;        // (dO * f / (1-p) * 2^e)^T
  uint32_t* mneg = doT + 8 * HPC * TSW;      // [HPC][Kp] half2(-m', -m')
  uint32_t* dneg = mneg + HPC * Kp;          // [HPC][Kp] half2(-D', -D')
  uint32_t* bp = dneg + HPC * Kp;            // [NBP]
  float* dbias = reinterpret_cast<float*>(bp + NBP);            // [HPC * KS][DBS], index (key - query) + KPA
  float* dqx = dbias + HPC * KS * DBS;                          // [HPC][2 (query-tile parity)][KS - 1][128]
  float* red = dqx + (KS > 1 ? HPC * 2 * (KS - 1) * 128 : 0);   // [64] block-reduction scratch
  uint32_t* keys = reinterpret_cast<uint32_t*>(red + 64);       // [HPC][Kp] (DROP && !KEEP)
  const int64_t b = blockIdx.x / (HA_H / HPC);
  const int hbase = (int)(blockIdx.x % (HA_H / HPC)) * HPC, dim0 = hbase * 8;      // first head / first column of this CTA
  const float* base = qkv + b * (int64_t)K * 3 * HA_D + dim0;
  const int64_t io_off = b * (int64_t)K * HA_D + dim0;          // element offset of this CTA's columns of row 0 in o / d_o
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const DropCtx dctx = drop_ctx(drop);
  // dO, the head-partial dot product dO . o (D = sum_j P_ij c_ij dP_ij = dO_i . o_i holds with the dropped-and-rescaled o the
  // forward wrote) and the log-sum-exp of the thread's units stay in registers across the barrier that agrees on the scale
  constexpr int RG = (Kp / 2 + 15) / 16, UPT = (RG * 16 * QD + NTH - 1) / NTH;
  float4 xd[UPT][2];
  float Ls[UPT][2], dro[UPT][2];
  float domax = 0.f;
#pragma unroll
  for (int it = 0; it < UPT; ++it) {
    const HaUnit un = ha_unit<Kp, QD>(threadIdx.x + it * NTH);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = 2 * un.rp + r;
      const bool ok = un.act && row < K;
      xd[it][r] = make_float4(0.f, 0.f, 0.f, 0.f); Ls[it][r] = 0.f; dro[it][r] = 0.f;
      if (ok) {
        xd[it][r] = ha_load4(d_o, io16 & 2, io_off + (int64_t)row * HA_D + 4 * un.c);
        const float4 oo = ha_load4(o, io16 & 1, io_off + (int64_t)row * HA_D + 4 * un.c);
        Ls[it][r] = lse[(b * HA_H + hbase + (un.c >> 1)) * (int64_t)K + row];
        dro[it][r] = xd[it][r].x * oo.x + xd[it][r].y * oo.y + xd[it][r].z * oo.z + xd[it][r].w * oo.w;
      }
    }
  }
  float vmax = in16 ? ha_stage_qkv16<NTH, Kp, QD>(qT, kT, vT, reinterpret_cast<const uint32_t*>(qkv) + b * (int64_t)K * (3 * HA_D / 2) + hbase * 4, K)
                    : ha_stage_qkv<NTH, Kp, QD>(qT, kT, vT, base, K, H2_QSCALE);
#pragma unroll
  for (int it = 0; it < UPT; ++it) domax = ha_absmax4(ha_absmax4(domax, xd[it][0]), xd[it][1]);
  vmax = warp_max(vmax); domax = warp_max(domax);
  if (lane == 0) { red[warp] = vmax; red[32 + warp] = domax; }
  ha_fill_bias<NTH>(bp, NBP, relmean, max_len, K);
  if (drelmean)
    for (int i = threadIdx.x; i < HPC * KS * DBS; i += NTH) dbias[i] = 0.f;
  if (DROP && !KEEP)
    for (int i = threadIdx.x; i < HPC * Kp; i += NTH) {
      const int hl = i / Kp, r = i % Kp;
      keys[i] = r < K ? drop_row_key(drop, ((uint64_t)b * HA_H + hbase + hl) * K + r) : 0u;
    }
  __syncthreads();
  // per-sample power of two: |dP'|, |D'| <= 8 max|dO'| max|V| are placed below 2^9, so |dS| < 2^10, a dbias slot (up to ten
  // dS) < 2^14 and a chain of four slots < 2^16 = fp16's top even in the worst case; a gradient of 0 or a non-finite one
  // leaves the scale at 1
  float spow = 1.f;
  {
    float vm = 0.f, dm = 0.f;
    for (int w = 0; w < NTH / 32; ++w) { vm = fmaxf(vm, red[w]); dm = fmaxf(dm, red[32 + w]); }
    const float bound = 8.2f * dm * vm * dctx.inv;
    if (bound > 0.f && bound < 3.0e38f) spow = scalbnf(1.f, min(max(8 - ilogbf(bound), -120), 120));
  }
  const float rspow = 1.f / spow;
  // dO' = dO * f * spow / (1-p), f = 2^(m' - lse*log2e), m' = the log-sum-exp rounded to fp16;  D' = dO' . o over the head
#pragma unroll
  for (int it = 0; it < UPT; ++it) {
    const HaUnit un = ha_unit<Kp, QD>(threadIdx.x + it * NTH);
    float m[2], f[2], d[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const bool ok = un.act && 2 * un.rp + r < K;
      const float L = Ls[it][r] * HA_LOG2E;
      m[r] = ok ? __half2float(__float2half_rn(L)) : 100.f;      // padded rows: e = 2^(bias - 100) ~ 0 and dO' = D' = 0
      f[r] = ok ? exp2f(m[r] - L) * spow * dctx.inv : 0.f;
      // dO' carries the extra 1/(1-p) of the mask scale, which D' must not: o already contains it
      d[r] = dro[it][r] * f[r] * (1.f / dctx.inv);
      d[r] += __shfl_xor_sync(0xffffffffu, d[r], 1);                // the other half of the head's 8 dims
    }
    if (un.act) {
      ha_store_T(doT, TSW, un.c, un.rp, xd[it][0], xd[it][1], f[0], f[1]);
      if ((un.c & 1) == 0) {
        const int hh = un.c >> 1, r0 = 2 * un.rp;          // local head
        mneg[hh * Kp + r0] = h2_pack(-m[0], -m[0]); mneg[hh * Kp + r0 + 1] = h2_pack(-m[1], -m[1]);
        dneg[hh * Kp + r0] = h2_pack(-d[0], -d[0]); dneg[hh * Kp + r0 + 1] = h2_pack(-d[1], -d[1]);
      }
    }
  }
  __syncthreads();
  const int hl = warp % HPC, part = warp / HPC, h = hbase + hl;  // (local head, key part), global head
  const int jb0 = part * NBW;                                   // first 16-key block of this warp
  const uint32_t thr2 = dctx.thr | (dctx.thr << 16);
  const uint32_t tphi = (uint32_t)t * 0x9E3779B9U;
  // ldmatrix lane addresses (bytes).  "ab": matrices (qT pos +0, qT pos +8, doT pos +0, doT pos +8); "kv": same for (kT, vT)
  const uint32_t lrow = (uint32_t)(((hl * 8 + (lane & 7)) * TS + 8 * ((lane >> 3) & 1)) * 2);
  const uint32_t ab_addr = (lane < 16 ? smem_u32(qT) : smem_u32(doT)) + lrow;
  const uint32_t kv_addr = (lane < 16 ? smem_u32(kT) : smem_u32(vT)) + lrow;
  uint32_t andm[2], orm_unused;
  ha_key_masks(8 * (NT - 2) + 2 * t, K, andm[0], orm_unused);
  ha_key_masks(8 * (NT - 1) + 2 * t, K, andm[1], orm_unused);
  float dk[NBW][4], dv[NBW][4];
#pragma unroll
  for (int jl = 0; jl < NBW; ++jl) { dk[jl][0] = dk[jl][1] = dk[jl][2] = dk[jl][3] = 0.f; dv[jl][0] = dv[jl][1] = dv[jl][2] = dv[jl][3] = 0.f; }
  // dbias: a thread's dS pair of local tile jt lies on the diagonals 8 (2 jb0 + jt - 2i) + (2t - g) {+0, +1} (rows g) and the
  // same shifted by -8 (rows g + 8): half2 slots indexed by the local tile (+1) that rotate by two per query tile; only the
  // two slots that leave the window are flushed per query tile (attn_tc.cu has the fp32 version of the same scheme).
  uint32_t A[NTW + 1];
#pragma unroll
  for (int sl = 0; sl <= NTW; ++sl) A[sl] = 0u;
  float* dbt = dbias + warp * DBS + KPA + 2 * t - g + 8 * (2 * jb0);    // + 8 * (local slot - 1 - 2i)
  // lanes (g, t), (g+2, t+1), (g+4, t+2), (g+6, t+3) = lane, lane+9, lane+18, lane+27 lie on the same diagonal pair:
  // two shuffle steps leave the chain sum in its first lane (t == 0 or g < 2); the even and the odd diagonal are written in
  // two steps because neighbouring chains share a word
  const bool ch1 = t < 3 && g < 6, ch2 = t < 2 && g < 4, chead = t == 0 || g < 2;
  auto dbias_add = [&](float* p, uint32_t v) {
    uint32_t x = __shfl_down_sync(0xffffffffu, v, 9);
    if (ch1) v = h2_add(v, x);
    x = __shfl_down_sync(0xffffffffu, v, 18);
    if (ch2) v = h2_add(v, x);
    if (chead) p[0] += h2_lo(v);
    __syncwarp();
    if (chead) p[1] += h2_hi(v);
    __syncwarp();
  };
  uint32_t kwn[NW];
  if (DROP && KEEP) {
#pragma unroll
    for (int w = 0; w < NW; ++w) kwn[w] = __ldg(keep + ha_keep_index(b, h, MT, 0, NW, w, lane));
  }
#pragma unroll 1
  for (int i = 0; i < MT; ++i) {
    const int qa = 16 * i + g, qb = qa + 8;
    uint32_t aq[2], ad[2], bq0, bq1, bd0, bd1;
    ldsm_x4_t(aq[0], aq[1], ad[0], ad[1], ab_addr + 32 * i);     // A fragments of S = Q.K^T and dP = dO'.V^T
    ldsm_x4(bq0, bq1, bd0, bd1, ab_addr + 32 * i);               // B fragments (16 queries x head dim) of dK += dS^T.Q, dV += P^T.dO'
    const uint32_t mna = mneg[hl * Kp + qa], mnb = mneg[hl * Kp + qb], dna = dneg[hl * Kp + qa], dnb = dneg[hl * Kp + qb];
    uint32_t kw[NW];
    if (DROP && KEEP) {      // this query tile's keep words were requested one iteration ago: the loop never waits on global memory
#pragma unroll
      for (int w = 0; w < NW; ++w) { kw[w] = kwn[w]; kwn[w] = __ldg(keep + ha_keep_index(b, h, MT, min(i + 1, MT - 1), NW, w, lane)); }
    }
    uint32_t keya = 0u, keyb = 0u;
    if (DROP && !KEEP) { keya = keys[hl * Kp + qa]; keyb = keys[hl * Kp + qb]; }
    const uint32_t* bpa = bp + 8 + 2 * t + (K - 1) - min(qa, K - 1) + 16 * jb0;     // + 8 * local tile
    uint32_t cprev = bpa[-8];
    float dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int jl = 0; jl < NBW; ++jl) {
      if (jb0 + jl < NB) {                                      // warp-uniform
        const int jb = jb0 + jl;
        uint32_t kb[2], vb[2], kq0, kq1;
        ldsm_x4_t(kb[0], kb[1], vb[0], vb[1], kv_addr + 32 * jb);
        ldsm_x2(kq0, kq1, kv_addr + 32 * jb);                    // lanes 0-15 address kT: (16 keys x head dim) of dQ += dS.K
        uint32_t ds[4], pd[4];                                   // (t0 row g, t0 row g+8, t1 row g, t1 row g+8)
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          const int jt = 2 * jl + tt;                           // local tile
          const int j = 2 * jb + tt;                            // global tile
          const uint32_t ca = bpa[8 * jt];
          uint32_t s[2], dp[2];
          mma_1688_h(s, aq, kb[tt], ca, cprev);
          cprev = ca;
          uint32_t ea = h2_ex2(h2_add(s[0], mna)), eb = h2_ex2(h2_add(s[1], mnb));
          if (j >= NT - 2) {                                    // warp-uniform: only the last two tiles can hold padded keys
            const uint32_t am = j == NT - 2 ? andm[0] : andm[1];
            ea &= am; eb &= am;
          }
          if (DROP) {
            uint32_t ma, mb;
            if (KEEP) {
              const int w = j >> 3;
              const uint32_t word = NW == 1 ? kw[0] : (w == 0 ? kw[0] : (w == 1 ? kw[NW > 1 ? 1 : 0] : kw[NW > 2 ? 2 : 0]));
              const int q = 2 * (j & 7);
              ma = h2_mask_from_signs(word << (15 - q));
              mb = h2_mask_from_signs(word << (14 - q));
            } else {
              const uint32_t cj = (uint32_t)(4 * j) * 0x9E3779B9U + tphi;      // (4j + t) * golden
              ma = drop_keep_mask2(drop_pair_hash_pre(keya, cj), thr2);
              mb = drop_keep_mask2(drop_pair_hash_pre(keyb, cj), thr2);
            }
            mma_1688_h(dp, ad, vb[tt], 0u, 0u);
            dp[0] = h2_add(dp[0] & ma, dna); dp[1] = h2_add(dp[1] & mb, dnb);
            pd[tt * 2] = ea & ma; pd[tt * 2 + 1] = eb & mb;
          } else {
            mma_1688_h(dp, ad, vb[tt], dna, dnb);
            pd[tt * 2] = ea; pd[tt * 2 + 1] = eb;
          }
          ds[tt * 2] = h2_mul(ea, dp[0]); ds[tt * 2 + 1] = h2_mul(eb, dp[1]);
          A[jt + 1] = h2_add(A[jt + 1], ds[tt * 2]); A[jt] = h2_add(A[jt], ds[tt * 2 + 1]);    // padded rows / keys contribute exact zeros
        }
        mma_16816_h(dq, ds[0], ds[1], ds[2], ds[3], kq0, kq1);
        // A = tile^T: rows = 16 keys of the block (even tile -> rows 0-7, odd tile -> rows 8-15), k = 16 queries
        mma_16816_h(dv[jl], movm_t(pd[0]), movm_t(pd[2]), movm_t(pd[1]), movm_t(pd[3]), bd0, bd1);
        mma_16816_h(dk[jl], movm_t(ds[0]), movm_t(ds[2]), movm_t(ds[1]), movm_t(ds[3]), bq0, bq1);
      }
    }
    {
      const float qs = rsqrtf(8.f) * rspow;
      if (KS > 1) {
        float2* xq = reinterpret_cast<float2*>(dqx + (size_t)((hl * 2 + (i & 1)) * (KS - 1)) * 128);
        if (part != 0) {
          xq[(part - 1) * 64 + lane] = make_float2(dq[0], dq[1]);            // row g, dims 2t, 2t+1
          xq[(part - 1) * 64 + 32 + lane] = make_float2(dq[2], dq[3]);       // row g + 8
        }
        asm volatile("bar.sync %0, %1;" :: "r"(hl + 1), "r"(32 * KS) : "memory");   // the KS warps of this head
        if (part == 0) {
#pragma unroll
          for (int pp = 0; pp < KS - 1; ++pp) {
            const float2 x0 = xq[pp * 64 + lane], x1 = xq[pp * 64 + 32 + lane];
            dq[0] += x0.x; dq[1] += x0.y; dq[2] += x1.x; dq[3] += x1.y;
          }
        }
      }
      if (part == 0) {
        if (qa < K) ha_store2(dqkv, dq16, (b * K + qa) * (int64_t)(3 * HA_D) + h * 8 + 2 * t, dq[0] * qs, dq[1] * qs);
        if (qb < K) ha_store2(dqkv, dq16, (b * K + qb) * (int64_t)(3 * HA_D) + h * 8 + 2 * t, dq[2] * qs, dq[3] * qs);
      }
    }
    if (drelmean) {
      // local slot sl holds the diagonal block u = 2*jb0 + sl - 1 - 2i
      if (i + 1 < MT) {
        dbias_add(dbt + 8 * (NTW - 1 - 2 * i), A[NTW]);
        dbias_add(dbt + 8 * (NTW - 2 - 2 * i), A[NTW - 1]);
#pragma unroll
        for (int sl = NTW; sl >= 2; --sl) A[sl] = A[sl - 2];
        A[0] = A[1] = 0u;
      } else {
#pragma unroll
        for (int sl = 0; sl <= NTW; ++sl) dbias_add(dbt + 8 * (sl - 1 - 2 * i), A[sl]);
      }
    }
  }
  {
    const float ksc = HA_LN2 * rspow;                             // dK = dS^T.(q log2e / sqrt(hd)) / log2e
#pragma unroll
    for (int jl = 0; jl < NBW; ++jl) {
      const int ka = 16 * (jb0 + jl) + g, kb = ka + 8;
      if (ka < K) {
        ha_store2(dqkv, dq16, (b * K + ka) * (int64_t)(3 * HA_D) + HA_D + h * 8 + 2 * t, dk[jl][0] * ksc, dk[jl][1] * ksc);
        ha_store2(dqkv, dq16, (b * K + ka) * (int64_t)(3 * HA_D) + 2 * HA_D + h * 8 + 2 * t, dv[jl][0] * rspow, dv[jl][1] * rspow);
      }
      if (kb < K) {
        ha_store2(dqkv, dq16, (b * K + kb) * (int64_t)(3 * HA_D) + HA_D + h * 8 + 2 * t, dk[jl][2] * ksc, dk[jl][3] * ksc);
        ha_store2(dqkv, dq16, (b * K + kb) * (int64_t)(3 * HA_D) + 2 * HA_D + h * 8 + 2 * t, dv[jl][2] * rspow, dv[jl][3] * rspow);
      }
    }
  }
  if (drelmean) {
    __syncthreads();
    for (int tt = threadIdx.x; tt < 2 * K - 1; tt += NTH) {
      const int d = tt - (K - 1);
      float acc = 0.f;
      for (int w = 0; w < HPC * KS; ++w) acc += dbias[w * DBS + d + KPA];
      // per-sample partial row (zero-initialised by the caller; |d| <= K - 1 < max_len: one writer per slot), summed in sample
      // order afterwards (b200ctr_colsum): no float atomics
      drelmean[blockIdx.x * (int64_t)(2 * max_len + 1) + min(max(d, -max_len), max_len) + max_len] = acc * rspow;
    }
  }
}

static int ha_set_smem(const void* fn, size_t bytes) {
  if (bytes > 227 * 1024) {
    set_error("attention (packed-half path): K too large for one CTA's shared memory (%zu bytes)", bytes);
    return B200CTR_E_UNSUPPORTED_SHAPE;
  }
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(attn_h16 smem)");
  }
  return B200CTR_OK;
}

template <int NT>
static int ha_launch_fwd(const float* qkv, int in16, const float* relmean, int max_len, int64_t B, int K, float* o, int o16, float* lse,
                         uint32_t* keep, b200ctr_dropout_t drop, cudaStream_t st) {
  constexpr int Kp = NT * 8, TSW = (Kp + 8) / 2;
  const bool on = drop.p > 0.f;
  const size_t smem = ((size_t)3 * 8 * HA_HPC * TSW + 2 * Kp + 16 + (on ? HA_HPC * Kp : 0)) * 4;
  int rc = on ? ha_set_smem((const void*)attn_h16_fwd_kernel<NT, true>, smem) : ha_set_smem((const void*)attn_h16_fwd_kernel<NT, false>, smem);
  if (rc) return rc;
  const unsigned grid = (unsigned)(B * (HA_H / HA_HPC));
  if (on) attn_h16_fwd_kernel<NT, true><<<grid, 32 * HA_HPC, smem, st>>>(qkv, in16, relmean, max_len, K, o, o16, lse, keep, drop);
  else attn_h16_fwd_kernel<NT, false><<<grid, 32 * HA_HPC, smem, st>>>(qkv, in16, relmean, max_len, K, o, o16, lse, keep, drop);
  return B200CTR_OK;
}

template <int NT, int KS, bool DROP, bool KEEP>
static int ha_launch_bwd_inst(size_t smem, const float* qkv, int in16, const float* relmean, int max_len, const float* o, const float* lse,
                              const float* d_o, int io16, int64_t B, int K, float* dqkv, int dq16, float* drelmean, const uint32_t* keep,
                              b200ctr_dropout_t drop, cudaStream_t st) {
  int rc = ha_set_smem((const void*)attn_h16_bwd_kernel<NT, KS, DROP, KEEP>, smem);
  if (rc) return rc;
  attn_h16_bwd_kernel<NT, KS, DROP, KEEP><<<(unsigned)(B * (HA_H / HA_HPC)), 32 * HA_HPC * KS, smem, st>>>(
      qkv, in16, relmean, max_len, o, lse, d_o, io16, K, dqkv, dq16, drelmean, keep, drop);
  return B200CTR_OK;
}

template <int NT>
static int ha_launch_bwd(const float* qkv, int in16, const float* relmean, int max_len, const float* o, const float* lse, const float* d_o,
                         int io16, int64_t B, int K, float* dqkv, int dq16, float* drelmean, const uint32_t* keep, b200ctr_dropout_t drop,
                         cudaStream_t st) {
  constexpr int KS = NT <= 8 ? 1 : 2;
  constexpr int Kp = NT * 8, TSW = (Kp + 8) / 2, NBW = (NT / 2 + KS - 1) / KS, KPA = 16 * NBW * KS;
  const bool on = drop.p > 0.f, use_keep = on && keep != nullptr;
  const size_t smem = ((size_t)4 * 8 * HA_HPC * TSW + 2 * HA_HPC * Kp + 2 * Kp + 16 + (size_t)(HA_HPC * KS) * (2 * KPA + 2) +
                       (KS > 1 ? HA_HPC * 2 * (KS - 1) * 128 : 0) + 64 + (on && !use_keep ? HA_HPC * Kp : 0)) * 4;
#define HA_ARGS smem, qkv, in16, relmean, max_len, o, lse, d_o, io16, B, K, dqkv, dq16, drelmean, keep, drop, st
  if (!on) return ha_launch_bwd_inst<NT, KS, false, false>(HA_ARGS);
  if (use_keep) return ha_launch_bwd_inst<NT, KS, true, true>(HA_ARGS);
  return ha_launch_bwd_inst<NT, KS, true, false>(HA_ARGS);
#undef HA_ARGS
}

}  // namespace b200ctr

using namespace b200ctr;

#define HA_DISPATCH(CALL)                                                            \
  switch ((K + 15) / 16) {                                                           \
    case 1: rc = CALL(2); break;   case 2: rc = CALL(4); break;   case 3: rc = CALL(6); break;   \
    case 4: rc = CALL(8); break;   case 5: rc = CALL(10); break;  case 6: rc = CALL(12); break;  \
    case 7: rc = CALL(14); break;  case 8: rc = CALL(16); break;  case 9: rc = CALL(18); break;  \
    case 10: rc = CALL(20); break;                                                   \
    default:                                                                         \
      set_error("attention (packed-half path): top_k=%d > 160 is not built", K);     \
      return B200CTR_E_UNSUPPORTED_SHAPE;                                            \
  }

extern "C" int64_t b200ctr_attn_h16_drel_rows(int64_t B) { return B * (HA_H / HA_HPC); }

extern "C" int b200ctr_attn_h16_fwd(const void* qkv, int qkv_f16, const float* relmean, int max_len, int64_t B, int K, int H, int D, void* o,
                                    int o_bf16, float* lse, uint32_t* keep, b200ctr_dropout_t drop, void* stream) {
  B200_REQUIRE(qkv && o && lse, B200CTR_E_INVALID_ARG, "attn_h16_fwd: null pointer");
  B200_REQUIRE(D == HA_D && H == HA_H, B200CTR_E_UNSUPPORTED_SHAPE,
               "attn_h16_fwd: only emb_dim=64 with 8 heads of head_dim=8 is built (got D=%d, H=%d)", D, H);
  B200_REQUIRE(drop.p < 0.96f, B200CTR_E_INVALID_ARG, "attn_h16_fwd: dropout p=%g >= 0.96", (double)drop.p);
  if (B == 0 || K == 0) return B200CTR_OK;
  int rc;
#define CALL(NT_) ha_launch_fwd<NT_>((const float*)qkv, qkv_f16, relmean, max_len, B, K, (float*)o, o_bf16, lse, keep, drop, as_stream(stream))
  HA_DISPATCH(CALL)
#undef CALL
  if (rc) return rc;
  B200_LAUNCH_CHECK("attn_h16_fwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_attn_h16_bwd(const void* qkv, int qkv_f16, const float* relmean, int max_len, const void* o, int o_bf16,
                                    const float* lse, const void* d_o, int do_bf16, int64_t B, int K, int H, int D, void* dqkv, int dqkv_bf16, float* drelmean,
                                    const uint32_t* keep, b200ctr_dropout_t drop, void* stream) {
  B200_REQUIRE(qkv && o && lse && d_o && dqkv, B200CTR_E_INVALID_ARG, "attn_h16_bwd: null pointer");
  B200_REQUIRE(D == HA_D && H == HA_H, B200CTR_E_UNSUPPORTED_SHAPE,
               "attn_h16_bwd: only emb_dim=64 with 8 heads of head_dim=8 is built (got D=%d, H=%d)", D, H);
  B200_REQUIRE(drop.p < 0.96f, B200CTR_E_INVALID_ARG, "attn_h16_bwd: dropout p=%g >= 0.96", (double)drop.p);
  if (B == 0 || K == 0) return B200CTR_OK;
  int rc;
#define CALL(NT_) ha_launch_bwd<NT_>((const float*)qkv, qkv_f16, relmean, max_len, (const float*)o, lse, (const float*)d_o, (o_bf16 ? 1 : 0) | (do_bf16 ? 2 : 0), B, K, (float*)dqkv, dqkv_bf16, drelmean, keep, drop, as_stream(stream))
  HA_DISPATCH(CALL)
#undef CALL
  if (rc) return rc;
  B200_LAUNCH_CHECK("attn_h16_bwd_kernel");
  return B200CTR_OK;
}
