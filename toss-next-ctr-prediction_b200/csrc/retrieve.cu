// K2: DARE target-aware retrieval (reference src/models/dare.py:116-138).
//
// forward (retrieve_fwd_warp_kernel), ONE WARP per sample, no CTA-wide barrier anywhere:
//   1. the (L,) token row is staged in the warp's slice of shared memory (coalesced);
//   2. every non-pad position's emb_att row (D=64 fp32 = 256 B) is fetched by one half-warp as 16 x 128-bit
//      no-allocate loads, 8 rows in flight per half-warp; pad positions are never fetched (score = -1e9);
//   3. score = dot(row, q) in a FIXED order (per-lane 4-term fma chain, then xor-tree 8,4,2,1) + rec[p];
//   4. scores and positions become 64-bit keys (~orderable(score) << 32 | position), ceil(L/32) per lane in
//      registers; the K-th smallest key is found by an MSB-first 8-bit radix select (warp-private histogram),
//      the K selected keys are compacted and sorted by a register/shuffle bitonic network => the canonical
//      order (score desc, position asc).  While one warp selects, the other warps of the SM keep gathering, so
//      the kernel stays on the HBM roofline instead of the old CTA-wide 512-key bitonic sort's issue limit;
//   5. only the K selected emb_rep rows are gathered and written with vals / idx / tok.
// (retrieve_fwd_kernel, one CTA per sample with a full bitonic sort, remains for max_len > 640 or top_k > 256.)
// Algorithmic bytes per sample: 8L (ids) + 4D*l (att rows) + 4D*Kv (rep rows) + 4D (q) + 4D*K + 12K (outputs).
//
// Row-sharded tables (data parallel, one process per GPU): row r lives on shard r % n at local row r / n; the
// shard base pointers may be PEER memory (NVLink), so the gather and the gradient scatter cross GPUs inside
// these kernels -- no staging copy, no collective (n is a power of two; n = 1 is the single-GPU layout).
//
// backward, one CTA per sample: dq = sum_k dvals_k * emb_att[tok_k]; table gradients are row-sparse
// (K rows per table per sample) and accumulated with 128-bit vector atomics (into the owner's shard).
#include "common.cuh"
#include <stdlib.h>
#include <algorithm>

namespace b200ctr {

constexpr int RT_D = 64;
constexpr int RT_THREADS = 256;
constexpr float RT_PAD_SCORE = -1e9f;
constexpr int RT_MAX_SHARDS = 8;

struct RtShards {
  float* p[RT_MAX_SHARDS];
  int mask, shift;                       // shard = row & mask ; local row = row >> shift
  int ld_mode;                           // 0: ld.global.nc.L1::no_allocate (streaming), 1: plain ld.global (peer shards)
  __device__ __forceinline__ float4* row4(int t) const { return reinterpret_cast<float4*>(p[t & mask]) + (int64_t)(t >> shift) * (RT_D / 4); }
};

__device__ __forceinline__ float4 rt_load_row(const RtShards& sh, int t, int hl) {
  const float4* p = sh.row4(t) + hl;
  if (sh.ld_mode == 0) return ldg_nc_f4(p);
  float4 r;
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

__device__ __forceinline__ uint32_t orderable(float s) {
  uint32_t u = __float_as_uint(s);
  return (u >> 31) ? ~u : (u ^ 0x80000000u);
}

template <typename IT>
__global__ void __launch_bounds__(RT_THREADS)
retrieve_fwd_kernel(const IT* __restrict__ seq, const float* __restrict__ q, const __grid_constant__ RtShards emb_att,
                    const __grid_constant__ RtShards emb_rep, const float* __restrict__ rec, int L, int LP, int K, int pad_id,
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
      if (t[u] != pad_id) r[u] = rt_load_row(emb_att, t[u], hl);
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
  float4* out4 = reinterpret_cast<float4*>(sel_rep + b * (int64_t)K * RT_D);
  for (int k0 = hw; k0 < K; k0 += NHW * 4) {
    float4 r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int k = k0 + u * NHW;
      r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) {
        int t = toks[(int)(keys[k] & 0xffffffffu)];
        if (t != pad_id) r[u] = rt_load_row(emb_rep, t, hl);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      int k = k0 + u * NHW;
      if (k < K) out4[k * (RT_D / 4) + hl] = r[u];
    }
  }
}


__device__ __forceinline__ float unorderable(uint32_t v) {      // inverse of orderable()
  return __uint_as_float((v >> 31) ? (v ^ 0x80000000u) : ~v);
}
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  const uint32_t lo = __shfl_xor_sync(0xffffffffu, (uint32_t)v, m), hi = __shfl_xor_sync(0xffffffffu, (uint32_t)(v >> 32), m);
  return ((unsigned long long)hi << 32) | lo;
}

// NKEYS = ceil(L / 32) keys per lane; R = sorted output slots per lane (32 * R >= K)
template <typename IT, int NKEYS, int R>
__global__ void __launch_bounds__(RT_THREADS)
retrieve_fwd_warp_kernel(const IT* __restrict__ seq, const float* __restrict__ q, const __grid_constant__ RtShards emb_att,
                         const __grid_constant__ RtShards emb_rep, const float* __restrict__ rec, int64_t B, int L, int K, int pad_id,
                         int64_t vocab, float* __restrict__ sel_rep, float* __restrict__ vals, int32_t* __restrict__ idx,
                         int32_t* __restrict__ tok_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int KP = 32 * R;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, hl = lane & 15, hw = lane >> 4;
  const int64_t b = (int64_t)blockIdx.x * (RT_THREADS / 32) + warp;
  if (b >= B) return;                                           // warp-uniform; no CTA barrier is used below
  const int LA = (L + 3) & ~3;
  unsigned char* mine = smem_raw + (size_t)warp * ((size_t)LA * 8 + 1024 + (size_t)KP * 8);
  int32_t* toks = reinterpret_cast<int32_t*>(mine);                                   // L
  float* scores = reinterpret_cast<float*>(mine + (size_t)LA * 4);                    // L
  int* hist = reinterpret_cast<int*>(mine + (size_t)LA * 8);                          // 256
  unsigned long long* selk = reinterpret_cast<unsigned long long*>(mine + (size_t)LA * 8 + 1024);   // KP

  for (int p = lane; p < L; p += 32) {
    int64_t t = (int64_t)seq[b * L + p];
    t = t < 0 ? 0 : (t >= vocab ? vocab - 1 : t);     // memory-safe clamp (reference would raise on OOB)
    toks[p] = (int32_t)t;
  }
  const float4 qv = *reinterpret_cast<const float4*>(q + b * RT_D + hl * 4);
  __syncwarp();

  // ---- gather + score: each half-warp owns 8 consecutive positions per iteration, all 8 row loads in flight ----
  for (int base = 0; base < L; base += 16) {
    float4 r[8];
    int t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = base + hw * 8 + u;
      t[u] = (p < L) ? toks[p] : pad_id;
      r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t[u] != pad_id) r[u] = rt_load_row(emb_att, t[u], hl);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int p = base + hw * 8 + u;
      float d = r[u].x * qv.x;
      d = fmaf(r[u].y, qv.y, d);
      d = fmaf(r[u].z, qv.z, d);
      d = fmaf(r[u].w, qv.w, d);
      d += __shfl_xor_sync(0xffffffffu, d, 8);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      if (hl == 0 && p < L) scores[p] = (t[u] != pad_id) ? d + rec[p] : RT_PAD_SCORE;
    }
  }
  __syncwarp();

  // ---- keys: position p = r*32 + lane; smaller key = better (score desc, position asc) ----
  unsigned long long key[NKEYS];
#pragma unroll
  for (int r = 0; r < NKEYS; ++r) {
    const int p = r * 32 + lane;
    key[r] = p < L ? (((unsigned long long)(~orderable(scores[p])) << 32) | (unsigned)p) : ~0ULL;
  }
  // ---- radix select of the K-th smallest key (MSB first, 8 bits per pass) ----
  unsigned long long prefix = 0ULL, thr = ~0ULL;
  int k_rem = K;
#pragma unroll 1
  for (int shift = 56; shift >= 0; shift -= 8) {
#pragma unroll
    for (int i = 0; i < 8; ++i) hist[lane * 8 + i] = 0;
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NKEYS; ++r) {
      const bool match = (shift == 56) || ((key[r] >> (shift + 8)) == (prefix >> (shift + 8)));
      if (match && key[r] != ~0ULL) atomicAdd(&hist[(int)((key[r] >> shift) & 255ULL)], 1);
    }
    __syncwarp();
    int c[8], sum = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { c[i] = hist[lane * 8 + i]; sum += c[i]; }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const int excl = incl - sum;
    const uint32_t hit = __ballot_sync(0xffffffffu, excl < k_rem && k_rem <= incl);
    const int src = __ffs(hit) - 1;                  // exactly one lane (K <= number of real keys)
    int bin = 0, below = excl, cnt = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (cnt == 0) {
        if (below + c[i] >= k_rem) { bin = lane * 8 + i; cnt = c[i]; }
        else below += c[i];
      }
    }
    bin = __shfl_sync(0xffffffffu, bin, src);
    below = __shfl_sync(0xffffffffu, below, src);
    cnt = __shfl_sync(0xffffffffu, cnt, src);
    k_rem -= below;
    prefix |= (unsigned long long)bin << shift;
    __syncwarp();
    if (cnt == k_rem) {                              // every key of this bin is selected: threshold found
      thr = shift ? (prefix | ((1ULL << shift) - 1ULL)) : prefix;
      break;
    }
  }
  // ---- compact the K selected keys, then sort them (ascending key) in registers ----
  for (int i = lane; i < KP; i += 32) selk[i] = ~0ULL;
  __syncwarp();
  int basec = 0;
#pragma unroll
  for (int r = 0; r < NKEYS; ++r) {
    const bool take = key[r] <= thr && key[r] != ~0ULL;
    const uint32_t bal = __ballot_sync(0xffffffffu, take);
    if (take) selk[basec + __popc(bal & ((1u << lane) - 1u))] = key[r];
    basec += __popc(bal);
  }
  __syncwarp();
  unsigned long long sk[R];                           // element index i = lane * R + r
#pragma unroll
  for (int r = 0; r < R; ++r) sk[r] = selk[lane * R + r];
#pragma unroll
  for (int k = 2; k <= KP; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= R) {                                   // partner in another lane
        const int lm = j / R;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const unsigned long long other = shfl_xor_u64(sk[r], lm);
          const int i = lane * R + r;
          const bool up = (i & k) == 0, lower = (i & j) == 0;
          const bool take_min = (up == lower);
          sk[r] = take_min ? (other < sk[r] ? other : sk[r]) : (other > sk[r] ? other : sk[r]);
        }
      } else {                                        // partner in this lane's registers
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if ((r & j) == 0) {
            const int i = lane * R + r;
            const bool up = (i & k) == 0;
            const unsigned long long a = sk[r], c2 = sk[r | j];
            if ((a > c2) == up) { sk[r] = c2; sk[r | j] = a; }
          }
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) selk[lane * R + r] = sk[r];
  __syncwarp();
  for (int k = lane; k < K; k += 32) {
    const unsigned long long kk = selk[k];
    const int p = (int)(kk & 0xffffffffULL);
    vals[b * K + k] = unorderable(~(uint32_t)(kk >> 32));
    idx[b * K + k] = p;
    tok_out[b * K + k] = toks[p];
  }
  float4* out4 = reinterpret_cast<float4*>(sel_rep + b * (int64_t)K * RT_D);
  for (int k0 = 0; k0 < K; k0 += 16) {
    float4 r[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + hw * 8 + u;
      r[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) {
        const int t = toks[(int)(selk[k] & 0xffffffffULL)];
        if (t != pad_id) r[u] = rt_load_row(emb_rep, t, hl);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + hw * 8 + u;
      if (k < K) out4[k * (RT_D / 4) + hl] = r[u];
    }
  }
}

// Sharded + staged: gradient rows of tokens owned by ANOTHER rank are not reduced remotely (fp32 vector atomics over
// NVLink measured 3.8 ms per step at N = 8); they are written with plain 128-bit stores into the owner's staging buffer
// stage[owner][my_rank][slot] and the owner adds them locally afterwards (stage_apply_kernel).
__global__ void __launch_bounds__(128)
retrieve_bwd_kernel(const int32_t* __restrict__ tok, const float* __restrict__ q, const __grid_constant__ RtShards emb_att,
                    const float* __restrict__ dsel_rep, const float* __restrict__ dvals, int K, int pad_id,
                    float* __restrict__ dq, const __grid_constant__ RtShards datt, const __grid_constant__ RtShards drep, int has_datt, int has_drep,
                    const __grid_constant__ RtShards stage_att, const __grid_constant__ RtShards stage_rep, int staged, int my_rank,
                    int64_t stage_cap) {
  __shared__ float4 red[128];
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x, hl = tid & 15, hw = tid >> 4;
  constexpr int NHW = 8;
  const float4 qv = *reinterpret_cast<const float4*>(q + b * RT_D + hl * 4);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int k = hw; k < K; k += NHW) {
    int t = tok[b * K + k];
    if (t == pad_id) continue;                      // masked_fill blocks the gradient; padding_idx row gets none
    float dv = dvals[b * K + k];
    float4 a = rt_load_row(emb_att, t, hl);
    acc.x = fmaf(dv, a.x, acc.x); acc.y = fmaf(dv, a.y, acc.y); acc.z = fmaf(dv, a.z, acc.z); acc.w = fmaf(dv, a.w, acc.w);
    const int owner = t & datt.mask;
    const bool remote = staged && (staged == 2 || owner != my_rank);     // staged == 2: the owner's own slots are staged too
    const int64_t srow = ((int64_t)my_rank * stage_cap + (b * K + k)) * (RT_D / 4) + hl;
    if (has_datt) {
      float4 g = make_float4(dv * qv.x, dv * qv.y, dv * qv.z, dv * qv.w);
      if (remote) reinterpret_cast<float4*>(stage_att.p[owner])[srow] = g;
      else if (staged != 2) atomicAdd(datt.row4(t) + hl, g);
    }
    if (has_drep) {
      float4 g = *reinterpret_cast<const float4*>(dsel_rep + (b * K + k) * (int64_t)RT_D + hl * 4);
      if (remote) reinterpret_cast<float4*>(stage_rep.p[owner])[srow] = g;
      else if (staged != 2) atomicAdd(drep.row4(t) + hl, g);
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

// owner side: add the rows the other ranks staged for tokens this rank owns (half-warp per (source rank, slot))
__global__ void __launch_bounds__(256)
stage_apply_kernel(const int32_t* __restrict__ gtok, int n_src, int64_t n_slots, int64_t cap, int pad_id, int mask, int shift,
                   int rank, const float4* __restrict__ stage_att, const float4* __restrict__ stage_rep,
                   float* __restrict__ g_att, float* __restrict__ g_rep) {
  const int hl = threadIdx.x & 15;
  int64_t hw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  const int64_t nhw = ((int64_t)gridDim.x * blockDim.x) >> 4;
  const int64_t total = (int64_t)n_src * n_slots;
  for (; hw < total; hw += nhw) {
    const int src = (int)(hw / n_slots);
    const int64_t i = hw - (int64_t)src * n_slots;
    if (src == rank) continue;
    const int t = gtok[hw];
    if (t == pad_id || (t & mask) != rank) continue;
    const int64_t lr = (int64_t)(t >> shift);
    const int64_t srow = ((int64_t)src * cap + i) * (RT_D / 4) + hl;
    if (g_att) atomicAdd(reinterpret_cast<float4*>(g_att + lr * RT_D) + hl, stage_att[srow]);
    if (g_rep) atomicAdd(reinterpret_cast<float4*>(g_rep + lr * RT_D) + hl, stage_rep[srow]);
  }
}

}  // namespace b200ctr

using namespace b200ctr;

static int make_shards(RtShards& out, const float* const* ptrs, int n, const char* what) {
  if (n < 1 || n > RT_MAX_SHARDS || (n & (n - 1))) {
    set_error("%s: n_shards=%d unsupported (power of two, 1..%d)", what, n, RT_MAX_SHARDS);
    return B200CTR_E_UNSUPPORTED_SHAPE;
  }
  int shift = 0;
  while ((1 << shift) < n) ++shift;
  out.mask = n - 1; out.shift = shift;
  static const int env_mode = getenv("B200CTR_RT_LDMODE") ? atoi(getenv("B200CTR_RT_LDMODE")) : -1;
  // measured on 2 x B200 (profiles/README.md): the non-coherent no-allocate path reaches only ~175 GB/s on PEER memory,
  // plain ld.global ~800 GB/s; locally the streaming path is the faster one
  out.ld_mode = env_mode >= 0 ? env_mode : (n > 1 ? 1 : 0);
  for (int i = 0; i < RT_MAX_SHARDS; ++i) out.p[i] = const_cast<float*>(ptrs && i < n ? ptrs[i] : nullptr);
  for (int i = 0; i < n; ++i)
    if (!out.p[i]) { set_error("%s: shard %d is NULL", what, i); return B200CTR_E_INVALID_ARG; }
  return B200CTR_OK;
}

template <typename IT>
static int launch_retrieve_fwd(const IT* seq, const float* q, const RtShards& att, const RtShards& rep, const float* rec, int64_t B,
                               int L, int K, int pad_id, int64_t vocab, float* sel_rep, float* vals, int32_t* idx, int32_t* tok,
                               cudaStream_t st) {
  const int nkeys = (L + 31) / 32;
  static const bool force_cta = getenv("B200CTR_RETRIEVE_CTA") != nullptr;
  if (nkeys <= 20 && K <= 256 && !force_cta) {
    const int R = K <= 64 ? 2 : (K <= 128 ? 4 : 8);
    const int LA = (L + 3) & ~3;
    const size_t smem = (size_t)(RT_THREADS / 32) * ((size_t)LA * 8 + 1024 + (size_t)32 * R * 8);
    const unsigned grid = (unsigned)ceil_div(B, (int64_t)(RT_THREADS / 32));
#define RT_LAUNCH(NK_, R_)                                                                                           \
    do {                                                                                                             \
      auto fn = retrieve_fwd_warp_kernel<IT, NK_, R_>;                                                               \
      if (smem > 48 * 1024) {                                                                                        \
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);           \
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(retrieve_fwd smem)");                        \
      }                                                                                                              \
      fn<<<grid, RT_THREADS, smem, st>>>(seq, q, att, rep, rec, B, L, K, pad_id, vocab, sel_rep, vals, idx, tok);    \
    } while (0)
#define RT_LAUNCH_R(NK_) do { if (R == 2) RT_LAUNCH(NK_, 2); else if (R == 4) RT_LAUNCH(NK_, 4); else RT_LAUNCH(NK_, 8); } while (0)
    if (nkeys <= 4) RT_LAUNCH_R(4);
    else if (nkeys <= 8) RT_LAUNCH_R(8);
    else if (nkeys <= 13) RT_LAUNCH_R(13);
    else if (nkeys <= 16) RT_LAUNCH_R(16);
    else RT_LAUNCH_R(20);
#undef RT_LAUNCH_R
#undef RT_LAUNCH
    return B200CTR_OK;
  }
  int LP = 2;
  while (LP < L) LP <<= 1;
  size_t smem = (size_t)LP * 8 + (size_t)L * 8;
  retrieve_fwd_kernel<IT><<<(unsigned)B, RT_THREADS, smem, st>>>(seq, q, att, rep, rec, L, LP, K, pad_id, vocab, sel_rep, vals, idx, tok);
  return B200CTR_OK;
}

extern "C" int b200ctr_retrieve_fwd_sharded(const void* seq, int idx_is_64, const float* q, const float* const* att_shards,
                                            const float* const* rep_shards, int n_shards, const float* rec, int64_t B, int L,
                                            int K, int D, int pad_id, int64_t vocab, float* sel_rep, float* vals, int32_t* idx,
                                            int32_t* tok, void* stream) {
  B200_REQUIRE(seq && q && att_shards && rep_shards && rec && sel_rep && vals && idx && tok, B200CTR_E_INVALID_ARG,
               "retrieve_fwd: null pointer");
  B200_REQUIRE(D == RT_D, B200CTR_E_UNSUPPORTED_SHAPE, "retrieve_fwd: emb_dim=%d unsupported (kernel is specialised for 64)", D);
  B200_REQUIRE(L >= 1 && L <= 4096, B200CTR_E_UNSUPPORTED_SHAPE, "retrieve_fwd: max_len=%d unsupported (1..4096)", L);
  B200_REQUIRE(K >= 1 && K <= L, B200CTR_E_INVALID_ARG, "retrieve_fwd: need 1 <= K <= L (K=%d, L=%d)", K, L);
  RtShards att, rep;
  int rc = make_shards(att, att_shards, n_shards, "retrieve_fwd(emb_att)");
  if (rc) return rc;
  rc = make_shards(rep, rep_shards, n_shards, "retrieve_fwd(emb_rep)");
  if (rc) return rc;
  if (B == 0) return B200CTR_OK;
  cudaStream_t st = as_stream(stream);
  rc = idx_is_64 ? launch_retrieve_fwd<int64_t>((const int64_t*)seq, q, att, rep, rec, B, L, K, pad_id, vocab, sel_rep, vals, idx, tok, st)
                 : launch_retrieve_fwd<int32_t>((const int32_t*)seq, q, att, rep, rec, B, L, K, pad_id, vocab, sel_rep, vals, idx, tok, st);
  if (rc) return rc;
  B200_LAUNCH_CHECK("retrieve_fwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_retrieve_fwd(const void* seq, int idx_is_64, const float* q, const float* emb_att,
                                    const float* emb_rep, const float* rec, int64_t B, int L, int K, int D, int pad_id,
                                    int64_t vocab, float* sel_rep, float* vals, int32_t* idx, int32_t* tok, void* stream) {
  B200_REQUIRE(emb_att && emb_rep, B200CTR_E_INVALID_ARG, "retrieve_fwd: null pointer");
  return b200ctr_retrieve_fwd_sharded(seq, idx_is_64, q, &emb_att, &emb_rep, 1, rec, B, L, K, D, pad_id, vocab, sel_rep, vals, idx,
                                      tok, stream);
}

extern "C" int b200ctr_retrieve_bwd_staged(const int32_t* tok, const float* q, const float* const* att_shards, int n_shards,
                                           const float* dsel_rep, const float* dvals, int64_t B, int K, int D, int pad_id,
                                           float* dq, float* const* datt_shards, float* const* drep_shards,
                                           float* const* stage_att, float* const* stage_rep, int my_rank, int64_t stage_cap,
                                           int stage_all, void* stream) {
  B200_REQUIRE(tok && q && att_shards && dvals && dq, B200CTR_E_INVALID_ARG, "retrieve_bwd: null pointer");
  B200_REQUIRE(!drep_shards || dsel_rep, B200CTR_E_INVALID_ARG, "retrieve_bwd: drep without dsel_rep");
  B200_REQUIRE(D == RT_D, B200CTR_E_UNSUPPORTED_SHAPE, "retrieve_bwd: emb_dim=%d unsupported", D);
  RtShards att, datt{}, drep{}, sa{}, sr{};
  int rc = make_shards(att, att_shards, n_shards, "retrieve_bwd(emb_att)");
  if (rc) return rc;
  if (datt_shards) { rc = make_shards(datt, datt_shards, n_shards, "retrieve_bwd(datt)"); if (rc) return rc; }
  if (drep_shards) { rc = make_shards(drep, drep_shards, n_shards, "retrieve_bwd(drep)"); if (rc) return rc; }
  else { datt.mask = datt_shards ? datt.mask : n_shards - 1; }
  const int staged = (stage_att || stage_rep) ? (stage_all ? 2 : 1) : 0;
  if (staged) {
    B200_REQUIRE(n_shards > 1 && my_rank >= 0 && my_rank < n_shards && stage_cap >= B * K, B200CTR_E_INVALID_ARG,
                 "retrieve_bwd: staging needs n_shards > 1, a valid rank and stage_cap >= B*K");
    B200_REQUIRE((!datt_shards || stage_att) && (!drep_shards || stage_rep), B200CTR_E_INVALID_ARG,
                 "retrieve_bwd: staging buffers missing for a requested gradient");
    if (stage_att) { rc = make_shards(sa, stage_att, n_shards, "retrieve_bwd(stage_att)"); if (rc) return rc; }
    if (stage_rep) { rc = make_shards(sr, stage_rep, n_shards, "retrieve_bwd(stage_rep)"); if (rc) return rc; }
  }
  if (!datt_shards) datt.mask = n_shards - 1;          // the owner test in the kernel reads datt.mask
  if (B == 0) return B200CTR_OK;
  retrieve_bwd_kernel<<<(unsigned)B, 128, 0, as_stream(stream)>>>(tok, q, att, dsel_rep, dvals, K, pad_id, dq, datt, drep,
                                                                   datt_shards ? 1 : 0, drep_shards ? 1 : 0, sa, sr, staged, my_rank,
                                                                   stage_cap);
  B200_LAUNCH_CHECK("retrieve_bwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_retrieve_bwd_sharded(const int32_t* tok, const float* q, const float* const* att_shards, int n_shards,
                                            const float* dsel_rep, const float* dvals, int64_t B, int K, int D, int pad_id,
                                            float* dq, float* const* datt_shards, float* const* drep_shards, void* stream) {
  return b200ctr_retrieve_bwd_staged(tok, q, att_shards, n_shards, dsel_rep, dvals, B, K, D, pad_id, dq, datt_shards, drep_shards,
                                     nullptr, nullptr, 0, 0, 0, stream);
}

extern "C" int b200ctr_stage_apply(const int32_t* gtok, int n_src, int64_t n_slots, int64_t stage_cap, int pad_id, int n_shards,
                                   int rank, const float* stage_att, const float* stage_rep, float* g_att, float* g_rep, int D,
                                   void* stream) {
  B200_REQUIRE(gtok && (!g_att || stage_att) && (!g_rep || stage_rep), B200CTR_E_INVALID_ARG, "stage_apply: null pointer");
  B200_REQUIRE(D == RT_D, B200CTR_E_UNSUPPORTED_SHAPE, "stage_apply: D=%d unsupported", D);
  B200_REQUIRE(n_shards >= 1 && n_shards <= RT_MAX_SHARDS && !(n_shards & (n_shards - 1)) && rank >= 0 && rank < n_shards &&
               stage_cap >= n_slots, B200CTR_E_INVALID_ARG, "stage_apply: bad shard arguments");
  if (n_src == 0 || n_slots == 0) return B200CTR_OK;
  int shift = 0;
  while ((1 << shift) < n_shards) ++shift;
  const int64_t total = (int64_t)n_src * n_slots;
  const int grid = (int)std::min<int64_t>(ceil_div(total, 16), (int64_t)num_sms() * 16);
  stage_apply_kernel<<<grid, 256, 0, as_stream(stream)>>>(gtok, n_src, n_slots, stage_cap, pad_id, n_shards - 1, shift, rank,
                                                           reinterpret_cast<const float4*>(stage_att),
                                                           reinterpret_cast<const float4*>(stage_rep), g_att, g_rep);
  B200_LAUNCH_CHECK("stage_apply_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_retrieve_bwd(const int32_t* tok, const float* q, const float* emb_att, const float* dsel_rep,
                                    const float* dvals, int64_t B, int K, int D, int pad_id, float* dq, float* datt_dense,
                                    float* drep_dense, void* stream) {
  B200_REQUIRE(emb_att, B200CTR_E_INVALID_ARG, "retrieve_bwd: null pointer");
  return b200ctr_retrieve_bwd_sharded(tok, q, &emb_att, 1, dsel_rep, dvals, B, K, D, pad_id, dq, datt_dense ? &datt_dense : nullptr,
                                      drep_dense ? &drep_dense : nullptr, stream);
}
