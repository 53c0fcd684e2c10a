// Row-lazy, replay-exact AdamW + EMA for the two seq_vocab x 64 sequence tables.
//
// The reference runs dense torch.optim.AdamW + ModelEMA over all 10,000,000 rows of both tables every step
// (src/train.py:138,193-199; src/utils/ema.py:91-131): weight decay, moment decay and the EMA recurrence touch
// 1.28 G parameters although only <= B*K rows receive a gradient (SURVEY hard part 2, ~46 GB of traffic/step even
// when fused).  Here every row carries `state = (last_step << 1) | has_moments` and is brought up to date lazily:
//   * rows about to be READ by a forward pass (every token of the batch) are caught up to step t-1 first, so the
//     forward sees exactly the values dense AdamW would have produced;
//   * rows that received a gradient get step t applied (g scaled by the global clip coefficient);
//   * a dense flush brings all rows to the current step before state_dict / evaluation.
// Catch-up of a row whose moments are zero (never touched) is closed-form from per-step prefix tables
// (C = prod(1 - lr*wd), S = EMA of C, logD = sum log d); a row with non-zero moments REPLAYS the skipped steps with the
// dense kernel's own per-element update (g = 0), so touched rows are bit-identical to the dense path.
// History entry per step s (float4): {1 - lr_s*wd, lr_s / (1 - b1^s), sqrt(1 - b2^s), ema_decay_s}.
#include "common.cuh"
#include <algorithm>

namespace b200ctr {

constexpr int RO_D = 64;
constexpr int RO_REPLAY_MAX = 256;     // skipped steps replayed exactly before the closed-form tail (see row_catchup)

struct RowHist {
  const float4* h;      // [steps+1]  per-step constants
  const float* C;       // [steps+1]  C[0] = 1
  const float* S;       // [steps+1]  S[0] = 1
  const float* logD;    // [steps+1]  logD[0] = 0
  float beta1, beta2, eps;
};

// bring one row (2 elements per lane) from `last` to `target` with zero gradient
__device__ __forceinline__ void row_catchup(float2& p, float2& m, float2& v, float2& e, bool has_ema, bool has_mom,
                                            int last, int target, const RowHist& H) {
  if (target <= last) return;
  if (!has_mom) {
    const float ca = H.C[last], cb = H.C[target];
    const float ratio = cb / ca;
    if (has_ema) {
      const float Dm = expf(H.logD[target] - H.logD[last]);
      const float w = (H.S[target] - H.S[last] * Dm) / ca;
      e.x = e.x * Dm + p.x * w;
      e.y = e.y * Dm + p.y * w;
    }
    p.x *= ratio; p.y *= ratio;
    return;
  }
  // Replay is exact for the first RO_REPLAY_MAX skipped steps.  Beyond that the moments have decayed by beta1^k with the
  // update ratio m/sqrt(v) shrinking as (beta1/sqrt(beta2))^k < 0.9005^k: the sum of ALL further updates is below
  // 3e-11 * lr (k > 256), far under one fp32 ulp of any weight, so the tail is the zero-moment closed form (weight decay
  // product + EMA prefix) plus the plain geometric decay of m and v.  Keeps a row that comes back after thousands of steps
  // (the Zipf tail in real training) from serialising its warp on an O(gap) loop of sqrt + divide.
  const int exact_to = (target - last > RO_REPLAY_MAX) ? last + RO_REPLAY_MAX : target;
  for (int s = last + 1; s <= exact_to; ++s) {
    const float4 c = __ldg(H.h + s);
    p.x *= c.x; p.y *= c.x;
    m.x = m.x + (0.f - m.x) * (1.f - H.beta1);   // lerp_(g = 0, 1 - b1), written exactly as the dense kernel computes it
    m.y = m.y + (0.f - m.y) * (1.f - H.beta1);
    v.x *= H.beta2; v.y *= H.beta2;
    p.x -= c.y * (m.x / (sqrtf(v.x) / c.z + H.eps));
    p.y -= c.y * (m.y / (sqrtf(v.y) / c.z + H.eps));
    if (has_ema) { e.x = e.x * c.w + p.x * (1.f - c.w); e.y = e.y * c.w + p.y * (1.f - c.w); }
  }
  if (exact_to < target) {
    const int rest = target - exact_to;
    const float ca = H.C[exact_to], cb = H.C[target];
    const float ratio = cb / ca;
    if (has_ema) {
      const float Dm = expf(H.logD[target] - H.logD[exact_to]);
      const float w = (H.S[target] - H.S[exact_to] * Dm) / ca;
      e.x = e.x * Dm + p.x * w;
      e.y = e.y * Dm + p.y * w;
    }
    p.x *= ratio; p.y *= ratio;
    const float d1 = powf(H.beta1, (float)rest), d2 = powf(H.beta2, (float)rest);
    m.x *= d1; m.y *= d1; v.x *= d2; v.y *= d2;
  }
}

__global__ void rowopt_push_kernel(float4* h, float* C, float* S, float* logD, int t, float4 entry, float c, float s, float ld) {
  h[t] = entry; C[t] = c; S[t] = s; logD[t] = ld;
}

template <typename IT>
__global__ void rows_claim_any_kernel(const IT* __restrict__ tok, int64_t n, int pad_id, int64_t vocab, int32_t* __restrict__ owner) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t t = (int64_t)tok[i];
  if (t != pad_id && t >= 0 && t < vocab) owner[t] = (int32_t)i;
}

// warp per slot; only the slot that owns its row does work
template <typename IT>
__global__ void __launch_bounds__(256)
rowopt_catchup_kernel(const IT* __restrict__ tok, int64_t n, int pad_id, int64_t vocab, const int32_t* __restrict__ owner,
                      int32_t* __restrict__ state, float* __restrict__ P, float* __restrict__ Mo, float* __restrict__ V,
                      float* __restrict__ E, int target, RowHist H) {
  const int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; w < n; w += nw) {
    const int64_t t = (int64_t)tok[w];
    if (t == pad_id || t < 0 || t >= vocab) continue;
    if (owner[t] != (int32_t)w) continue;
    const int32_t st = state[t];
    const int last = st >> 1;
    if (last >= target) continue;
    const bool has_mom = st & 1;
    float2* p2 = reinterpret_cast<float2*>(P + t * RO_D) + lane;
    float2 p = *p2, m = make_float2(0.f, 0.f), v = m, e = m;
    if (has_mom) { m = reinterpret_cast<float2*>(Mo + t * RO_D)[lane]; v = reinterpret_cast<float2*>(V + t * RO_D)[lane]; }
    if (E) e = reinterpret_cast<float2*>(E + t * RO_D)[lane];
    row_catchup(p, m, v, e, E != nullptr, has_mom, last, target, H);
    *p2 = p;
    if (has_mom) { reinterpret_cast<float2*>(Mo + t * RO_D)[lane] = m; reinterpret_cast<float2*>(V + t * RO_D)[lane] = v; }
    if (E) reinterpret_cast<float2*>(E + t * RO_D)[lane] = e;
    __syncwarp();
    if (lane == 0) state[t] = (target << 1) | (st & 1);
  }
}

// apply step t to every distinct row listed in tok (owner slot only): catch up to t-1, then the AdamW step with the
// accumulated gradient row (scaled by the clip coefficient), EMA, and re-zero the gradient row.
__global__ void __launch_bounds__(256)
rowopt_update_kernel(const int32_t* __restrict__ tok, int64_t n, int pad_id, const int32_t* __restrict__ owner,
                     int32_t* __restrict__ state, float* __restrict__ P, float* __restrict__ G, float* __restrict__ Mo,
                     float* __restrict__ V, float* __restrict__ E, int t_step, const float* __restrict__ coef, RowHist H) {
  const int lane = threadIdx.x & 31;
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float c = coef ? coef[0] : 1.f;
  const float4 hc = __ldg(H.h + t_step);
  for (; w < n; w += nw) {
    const int t = tok[w];
    if (t == pad_id || owner[t] != (int32_t)w) continue;
    const int32_t st = state[t];
    const bool has_mom = st & 1;
    float2* p2 = reinterpret_cast<float2*>(P + (int64_t)t * RO_D) + lane;
    float2* g2 = reinterpret_cast<float2*>(G + (int64_t)t * RO_D) + lane;
    float2* m2 = reinterpret_cast<float2*>(Mo + (int64_t)t * RO_D) + lane;
    float2* v2 = reinterpret_cast<float2*>(V + (int64_t)t * RO_D) + lane;
    float2 p = *p2, g = *g2, m = make_float2(0.f, 0.f), v = m, e = m;
    if (has_mom) { m = *m2; v = *v2; }
    if (E) e = reinterpret_cast<float2*>(E + (int64_t)t * RO_D)[lane];
    row_catchup(p, m, v, e, E != nullptr, has_mom, st >> 1, t_step - 1, H);
    g.x *= c; g.y *= c;
    p.x *= hc.x; p.y *= hc.x;
    m.x = m.x + (g.x - m.x) * (1.f - H.beta1); m.y = m.y + (g.y - m.y) * (1.f - H.beta1);
    v.x = v.x * H.beta2 + (1.f - H.beta2) * g.x * g.x; v.y = v.y * H.beta2 + (1.f - H.beta2) * g.y * g.y;
    p.x -= hc.y * (m.x / (sqrtf(v.x) / hc.z + H.eps));
    p.y -= hc.y * (m.y / (sqrtf(v.y) / hc.z + H.eps));
    *p2 = p; *m2 = m; *v2 = v;
    *g2 = make_float2(0.f, 0.f);
    if (E) {
      e.x = e.x * hc.w + p.x * (1.f - hc.w); e.y = e.y * hc.w + p.y * (1.f - hc.w);
      reinterpret_cast<float2*>(E + (int64_t)t * RO_D)[lane] = e;
    }
    __syncwarp();
    if (lane == 0) state[t] = (t_step << 1) | 1;
  }
}

// dense flush: every row to `target`
__global__ void __launch_bounds__(256)
rowopt_flush_kernel(int32_t* __restrict__ state, float* __restrict__ P, float* __restrict__ Mo, float* __restrict__ V,
                    float* __restrict__ E, int64_t n_rows, int target, RowHist H) {
  const int lane = threadIdx.x & 31;
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (; r < n_rows; r += nw) {
    const int32_t st = state[r];
    const int last = st >> 1;
    if (last >= target) continue;
    const bool has_mom = st & 1;
    float2* p2 = reinterpret_cast<float2*>(P + r * RO_D) + lane;
    float2 p = *p2, m = make_float2(0.f, 0.f), v = m, e = m;
    if (has_mom) { m = reinterpret_cast<float2*>(Mo + r * RO_D)[lane]; v = reinterpret_cast<float2*>(V + r * RO_D)[lane]; }
    if (E) e = reinterpret_cast<float2*>(E + r * RO_D)[lane];
    row_catchup(p, m, v, e, E != nullptr, has_mom, last, target, H);
    *p2 = p;
    if (has_mom) { reinterpret_cast<float2*>(Mo + r * RO_D)[lane] = m; reinterpret_cast<float2*>(V + r * RO_D)[lane] = v; }
    if (E) reinterpret_cast<float2*>(E + r * RO_D)[lane] = e;
    __syncwarp();
    if (lane == 0) state[r] = (target << 1) | (st & 1);
  }
}

static RowHist make_hist(const void* h, const float* C, const float* S, const float* logD, float b1, float b2, float eps) {
  RowHist H;
  H.h = reinterpret_cast<const float4*>(h); H.C = C; H.S = S; H.logD = logD; H.beta1 = b1; H.beta2 = b2; H.eps = eps;
  return H;
}


// =============================================================================================================
// Bitmap-driven row sets.  "Rows the batch reads" and "rows that received a gradient" are kept as one bit per GLOBAL
// row (10 M rows = 1.25 MB, L2 resident): marking is one atomicOr per slot, the sets of all data-parallel ranks are
// exchanged as bitmaps (n_maps x words), and the owner of a row walks only ITS bits -- the cost no longer grows with
// the number of (duplicate-heavy) token slots of the whole job, and no per-row owner-slot election is needed.
// Row r lives on rank r % n_shards at local row r / n_shards (n_shards in {1,2,4,8}, so the owner of bit b of a word is
// b % n_shards).
// =============================================================================================================
template <typename IT>
__global__ void bitmap_mark_kernel(const IT* __restrict__ tok, int64_t n, int pad_id, int64_t vocab, uint32_t* __restrict__ bm) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t t = (int64_t)tok[i];
  if (t == pad_id) return;
  t = t < 0 ? 0 : (t >= vocab ? vocab - 1 : t);     // the same memory-safe clamp retrieve_fwd applies: the row it READS is the row marked
  if (t == pad_id) return;
  const uint32_t bit = 1u << (t & 31);
  uint32_t* w = bm + (t >> 5);
  if (!(*w & bit)) atomicOr(w, bit);              // hot rows are marked thousands of times: test first
}

__device__ __forceinline__ uint32_t own_mask(int n_shards, int rank) {
  uint32_t m = 0;
  for (int b = rank; b < 32; b += n_shards) m |= 1u << b;
  return m;
}

// Calls f(local_row, hl) for every owned row whose bit is set in any of the n_maps bitmaps.  Each HALF-warp gets its own
// row (hl = lane within the half-warp, 16 lanes x float4 = one 256-byte row), so a warp keeps two rows' dependent
// load chains (state -> p/m/v/ema) in flight; local_row < 0 = this half-warp has no row in this round.
template <typename F>
__device__ __forceinline__ void for_each_owned_row(const uint32_t* __restrict__ bm, int n_maps, int64_t words, int n_shards,
                                                   int rank, F f) {
  const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t om = own_mask(n_shards, rank);
  int shift = 0;
  while ((1 << shift) < n_shards) ++shift;
  // word (j * nw + warp) belongs to this warp: consecutive bitmap words go to different warps, so a dense run of ids
  // (frequency-sorted vocabularies put all the hot rows together) is spread over the whole grid instead of
  // serialising inside a few warps
  for (int64_t j0 = 0; j0 * nw < words; j0 += 32) {
    const int64_t wi = (j0 + lane) * nw + warp;
    uint32_t bits = 0;
    if (wi < words)
      for (int k = 0; k < n_maps; ++k) bits |= bm[(int64_t)k * words + wi];
    bits &= om;
    uint32_t active = __ballot_sync(0xffffffffu, bits != 0);
    while (active) {
      const int src = __ffs(active) - 1;
      active &= active - 1;
      uint32_t wb = __shfl_sync(0xffffffffu, bits, src);
      const int64_t base_row = ((j0 + src) * nw + warp) * 32;
      while (wb) {
        const int b0 = __ffs(wb) - 1;
        wb &= wb - 1;
        int b1 = -1;
        if (wb) { b1 = __ffs(wb) - 1; wb &= wb - 1; }
        const int bsel = half ? b1 : b0;
        f(bsel >= 0 ? ((base_row + bsel) >> shift) : (int64_t)-1, hl);
      }
    }
  }
}

// 4 elements per lane = two of the 2-element catch-up updates (identical per-element arithmetic)
__device__ __forceinline__ void row_catchup4(float4& p, float4& m, float4& v, float4& e, bool has_ema, bool has_mom, int last,
                                             int target, const RowHist& H) {
  float2 p0 = make_float2(p.x, p.y), p1 = make_float2(p.z, p.w), m0 = make_float2(m.x, m.y), m1 = make_float2(m.z, m.w);
  float2 v0 = make_float2(v.x, v.y), v1 = make_float2(v.z, v.w), e0 = make_float2(e.x, e.y), e1 = make_float2(e.z, e.w);
  row_catchup(p0, m0, v0, e0, has_ema, has_mom, last, target, H);
  row_catchup(p1, m1, v1, e1, has_ema, has_mom, last, target, H);
  p = make_float4(p0.x, p0.y, p1.x, p1.y); m = make_float4(m0.x, m0.y, m1.x, m1.y);
  v = make_float4(v0.x, v0.y, v1.x, v1.y); e = make_float4(e0.x, e0.y, e1.x, e1.y);
}

__global__ void __launch_bounds__(256, 4)
rowopt_catchup_bm_kernel(const uint32_t* __restrict__ bm, int n_maps, int64_t words, int n_shards, int rank,
                         int32_t* __restrict__ state, float* __restrict__ P, float* __restrict__ Mo, float* __restrict__ V,
                         float* __restrict__ E, int target, RowHist H) {
  for_each_owned_row(bm, n_maps, words, n_shards, rank, [&](int64_t t, int hl) {
    if (t < 0) return;
    const int32_t st = state[t];
    const int last = st >> 1;
    if (last >= target) return;
    const bool has_mom = st & 1;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* p4 = reinterpret_cast<float4*>(P + t * RO_D) + hl;
    float4 p = *p4, m = z4, v = z4, e = z4;
    if (has_mom) { m = reinterpret_cast<float4*>(Mo + t * RO_D)[hl]; v = reinterpret_cast<float4*>(V + t * RO_D)[hl]; }
    if (E) e = reinterpret_cast<float4*>(E + t * RO_D)[hl];
    row_catchup4(p, m, v, e, E != nullptr, has_mom, last, target, H);
    *p4 = p;
    if (has_mom) { reinterpret_cast<float4*>(Mo + t * RO_D)[hl] = m; reinterpret_cast<float4*>(V + t * RO_D)[hl] = v; }
    if (E) reinterpret_cast<float4*>(E + t * RO_D)[hl] = e;
    // all 16 lanes of the half-warp must have READ state[t] before lane 0 of the half publishes the new step (independent
    // thread scheduling gives no such order by itself; the two halves work on different rows and may have diverged)
    __syncwarp(0xffffu << (threadIdx.x & 16));
    if (hl == 0) state[t] = (target << 1) | (st & 1);
  });
}

__global__ void __launch_bounds__(256)
rows_sqnorm_bm_kernel(const uint32_t* __restrict__ bm, int n_maps, int64_t words, int n_shards, int rank,
                      const float* __restrict__ G, double* __restrict__ acc) {
  __shared__ double red[8];
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for_each_owned_row(bm, n_maps, words, n_shards, rank, [&](int64_t t, int hl) {
    if (t < 0) return;
    const float4 g = reinterpret_cast<const float4*>(G + t * RO_D)[hl];
    s += (double)(g.x * g.x + g.y * g.y + g.z * g.z + g.w * g.w);
  });
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; ++k) t += red[k];
    norm_acc_add(acc, t);
  }
}

__global__ void __launch_bounds__(256)
rowopt_update_bm_kernel(const uint32_t* __restrict__ bm, int n_maps, int64_t words, int n_shards, int rank,
                        int32_t* __restrict__ state, float* __restrict__ P, float* __restrict__ G, float* __restrict__ Mo,
                        float* __restrict__ V, float* __restrict__ E, int t_step, const float* __restrict__ coef, RowHist H) {
  const float c = coef ? coef[0] : 1.f;
  const float4 hc = __ldg(H.h + t_step);
  for_each_owned_row(bm, n_maps, words, n_shards, rank, [&](int64_t t, int hl) {
    if (t < 0) return;
    const int32_t st = state[t];
    const bool has_mom = st & 1;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    float4* p4 = reinterpret_cast<float4*>(P + t * RO_D) + hl;
    float4* g4 = reinterpret_cast<float4*>(G + t * RO_D) + hl;
    float4* m4 = reinterpret_cast<float4*>(Mo + t * RO_D) + hl;
    float4* v4 = reinterpret_cast<float4*>(V + t * RO_D) + hl;
    float4 p = *p4, g = *g4, m = z4, v = z4, e = z4;
    if (has_mom) { m = *m4; v = *v4; }
    if (E) e = reinterpret_cast<float4*>(E + t * RO_D)[hl];
    row_catchup4(p, m, v, e, E != nullptr, has_mom, st >> 1, t_step - 1, H);
    float pp[4] = {p.x, p.y, p.z, p.w}, gg[4] = {g.x, g.y, g.z, g.w}, mm[4] = {m.x, m.y, m.z, m.w}, vv[4] = {v.x, v.y, v.z, v.w},
          ee[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gk = gg[k] * c;
      pp[k] *= hc.x;
      mm[k] = mm[k] + (gk - mm[k]) * (1.f - H.beta1);
      vv[k] = vv[k] * H.beta2 + (1.f - H.beta2) * gk * gk;
      pp[k] -= hc.y * (mm[k] / (sqrtf(vv[k]) / hc.z + H.eps));
      ee[k] = ee[k] * hc.w + pp[k] * (1.f - hc.w);
    }
    *p4 = make_float4(pp[0], pp[1], pp[2], pp[3]);
    *m4 = make_float4(mm[0], mm[1], mm[2], mm[3]);
    *v4 = make_float4(vv[0], vv[1], vv[2], vv[3]);
    *g4 = z4;
    if (E) reinterpret_cast<float4*>(E + t * RO_D)[hl] = make_float4(ee[0], ee[1], ee[2], ee[3]);
    __syncwarp(0xffffu << (threadIdx.x & 16));
    if (hl == 0) state[t] = (t_step << 1) | 1;
  });
}

// global row id -> local row of this rank's shard (row r lives on rank r % n at local row r / n), -1 for pad / foreign rows
template <typename IT>
__global__ void rows_localize_kernel(const IT* __restrict__ tok, int64_t n, int pad_id, int mask, int shift, int rank,
                                     int32_t* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t t = (int64_t)tok[i];
  out[i] = (t == pad_id || t < 0 || (int)(t & mask) != rank) ? -1 : (int32_t)(t >> shift);
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_rowopt_push(void* hist4, float* C, float* S, float* logD, int64_t capacity, int64_t t,
                                   float decay_mul, float step_size, float bc2_sqrt, float ema_d, float C_t, float S_t,
                                   float logD_t, void* stream) {
  B200_REQUIRE(hist4 && C && S && logD, B200CTR_E_INVALID_ARG, "rowopt_push: null pointer");
  B200_REQUIRE(t >= 0 && t < capacity, B200CTR_E_INVALID_ARG, "rowopt_push: step %lld outside history capacity %lld", (long long)t, (long long)capacity);
  rowopt_push_kernel<<<1, 1, 0, as_stream(stream)>>>((float4*)hist4, C, S, logD, (int)t, make_float4(decay_mul, step_size, bc2_sqrt, ema_d), C_t, S_t, logD_t);
  B200_LAUNCH_CHECK("rowopt_push_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_rows_claim_any(const void* tok, int idx_is_64, int64_t n_slots, int pad_id, int64_t vocab, int32_t* owner,
                                      void* stream) {
  B200_REQUIRE(tok && owner, B200CTR_E_INVALID_ARG, "rows_claim_any: null pointer");
  B200_REQUIRE(n_slots < (1LL << 31), B200CTR_E_UNSUPPORTED_SHAPE, "rows_claim_any: too many slots");
  if (n_slots == 0) return B200CTR_OK;
  unsigned grid = (unsigned)ceil_div(n_slots, 256);
  if (idx_is_64) rows_claim_any_kernel<int64_t><<<grid, 256, 0, as_stream(stream)>>>((const int64_t*)tok, n_slots, pad_id, vocab, owner);
  else rows_claim_any_kernel<int32_t><<<grid, 256, 0, as_stream(stream)>>>((const int32_t*)tok, n_slots, pad_id, vocab, owner);
  B200_LAUNCH_CHECK("rows_claim_any_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_rowopt_catchup(const void* tok, int idx_is_64, int64_t n_slots, int pad_id, int64_t vocab,
                                      const int32_t* owner, int32_t* state, float* p, float* m, float* v, float* ema, int D,
                                      int64_t target_step, const void* hist4, const float* C, const float* S, const float* logD,
                                      float beta1, float beta2, float eps, void* stream) {
  B200_REQUIRE(tok && owner && state && p && m && v && hist4 && C && S && logD, B200CTR_E_INVALID_ARG, "rowopt_catchup: null pointer");
  B200_REQUIRE(D == RO_D, B200CTR_E_UNSUPPORTED_SHAPE, "rowopt_catchup: D=%d unsupported (64 only)", D);
  if (n_slots == 0 || target_step <= 0) return B200CTR_OK;
  RowHist H = make_hist(hist4, C, S, logD, beta1, beta2, eps);
  int grid = (int)std::min<int64_t>(ceil_div(n_slots, 8), (int64_t)num_sms() * 16);
  if (idx_is_64) rowopt_catchup_kernel<int64_t><<<grid, 256, 0, as_stream(stream)>>>((const int64_t*)tok, n_slots, pad_id, vocab, owner, state, p, m, v, ema, (int)target_step, H);
  else rowopt_catchup_kernel<int32_t><<<grid, 256, 0, as_stream(stream)>>>((const int32_t*)tok, n_slots, pad_id, vocab, owner, state, p, m, v, ema, (int)target_step, H);
  B200_LAUNCH_CHECK("rowopt_catchup_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_rowopt_update(const int32_t* tok, int64_t n_slots, int pad_id, const int32_t* owner, int32_t* state,
                                     float* p, float* g, float* m, float* v, float* ema, int D, int64_t step, const float* coef,
                                     const void* hist4, const float* C, const float* S, const float* logD, float beta1,
                                     float beta2, float eps, void* stream) {
  B200_REQUIRE(tok && owner && state && p && g && m && v && hist4 && C && S && logD, B200CTR_E_INVALID_ARG, "rowopt_update: null pointer");
  B200_REQUIRE(D == RO_D, B200CTR_E_UNSUPPORTED_SHAPE, "rowopt_update: D=%d unsupported (64 only)", D);
  B200_REQUIRE(step >= 1, B200CTR_E_INVALID_ARG, "rowopt_update: step must be >= 1");
  if (n_slots == 0) return B200CTR_OK;
  RowHist H = make_hist(hist4, C, S, logD, beta1, beta2, eps);
  int grid = (int)std::min<int64_t>(ceil_div(n_slots, 8), (int64_t)num_sms() * 16);
  rowopt_update_kernel<<<grid, 256, 0, as_stream(stream)>>>(tok, n_slots, pad_id, owner, state, p, g, m, v, ema, (int)step, coef, H);
  B200_LAUNCH_CHECK("rowopt_update_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_rowopt_flush(int32_t* state, float* p, float* m, float* v, float* ema, int64_t n_rows, int D,
                                    int64_t target_step, const void* hist4, const float* C, const float* S, const float* logD,
                                    float beta1, float beta2, float eps, void* stream) {
  B200_REQUIRE(state && p && m && v && hist4 && C && S && logD, B200CTR_E_INVALID_ARG, "rowopt_flush: null pointer");
  B200_REQUIRE(D == RO_D, B200CTR_E_UNSUPPORTED_SHAPE, "rowopt_flush: D=%d unsupported (64 only)", D);
  if (n_rows == 0 || target_step <= 0) return B200CTR_OK;
  RowHist H = make_hist(hist4, C, S, logD, beta1, beta2, eps);
  int grid = (int)std::min<int64_t>(ceil_div(n_rows, 8), (int64_t)num_sms() * 32);
  rowopt_flush_kernel<<<grid, 256, 0, as_stream(stream)>>>(state, p, m, v, ema, n_rows, (int)target_step, H);
  B200_LAUNCH_CHECK("rowopt_flush_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_rows_localize(const void* tok, int idx_is_64, int64_t n_slots, int pad_id, int n_shards, int rank,
                                     int32_t* out, void* stream) {
  B200_REQUIRE(tok && out, B200CTR_E_INVALID_ARG, "rows_localize: null pointer");
  B200_REQUIRE(n_shards >= 1 && n_shards <= 8 && !(n_shards & (n_shards - 1)) && rank >= 0 && rank < n_shards,
               B200CTR_E_INVALID_ARG, "rows_localize: n_shards=%d rank=%d", n_shards, rank);
  if (n_slots == 0) return B200CTR_OK;
  int shift = 0;
  while ((1 << shift) < n_shards) ++shift;
  const unsigned grid = (unsigned)ceil_div(n_slots, 256);
  if (idx_is_64) rows_localize_kernel<int64_t><<<grid, 256, 0, as_stream(stream)>>>((const int64_t*)tok, n_slots, pad_id, n_shards - 1, shift, rank, out);
  else rows_localize_kernel<int32_t><<<grid, 256, 0, as_stream(stream)>>>((const int32_t*)tok, n_slots, pad_id, n_shards - 1, shift, rank, out);
  B200_LAUNCH_CHECK("rows_localize_kernel");
  return B200CTR_OK;
}

static int bm_check(const void* bm, int n_maps, int64_t words, int n_shards, int rank, const char* what) {
  if (!bm || n_maps < 1 || words < 1 || n_shards < 1 || n_shards > 8 || (n_shards & (n_shards - 1)) || rank < 0 || rank >= n_shards) {
    set_error("%s: bad bitmap arguments (n_maps=%d words=%lld n_shards=%d rank=%d)", what, n_maps, (long long)words, n_shards, rank);
    return B200CTR_E_INVALID_ARG;
  }
  return B200CTR_OK;
}
static int bm_grid(int64_t words) { return (int)std::min<int64_t>(ceil_div(words, 32 * 8), (int64_t)num_sms() * 8); }

extern "C" int b200ctr_bitmap_mark(const void* tok, int idx_is_64, int64_t n_slots, int pad_id, int64_t vocab, uint32_t* bitmap,
                                   void* stream) {
  B200_REQUIRE(tok && bitmap, B200CTR_E_INVALID_ARG, "bitmap_mark: null pointer");
  if (n_slots == 0) return B200CTR_OK;
  const unsigned grid = (unsigned)ceil_div(n_slots, 256);
  if (idx_is_64) bitmap_mark_kernel<int64_t><<<grid, 256, 0, as_stream(stream)>>>((const int64_t*)tok, n_slots, pad_id, vocab, bitmap);
  else bitmap_mark_kernel<int32_t><<<grid, 256, 0, as_stream(stream)>>>((const int32_t*)tok, n_slots, pad_id, vocab, bitmap);
  B200_LAUNCH_CHECK("bitmap_mark_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_rowopt_catchup_bm(const uint32_t* bitmaps, int n_maps, int64_t words, int n_shards, int rank, int32_t* state,
                                         float* p, float* m, float* v, float* ema, int D, int64_t target_step, const void* hist4,
                                         const float* C, const float* S, const float* logD, float beta1, float beta2, float eps,
                                         void* stream) {
  int rc = bm_check(bitmaps, n_maps, words, n_shards, rank, "rowopt_catchup_bm");
  if (rc) return rc;
  B200_REQUIRE(state && p && m && v && hist4 && C && S && logD, B200CTR_E_INVALID_ARG, "rowopt_catchup_bm: null pointer");
  B200_REQUIRE(D == RO_D, B200CTR_E_UNSUPPORTED_SHAPE, "rowopt_catchup_bm: D=%d unsupported", D);
  rowopt_catchup_bm_kernel<<<bm_grid(words), 256, 0, as_stream(stream)>>>(bitmaps, n_maps, words, n_shards, rank, state, p, m, v, ema,
                                                                           (int)target_step, make_hist(hist4, C, S, logD, beta1, beta2, eps));
  B200_LAUNCH_CHECK("rowopt_catchup_bm_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_rows_sqnorm_bm(const uint32_t* bitmaps, int n_maps, int64_t words, int n_shards, int rank, const float* g,
                                      int D, double* acc, void* stream) {
  int rc = bm_check(bitmaps, n_maps, words, n_shards, rank, "rows_sqnorm_bm");
  if (rc) return rc;
  B200_REQUIRE(g && acc, B200CTR_E_INVALID_ARG, "rows_sqnorm_bm: null pointer");
  B200_REQUIRE(D == RO_D, B200CTR_E_UNSUPPORTED_SHAPE, "rows_sqnorm_bm: D=%d unsupported", D);
  rows_sqnorm_bm_kernel<<<bm_grid(words), 256, 0, as_stream(stream)>>>(bitmaps, n_maps, words, n_shards, rank, g, acc);
  B200_LAUNCH_CHECK("rows_sqnorm_bm_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_rowopt_update_bm(const uint32_t* bitmaps, int n_maps, int64_t words, int n_shards, int rank, int32_t* state,
                                        float* p, float* g, float* m, float* v, float* ema, int D, int64_t step, const float* coef,
                                        const void* hist4, const float* C, const float* S, const float* logD, float beta1,
                                        float beta2, float eps, void* stream) {
  int rc = bm_check(bitmaps, n_maps, words, n_shards, rank, "rowopt_update_bm");
  if (rc) return rc;
  B200_REQUIRE(state && p && g && m && v && hist4 && C && S && logD, B200CTR_E_INVALID_ARG, "rowopt_update_bm: null pointer");
  B200_REQUIRE(D == RO_D, B200CTR_E_UNSUPPORTED_SHAPE, "rowopt_update_bm: D=%d unsupported", D);
  rowopt_update_bm_kernel<<<bm_grid(words), 256, 0, as_stream(stream)>>>(bitmaps, n_maps, words, n_shards, rank, state, p, g, m, v, ema,
                                                                          (int)step, coef, make_hist(hist4, C, S, logD, beta1, beta2, eps));
  B200_LAUNCH_CHECK("rowopt_update_bm_kernel");
  return B200CTR_OK;
}
