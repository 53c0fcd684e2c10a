// K1: numeric / binary feature embedding (rank-1 folded) and hashed-categorical gather + projection,
// forward and backward.  Reference: src/models/feature_embed.py:19-27,42-48 and src/models/wrapper.py:106-136.
#include "common.cuh"
#include <algorithm>

namespace b200ctr {

// ------------------------------------------------------------------------------------------------------
// numeric / binary: out[b,f,:] = x[b,f] * P[f,:] + Q[f,:],  P = w @ w_out^T, Q = b @ w_out^T
// ------------------------------------------------------------------------------------------------------
__global__ void numbin_fold_kernel(const float* __restrict__ w, const float* __restrict__ b,
                                   const float* __restrict__ w_out, float* __restrict__ P, float* __restrict__ Q,
                                   int F, int fe, int D) {
  int f = blockIdx.x, d = threadIdx.x;
  if (d >= D) return;
  float p = 0.f, q = 0.f;
  for (int e = 0; e < fe; ++e) {
    float wo = w_out[d * fe + e];
    p = fmaf(w[f * fe + e], wo, p);
    if (b) q = fmaf(b[f * fe + e], wo, q);
  }
  P[f * D + d] = p;
  Q[f * D + d] = q;
}

// one block per sample (grid-stride); D/4 float4 lanes per feature row, blockDim/(D/4) rows in flight
__global__ void __launch_bounds__(256)
numbin_fwd_kernel(const float* __restrict__ x, const float* __restrict__ P, const float* __restrict__ Q,
                  float* __restrict__ out, int64_t ld_b, float* __restrict__ mean_out, int64_t ld_mean,
                  int64_t B, int F, int D) {
  __shared__ float4 red[256];
  const int D4 = D >> 2;
  const int lane = threadIdx.x % D4, rg = threadIdx.x / D4, nrg = blockDim.x / D4;
  const float4* P4 = reinterpret_cast<const float4*>(P);
  const float4* Q4 = reinterpret_cast<const float4*>(Q);
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rg < nrg) {
      for (int f = rg; f < F; f += nrg) {
        float xv = x[b * F + f];
        float4 p = __ldg(P4 + f * D4 + lane), q = __ldg(Q4 + f * D4 + lane);
        float4 o = make_float4(fmaf(xv, p.x, q.x), fmaf(xv, p.y, q.y), fmaf(xv, p.z, q.z), fmaf(xv, p.w, q.w));
        if (out) *reinterpret_cast<float4*>(out + b * ld_b + (int64_t)f * D + lane * 4) = o;
        acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
      }
    }
    if (mean_out) {
      red[threadIdx.x] = acc;
      __syncthreads();
      if (rg == 0) {
        float4 s = red[lane];
        for (int r = 1; r < nrg; ++r) {
          float4 t = red[r * D4 + lane];
          s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
        }
        float inv = 1.f / (float)F;
        *reinterpret_cast<float4*>(mean_out + b * ld_mean + lane * 4) = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
      }
      __syncthreads();
    }
  }
}

// dP[f,:] += sum_b x[b,f]*g, dQ[f,:] += sum_b g ; each block owns a contiguous chunk of samples,
// each thread a fixed set of (f, d4) pairs; one atomic per pair per block.
__global__ void __launch_bounds__(256)
numbin_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dout, int64_t ld_b,
                  const float* __restrict__ dmean, int64_t ld_dmean, float* __restrict__ dP, float* __restrict__ dQ,
                  int64_t B, int F, int D, int64_t chunk) {
  const int D4 = D >> 2;
  const int npairs = F * D4;
  const int64_t b0 = (int64_t)blockIdx.x * chunk, b1 = min(B, b0 + chunk);
  const float invF = 1.f / (float)F;
  // grid.y slices the (f, d4) pairs, grid.x the samples: one pair per thread and 64 samples per CTA keep the serial
  // chain short (it was 5 pairs x 32 samples per thread at 19 % occupancy) without adding atomics
  for (int pr = blockIdx.y * blockDim.x + threadIdx.x; pr < npairs; pr += gridDim.y * blockDim.x) {
    int f = pr / D4, l = pr % D4;
    float4 ap = make_float4(0.f, 0.f, 0.f, 0.f), aq = ap;
    auto grad = [&](int64_t b) {
      float4 g = dout ? *reinterpret_cast<const float4*>(dout + b * ld_b + (int64_t)f * D + l * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (dmean) {
        float4 m = *reinterpret_cast<const float4*>(dmean + b * ld_dmean + l * 4);
        g.x = fmaf(m.x, invF, g.x); g.y = fmaf(m.y, invF, g.y); g.z = fmaf(m.z, invF, g.z); g.w = fmaf(m.w, invF, g.w);
      }
      return g;
    };
    auto accum = [&](const float4& g, float xv) {
      ap.x = fmaf(xv, g.x, ap.x); ap.y = fmaf(xv, g.y, ap.y); ap.z = fmaf(xv, g.z, ap.z); ap.w = fmaf(xv, g.w, ap.w);
      aq.x += g.x; aq.y += g.y; aq.z += g.z; aq.w += g.w;
    };
    int64_t b = b0;
    for (; b + 4 <= b1; b += 4) {                 // 4 samples' loads in flight
      const float4 g0 = grad(b), g1 = grad(b + 1), g2 = grad(b + 2), g3 = grad(b + 3);
      const float x0 = x[b * F + f], x1 = x[(b + 1) * F + f], x2 = x[(b + 2) * F + f], x3 = x[(b + 3) * F + f];
      accum(g0, x0); accum(g1, x1); accum(g2, x2); accum(g3, x3);
    }
    for (; b < b1; ++b) accum(grad(b), x[b * F + f]);
    // one partial row per sample chunk (= blockIdx.x), summed in chunk order afterwards (b200ctr_colsum): no float atomics
    const int64_t prow = (int64_t)blockIdx.x * F * D;
    *reinterpret_cast<float4*>(dP + prow + f * D + l * 4) = ap;
    *reinterpret_cast<float4*>(dQ + prow + f * D + l * 4) = aq;
  }
}

// blocks [0,F): dw[f,:], db[f,:] ; blocks [F, F+D): dw_out[d,:]
__global__ void numbin_unfold_kernel(const float* __restrict__ dP, const float* __restrict__ dQ, const float* __restrict__ w,
                                     const float* __restrict__ b, const float* __restrict__ w_out, float* __restrict__ dw,
                                     float* __restrict__ db, float* __restrict__ dw_out, int F, int fe, int D) {
  int e = threadIdx.x;
  if (e >= fe) return;
  if ((int)blockIdx.x < F) {
    int f = blockIdx.x;
    float a = 0.f, c = 0.f;
    for (int d = 0; d < D; ++d) {
      float wo = w_out[d * fe + e];
      a = fmaf(dP[f * D + d], wo, a);
      c = fmaf(dQ[f * D + d], wo, c);
    }
    dw[f * fe + e] = a;
    if (db) db[f * fe + e] = c;
  } else {
    int d = blockIdx.x - F;
    float a = 0.f;
    for (int f = 0; f < F; ++f) {
      a = fmaf(dP[f * D + d], w[f * fe + e], a);
      if (b) a = fmaf(dQ[f * D + d], b[f * fe + e], a);
    }
    dw_out[d * fe + e] = a;
  }
}

// ------------------------------------------------------------------------------------------------------
// categorical gather + projection.  D == 64 (two outputs per lane).
// ------------------------------------------------------------------------------------------------------
constexpr int CAT_D = 64;
constexpr int CAT_MAXDC = 64;
constexpr int CAT_SPW = 4;            // samples per warp per block pass
constexpr int CAT_WARPS = 8;

// One CTA = CAT_WARPS x CAT_SPW samples, walking the Fc columns.  Column c + 1's operands (its W_c elements and the gathered
// table rows) are loaded into registers BEFORE column c's products are issued and stored to shared memory after them: the
// global-load latency of a column sits under the previous column's arithmetic (ncu r02ai: 65 % of the kernel's stall samples
// were the W loads of the staging loop, one exposed round trip per element; 21 % of the warp slots active).
template <typename IT>
__global__ void __launch_bounds__(CAT_WARPS * 32)
cat_fwd_kernel(const IT* __restrict__ x_cat, const float* const* __restrict__ tables, const float* const* __restrict__ projs,
               const int32_t* __restrict__ dims, const int32_t* __restrict__ cards, const int32_t* __restrict__ lds,
               float* __restrict__ out, int64_t ld_b,
               float* __restrict__ mean_out, int64_t ld_mean, int q_idx, float* __restrict__ q_out,
               int64_t B, int Fc, b200ctr_dropout_t drop) {
  __shared__ float Wt[CAT_MAXDC][CAT_D + 1];               // W_c transposed: Wt[j][n]
  __shared__ __align__(16) float rows[CAT_WARPS][CAT_SPW][CAT_MAXDC];
  constexpr int NT = CAT_WARPS * 32;
  constexpr int WPT = CAT_D * CAT_MAXDC / NT;              // W elements per thread (widest column)
  constexpr int RPT = CAT_MAXDC / 32;                      // row elements per lane and sample
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t b_base = ((int64_t)blockIdx.x * CAT_WARPS + warp) * CAT_SPW;
  float m0[CAT_SPW], m1[CAT_SPW];
#pragma unroll
  for (int s = 0; s < CAT_SPW; ++s) m0[s] = m1[s] = 0.f;
  float wreg[WPT], rreg[CAT_SPW][RPT];

  auto fetch = [&](int c) {                                // global loads of column c, all independent
    const int dc = dims[c];
    const float* W = projs[c];
    const float* T = tables[c];
    const int card = cards[c];
    const int64_t ldt = lds ? lds[c] : dc;                  // row stride of table c (row-lazy tables are stored 64 wide)
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
      const int i = (int)threadIdx.x + k * NT;
      wreg[k] = i < CAT_D * dc ? W[i] : 0.f;
    }
#pragma unroll
    for (int s = 0; s < CAT_SPW; ++s) {
      const int64_t b = b_base + s;
      int64_t id = 0;
      if (b < B) {
        id = (int64_t)x_cat[b * Fc + c];
        id = id < 0 ? 0 : (id >= card ? card - 1 : id);     // memory-safe clamp (reference would raise)
      }
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        const int j = lane + 32 * k;
        rreg[s][k] = (b < B && j < dc) ? T[id * ldt + j] : 0.f;
      }
    }
  };
  auto stash = [&](int dc) {                               // registers -> shared memory
    // W_c[n][j] -> Wt[j][n]; (n, j) of element i advance incrementally (a division per element cost as much as the products)
    const int step_n = NT / dc, step_j = NT % dc;
    int n = (int)threadIdx.x / dc, j = (int)threadIdx.x % dc;
#pragma unroll
    for (int k = 0; k < WPT; ++k) {
      if ((int)threadIdx.x + k * NT < CAT_D * dc) Wt[j][n] = wreg[k];
      n += step_n; j += step_j;
      if (j >= dc) { j -= dc; ++n; }
    }
#pragma unroll
    for (int s = 0; s < CAT_SPW; ++s)
#pragma unroll
      for (int k = 0; k < RPT; ++k)
        if (lane + 32 * k < dc) rows[warp][s][lane + 32 * k] = rreg[s][k];
  };

  if (Fc > 0) { fetch(0); stash(dims[0]); }
  __syncthreads();
  for (int c = 0; c < Fc; ++c) {
    const int dc = dims[c];
    if (c + 1 < Fc) fetch(c + 1);
    // the warp's CAT_SPW samples share every W element read: per 4 table columns 8 W loads + CAT_SPW 16-byte row loads feed
    // 8 * CAT_SPW FMAs (sample by sample it was 3 loads per 2 FMAs); each output still accumulates j in ascending order
    float a0[CAT_SPW], a1[CAT_SPW];
#pragma unroll
    for (int s = 0; s < CAT_SPW; ++s) a0[s] = a1[s] = 0.f;
    int j = 0;
    for (; j + 4 <= dc; j += 4) {
      float w0[4], w1[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) { w0[k] = Wt[j + k][lane]; w1[k] = Wt[j + k][lane + 32]; }
#pragma unroll
      for (int s = 0; s < CAT_SPW; ++s) {
        const float4 r = *reinterpret_cast<const float4*>(&rows[warp][s][j]);
        a0[s] = fmaf(r.x, w0[0], a0[s]); a1[s] = fmaf(r.x, w1[0], a1[s]);
        a0[s] = fmaf(r.y, w0[1], a0[s]); a1[s] = fmaf(r.y, w1[1], a1[s]);
        a0[s] = fmaf(r.z, w0[2], a0[s]); a1[s] = fmaf(r.z, w1[2], a1[s]);
        a0[s] = fmaf(r.w, w0[3], a0[s]); a1[s] = fmaf(r.w, w1[3], a1[s]);
      }
    }
    for (; j < dc; ++j) {
      const float w0 = Wt[j][lane], w1 = Wt[j][lane + 32];
#pragma unroll
      for (int s = 0; s < CAT_SPW; ++s) {
        const float r = rows[warp][s][j];
        a0[s] = fmaf(r, w0, a0[s]); a1[s] = fmaf(r, w1, a1[s]);
      }
    }
#pragma unroll
    for (int s = 0; s < CAT_SPW; ++s) {
      int64_t b = b_base + s;
      if (b >= B) continue;
      m0[s] += a0[s]; m1[s] += a1[s];
      if (q_out && c == q_idx) { q_out[b * CAT_D + lane] = a0[s]; q_out[b * CAT_D + lane + 32] = a1[s]; }
      uint64_t eid = ((uint64_t)b * Fc + c) * CAT_D;
      float* o = out + b * ld_b + (int64_t)c * CAT_D;
      o[lane] = a0[s] * drop_scale(drop, eid + lane);
      o[lane + 32] = a1[s] * drop_scale(drop, eid + lane + 32);
    }
    __syncthreads();                                       // column c's operands have been read by every warp
    if (c + 1 < Fc) stash(dims[c + 1]);
    __syncthreads();
  }
  if (mean_out) {
    float inv = 1.f / (float)Fc;
#pragma unroll
    for (int s = 0; s < CAT_SPW; ++s) {
      int64_t b = b_base + s;
      if (b < B) { mean_out[b * ld_mean + lane] = m0[s] * inv; mean_out[b * ld_mean + lane + 32] = m1[s] * inv; }
    }
  }
}

// grid (NBX, Fc).  Block (bx, c) walks sample tiles of 32 with stride NBX: stages g and the gathered rows,
// accumulates dW_c in registers (one atomic per element per block at the end) and scatters dE_c rows.
constexpr int CATB_TS = 32;
// The two small products of one (32-sample tile, column) pair on register tiles: the thread owns n0 .. n0+3 x columns jg, jg+16, ..
// of dW_c and samples s0, s0+1 x the same columns of dE_c (thread per output element cost two shared-memory loads per FMA).  MM =
// ceil(d_c / 16) is uniform over the CTA and a template parameter: with the column count as a per-thread predicate every warp
// issued all four column steps -- 3.8x the useful FMAs over the 35 shipped columns, 22 of which are 8 wide (ncu r02ai: 170 M warp
// instructions, issue 74 %).  Columns >= d_c of the last step read don't-care shared memory and are never stored.  Every output
// accumulates its terms in ascending order (bit-identical to the element-per-thread form).
template <int MM>
__device__ __forceinline__ void cat_bwd_products(const float (*gs)[CAT_D + 4], const float (*rs)[CAT_MAXDC + 1], const float (*Ws)[CAT_MAXDC + 1],
                                                 int jg, int n0, int s0, float (&acc)[4][4], float (&de)[2][4]) {
  // dW[n][j] += sum_s g[s][n] * row[s][j]
#pragma unroll 4
  for (int s = 0; s < CATB_TS; ++s) {
    const float4 g4 = *reinterpret_cast<const float4*>(&gs[s][n0]);
    const float gv[4] = {g4.x, g4.y, g4.z, g4.w};
#pragma unroll
    for (int m = 0; m < MM; ++m) {
      const float r = rs[s][jg + 16 * m];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i][m] = fmaf(gv[i], r, acc[i][m]);
    }
  }
  // dE[id][j] += sum_n W[n][j] * g[s][n]
#pragma unroll
  for (int q = 0; q < 2; ++q)
#pragma unroll
    for (int m = 0; m < 4; ++m) de[q][m] = 0.f;
#pragma unroll 2
  for (int n = 0; n < CAT_D; n += 4) {
    const float4 ga = *reinterpret_cast<const float4*>(&gs[s0][n]);
    const float4 gb = *reinterpret_cast<const float4*>(&gs[s0 + 1][n]);
    const float g0[4] = {ga.x, ga.y, ga.z, ga.w}, g1[4] = {gb.x, gb.y, gb.z, gb.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int m = 0; m < MM; ++m) {
        const float w = Ws[n + k][jg + 16 * m];
        de[0][m] = fmaf(w, g0[k], de[0][m]);
        de[1][m] = fmaf(w, g1[k], de[1][m]);
      }
    }
  }
}

// grid (NBX, Fc).  Block (bx, c) walks sample tiles of 32 with stride NBX.  Per tile: every global load (the gradient rows, the
// context-mean gradient, the ids and the gathered table rows) is issued before the first use, ONE barrier, then the products.
template <typename IT>
__global__ void __launch_bounds__(256, 3)
cat_bwd_kernel(const IT* __restrict__ x_cat, const float* const* __restrict__ tables, const float* const* __restrict__ projs,
               const int32_t* __restrict__ dims, const int32_t* __restrict__ cards, const int32_t* __restrict__ lds,
               const float* __restrict__ dout, int64_t ld_b,
               const float* __restrict__ dmean, int64_t ld_dmean, int q_idx, const float* __restrict__ dq,
               float* __restrict__ rowg, int32_t* __restrict__ keys, float* __restrict__ dwpart, int sumdc, int64_t B, int Fc,
               b200ctr_dropout_t drop, const int32_t* __restrict__ defer_slot, int n_defer, float* __restrict__ defer_rows,
               int32_t* __restrict__ defer_keys, int32_t sentinel) {
  __shared__ float Ws[CAT_D][CAT_MAXDC + 1];           // natural layout W[n][j]
  __shared__ __align__(16) float gs[CATB_TS][CAT_D + 4];   // rows 16-byte aligned: four n per load
  __shared__ float rs[CATB_TS][CAT_MAXDC + 1];
  constexpr int GPT = CATB_TS * CAT_D / 256;            // gradient elements per thread and tile
  constexpr int RPT = CATB_TS * CAT_MAXDC / 256;        // gathered row elements per thread and tile (widest column)
  const int c = blockIdx.y;
  const int dc = dims[c];
  const int card = cards[c];
  const float* W = projs[c];
  const float* T = tables[c];
  int off_c = 0, row_base = 0;                            // column offset of dW_c in a partial; first global row id of table c
  for (int cc = 0; cc < c; ++cc) { off_c += dims[cc]; row_base += cards[cc]; }
  const int64_t ldt = lds ? lds[c] : dc;
  const int dslot = defer_slot ? defer_slot[c] : -1;
  const int tid = threadIdx.x;
  const bool use_dq = dq != nullptr && c == q_idx;
  // (row, column) of element i of a (., dc) array advance incrementally over the strided loops (no division per element)
  const int step_r = 256 / dc, step_j = 256 % dc, r_first = tid / dc, j_first = tid % dc;
  {
    int n = r_first, j = j_first;
    for (int i = tid; i < CAT_D * dc; i += 256) {
      Ws[n][j] = W[i];
      n += step_r; j += step_j;
      if (j >= dc) { j -= dc; ++n; }
    }
  }
  const int jg = tid & 15, n0 = (tid >> 4) * 4, s0 = (tid >> 4) * 2;
  const int mm = (dc + 15) >> 4;                          // column steps of this table (uniform over the CTA)
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[i][m] = 0.f;
  const float invFc = 1.f / (float)Fc;
  const int64_t ntiles = ceil_div(B, (int64_t)CATB_TS);
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t b0 = t * CATB_TS;
    // ---- all global loads of the tile ----
    float gv[GPT], dmv[GPT], dqv[GPT], rv[RPT];
#pragma unroll
    for (int k = 0; k < GPT; ++k) {
      const int i = tid + k * 256, sgl = i / CAT_D, n = i % CAT_D;
      const int64_t b = b0 + sgl;
      gv[k] = dmv[k] = dqv[k] = 0.f;
      if (b < B) {
        if (dout) gv[k] = dout[b * ld_b + (int64_t)c * CAT_D + n];
        if (dmean) dmv[k] = dmean[b * ld_dmean + n];
        if (use_dq) dqv[k] = dq[b * CAT_D + n];
      }
    }
    {
      int sr = r_first, j = j_first;
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        rv[k] = 0.f;
        if (tid + k * 256 < CATB_TS * dc && b0 + sr < B) {
          int64_t id = (int64_t)x_cat[(b0 + sr) * Fc + c];
          id = id < 0 ? 0 : (id >= card ? card - 1 : id);
          rv[k] = T[id * ldt + j];
        }
        sr += step_r; j += step_j;
        if (j >= dc) { j -= dc; ++sr; }
      }
    }
    __syncthreads();                                      // the previous tile's products have read gs / rs
#pragma unroll
    for (int k = 0; k < GPT; ++k) {
      const int i = tid + k * 256, sgl = i / CAT_D, n = i % CAT_D;
      const int64_t b = b0 + sgl;
      float g = 0.f;
      if (b < B) {
        if (dout) g = gv[k] * drop_scale(drop, ((uint64_t)b * Fc + c) * CAT_D + n);
        if (dmean) g = fmaf(dmv[k], invFc, g);
        if (use_dq) g += dqv[k];
      }
      gs[sgl][n] = g;
    }
    {
      int sr = r_first, j = j_first;
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        if (tid + k * 256 < CATB_TS * dc) rs[sr][j] = rv[k];
        sr += step_r; j += step_j;
        if (j >= dc) { j -= dc; ++sr; }
      }
    }
    __syncthreads();
    float de[2][4];
    switch (mm) {
      case 1: cat_bwd_products<1>(gs, rs, Ws, jg, n0, s0, acc, de); break;
      case 2: cat_bwd_products<2>(gs, rs, Ws, jg, n0, s0, acc, de); break;
      case 3: cat_bwd_products<3>(gs, rs, Ws, jg, n0, s0, acc, de); break;
      default: cat_bwd_products<4>(gs, rs, Ws, jg, n0, s0, acc, de); break;
    }
    // a sample's contribution to dE_c[id], as a 64-wide row (columns >= d_c stay zero) of slot (b, c): the slots are
    // sorted by global table row and reduced in slot order by b200ctr_scatter_rows (no float atomics)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int64_t b = b0 + s0 + q;
      if (b >= B) continue;
      float* dst = dslot >= 0 ? defer_rows + (b * n_defer + dslot) * (int64_t)CAT_D : rowg + (b * Fc + c) * (int64_t)CAT_D;
#pragma unroll
      for (int m = 0; m < 4; ++m)
        if (jg + 16 * m < dc) dst[jg + 16 * m] = de[q][m];
    }
    if (tid < CATB_TS && b0 + tid < B) {
      // deferred column (data parallel, replicated row-lazy table): the slot leaves this call unreduced -- the trainer gathers
      // every rank's slots and reduces them once, identically on every replica; the local sort skips it (sentinel key)
      int64_t id = (int64_t)x_cat[(b0 + tid) * Fc + c];
      id = id < 0 ? 0 : (id >= card ? card - 1 : id);
      const int32_t key = row_base + (int32_t)id;
      keys[(b0 + tid) * Fc + c] = dslot >= 0 ? sentinel : key;
      if (dslot >= 0) defer_keys[(b0 + tid) * n_defer + dslot] = key;
    }
  }
  // this CTA's partial dW_c; the partials are added in CTA order by cat_dw_reduce_kernel
  float* dwp = dwpart + (int64_t)blockIdx.x * CAT_D * sumdc + (int64_t)CAT_D * off_c;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int m = 0; m < 4; ++m)
      if (jg + 16 * m < dc) dwp[(n0 + i) * dc + jg + 16 * m] = acc[i][m];
}

// dE_c[r, 0:d_c] += temp[row_base_c + r, 0:d_c]  (temp: the 64-wide reduced rows of all tables, one global row id per table row).
// 1-D over the 16-byte quads of temp, so the cost follows the total row count (a (max_card, Fc) grid launched 35 M mostly idle
// blocks with one 4 M-row table); the table of a global row is found by a search over the <= 64 row bases in shared memory.
constexpr int CAT_MAXF = 64;
__global__ void __launch_bounds__(256)
cat_unpad_kernel(float* __restrict__ temp, const int32_t* __restrict__ dims, const int32_t* __restrict__ cards,
                 const int32_t* __restrict__ keep_rows, int rezero, float* const* __restrict__ dtables, int Fc, int64_t sum_cards) {
  __shared__ int64_t base[CAT_MAXF + 1];
  if (threadIdx.x == 0) {
    int64_t b = 0;
    for (int c = 0; c < Fc; ++c) { base[c] = b; b += cards[c]; }
    base[Fc] = b;
  }
  __syncthreads();
  // grid-stride over the tables' quads; a row-lazy table's whole range is stepped over in one go
  const int64_t nq = sum_cards * (CAT_D / 4), stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {   // quad: global row q / 16, columns 4 (q % 16) ..
    const int64_t g = q >> 4;
    const int j0 = (int)(q & 15) * 4;
    int lo = 0, hi = Fc - 1;                                                     // largest c with base[c] <= g
    while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (base[mid] <= g) lo = mid; else hi = mid - 1; }
    const int c = lo;
    if (keep_rows && keep_rows[c]) {                     // row-lazy table: its reduced rows in `temp` ARE its gradient rows
      const int64_t q_end = base[c + 1] * (CAT_D / 4);   // first quad after this table: jump there, keeping this thread's phase
      q += ((q_end - q + stride - 1) / stride - 1) * stride;
      continue;
    }
    const int dc = dims[c];
    if (j0 >= dc) continue;
    float4* t4 = reinterpret_cast<float4*>(temp) + q;
    const float4 v = *t4;
    if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
    float* dst = dtables[c] + (g - base[c]) * dc + j0;                           // single writer per element
    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (j0 + k < dc && vv[k] != 0.f) dst[k] += vv[k];
    if (rezero) *t4 = make_float4(0.f, 0.f, 0.f, 0.f);   // caller-owned buffer: back to all-zero for the next step
  }
}

__global__ void __launch_bounds__(256)
cat_dw_reduce_kernel(const float* __restrict__ dwpart, int nparts, const int32_t* __restrict__ dims, int sumdc,
                     float* const* __restrict__ dprojs) {
  const int c = blockIdx.y, dc = dims[c];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= CAT_D * dc) return;
  int off_c = 0;
  for (int cc = 0; cc < c; ++cc) off_c += dims[cc];
  float a = 0.f;
  for (int p = 0; p < nparts; ++p) a += dwpart[(int64_t)p * CAT_D * sumdc + (int64_t)CAT_D * off_c + e];
  dprojs[c][e] += a;
}

__global__ void axpby_kernel(const float* __restrict__ x, const float* __restrict__ y, float a, float b,
                             float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = a * (x ? x[i] : 0.f) + b * (y ? y[i] : 0.f);
}

__global__ void relu_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float scale,
                                float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = y[i] > 0.f ? scale * dy[i] : 0.f;
}
__global__ void sigmoid_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = sigmoidf_(x[i]);
}
__global__ void add2d_kernel(float* __restrict__ dst, int64_t ld_dst, const float* __restrict__ src, int64_t ld_src,
                             int64_t rows, int64_t cols) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  int64_t r = i / cols, c = i % cols;
  dst[r * ld_dst + c] += src[r * ld_src + c];
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int b200ctr_relu_bwd(const float* dy, const float* y, float scale, float* out, int64_t n, void* stream) {
  B200_REQUIRE(dy && y && out, B200CTR_E_INVALID_ARG, "relu_bwd: null pointer");
  if (n == 0) return B200CTR_OK;
  relu_bwd_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(dy, y, scale, out, n);
  B200_LAUNCH_CHECK("relu_bwd_kernel");
  return B200CTR_OK;
}
extern "C" int b200ctr_sigmoid(const float* x, float* out, int64_t n, void* stream) {
  B200_REQUIRE(x && out, B200CTR_E_INVALID_ARG, "sigmoid: null pointer");
  if (n == 0) return B200CTR_OK;
  sigmoid_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(x, out, n);
  B200_LAUNCH_CHECK("sigmoid_kernel");
  return B200CTR_OK;
}
extern "C" int b200ctr_add2d(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int64_t rows, int64_t cols, void* stream) {
  B200_REQUIRE(dst && src, B200CTR_E_INVALID_ARG, "add2d: null pointer");
  if (rows * cols == 0) return B200CTR_OK;
  add2d_kernel<<<(unsigned)ceil_div(rows * cols, 256), 256, 0, as_stream(stream)>>>(dst, ld_dst, src, ld_src, rows, cols);
  B200_LAUNCH_CHECK("add2d_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_numbin_fold(const float* w, const float* b, const float* w_out, float* P, float* Q, int F, int fe,
                                   int D, void* stream) {
  B200_REQUIRE(w && w_out && P && Q, B200CTR_E_INVALID_ARG, "numbin_fold: null pointer");
  B200_REQUIRE(D > 0 && D <= 1024, B200CTR_E_UNSUPPORTED_SHAPE, "numbin_fold: D=%d unsupported", D);
  if (F == 0) return B200CTR_OK;
  numbin_fold_kernel<<<F, D, 0, as_stream(stream)>>>(w, b, w_out, P, Q, F, fe, D);
  B200_LAUNCH_CHECK("numbin_fold_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_numbin_fwd(const float* x, const float* P, const float* Q, float* out, int64_t ld_b,
                                  float* mean_out, int64_t ld_mean, int64_t B, int F, int D, void* stream) {
  B200_REQUIRE(x && P && Q && (out || mean_out), B200CTR_E_INVALID_ARG, "numbin_fwd: null pointer");
  B200_REQUIRE(D % 4 == 0 && D >= 4 && D <= 256 && 256 % (D / 4) == 0, B200CTR_E_UNSUPPORTED_SHAPE, "numbin_fwd: D=%d unsupported", D);
  B200_REQUIRE(ld_b % 4 == 0 && (!mean_out || ld_mean % 4 == 0), B200CTR_E_INVALID_ARG, "numbin_fwd: leading dims must be multiples of 4");
  if (B == 0 || F == 0) return B200CTR_OK;
  int grid = (int)std::min<int64_t>(B, (int64_t)num_sms() * 8);
  numbin_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, P, Q, out, ld_b, mean_out, ld_mean, B, F, D);
  B200_LAUNCH_CHECK("numbin_fwd_kernel");
  return B200CTR_OK;
}

constexpr int64_t NUMBIN_CHUNK = 64;
extern "C" int64_t b200ctr_numbin_bwd_rows(int64_t B) { return std::max<int64_t>(1, ceil_div(B, NUMBIN_CHUNK)); }

extern "C" int b200ctr_numbin_bwd(const float* x, const float* dout, int64_t ld_b, const float* dmean, int64_t ld_dmean,
                                  float* dP, float* dQ, int64_t B, int F, int D, void* stream) {
  B200_REQUIRE(x && dP && dQ && (dout || dmean), B200CTR_E_INVALID_ARG, "numbin_bwd: null pointer");
  B200_REQUIRE(D % 4 == 0, B200CTR_E_UNSUPPORTED_SHAPE, "numbin_bwd: D=%d unsupported", D);
  cudaStream_t st = as_stream(stream);
  if (B == 0 || F == 0) return B200CTR_OK;
  int64_t chunk = NUMBIN_CHUNK;
  dim3 grid((unsigned)ceil_div(B, chunk), (unsigned)std::min<int64_t>(ceil_div((int64_t)F * (D / 4), (int64_t)256), 8));
  numbin_bwd_kernel<<<grid, 256, 0, st>>>(x, dout, ld_b, dmean, ld_dmean, dP, dQ, B, F, D, chunk);
  B200_LAUNCH_CHECK("numbin_bwd_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_numbin_unfold(const float* dP, const float* dQ, const float* w, const float* b, const float* w_out,
                                     float* dw, float* db, float* dw_out, int F, int fe, int D, void* stream) {
  B200_REQUIRE(dP && dQ && w && w_out && dw && dw_out, B200CTR_E_INVALID_ARG, "numbin_unfold: null pointer");
  B200_REQUIRE(fe > 0 && fe <= 1024, B200CTR_E_UNSUPPORTED_SHAPE, "numbin_unfold: fe=%d unsupported", fe);
  numbin_unfold_kernel<<<F + D, fe, 0, as_stream(stream)>>>(dP, dQ, w, b, w_out, dw, db, dw_out, F, fe, D);
  B200_LAUNCH_CHECK("numbin_unfold_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_cat_fwd(const void* x_cat, int idx_is_64, const float* const* tables, const float* const* projs,
                               const int32_t* dims, const int32_t* cards, const int32_t* lds, float* out, int64_t ld_b, float* mean_out,
                               int64_t ld_mean, int q_idx, float* q_out, int64_t B, int Fc, int D, b200ctr_dropout_t drop,
                               void* stream) {
  B200_REQUIRE(x_cat && tables && projs && dims && cards && out, B200CTR_E_INVALID_ARG, "cat_fwd: null pointer");
  B200_REQUIRE(D == CAT_D, B200CTR_E_UNSUPPORTED_SHAPE, "cat_fwd: emb_dim=%d unsupported (kernel is specialised for 64)", D);
  if (B == 0 || Fc == 0) return B200CTR_OK;
  int grid = (int)ceil_div(B, (int64_t)CAT_WARPS * CAT_SPW);
  cudaStream_t st = as_stream(stream);
  if (idx_is_64)
    cat_fwd_kernel<int64_t><<<grid, CAT_WARPS * 32, 0, st>>>((const int64_t*)x_cat, tables, projs, dims, cards, lds, out, ld_b,
                                                              mean_out, ld_mean, q_idx, q_out, B, Fc, drop);
  else
    cat_fwd_kernel<int32_t><<<grid, CAT_WARPS * 32, 0, st>>>((const int32_t*)x_cat, tables, projs, dims, cards, lds, out, ld_b,
                                                              mean_out, ld_mean, q_idx, q_out, B, Fc, drop);
  B200_LAUNCH_CHECK("cat_fwd_kernel");
  return B200CTR_OK;
}

static size_t cat_al(size_t x) { return (x + 255) & ~(size_t)255; }
extern "C" int64_t b200ctr_cat_bwd_workspace_bytes(int64_t B, int Fc, int sum_dims, int64_t sum_cards) {
  const int64_t n = std::max<int64_t>(B, 1) * Fc;
  size_t t = cat_al((size_t)n * 4);                                       // keys: global table row of slot (b, c)
  t += cat_al((size_t)n * CAT_D * 4);                                     // 64-wide contribution rows of the slots
  t += cat_al((size_t)sum_cards * CAT_D * 4);                             // reduced rows of all tables
  t += cat_al((size_t)64 * CAT_D * sum_dims * 4);                         // per-CTA partials of all dW_c
  t += (size_t)b200ctr_scatter_rows_workspace_bytes(n, sum_cards);
  return (int64_t)t;
}

extern "C" int b200ctr_cat_bwd(const void* x_cat, int idx_is_64, const float* const* tables, const float* const* projs,
                               const int32_t* dims, const int32_t* cards, const int32_t* lds, const float* dout, int64_t ld_b,
                               const float* dmean, int64_t ld_dmean, int q_idx, const float* dq, float* const* dtables,
                               float* const* dprojs, int64_t B, int Fc, int D, int sum_dims, int64_t sum_cards, int max_card,
                               b200ctr_dropout_t drop, float* rows_acc, const int32_t* keep_rows, const int32_t* defer_slot,
                               int n_defer, float* defer_rows, int32_t* defer_keys, void* workspace,
                               int64_t workspace_bytes, void* stream) {
  B200_REQUIRE(!defer_slot || (n_defer > 0 && defer_rows && defer_keys), B200CTR_E_INVALID_ARG, "cat_bwd: deferred columns need their output buffers");
  B200_REQUIRE(x_cat && tables && projs && dims && cards && dtables && dprojs && workspace, B200CTR_E_INVALID_ARG, "cat_bwd: null pointer");
  B200_REQUIRE(D == CAT_D, B200CTR_E_UNSUPPORTED_SHAPE, "cat_bwd: emb_dim=%d unsupported (kernel is specialised for 64)", D);
  B200_REQUIRE(sum_dims > 0 && sum_cards > 0 && max_card > 0 && workspace_bytes >= b200ctr_cat_bwd_workspace_bytes(B, Fc, sum_dims, sum_cards),
               B200CTR_E_INVALID_ARG, "cat_bwd: workspace too small (sum_dims=%d sum_cards=%lld)", sum_dims, (long long)sum_cards);
  if (B == 0 || Fc == 0) return B200CTR_OK;
  dim3 grid((unsigned)std::min<int64_t>(ceil_div(B, (int64_t)CATB_TS), 64), (unsigned)Fc);     // wide columns (d_c up to 64) set the tail: many short CTAs
  cudaStream_t st = as_stream(stream);
  const int64_t n = B * Fc;
  char* w = (char*)workspace;
  int32_t* keys = (int32_t*)w; w += cat_al((size_t)n * 4);
  float* rowg = (float*)w; w += cat_al((size_t)n * CAT_D * 4);
  float* temp = (float*)w; w += cat_al((size_t)sum_cards * CAT_D * 4);
  float* dwpart = (float*)w; w += cat_al((size_t)64 * CAT_D * sum_dims * 4);
  const int64_t sc_bytes = b200ctr_scatter_rows_workspace_bytes(n, sum_cards);
  cudaError_t e = cudaMemsetAsync(rowg, 0, (size_t)n * CAT_D * 4, st);
  // rows_acc: a caller-owned, ALL-ZERO (sum_cards, 64) buffer that outlives the call (the trainer's: the rows of its row-lazy
  // tables are their gradient rows, the other tables' rows are zeroed again below); NULL: scratch from the workspace
  if (defer_slot && e == cudaSuccess) e = cudaMemsetAsync(defer_rows, 0, (size_t)B * n_defer * CAT_D * 4, st);
  const int rezero = rows_acc != nullptr;
  if (rows_acc) temp = rows_acc;
  else if (e == cudaSuccess) e = cudaMemsetAsync(temp, 0, (size_t)sum_cards * CAT_D * 4, st);
  if (e != cudaSuccess) return cuda_fail(e, "cat_bwd memset");
  if (idx_is_64)
    cat_bwd_kernel<int64_t><<<grid, 256, 0, st>>>((const int64_t*)x_cat, tables, projs, dims, cards, lds, dout, ld_b, dmean,
                                                   ld_dmean, q_idx, dq, rowg, keys, dwpart, sum_dims, B, Fc, drop, defer_slot, n_defer, defer_rows, defer_keys,
                                                   (int32_t)sum_cards);
  else
    cat_bwd_kernel<int32_t><<<grid, 256, 0, st>>>((const int32_t*)x_cat, tables, projs, dims, cards, lds, dout, ld_b, dmean,
                                                   ld_dmean, q_idx, dq, rowg, keys, dwpart, sum_dims, B, Fc, drop, defer_slot, n_defer, defer_rows, defer_keys,
                                                   (int32_t)sum_cards);
  B200_LAUNCH_CHECK("cat_bwd_kernel");
  // every distinct table row is reduced once, its slots in (sample, column) order
  int rc = b200ctr_scatter_rows(keys, 1, n, -1, sum_cards, 1, nullptr, nullptr, rowg, nullptr, nullptr, 0, 1, 0, nullptr, temp, w, sc_bytes, stream);
  if (rc) return rc;
  B200_REQUIRE(Fc <= CAT_MAXF, B200CTR_E_UNSUPPORTED_SHAPE, "cat_bwd: more than %d categorical columns", CAT_MAXF);
  (void)max_card;
  const int ublocks = (int)std::min<int64_t>(ceil_div(sum_cards * (CAT_D / 4), (int64_t)256), (int64_t)num_sms() * 16);
  cat_unpad_kernel<<<ublocks, 256, 0, st>>>(temp, dims, cards, keep_rows, rezero, dtables, Fc, sum_cards);
  B200_LAUNCH_CHECK("cat_unpad_kernel");
  dim3 wgrid((unsigned)ceil_div((int64_t)CAT_D * CAT_MAXDC, (int64_t)256), (unsigned)Fc);
  cat_dw_reduce_kernel<<<wgrid, 256, 0, st>>>(dwpart, (int)grid.x, dims, sum_dims, dprojs);
  B200_LAUNCH_CHECK("cat_dw_reduce_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_axpby(const float* x, const float* y, float a, float b, float* out, int64_t n, void* stream) {
  B200_REQUIRE(out && (x || y), B200CTR_E_INVALID_ARG, "axpby: null pointer");
  if (n == 0) return B200CTR_OK;
  axpby_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(x, y, a, b, out, n);
  B200_LAUNCH_CHECK("axpby_kernel");
  return B200CTR_OK;
}
