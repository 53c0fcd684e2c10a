// Deterministic scatter-add of the sequence-table gradients ("K2a bwd": sort once, segmented reduction).
//
// Only the K selected tokens of a sample receive gradient (top-k backward, src/models/dare.py:135-137 + the embedding backward of
// dare.py:122-123): slot s = (b, k) contributes  d emb_rep[tok_s] += dsel_rep[s]  and  d emb_att[tok_s] += dvals[s] * q[b].
// The first version added these B*K rows with fp32 vector atomics, so a row that several slots hit was summed in whatever
// order the atomics landed: results differed from run to run at the 1e-7 level.  Here the (token, slot) pairs are radix-
// sorted by token (stable: equal tokens stay in slot order), and every distinct row is reduced ONCE, in slot order, by plain
// loads and one plain store:
//   1. keys       key_i = token (or a sentinel for pad slots / rows another rank owns), val_i = i
//   2. sort       cub::DeviceRadixSort::SortPairs over ceil(log2(vocab + 1)) bits
//   3. reduce     one half-warp per CHUNK of 32 sorted slots walks its slots in order; a run of equal tokens that begins and
//                 ends inside the chunk is stored straight to the gradient row; a run cut by a chunk border goes to one of the
//                 chunk's two partial slots (HEAD: continues a row from the previous chunk, TAIL: starts a row that goes on)
//   4. stitch     the half-warp of the chunk that holds a row's TAIL adds the HEAD partials of the following chunks in chunk
//                 order and stores the row
// Every sum has one fixed association order that depends only on (token, slot), so two runs are bit-identical; hot tokens
// (thousands of slots under a Zipf stream) are spread over chunks instead of serialising one warp.
// Two source forms: DIRECT (one GPU / replicated tables: contributions are read from dsel_rep, dvals and q) and STAGED
// (row-sharded tables: every rank stored its slots' rows into the owner's peer-mapped staging buffer, retrieve.cu; the
// owner reduces rows of ALL ranks' slots in (source rank, slot) order).
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>

namespace b200ctr {

constexpr int SC_D = 64;
constexpr int SC_CHUNK = 32;

struct ScSrc {
  // direct
  const float* q;          // (B, 64)
  const float* dvals;      // (n_slots)
  const float* rows_rep;   // (n_slots, 64)
  int K;                   // slots per sample
  // staged
  const float4* stage_att; // (n_src, cap, 16) float4
  const float4* stage_rep;
  int64_t n_slots, cap;    // slots per source rank, staging capacity per source rank
  int staged;
};

__global__ void sc_keys_kernel(const int32_t* __restrict__ tok, int64_t n, int pad_id, int mask, int rank, uint32_t sentinel,
                               uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int t = tok[i];
  const bool mine = t != pad_id && t >= 0 && (uint32_t)t < sentinel && (t & mask) == rank;
  keys[i] = mine ? (uint32_t)t : sentinel;
  vals[i] = (uint32_t)i;
}

__device__ __forceinline__ void sc_load(const ScSrc& S, uint32_t slot, int hl, float4& a, float4& r, bool want_att, bool want_rep) {
  if (S.staged) {
    const int64_t src = slot / S.n_slots, i = slot - src * S.n_slots;
    const int64_t row = (src * S.cap + i) * (SC_D / 4) + hl;
    if (want_att) a = S.stage_att[row];
    if (want_rep) r = S.stage_rep[row];
  } else {
    if (want_att) {
      const float dv = S.dvals[slot];
      const float4 qv = *reinterpret_cast<const float4*>(S.q + (int64_t)(slot / S.K) * SC_D + hl * 4);
      a = make_float4(dv * qv.x, dv * qv.y, dv * qv.z, dv * qv.w);          // the same products the atomic version added
    }
    if (want_rep) r = ldg_nc_f4(reinterpret_cast<const float4*>(S.rows_rep + (int64_t)slot * SC_D) + hl);
  }
}
__device__ __forceinline__ void f4_add(float4& x, const float4& y) { x.x += y.x; x.y += y.y; x.z += y.z; x.w += y.w; }

// meta[c] = {head_key, tail_key, head_continues, 0}; partial rows: part[(c * 2 + which) * 2 + table]
__global__ void __launch_bounds__(256, 3)
sc_reduce_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ vals, int64_t n, uint32_t sentinel, int shift,
                 ScSrc S, float* __restrict__ g_att, float* __restrict__ g_rep, float4* __restrict__ part, int4* __restrict__ meta,
                 int64_t n_chunks) {
  const int hl = threadIdx.x & 15, lane = threadIdx.x & 31, hbase = lane & 16;
  const uint32_t hmask = 0xffffu << hbase;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  if (c >= n_chunks) return;
  const int64_t p0 = c * SC_CHUNK;
  const int cnt = (int)min((int64_t)SC_CHUNK, n - p0);
  // lane hl holds the keys / slots of chunk positions hl and hl + 16
  const uint32_t k_lo = hl < cnt ? keys[p0 + hl] : sentinel, k_hi = hl + 16 < cnt ? keys[p0 + hl + 16] : sentinel;
  const uint32_t v_lo = hl < cnt ? vals[p0 + hl] : 0u, v_hi = hl + 16 < cnt ? vals[p0 + hl + 16] : 0u;
  const uint32_t k_prev = p0 > 0 ? keys[p0 - 1] : sentinel;                         // sentinel never equals a live key
  const uint32_t k_next = p0 + cnt < n ? keys[p0 + cnt] : sentinel;
  const bool want_att = g_att != nullptr, want_rep = g_rep != nullptr;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 aa = z4, ar = z4;
  int4 m = make_int4(-1, -1, 0, 0);
  uint32_t cur = __shfl_sync(hmask, k_lo, hbase);
  int run_start = 0;
  auto flush = [&](uint32_t key, int a, int b) {        // run [a, b) of `key` inside this chunk (all 16 lanes take the same path)
    if (key == sentinel) return;
    const bool starts = a > 0 || key != k_prev;
    const bool ends = b < cnt || key != k_next;
    if (starts && ends) {
      const int64_t row = (int64_t)(key >> shift) * (SC_D / 4) + hl;
      if (want_att) reinterpret_cast<float4*>(g_att)[row] = aa;
      if (want_rep) reinterpret_cast<float4*>(g_rep)[row] = ar;
    } else {
      const int which = starts ? 1 : 0;                 // 1 = TAIL (the row starts here and goes on), 0 = HEAD (continues a row)
      float4* dst = part + ((c * 2 + which) * 2) * (SC_D / 4) + hl;
      if (want_att) dst[0] = aa;
      if (want_rep) dst[SC_D / 4] = ar;
      if (which) m.y = (int)key;
      else { m.x = (int)key; m.z = ends ? 0 : 1; m.w = (!ends && a == 0 && b == SC_CHUNK) ? 1 : 0; }   // w: the WHOLE chunk is one row that goes on
    }
  };
  // The rows of SC_PF consecutive slots are fetched before any of them is added (independent loads in flight; the adds stay
  // in slot order, so the association order -- and the result bits -- do not change).
  constexpr int SC_PF = 4;
  bool done = false;
  for (int i0 = 0; i0 < cnt && !done; i0 += SC_PF) {
    uint32_t kk[SC_PF];
    float4 pa[SC_PF], pr[SC_PF];
#pragma unroll
    for (int u = 0; u < SC_PF; ++u) {
      const int i = i0 + u;
      kk[u] = __shfl_sync(hmask, i < 16 ? k_lo : k_hi, hbase + (i & 15));
      const uint32_t v = __shfl_sync(hmask, i < 16 ? v_lo : v_hi, hbase + (i & 15));
      if (i >= cnt) kk[u] = sentinel;
      pa[u] = z4; pr[u] = z4;
      if (kk[u] != sentinel) sc_load(S, v, hl, pa[u], pr[u], want_att, want_rep);
    }
#pragma unroll
    for (int u = 0; u < SC_PF; ++u) {
      const int i = i0 + u;
      if (i >= cnt) break;
      const uint32_t k = kk[u];
      if (k != cur) {
        flush(cur, run_start, i);
        cur = k; run_start = i; aa = z4; ar = z4;
      }
      if (k == sentinel) { done = true; break; }        // sorted: only pad / foreign slots follow
      f4_add(aa, pa[u]); f4_add(ar, pr[u]);
    }
  }
  if (cur != sentinel) {
    int end = cnt;
    // the loop may have stopped at a sentinel: then the run ended at that position and was flushed above (cur == sentinel)
    flush(cur, run_start, end);
  }
  if (hl == 0) meta[c] = m;
}

// Level 2: a row that collects thousands of slots (the hot tokens of a Zipf stream: 8 % of all slots for the hottest one) spans
// hundreds of chunks, and stitching them one by one would serialise one half-warp for a millisecond.  One half-warp per GROUP of
// 32 chunks: if all 32 are FULL chunks of the same row, their HEAD partials are added in chunk order into one group partial, and
// the stitch below steps over the whole group at once.  The association order still depends on the sorted data only.
constexpr int SC_GROUP = 32;
__global__ void __launch_bounds__(256)
sc_group_kernel(const float4* __restrict__ part, const int4* __restrict__ meta, int64_t n_chunks, int64_t n_groups, bool want_att,
                bool want_rep, float4* __restrict__ gpart, int32_t* __restrict__ gmeta) {
  const int hl = threadIdx.x & 15, hbase = threadIdx.x & 16;
  const uint32_t hmask = 0xffffu << hbase;
  const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  if (g >= n_groups) return;
  const int64_t c0 = g * SC_GROUP;
  bool solid = c0 + SC_GROUP <= n_chunks;
  int key = -1;
  if (solid) {
    const int4 m0 = meta[c0 + hl], m1 = meta[c0 + hl + 16];
    key = __shfl_sync(hmask, m0.x, hbase);
    const bool ok = m0.w && m1.w && m0.x == key && m1.x == key && key >= 0;
    solid = __all_sync(hmask, ok);
  }
  if (!solid) { if (hl == 0) gmeta[g] = -1; return; }
  float4 aa = make_float4(0.f, 0.f, 0.f, 0.f), ar = aa;
#pragma unroll 8
  for (int i = 0; i < SC_GROUP; ++i) {
    const float4* h = part + (((c0 + i) * 2 + 0) * 2) * (SC_D / 4) + hl;
    if (want_att) f4_add(aa, h[0]);
    if (want_rep) f4_add(ar, h[SC_D / 4]);
  }
  float4* dst = gpart + (g * 2) * (SC_D / 4) + hl;
  if (want_att) dst[0] = aa;
  if (want_rep) dst[SC_D / 4] = ar;
  if (hl == 0) gmeta[g] = key;
}

__global__ void __launch_bounds__(256)
sc_stitch_kernel(const float4* __restrict__ part, const int4* __restrict__ meta, const float4* __restrict__ gpart,
                 const int32_t* __restrict__ gmeta, int64_t n_chunks, int64_t n_groups, int shift,
                 float* __restrict__ g_att, float* __restrict__ g_rep) {
  const int hl = threadIdx.x & 15;
  const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
  if (c >= n_chunks) return;
  const int4 m = meta[c];
  if (m.y < 0) return;                                   // no row starts in this chunk and continues
  const int key = m.y;
  const bool want_att = g_att != nullptr, want_rep = g_rep != nullptr;
  const float4* src = part + ((c * 2 + 1) * 2) * (SC_D / 4) + hl;
  float4 aa = want_att ? src[0] : make_float4(0.f, 0.f, 0.f, 0.f), ar = want_rep ? src[SC_D / 4] : make_float4(0.f, 0.f, 0.f, 0.f);
  int64_t cc = c + 1;
  while (cc < n_chunks) {
    if ((cc % SC_GROUP) == 0 && cc / SC_GROUP < n_groups && gmeta[cc / SC_GROUP] == key) {
      const float4* gp = gpart + ((cc / SC_GROUP) * 2) * (SC_D / 4) + hl;   // 32 full chunks of this row at once
      if (want_att) f4_add(aa, gp[0]);
      if (want_rep) f4_add(ar, gp[SC_D / 4]);
      cc += SC_GROUP;                                    // a solid group always goes on into the next chunk
      continue;
    }
    const int4 mm = meta[cc];
    if (mm.x != key) break;
    const float4* h = part + ((cc * 2 + 0) * 2) * (SC_D / 4) + hl;
    if (want_att) f4_add(aa, h[0]);
    if (want_rep) f4_add(ar, h[SC_D / 4]);
    if (!mm.z) break;
    ++cc;
  }
  const int64_t row = (int64_t)(key >> shift) * (SC_D / 4) + hl;
  if (want_att) reinterpret_cast<float4*>(g_att)[row] = aa;
  if (want_rep) reinterpret_cast<float4*>(g_rep)[row] = ar;
}

static int key_bits(int64_t vocab) {
  int b = 1;
  while ((1LL << b) <= vocab) ++b;                        // the sentinel `vocab` itself must be representable
  return b;
}

static size_t sort_temp_bytes(int64_t n, int bits) {
  size_t t = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, t, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (int)n, 0, bits);
  return t;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int64_t b200ctr_scatter_rows_workspace_bytes(int64_t n_total_slots, int64_t vocab) {
  if (n_total_slots <= 0) return 256;
  const int64_t n_chunks = ceil_div(n_total_slots, (int64_t)SC_CHUNK);
  size_t total = 4 * align256((size_t)n_total_slots * 4);                                 // keys, vals, sorted keys, sorted vals
  total += align256(sort_temp_bytes(n_total_slots, key_bits(vocab)));
  total += align256((size_t)n_chunks * 2 * 2 * SC_D * 4);                                 // partial rows
  total += align256((size_t)n_chunks * sizeof(int4));
  const int64_t n_groups = ceil_div(n_chunks, (int64_t)SC_GROUP);
  total += align256((size_t)n_groups * 2 * SC_D * 4) + align256((size_t)n_groups * 4);      // group partials + group keys
  return (int64_t)total;
}

// tok: n_src * n_slots token ids (n_src = 1 in the direct form).  Rows of tokens t with (t & (n_shards-1)) == rank are
// reduced into g_att / g_rep (local row t >> log2(n_shards)); the rows must be zero on entry (each touched row is WRITTEN).
extern "C" int b200ctr_scatter_rows(const int32_t* tok, int n_src, int64_t n_slots, int pad_id, int64_t vocab, int K,
                                    const float* q, const float* dvals, const float* rows_rep, const float* stage_att,
                                    const float* stage_rep, int64_t stage_cap, int n_shards, int rank, float* g_att,
                                    float* g_rep, void* workspace, int64_t workspace_bytes, void* stream) {
  B200_REQUIRE(tok && workspace && n_src >= 1 && n_slots >= 0 && K >= 1, B200CTR_E_INVALID_ARG, "scatter_rows: bad argument");
  B200_REQUIRE(n_shards >= 1 && n_shards <= 8 && !(n_shards & (n_shards - 1)) && rank >= 0 && rank < n_shards,
               B200CTR_E_INVALID_ARG, "scatter_rows: n_shards=%d rank=%d", n_shards, rank);
  const bool staged = stage_att || stage_rep;
  B200_REQUIRE(staged || n_src == 1, B200CTR_E_INVALID_ARG, "scatter_rows: several source ranks need the staged form");
  B200_REQUIRE(!g_att || (staged ? stage_att != nullptr : (q && dvals)), B200CTR_E_INVALID_ARG, "scatter_rows: emb_att sources missing");
  B200_REQUIRE(!g_rep || (staged ? stage_rep != nullptr : rows_rep != nullptr), B200CTR_E_INVALID_ARG, "scatter_rows: emb_rep sources missing");
  B200_REQUIRE(!staged || stage_cap >= n_slots, B200CTR_E_INVALID_ARG, "scatter_rows: stage_cap < n_slots");
  const int64_t n = (int64_t)n_src * n_slots;
  B200_REQUIRE(n < (1LL << 31) && vocab < (1LL << 31), B200CTR_E_UNSUPPORTED_SHAPE, "scatter_rows: too many slots / rows");
  if (n == 0 || (!g_att && !g_rep)) return B200CTR_OK;
  B200_REQUIRE(workspace_bytes >= b200ctr_scatter_rows_workspace_bytes(n, vocab), B200CTR_E_INVALID_ARG, "scatter_rows: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int bits = key_bits(vocab);
  const int64_t n_chunks = ceil_div(n, (int64_t)SC_CHUNK);
  char* w = (char*)workspace;
  uint32_t* keys = (uint32_t*)w; w += align256((size_t)n * 4);
  uint32_t* vals = (uint32_t*)w; w += align256((size_t)n * 4);
  uint32_t* skeys = (uint32_t*)w; w += align256((size_t)n * 4);
  uint32_t* svals = (uint32_t*)w; w += align256((size_t)n * 4);
  size_t tb = sort_temp_bytes(n, bits);
  void* temp = w; w += align256(tb);
  float4* part = (float4*)w; w += align256((size_t)n_chunks * 2 * 2 * SC_D * 4);
  int4* meta = (int4*)w; w += align256((size_t)n_chunks * sizeof(int4));
  const int64_t n_groups = ceil_div(n_chunks, (int64_t)SC_GROUP);
  float4* gpart = (float4*)w; w += align256((size_t)n_groups * 2 * SC_D * 4);
  int32_t* gmeta = (int32_t*)w;
  int shift = 0;
  while ((1 << shift) < n_shards) ++shift;
  sc_keys_kernel<<<(unsigned)ceil_div(n, (int64_t)256), 256, 0, st>>>(tok, n, pad_id, n_shards - 1, rank, (uint32_t)vocab, keys, vals);
  B200_LAUNCH_CHECK("sc_keys_kernel");
  cudaError_t e = cub::DeviceRadixSort::SortPairs(temp, tb, keys, skeys, vals, svals, (int)n, 0, bits, st);
  if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceRadixSort::SortPairs");
  ScSrc S{};
  S.q = q; S.dvals = dvals; S.rows_rep = rows_rep; S.K = K;
  S.stage_att = (const float4*)stage_att; S.stage_rep = (const float4*)stage_rep; S.n_slots = n_slots; S.cap = stage_cap;
  S.staged = staged ? 1 : 0;
  const unsigned grid = (unsigned)ceil_div(n_chunks * 16, (int64_t)256);
  sc_reduce_kernel<<<grid, 256, 0, st>>>(skeys, svals, n, (uint32_t)vocab, shift, S, g_att, g_rep, part, meta, n_chunks);
  B200_LAUNCH_CHECK("sc_reduce_kernel");
  sc_group_kernel<<<(unsigned)ceil_div(n_groups * 16, (int64_t)256), 256, 0, st>>>(part, meta, n_chunks, n_groups, g_att != nullptr,
                                                                                    g_rep != nullptr, gpart, gmeta);
  B200_LAUNCH_CHECK("sc_group_kernel");
  sc_stitch_kernel<<<grid, 256, 0, st>>>(part, meta, gpart, gmeta, n_chunks, n_groups, shift, g_att, g_rep);
  B200_LAUNCH_CHECK("sc_stitch_kernel");
  return B200CTR_OK;
}
