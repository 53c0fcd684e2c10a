// On-GPU cache builder: the byte / integer kernels that turn Arrow column buffers into the App. D shard
// arrays the hot path trains from.  Reference: src/data/build_cache_v1.py:78-173 (`_process_arrow_batch`: categorical hash
// 123-132, NaN mask + median imputation 135-146, `seq` string parse 149-158, groups 97-113) and 28-75 (global medians).
//
// The reference hashes with polars.Series.hash(seed=2025), which Polars documents as unstable across versions and which is
// absent from this image ("parity unpinned"), so this builder defines cache format v2 with a PINNED hash:
//     h = splitmix64_finalize( fnv1a64(bytes, basis = 0xcbf29ce484222325 ^ seed) )        (seed 2025; "NA" for nulls)
// over the UTF-8 bytes of the value's string form (integers: their decimal digits) -- everything else (mask, imputation,
// nan_to_num, right-aligned last-max_len tokens, dtypes, manifest) is the reference's semantics bit for bit.
// All kernels are HBM-bound byte work: coalesced column reads, shared-memory transposes to the row-major outputs.
#include "common.cuh"
#include <math.h>
#include <algorithm>

namespace b200ctr {

// ---- pinned 64-bit hash ------------------------------------------------------------------------------------------------
constexpr uint64_t CB_FNV_BASIS = 0xcbf29ce484222325ULL, CB_FNV_PRIME = 0x100000001b3ULL;
__host__ __device__ __forceinline__ uint64_t cb_fnv_step(uint64_t h, uint8_t b) { return (h ^ (uint64_t)b) * CB_FNV_PRIME; }
__host__ __device__ __forceinline__ uint64_t cb_finalize(uint64_t z) {          // splitmix64 finalizer
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ bool cb_valid(const uint8_t* __restrict__ validity, int64_t r) {
  return validity == nullptr || ((validity[r >> 3] >> (r & 7)) & 1);
}
__device__ __forceinline__ void cb_store(void* out, int out32, int64_t idx, uint64_t v) {
  if (out32) reinterpret_cast<int32_t*>(out)[idx] = (int32_t)v;
  else reinterpret_cast<int64_t*>(out)[idx] = (int64_t)v;
}

// one thread per row: values are a few bytes long
__global__ void cache_hash_strings_kernel(const uint8_t* __restrict__ bytes, const void* __restrict__ offsets, int off64,
                                          const uint8_t* __restrict__ validity, int64_t rows, uint64_t mod, uint64_t seed,
                                          void* __restrict__ out, int64_t out_stride, int out32) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  uint64_t h = CB_FNV_BASIS ^ seed;
  if (cb_valid(validity, r)) {
    const int64_t s = off64 ? reinterpret_cast<const int64_t*>(offsets)[r] : reinterpret_cast<const int32_t*>(offsets)[r];
    const int64_t e = off64 ? reinterpret_cast<const int64_t*>(offsets)[r + 1] : reinterpret_cast<const int32_t*>(offsets)[r + 1];
    for (int64_t i = s; i < e; ++i) h = cb_fnv_step(h, bytes[i]);
  } else {
    h = cb_fnv_step(cb_fnv_step(h, 'N'), 'A');
  }
  cb_store(out, out32, r * out_stride, cb_finalize(h) % mod);
}

// integers hash as their decimal string (what `cast(pl.Utf8)` yields): '-' then the digits, most significant first
__global__ void cache_hash_int64_kernel(const int64_t* __restrict__ vals, const uint8_t* __restrict__ validity, int64_t rows,
                                        uint64_t mod, uint64_t seed, void* __restrict__ out, int64_t out_stride, int out32) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  uint64_t h = CB_FNV_BASIS ^ seed;
  if (cb_valid(validity, r)) {
    const int64_t v = vals[r];
    uint64_t a = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    uint8_t dig[20];
    int n = 0;
    do { dig[n++] = (uint8_t)('0' + a % 10); a /= 10; } while (a);
    if (v < 0) h = cb_fnv_step(h, '-');
    for (int i = n - 1; i >= 0; --i) h = cb_fnv_step(h, dig[i]);
  } else {
    h = cb_fnv_step(cb_fnv_step(h, 'N'), 'A');
  }
  cb_store(out, out32, r * out_stride, cb_finalize(h) % mod);
}

// ---- seq: "12,345,6789" -> the last max_len tokens, right-aligned, pad_id on the left  (build_cache_v1.py:149-158) ------
// One warp per row, two scans of the row's bytes 32 at a time: a token starts at a digit whose predecessor is not a digit;
// scan 1 counts the tokens, scan 2 numbers them (ballot + popc) and the lane that owns a surviving token's first byte parses
// its digits.  Bytes other than digits and ',' are counted in n_bad (Python's int() would raise; the host raises), so are
// tokens above INT32_MAX.
__global__ void cache_parse_seq_kernel(const uint8_t* __restrict__ bytes, const void* __restrict__ offsets, int off64,
                                       const uint8_t* __restrict__ validity, int64_t rows, int max_len, int pad_id,
                                       int32_t* __restrict__ seq_out, unsigned long long* __restrict__ n_bad) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (r >= rows) return;
  int64_t s = 0, e = 0;
  if (cb_valid(validity, r)) {
    s = off64 ? reinterpret_cast<const int64_t*>(offsets)[r] : reinterpret_cast<const int32_t*>(offsets)[r];
    e = off64 ? reinterpret_cast<const int64_t*>(offsets)[r + 1] : reinterpret_cast<const int32_t*>(offsets)[r + 1];
  }
  const int64_t n = e - s;
  const uint8_t* p = bytes + s;
  int n_tok = 0, bad = 0;
  bool carry = false;
#pragma unroll 4
  for (int64_t base = 0; base < n; base += 32) {
    const int64_t i = base + lane;
    const uint8_t c = i < n ? p[i] : (uint8_t)',';
    const bool d = c >= '0' && c <= '9';
    bool pd = __shfl_up_sync(0xffffffffu, d, 1);
    if (lane == 0) pd = carry;
    bad += (!d && c != ',') ? 1 : 0;
    n_tok += __popc(__ballot_sync(0xffffffffu, d && !pd));
    carry = __shfl_sync(0xffffffffu, d, 31);
  }
  const int keep = min(n_tok, max_len), skip = n_tok - keep;
  int32_t* orow = seq_out + r * (int64_t)max_len;
  for (int j = lane; j < max_len - keep; j += 32) orow[j] = pad_id;
  int seen = 0;
  carry = false;
#pragma unroll 2
  for (int64_t base = 0; base < n && keep > 0; base += 32) {
    const int64_t i = base + lane;
    const uint8_t c = i < n ? p[i] : (uint8_t)',';
    const bool d = c >= '0' && c <= '9';
    bool pd = __shfl_up_sync(0xffffffffu, d, 1);
    if (lane == 0) pd = carry;
    const bool start = d && !pd;
    const uint32_t m = __ballot_sync(0xffffffffu, start);
    const int k = seen + __popc(m & ((1u << lane) - 1u));
    if (start && k >= skip) {
      unsigned long long v = 0;
      int nd = 0;
      for (int64_t j = i; j < n; ++j) {
        const uint8_t cj = p[j];
        if (cj < '0' || cj > '9') break;
        v = v * 10ull + (unsigned long long)(cj - '0');
        if (++nd > 18) break;
      }
      if (v > 2147483647ull || nd > 18) { ++bad; v = 0; }
      orow[max_len - (n_tok - k)] = (int32_t)v;
    }
    seen += __popc(m);
    carry = __shfl_sync(0xffffffffu, d, 31);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if (lane == 0 && bad) atomicAdd(n_bad, (unsigned long long)bad);
}

// ---- numeric block: column buffers -> row-major float32 + NaN mask, imputed and clamped  (build_cache_v1.py:135-146) -----
// dtype codes of the column buffers
enum { CB_F32 = 0, CB_F64 = 1, CB_I32 = 2, CB_I64 = 3, CB_U8 = 4 };
__device__ __forceinline__ float cb_load_num(const void* col, int dtype, const uint8_t* validity, int64_t r) {
  if (!cb_valid(validity, r)) return nanf("");
  switch (dtype) {
    case CB_F32: return reinterpret_cast<const float*>(col)[r];
    case CB_F64: return (float)reinterpret_cast<const double*>(col)[r];
    case CB_I32: return (float)(double)reinterpret_cast<const int32_t*>(col)[r];
    case CB_I64: return (float)(double)reinterpret_cast<const int64_t*>(col)[r];
    default: return (float)reinterpret_cast<const uint8_t*>(col)[r];
  }
}
// float32 copy of one column with nulls as NaN (the resident column-major matrix the medians are selected from)
__global__ void cache_col_to_f32_kernel(const void* __restrict__ col, int dtype, const uint8_t* __restrict__ validity, int64_t rows,
                                        float* __restrict__ out) {
  const int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (r < rows) out[r] = cb_load_num(col, dtype, validity, r);
}
// x: [Fn][ld] column-major float32 (NaN = missing), rows r0 .. r0 + rows of it -> X_num / X_mask [rows][Fn] row-major
__global__ void cache_numeric_kernel(const float* __restrict__ x, int64_t ld, int64_t r0, int64_t rows, int Fn,
                                     const float* __restrict__ med, float* __restrict__ X_num, uint8_t* __restrict__ X_mask) {
  __shared__ float tile[32][33];
  __shared__ uint8_t mt[32][33];
  const int64_t rb = (int64_t)blockIdx.x * 32;
  const int cb = blockIdx.y * 32;
  for (int cc = threadIdx.y; cc < 32; cc += blockDim.y) {         // lanes along rows: coalesced reads of one column
    const int c = cb + cc;
    const int64_t r = rb + threadIdx.x;
    float v = 0.f;
    uint8_t mk = 0;
    if (c < Fn && r < rows) {
      v = x[(int64_t)c * ld + r0 + r];
      mk = isnan(v) ? 1 : 0;
      if (mk) v = med[c];
      if (isnan(v)) v = 0.f;                                      // np.nan_to_num(nan=0, posinf=1e6, neginf=-1e6)
      else if (isinf(v)) v = v > 0.f ? 1e6f : -1e6f;
    }
    tile[cc][threadIdx.x] = v;
    mt[cc][threadIdx.x] = mk;
  }
  __syncthreads();
  for (int rr = threadIdx.y; rr < 32; rr += blockDim.y) {         // lanes along columns: coalesced writes of one row
    const int64_t r = rb + rr;
    const int c = cb + threadIdx.x;
    if (r < rows && c < Fn) {
      X_num[r * Fn + c] = tile[threadIdx.x][rr];
      X_mask[r * Fn + c] = mt[threadIdx.x][rr];
    }
  }
}

// ---- exact medians by radix select  (build_cache_v1.py:55-66: pl.col(c).median(), missing values ignored) -----------------
// Order-preserving key of a float; one 8-bit histogram pass per call over the elements whose key matches `prefix` above the
// current digit.  Query q = (column, prefix, shift); NaNs never match (they are counted separately in pass 0 via n_valid).
__device__ __forceinline__ uint32_t cb_key(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__global__ void cache_select_hist_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, const int* __restrict__ q_col,
                                         const uint32_t* __restrict__ q_prefix, int shift, unsigned long long* __restrict__ hist) {
  __shared__ unsigned int h[256];
  const int q = blockIdx.y;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) h[i] = 0u;
  __syncthreads();
  const float* col = x + (int64_t)q_col[q] * ld;
  const uint32_t prefix = q_prefix[q];
  const uint32_t himask = shift >= 24 ? 0u : (0xffffffffu << (shift + 8));
  for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < rows; r += (int64_t)gridDim.x * blockDim.x) {
    const float v = col[r];
    if (isnan(v)) continue;
    const uint32_t k = cb_key(v);
    if ((k & himask) == (prefix & himask)) atomicAdd(&h[(k >> shift) & 255u], 1u);        // integer atomics: order-independent
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 256; i += blockDim.x)
    if (h[i]) atomicAdd(&hist[(int64_t)q * 256 + i], (unsigned long long)h[i]);
}

}  // namespace b200ctr

using namespace b200ctr;

extern "C" int64_t b200ctr_cache_hash_bytes_host(const void* bytes, int64_t n, uint64_t seed) {
  uint64_t h = CB_FNV_BASIS ^ seed;
  for (int64_t i = 0; i < n; ++i) h = cb_fnv_step(h, reinterpret_cast<const uint8_t*>(bytes)[i]);
  return (int64_t)cb_finalize(h);
}

extern "C" int b200ctr_cache_hash_strings(const void* bytes, const void* offsets, int offsets64, const void* validity, int64_t rows,
                                          uint64_t mod, uint64_t seed, void* out, int64_t out_stride, int out_int32, void* stream) {
  B200_REQUIRE(offsets && out && mod > 0, B200CTR_E_INVALID_ARG, "cache_hash_strings: null pointer or zero modulus");
  if (rows == 0) return B200CTR_OK;
  cache_hash_strings_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, as_stream(stream)>>>(
      (const uint8_t*)bytes, offsets, offsets64, (const uint8_t*)validity, rows, mod, seed, out, out_stride, out_int32);
  B200_LAUNCH_CHECK("cache_hash_strings_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_cache_hash_int64(const void* vals, const void* validity, int64_t rows, uint64_t mod, uint64_t seed, void* out,
                                        int64_t out_stride, int out_int32, void* stream) {
  B200_REQUIRE(vals && out && mod > 0, B200CTR_E_INVALID_ARG, "cache_hash_int64: null pointer or zero modulus");
  if (rows == 0) return B200CTR_OK;
  cache_hash_int64_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, as_stream(stream)>>>(
      (const int64_t*)vals, (const uint8_t*)validity, rows, mod, seed, out, out_stride, out_int32);
  B200_LAUNCH_CHECK("cache_hash_int64_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_cache_parse_seq(const void* bytes, const void* offsets, int offsets64, const void* validity, int64_t rows,
                                       int max_len, int pad_id, int32_t* seq_out, void* n_bad, void* stream) {
  B200_REQUIRE(offsets && seq_out && n_bad && max_len > 0, B200CTR_E_INVALID_ARG, "cache_parse_seq: null pointer or max_len <= 0");
  if (rows == 0) return B200CTR_OK;
  cache_parse_seq_kernel<<<(unsigned)ceil_div(rows * 32, 256), 256, 0, as_stream(stream)>>>(
      (const uint8_t*)bytes, offsets, offsets64, (const uint8_t*)validity, rows, max_len, pad_id, seq_out, (unsigned long long*)n_bad);
  B200_LAUNCH_CHECK("cache_parse_seq_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_cache_col_to_f32(const void* col, int dtype, const void* validity, int64_t rows, float* out, void* stream) {
  B200_REQUIRE(col && out, B200CTR_E_INVALID_ARG, "cache_col_to_f32: null pointer");
  B200_REQUIRE(dtype >= CB_F32 && dtype <= CB_U8, B200CTR_E_INVALID_ARG, "cache_col_to_f32: dtype code %d", dtype);
  if (rows == 0) return B200CTR_OK;
  cache_col_to_f32_kernel<<<(unsigned)ceil_div(rows, 256), 256, 0, as_stream(stream)>>>(col, dtype, (const uint8_t*)validity, rows, out);
  B200_LAUNCH_CHECK("cache_col_to_f32_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_cache_numeric(const float* x, int64_t ld, int64_t r0, int64_t rows, int Fn, const float* med, float* X_num,
                                     void* X_mask, void* stream) {
  B200_REQUIRE(x && med && X_num && X_mask, B200CTR_E_INVALID_ARG, "cache_numeric: null pointer");
  if (rows == 0 || Fn == 0) return B200CTR_OK;
  dim3 grid((unsigned)ceil_div(rows, 32), (unsigned)ceil_div(Fn, 32));
  cache_numeric_kernel<<<grid, dim3(32, 8), 0, as_stream(stream)>>>(x, ld, r0, rows, Fn, med, X_num, (uint8_t*)X_mask);
  B200_LAUNCH_CHECK("cache_numeric_kernel");
  return B200CTR_OK;
}

extern "C" int b200ctr_cache_select_hist(const float* x, int64_t ld, int64_t rows, const int* q_col, const uint32_t* q_prefix, int n_q,
                                         int shift, void* hist, void* stream) {
  B200_REQUIRE(x && q_col && q_prefix && hist, B200CTR_E_INVALID_ARG, "cache_select_hist: null pointer");
  B200_REQUIRE(shift == 0 || shift == 8 || shift == 16 || shift == 24, B200CTR_E_INVALID_ARG, "cache_select_hist: shift %d", shift);
  if (rows == 0 || n_q == 0) return B200CTR_OK;
  const int bx = (int)std::min<int64_t>(ceil_div(rows, 256 * 8), (int64_t)num_sms() * 4);
  cache_select_hist_kernel<<<dim3((unsigned)std::max(bx, 1), (unsigned)n_q), 256, 0, as_stream(stream)>>>(
      x, ld, rows, q_col, q_prefix, shift, (unsigned long long*)hist);
  B200_LAUNCH_CHECK("cache_select_hist_kernel");
  return B200CTR_OK;
}
