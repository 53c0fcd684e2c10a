// Host-side row gather for the batched shard reader (the "feed the GPU" row of SURVEY 8f-1) + the device-side widening of
// the narrow on-disk dtypes.
//
// The reference reads one row at a time in Python (src/data/dataset.py:28-43 ShardedNPY.get_row over np.load(mmap_mode="r"))
// and stacks B rows per batch in collate_sharded (dataset.py:98-124): ~10^3-10^4 samples/s per worker.  Here a batch is
// gathered by a handful of host threads straight from the memory-mapped .npy payloads into pinned staging buffers, in the
// on-disk dtypes (uint8 masks, int32 ids, int8 labels), copied to the device as they are, and widened there.
#include "common.cuh"
#include <cstring>
#include <thread>
#include <vector>
#include <algorithm>

namespace b200ctr {

template <typename TI, typename TO>
__global__ void widen_kernel(const TI* __restrict__ in, int64_t n, TO* __restrict__ out) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    TO v0 = (TO)in[i], v1 = (TO)in[i + 1], v2 = (TO)in[i + 2], v3 = (TO)in[i + 3];
    out[i] = v0; out[i + 1] = v1; out[i + 2] = v2; out[i + 3] = v3;
  } else {
    for (int64_t j = i; j < n; ++j) out[j] = (TO)in[j];
  }
}

}  // namespace b200ctr

using namespace b200ctr;

// dst row (dst_idx ? dst_idx[i] : i) <- src row src_idx[i], i in [0, n).  Plain memory, no CUDA call: usable on any host.
extern "C" int b200ctr_gather_rows(const void* src, int64_t row_bytes, const int64_t* src_idx, const int64_t* dst_idx, int64_t n,
                                   void* dst, int n_threads) {
  B200_REQUIRE(src && src_idx && dst && row_bytes > 0 && n >= 0, B200CTR_E_INVALID_ARG, "gather_rows: bad argument");
  if (n == 0) return B200CTR_OK;
  const char* s = (const char*)src;
  char* d = (char*)dst;
  auto work = [=](int64_t lo, int64_t hi) {
    for (int64_t i = lo; i < hi; ++i)
      std::memcpy(d + (dst_idx ? dst_idx[i] : i) * row_bytes, s + src_idx[i] * row_bytes, (size_t)row_bytes);
  };
  int nt = std::max(1, std::min<int>(n_threads, (int)std::min<int64_t>(64, (n * row_bytes) / (256 * 1024) + 1)));
  if (nt == 1) { work(0, n); return B200CTR_OK; }
  std::vector<std::thread> th;
  th.reserve(nt);
  const int64_t per = (n + nt - 1) / nt;
  for (int t = 0; t < nt; ++t) {
    const int64_t lo = t * per, hi = std::min(n, lo + per);
    if (lo < hi) th.emplace_back(work, lo, hi);
  }
  for (auto& t : th) t.join();
  return B200CTR_OK;
}

// kind 0: uint8 -> float32 (X_mask), 1: int8 -> float32 (y), 2: int32 -> int64 (ids, only for callers that insist on the
// collate contract's int64 on the device; the b200ctr kernels read int32 ids directly)
extern "C" int b200ctr_widen(int kind, const void* in, int64_t n, void* out, void* stream) {
  B200_REQUIRE(in && out && n >= 0, B200CTR_E_INVALID_ARG, "widen: bad argument");
  if (n == 0) return B200CTR_OK;
  const unsigned grid = (unsigned)ceil_div(ceil_div(n, (int64_t)4), (int64_t)256);
  cudaStream_t st = as_stream(stream);
  if (kind == 0) widen_kernel<uint8_t, float><<<grid, 256, 0, st>>>((const uint8_t*)in, n, (float*)out);
  else if (kind == 1) widen_kernel<int8_t, float><<<grid, 256, 0, st>>>((const int8_t*)in, n, (float*)out);
  else if (kind == 2) widen_kernel<int32_t, int64_t><<<grid, 256, 0, st>>>((const int32_t*)in, n, (int64_t*)out);
  else { set_error("widen: unknown kind %d", kind); return B200CTR_E_INVALID_ARG; }
  B200_LAUNCH_CHECK("widen_kernel");
  return B200CTR_OK;
}
