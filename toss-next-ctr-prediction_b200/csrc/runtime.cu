// Error plumbing, device check, version.
#include "common.cuh"
#include <string.h>

namespace b200ctr {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
  set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
  return B200CTR_E_CUDA;
}
int num_sms() {
  static thread_local int cached = 0;
  if (cached) return cached;
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  cached = n > 0 ? n : 148;
  return cached;
}
}  // namespace b200ctr

extern "C" int b200ctr_version(void) { return B200CTR_ABI_VERSION; }
extern "C" const char* b200ctr_last_error(void) { return b200ctr::g_err; }
extern "C" int b200ctr_check_device(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return b200ctr::cuda_fail(e, "cudaGetDevice");
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    b200ctr::set_error("device %d is sm_%d%d; libb200ctr is built for sm_100a only (no fallback path)", dev, major, minor);
    return B200CTR_E_ARCH;
  }
  return B200CTR_OK;
}

// Row-sharded tables: kernels on the current device dereference peer-mapped shards (CUDA IPC) of `peer_device`.
extern "C" int b200ctr_enable_peer_access(int peer_device) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return b200ctr::cuda_fail(e, "cudaGetDevice");
  if (dev == peer_device) return B200CTR_OK;
  int can = 0;
  e = cudaDeviceCanAccessPeer(&can, dev, peer_device);
  if (e != cudaSuccess) return b200ctr::cuda_fail(e, "cudaDeviceCanAccessPeer");
  if (!can) {
    b200ctr::set_error("device %d cannot access device %d as a peer (row-sharded tables need NVLink/P2P)", dev, peer_device);
    return B200CTR_E_UNSUPPORTED_SHAPE;
  }
  e = cudaDeviceEnablePeerAccess(peer_device, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); return B200CTR_OK; }
  if (e != cudaSuccess) return b200ctr::cuda_fail(e, "cudaDeviceEnablePeerAccess");
  return B200CTR_OK;
}

// ---- CUDA IPC export / import of table shards (one process per GPU; remote shards are read and written in place) ----
// export: handle of the allocation that contains `ptr` (cudaMalloc granularity: the caching allocator of the host
// framework may pack several tensors into one allocation) + the byte offset of `ptr` inside it.
extern "C" int b200ctr_ipc_export(const void* ptr, void* handle_out64, int64_t* offset_out) {
  B200_REQUIRE(ptr && handle_out64 && offset_out, B200CTR_E_INVALID_ARG, "ipc_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  typedef int (*range_fn_t)(unsigned long long*, size_t*, unsigned long long);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || !fn) return b200ctr::cuda_fail(e, "cudaGetDriverEntryPoint(cuMemGetAddressRange)");
  unsigned long long base = 0;
  size_t size = 0;
  int rc = ((range_fn_t)fn)(&base, &size, (unsigned long long)(uintptr_t)ptr);
  if (rc != 0) {
    b200ctr::set_error("ipc_export: cuMemGetAddressRange failed with driver error %d", rc);
    return B200CTR_E_CUDA;
  }
  e = cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle_out64, (void*)(uintptr_t)base);
  if (e != cudaSuccess) return b200ctr::cuda_fail(e, "cudaIpcGetMemHandle");
  *offset_out = (int64_t)((unsigned long long)(uintptr_t)ptr - base);
  return B200CTR_OK;
}
// import on the CURRENT device: maps the peer allocation and (lazily) enables peer access to its device
extern "C" int b200ctr_ipc_import(const void* handle64, void** base_out) {
  B200_REQUIRE(handle64 && base_out, B200CTR_E_INVALID_ARG, "ipc_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(base_out, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return b200ctr::cuda_fail(e, "cudaIpcOpenMemHandle");
  return B200CTR_OK;
}
