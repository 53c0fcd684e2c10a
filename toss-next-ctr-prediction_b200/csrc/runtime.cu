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
