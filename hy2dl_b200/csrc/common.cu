#include "common.cuh"
#include <string.h>

namespace hy2dl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  if (cached[dev] == 0) {
    int n = 0;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cached[dev] = n > 0 ? n : 148;
  }
  return cached[dev];
}

}  // namespace hy2dl

extern "C" {

int hy2dl_abi_version(void) { return HY2DL_ABI_VERSION; }

const char* hy2dl_last_error(void) { return hy2dl::g_err; }

int hy2dl_check_device(int dev) {
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
    hy2dl::set_error("hy2dl_check_device: cannot query device %d: %s", dev, cudaGetErrorString(cudaGetLastError()));
    return HY2DL_ERR_ARCH;
  }
  if (major != 10) {
    hy2dl::set_error("hy2dl_check_device: device %d is sm_%d%d; this library is built for sm_100a only", dev, major, minor);
    return HY2DL_ERR_ARCH;
  }
  return HY2DL_OK;
}

}  // extern "C"
