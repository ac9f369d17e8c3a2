// Library-wide pieces of the C ABI (include/puzzlenet_b200.h): version, error text, device probe.
#include "pz_common.cuh"

namespace pz {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

}  // namespace pz

extern "C" int pz_abi_version(void) { return PZ_ABI_VERSION; }

extern "C" const char* pz_last_error(void) { return pz::error_buffer(); }

extern "C" int pz_device_arch(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return pz::fail(-1000, "cudaGetDevice failed (no CUDA device?)");
  int major = 0, minor = 0;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess)
    return pz::fail(-1000, "cudaDeviceGetAttribute failed");
  return major * 10 + minor;
}
