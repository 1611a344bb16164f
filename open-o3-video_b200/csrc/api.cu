// Version / error strings / device check of the C ABI (include/o3v.h).
#include "common.cuh"

namespace o3v {
int check_device() {
  int dev = 0;
  O3V_CUDA_TRY(cudaGetDevice(&dev));
  static int cached[64] = {0};  // 0 unknown, 1 ok, 2 bad
  if (dev >= 0 && dev < 64 && cached[dev]) return cached[dev] == 1 ? O3V_OK : O3V_ERR_UNSUPPORTED_ARCH;
  int major = 0, minor = 0;
  O3V_CUDA_TRY(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  O3V_CUDA_TRY(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  bool ok = (major == 10 && minor == 0);
  if (dev >= 0 && dev < 64) cached[dev] = ok ? 1 : 2;
  return ok ? O3V_OK : O3V_ERR_UNSUPPORTED_ARCH;
}
int num_sms() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) != cudaSuccess) return n;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}
}  // namespace o3v

extern "C" int o3v_version(void) { return O3V_VERSION; }

extern "C" int o3v_check_device(void) { return o3v::check_device(); }

extern "C" const char* o3v_strerror(int code) {
  switch (code) {
    case O3V_OK: return "ok";
    case O3V_ERR_INVALID_ARG: return "o3v: invalid argument (null pointer, non-positive size or bad flag)";
    case O3V_ERR_ALIGNMENT: return "o3v: pointer or leading dimension not 16-byte aligned";
    case O3V_ERR_UNSUPPORTED_ARCH: return "o3v: device is not sm_100 (B200); no fallback path exists";
    case O3V_ERR_WORKSPACE: return "o3v: workspace too small";
    case O3V_ERR_DRIVER: return "o3v: cuTensorMapEncodeTiled / driver entry point failed";
    case O3V_ERR_SHAPE: return "o3v: unsupported shape";
    case O3V_ERR_UNSUPPORTED_MODE: return "o3v: entry point not available in the tile mode selected with o3v_set_tunable";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "o3v: unknown error";
}
