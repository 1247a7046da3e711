// Library-level entry points: error reporting, version, device check.
#include "common.cuh"

namespace es {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
}  // namespace es

extern "C" const char* es_last_error(void) { return es::g_last_error.c_str(); }
extern "C" int es_version(void) { return 100; }

extern "C" int es_device_ok(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    es::set_error("no CUDA device");
    return 0;
  }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    es::set_error("libexpertsim_b200 contains sm_100a code only; device compute capability major is " + std::to_string(major));
    return 0;
  }
  return 1;
}
