// Library-level entry points: error reporting, version, device check.
#include <mutex>

#include "common.cuh"

namespace es {
static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }

// Pipeline-timeout flag of the tensor-core kernels (tc_ptx.cuh: mbar_wait writes its tag, fences system-wide, traps).
// ONE flag PER DEVICE, in host-mapped pinned memory: the tag stays readable by the host after the trap has poisoned the
// context, and a process that drives several GPUs never hands one device's kernels a pointer into another device's memory.
namespace {
constexpr int kMaxDevices = 64;
std::mutex g_flag_mutex;
int* g_flag_host[kMaxDevices] = {};
int* g_flag_dev[kMaxDevices] = {};
}  // namespace

int* pipeline_err_flag() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return nullptr;
  std::lock_guard<std::mutex> lock(g_flag_mutex);
  if (!g_flag_dev[dev]) {
    int* h = nullptr;
    if (cudaHostAlloc(&h, sizeof(int), cudaHostAllocMapped | cudaHostAllocPortable) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    *h = 0;
    int* d = nullptr;
    if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) {
      cudaGetLastError();
      cudaFreeHost(h);
      return nullptr;
    }
    g_flag_host[dev] = h;
    g_flag_dev[dev] = d;
  }
  return g_flag_dev[dev];
}

// tags: 1 gather waits for a free stage, 2 MMA waits for a full stage, 3 epilogue waits for the accumulator,
//       4 TMA producer waits for a free stage, 5 MMA waits for a drained accumulator, 6/7 resident-weight barriers
std::string pipeline_err_report() {
  std::string r;
  std::lock_guard<std::mutex> lock(g_flag_mutex);
  for (int d = 0; d < kMaxDevices; ++d)
    if (g_flag_host[d] && *(volatile int*)g_flag_host[d] != 0)
      r += " [device " + std::to_string(d) + ": tensor-core pipeline timed out at barrier tag " +
           std::to_string(*(volatile int*)g_flag_host[d]) + "]";
  return r;
}
}  // namespace es

extern "C" const char* es_last_error(void) {
  static thread_local std::string buf;
  buf = es::g_last_error + es::pipeline_err_report();
  return buf.c_str();
}
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
