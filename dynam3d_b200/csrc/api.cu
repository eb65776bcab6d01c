// Library-level entry points: error string, version, device check.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[1024] = "";

void d3d_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int d3d_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

unsigned long long g_d3d_launches = 0;
extern "C" long long d3d_launch_count(void) { return (long long)g_d3d_launches; }
extern "C" const char* d3d_last_error(void) { return g_err; }
extern "C" int d3d_version(void) { return 100; }
extern "C" int d3d_sm_count(void) { return d3d_num_sms(); }

extern "C" int d3d_check_device(int dev) {
  int major = 0, minor = 0;
  D3D_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  D3D_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  if (major != 10) {
    d3d_set_error("device %d is sm_%d%d; libdynam3d_b200 is built for sm_100a only and has no fallback path", dev, major, minor);
    return D3D_EARCH;
  }
  return 0;
}
