#include <stdlib.h>
// Error plumbing, device checks, version.
#include <stdarg.h>

#include "../../include/b200asr.h"
#include "common.cuh"

namespace b200asr {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static unsigned long long g_launches = 0;   // kernels launched by this library (process-wide, approximate under threads)
void note_launch(int n) { g_launches += (unsigned long long)n; }

int check_launch(const char* what) {
  g_launches += 1;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: CUDA error %d (%s)", what, (int)e, cudaGetErrorString(e));
    return B200ASR_CUDA_ERROR;
  }
  return B200ASR_OK;
}

int ensure_dynamic_smem(const void* kernel, int bytes, bool (&done)[kMaxDevices], const char* what) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { set_error("%s: cannot query the current device", what); return B200ASR_CUDA_ERROR; }
  if (done[dev]) return B200ASR_OK;
  cudaError_t r = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (r != cudaSuccess) { set_error("%s: cannot reserve %d bytes of shared memory: %s", what, bytes, cudaGetErrorString(r)); return B200ASR_CUDA_ERROR; }
  done[dev] = true;
  return B200ASR_OK;
}

bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("B200ASR_PDL"); return !(e && e[0] == '0'); }();
  return on;
}

int device_sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    cached = p.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

int ensure_sm100() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    set_error("no CUDA device: %s", cudaGetErrorString(e));
    return B200ASR_UNSUPPORTED_ARCH;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    set_error("libb200asr is built for sm_100a only; device %d is sm_%d%d", dev, major, minor);
    return B200ASR_UNSUPPORTED_ARCH;
  }
  return B200ASR_OK;
}

}  // namespace b200asr

extern "C" {
int b200asr_version(void) { return 100; }
const char* b200asr_last_error(void) { return b200asr::g_err; }
int b200asr_device_check(void) { return b200asr::ensure_sm100(); }
unsigned long long b200asr_launch_count(void) { return b200asr::g_launches; }
}
