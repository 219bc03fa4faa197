// api.cu -- version, error reporting and device check of libdinomc.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "dmc_common.cuh"

namespace dmc {
namespace {
thread_local char g_err[512] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_status(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
  return static_cast<int>(e);
}

namespace {
int& pdl_flag() {
  static int on = [] {
    const char* e = getenv("DMC_PDL");
    return (e == nullptr || atoi(e) != 0) ? 1 : 0;
  }();
  return on;
}
}  // namespace

bool pdl_enabled() { return pdl_flag() != 0; }

void prefer_max_smem_carveout(const void* func) {
  // tiny open-addressed set of kernels already configured (benign races: at worst the attribute is set twice)
  static const void* seen[256] = {nullptr};
  static const bool enabled = [] { const char* e = getenv("DMC_MAX_SMEM_CARVEOUT"); return e == nullptr || atoi(e) != 0; }();
  if (!enabled) return;
  size_t h = (reinterpret_cast<uintptr_t>(func) >> 4) & 255;
  for (int probe = 0; probe < 256; ++probe, h = (h + 1) & 255) {
    if (seen[h] == func) return;
    if (seen[h] == nullptr) {
      seen[h] = func;
      cudaFuncSetAttribute(func, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
      cudaGetLastError();                   // a hint: never fail a launch because of it
      return;
    }
  }
}

namespace {
thread_local int g_streaming_ctas = 0;
}
int streaming_ctas() { return g_streaming_ctas; }
}  // namespace dmc

extern "C" int dmc_set_streaming_ctas(int n) {
  const int prev = dmc::g_streaming_ctas;
  dmc::g_streaming_ctas = n > 0 ? n : 0;
  return prev;
}

extern "C" int dmc_version(void) { return DMC_VERSION; }

extern "C" const char* dmc_last_error_string(void) { return dmc::g_err; }

extern "C" int dmc_set_pdl(int enabled) {
  int& f = dmc::pdl_flag();
  const int prev = f;
  f = enabled ? 1 : 0;
  return prev;
}

extern "C" int dmc_device_check(int device) {
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);
  if (e != cudaSuccess) return dmc::cuda_status(e, "cudaGetDeviceProperties");
  if (prop.major != 10) {
    dmc::set_error("libdinomc is built for sm_100a (B200); device %d is sm_%d%d", device, prop.major, prop.minor);
    return -1;
  }
  return 0;
}
