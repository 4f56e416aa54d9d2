// Error plumbing and device checks for the hcir_b200 C ABI (include/hcir_b200.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "hcir_common.cuh"

namespace hcir {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", static_cast<int>(e), cudaGetErrorString(e), what);
  return HCIR_ECUDA;
}

bool pdl_enabled() {
  static const bool on = [] {
    const char* e = getenv("HCIR_PDL");
    return !(e != nullptr && e[0] == '0');
  }();
  return on;
}

int check_device() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
  int major = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e != cudaSuccess) return cuda_fail(e, "cudaDeviceGetAttribute");
  if (major != 10) {
    set_error("hcir_b200 kernels are built for sm_100a only; device %d has compute capability %d.x "
              "(no CPU fallback, no other backend)", dev, major);
    return HCIR_EARCH;
  }
  return HCIR_OK;
}

}  // namespace hcir

extern "C" {

int hcir_abi_version(void) { return HCIR_ABI_VERSION; }
const char* hcir_last_error(void) { return hcir::g_err; }
int hcir_device_supported(void) { return hcir::check_device() == HCIR_OK ? 1 : 0; }
int hcir_padded_dim(int d) { return d <= 0 ? 0 : hcir::round_up_int(d, 64); }

}  // extern "C"
