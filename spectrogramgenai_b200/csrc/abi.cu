// C-ABI plumbing of libsgb200.so: version, thread-local error string, device check.
#include <stdarg.h>
#include <string.h>

#include <stdlib.h>

#include "common.cuh"

namespace sg {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int launch_status(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return SG_ERR_LAUNCH;
  }
  return SG_OK;
}

// launch policy of the calling thread (like the current device, it is per-thread state of the caller, not of the library's
// kernels): 2 = programmatic dependent launch for grids of at most pdl_max_ctas() CTAs
static thread_local int g_pdl_mode = 2;

int pdl_mode() { return g_pdl_mode; }

// measured: grids larger than 4 x #SM CTAs gain nothing from the early launch and the programmatic edge costs ~5 us
int pdl_max_ctas() { return 4 * num_sms(); }

int num_sms() {
  static int n[SG_MAX_DEVICES] = {};
  const int dev = current_device();
  if (n[dev] == 0) {
    cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

}  // namespace sg

extern "C" {

int sg_abi_version(void) { return SG_ABI_VERSION; }

int sg_set_pdl(int mode) {
  SG_REQUIRE(mode >= 0 && mode <= 2, "sg_set_pdl: mode %d not in 0..2", mode);
  sg::g_pdl_mode = mode;
  return SG_OK;
}

const char* sg_last_error(void) { return sg::g_err; }

int sg_device_check(int device) {
  cudaDeviceProp p;
  cudaError_t e = cudaGetDeviceProperties(&p, device);
  if (e != cudaSuccess) {
    sg::set_error("cudaGetDeviceProperties(%d): %s", device, cudaGetErrorString(e));
    return SG_ERR_LAUNCH;
  }
  if (p.major != 10 || p.minor != 0) {
    sg::set_error("device %d is sm_%d%d; libsgb200 ships only an sm_100a image (B200) and has no fallback", device,
                  p.major, p.minor);
    return SG_ERR_ARCH;
  }
  return SG_OK;
}

int sg_set_device(int device) {
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) {
    sg::set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    return SG_ERR_LAUNCH;
  }
  return SG_OK;
}

}  // extern "C"
