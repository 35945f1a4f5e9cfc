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

static int g_pdl_mode = -1;     // -1: not initialised
static bool g_pdl_env = false;  // SGB200_PDL was given: sg_set_pdl does not override it

int pdl_mode() {
  if (g_pdl_mode < 0) {
    const char* e = getenv("SGB200_PDL");
    g_pdl_env = e != nullptr;
    g_pdl_mode = e ? atoi(e) : 2;
    if (g_pdl_mode < 0 || g_pdl_mode > 2) g_pdl_mode = 2;
  }
  return g_pdl_mode;
}

int pdl_max_ctas() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("SGB200_PDL_MAX_CTAS");
    v = e ? atoi(e) : 4 * num_sms();  // measured: larger grids gain nothing from the early launch and the programmatic edge costs ~5 us
  }
  return v;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace sg

extern "C" {

int sg_abi_version(void) { return SG_ABI_VERSION; }

int sg_set_pdl(int mode) {
  SG_REQUIRE(mode >= 0 && mode <= 2, "sg_set_pdl: mode %d not in 0..2", mode);
  (void)sg::pdl_mode();  // reads the environment once
  if (!sg::g_pdl_env) sg::g_pdl_mode = mode;
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
  if (p.major != 10) {
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
