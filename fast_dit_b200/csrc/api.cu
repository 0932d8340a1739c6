// api.cu — process-level state of libditb200: init, error reporting, driver entry points.
#include <cstdarg>
#include <cstdio>
#include <mutex>

#include "common.cuh"

namespace ditb200 {

static thread_local char g_err[512] = "";
static std::mutex g_init_mu;
static bool g_init = false;
static int g_device = -1;
static int g_sms = 0;
static EncodeTiledFn g_encode = nullptr;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return 0;
  set_error("%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
  return (int)e;
}

int num_sms() { return g_sms; }
bool is_initialised() { return g_init; }
EncodeTiledFn encode_tiled_fn() { return g_encode; }

}  // namespace ditb200

using namespace ditb200;

extern "C" int ditb200_abi_version(void) { return DITB200_ABI_VERSION; }

extern "C" const char* ditb200_last_error(void) { return g_err; }

extern "C" int ditb200_sm_count(void) { return g_sms; }

extern "C" int ditb200_init(int device) {
  std::lock_guard<std::mutex> lk(g_init_mu);
  if (g_init && g_device == device) return 0;
  // One device per process (one process per GPU is how the path is deployed: sample_ddp.py:54-58): the function
  // attributes, SM count and cluster occupancy the launchers cache are per device.
  if (g_init) {
    set_error("libditb200 is bound to device %d in this process; refusing device %d (one process per GPU)", g_device, device);
    return DITB200_EINVAL;
  }
  cudaDeviceProp prop;
  cudaError_t e = cudaGetDeviceProperties(&prop, device);  // does not change the caller's current device
  if (e != cudaSuccess) return check_cuda(e, "cudaGetDeviceProperties");
  if (prop.major != 10) {
    set_error("libditb200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major,
              prop.minor);
    return DITB200_EINVAL;
  }
  g_sms = prop.multiProcessorCount;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess) return check_cuda(e, "cudaGetDriverEntryPoint(cuTensorMapEncodeTiled)");
  if (fn == nullptr || qres != cudaDriverEntryPointSuccess) {
    set_error("driver does not export cuTensorMapEncodeTiled");
    return DITB200_EINVAL;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  g_device = device;
  g_init = true;
  return 0;
}
