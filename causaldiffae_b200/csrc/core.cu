// Library-level entry points: version, error string, device check.
#include <stdarg.h>

#include "sm100.cuh"

namespace cdae {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace cdae

extern "C" int cdae_version(void) { return 100; }
extern "C" const char* cdae_last_error(void) { return cdae::g_err; }

extern "C" int cdae_init(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { cdae::set_error("cudaGetDevice: %s", cudaGetErrorString(e)); return CDAE_ERR_CUDA; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { cdae::set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return CDAE_ERR_CUDA; }
  if (prop.major != 10) {
    cdae::set_error("libcdae is built for sm_100a only; device %d is sm_%d%d (no fallback path exists)", dev, prop.major,
                    prop.minor);
    return CDAE_ERR_ARCH;
  }
  if (!cdae::get_encode_tiled()) { cdae::set_error("driver entry point cuTensorMapEncodeTiled not found"); return CDAE_ERR_CUDA; }
  return CDAE_OK;
}
