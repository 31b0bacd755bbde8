// Error channel, version and device queries of the C ABI.
#include <stdarg.h>

#include "common.cuh"

namespace dc {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int fail(int status, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return status;
}

int sm_count() {
  static int cached = 0;
  if (cached == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      cached = n;
    else
      return 148;
  }
  return cached;
}

namespace {
int g_stream_overlap = 1;
}
bool stream_overlap() { return g_stream_overlap != 0; }

}  // namespace dc

extern "C" {

int dc_set_stream_overlap(int on) {
  const int prev = dc::g_stream_overlap;
  dc::g_stream_overlap = on ? 1 : 0;
  return prev;
}

int dc_abi_version(void) { return DC_ABI_VERSION; }

const char* dc_last_error(void) { return dc::error_buffer(); }

int dc_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes) {
  int dev = 0;
  DC_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  DC_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (l2_bytes) *l2_bytes = (size_t)p.l2CacheSize;
  return DC_OK;
}

}  // extern "C"
