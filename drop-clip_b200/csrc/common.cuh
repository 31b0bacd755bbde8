// Shared helpers of libdropclip (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/dropclip.h"

namespace dc {

// thread-local error message returned by dc_last_error()
char* error_buffer();
int fail(int status, const char* fmt, ...);

#define DC_CHECK_ARG(cond, ...)                                         \
  do {                                                                  \
    if (!(cond)) return ::dc::fail(DC_ERR_INVALID, __VA_ARGS__);        \
  } while (0)

#define DC_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess)                                                                 \
      return ::dc::fail(DC_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                                \
  } while (0)

#define DC_LAUNCH_CHECK() DC_CUDA(cudaGetLastError())

inline cudaStream_t as_stream(dc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // cached per process (device 0 of the current context)

// dc_set_stream_overlap(): the histogram ring kernel and the visibility filter share the SMs (two streams). Both then
// ask for the same shared-memory carve-out - an SM cannot change it while CTAs are resident. 58 % of 228 KB is rounded up
// by the driver to the 164 KB configuration (ncu: launch__shared_mem_config_size 167.9 KB): room for the ring CTA (50 KB)
// and two filter CTAs (34 KB each at <= 96 views; registers allow no third), and ~90 KB of L1 left for the filter's depth
// gathers (70 % L1 hit rate). Measured on the headline step (profiles/r02_seg_ring.md): this setting 2.75-2.78 ms, the
// 196 KB configuration with 3-slot rings 2.80-2.82 ms, 228 KB 3.3-3.6 ms (the filter alone runs 24 % slower without its
// L1); one stream 3.58 ms.
bool stream_overlap();
constexpr int kOverlapCarveoutPct = 58;

// cudaFuncSetAttribute only when the value differs from the one last set for this kernel on the current device
// (the attribute calls of a step add up on the host: the two-stream step is enqueued in ~1 ms)
struct FuncAttrCache {
  static constexpr int kMaxDevices = 32;
  int value[kMaxDevices];
  FuncAttrCache() {
    for (int i = 0; i < kMaxDevices; ++i) value[i] = -12345;
  }
  template <typename K>
  cudaError_t set(K kernel, cudaFuncAttribute attr, int v) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices && value[dev] == v) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, attr, v);
    if (e == cudaSuccess && dev >= 0 && dev < kMaxDevices) value[dev] = v;
    return e;
  }
};

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// torch-style min/max propagate NaN; fminf/fmaxf do not. Used where the reference calls
// tensor.min()/max().
__device__ __forceinline__ float nan_min(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b); }
__device__ __forceinline__ float nan_max(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b); }

// streaming (read-once / write-once) 128-bit accesses that do not pollute L1
__device__ __forceinline__ int4 ld_stream(const int4* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(int4* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

}  // namespace dc
