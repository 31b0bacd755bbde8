// Grounding / segmentation metric counts: the reductions that directly follow the
// grounding kernel in the reference's evaluation loops.
//
//   trainMetricPC            utils/misc.py:21-50   per instance: binarise the score at `threshold`
//                            (in place when the caller's tensor is passed un-sigmoided), intersection
//                            and union counts against the ground-truth mask
//   intersectionAndUnionGPU  utils/misc.py:186-199 K-class intersection / output / target histograms
//                            (the reference bounces to the CPU for torch.histc; here one pass on the
//                            device), `output[target == ignore_index] = ignore_index` in place
// Integer counts are exact; everything floating point that follows (iou = inter / (union + 1e-6),
// Pr@k, means) stays in the Python mirror `dropclip_b200/metrics.py`, written like the reference.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ bool truthy(const void* p, int dtype, int64_t i) {
  switch (dtype) {
    case DC_U8: return reinterpret_cast<const uint8_t*>(p)[i] != 0;
    case DC_I32: return reinterpret_cast<const int32_t*>(p)[i] != 0;
    case DC_I64: return reinterpret_cast<const long long*>(p)[i] != 0;
    case DC_F32: return reinterpret_cast<const float*>(p)[i] != 0.f;  // NaN is truthy, like tensor.bool()
    default: return false;
  }
}

// grid (chunks, n_instances): ragged instances, counts accumulated with 64-bit atomics
__global__ void __launch_bounds__(kThreads) binary_iou_kernel(float* __restrict__ pred, const void* __restrict__ gt, int gt_dtype,
                                                              const int64_t* __restrict__ off, float threshold, int sigmoid,
                                                              int write_back, unsigned long long* __restrict__ inter,
                                                              unsigned long long* __restrict__ uni) {
  const int inst = blockIdx.y;
  const int64_t b = off[inst], e = off[inst + 1];
  unsigned li = 0, lu = 0;
  for (int64_t i = b + (int64_t)blockIdx.x * kThreads + threadIdx.x; i < e; i += (int64_t)gridDim.x * kThreads) {
    float s = pred[i];
    if (sigmoid) s = 1.f / (1.f + expf(-s));  // torch.sigmoid makes a copy: nothing is written back
    // pred[pred < thr] = 0; pred[pred >= thr] = 1  (NaN satisfies neither and stays NaN -> truthy)
    const float bin = (s < threshold) ? 0.f : ((s >= threshold) ? 1.f : s);
    if (write_back && !sigmoid) pred[i] = bin;
    const bool p = bin != 0.f;
    const bool g = truthy(gt, gt_dtype, i);
    li += (p && g);
    lu += (p || g);
  }
  li = __reduce_add_sync(0xffffffffu, li);
  lu = __reduce_add_sync(0xffffffffu, lu);
  __shared__ unsigned s_i[kThreads / 32], s_u[kThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    s_i[threadIdx.x >> 5] = li;
    s_u[threadIdx.x >> 5] = lu;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned a = 0, c = 0;
    for (int w = 0; w < kThreads / 32; ++w) {
      a += s_i[w];
      c += s_u[w];
    }
    if (a) atomicAdd(inter + inst, (unsigned long long)a);
    if (c) atomicAdd(uni + inst, (unsigned long long)c);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) class_hist_kernel(T* __restrict__ output, const T* __restrict__ target, int64_t n, int k,
                                                              long long ignore_index, unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned s_hist[];  // [3][k]: intersection, output, target
  for (int i = threadIdx.x; i < 3 * k; i += kThreads) s_hist[i] = 0;
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const long long t = (long long)target[i];
    long long o = (long long)output[i];
    if (t == ignore_index) {
      o = ignore_index;
      output[i] = (T)ignore_index;  // in place, like the reference
    }
    // torch.histc(bins=K, min=0, max=K-1) maps the integer class c in [0, K-1] to bin c and ignores the rest
    if (o >= 0 && o < k) {
      atomicAdd(s_hist + k + (int)o, 1u);
      if (o == t) atomicAdd(s_hist + (int)o, 1u);
    }
    if (t >= 0 && t < k) atomicAdd(s_hist + 2 * k + (int)t, 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * k; i += kThreads)
    if (s_hist[i]) atomicAdd(hist + i, (unsigned long long)s_hist[i]);
}

__global__ void hist_finish_kernel(const unsigned long long* __restrict__ hist, int k, float* __restrict__ area_inter,
                                   float* __restrict__ area_union, float* __restrict__ area_target) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= k) return;
  const float i = (float)hist[c], o = (float)hist[k + c], t = (float)hist[2 * k + c];
  area_inter[c] = i;
  area_target[c] = t;
  area_union[c] = o + t - i;  // fp32 arithmetic on the histc outputs, as in the reference
}

}  // namespace

extern "C" {

int dc_binary_iou_counts(float* pred, const void* gt, int gt_dtype, const int64_t* inst_off, int n_instances,
                         int64_t max_points_per_instance, float threshold, int apply_sigmoid, int binarize_in_place,
                         int64_t* inter, int64_t* uni, dc_stream_t stream) {
  DC_CHECK_ARG(pred && gt && inst_off && inter && uni, "dc_binary_iou_counts: null pointer argument");
  DC_CHECK_ARG(gt_dtype == DC_U8 || gt_dtype == DC_I32 || gt_dtype == DC_I64 || gt_dtype == DC_F32,
               "dc_binary_iou_counts: ground truth must be u8 (bool), i32, i64 or f32");
  if (n_instances <= 0) return DC_OK;
  DC_CHECK_ARG(n_instances <= 65535, "dc_binary_iou_counts: at most 65535 instances per call");
  cudaStream_t st = dc::as_stream(stream);
  DC_CUDA(cudaMemsetAsync(inter, 0, sizeof(int64_t) * (size_t)n_instances, st));
  DC_CUDA(cudaMemsetAsync(uni, 0, sizeof(int64_t) * (size_t)n_instances, st));
  if (max_points_per_instance <= 0) return DC_OK;
  int64_t chunks = dc::ceil_div<int64_t>(max_points_per_instance, kThreads * 8);
  const int64_t cap = dc::ceil_div<int64_t>((int64_t)dc::sm_count() * 8, n_instances);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  dim3 grid((unsigned)chunks, (unsigned)n_instances);
  binary_iou_kernel<<<grid, kThreads, 0, st>>>(pred, gt, gt_dtype, inst_off, threshold, apply_sigmoid, binarize_in_place,
                                              reinterpret_cast<unsigned long long*>(inter),
                                              reinterpret_cast<unsigned long long*>(uni));
  DC_LAUNCH_CHECK();
  return DC_OK;
}

size_t dc_class_iou_workspace(int n_classes) { return sizeof(unsigned long long) * 3 * (size_t)(n_classes > 0 ? n_classes : 1); }

int dc_class_iou_hist(void* output, const void* target, int dtype, int64_t n, int n_classes, int64_t ignore_index,
                      float* area_intersection, float* area_union, float* area_target, void* workspace, size_t workspace_bytes,
                      dc_stream_t stream) {
  DC_CHECK_ARG(output && target && area_intersection && area_union && area_target && workspace,
               "dc_class_iou_hist: null pointer argument");
  DC_CHECK_ARG(dtype == DC_I32 || dtype == DC_I64 || dtype == DC_U8, "dc_class_iou_hist: labels must be u8, i32 or i64");
  DC_CHECK_ARG(n_classes >= 1 && n_classes <= 4096, "dc_class_iou_hist: 1..4096 classes");
  if (workspace_bytes < dc_class_iou_workspace(n_classes))
    return dc::fail(DC_ERR_WORKSPACE, "dc_class_iou_hist: workspace %zu < %zu", workspace_bytes, dc_class_iou_workspace(n_classes));
  cudaStream_t st = dc::as_stream(stream);
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(workspace);
  DC_CUDA(cudaMemsetAsync(hist, 0, dc_class_iou_workspace(n_classes), st));
  if (n > 0) {
    int64_t blocks = dc::ceil_div<int64_t>(n, kThreads * 8);
    if (blocks > (int64_t)dc::sm_count() * 8) blocks = (int64_t)dc::sm_count() * 8;
    const size_t smem = sizeof(unsigned) * 3 * (size_t)n_classes;
    if (dtype == DC_I64)
      class_hist_kernel<long long><<<(unsigned)blocks, kThreads, smem, st>>>((long long*)output, (const long long*)target, n,
                                                                            n_classes, (long long)ignore_index, hist);
    else if (dtype == DC_I32)
      class_hist_kernel<int><<<(unsigned)blocks, kThreads, smem, st>>>((int*)output, (const int*)target, n, n_classes,
                                                                      (long long)ignore_index, hist);
    else
      class_hist_kernel<uint8_t><<<(unsigned)blocks, kThreads, smem, st>>>((uint8_t*)output, (const uint8_t*)target, n,
                                                                          n_classes, (long long)ignore_index, hist);
  }
  hist_finish_kernel<<<dc::ceil_div(n_classes, 128), 128, 0, st>>>(hist, n_classes, area_intersection, area_union, area_target);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // extern "C"
