// Training-sample assembly: the deterministic part of MVDistilDataset.__getitem__
// data/dataset_blender.py:330-362,400-414 for a ragged batch of samples, on the device.
//
// The reference materialises feat = per_obj[label] for the WHOLE cloud on the CPU (N x 768 fp32 =
// 307 MB per sample at N = 100 k), filters it by the visibility of k random views, then keeps
// MAX_POINTS = 10 000 random points. Here the selection comes first and only the selected rows are
// ever gathered:
//   1. keep[j]   = OR over the sample's chosen views of vis_mask[v, j]      (dataset_blender.py:338-351)
//   2. kept list = indices of kept points (dc_compact_scan + kept_list_kernel)
//   3. rows      = kept[indices[i]]; xyz rows gathered in fp64, centred with the sequential
//                  column mean numpy's xyz.mean(0) computes (:364), then cast to fp32;
//                  rgb fp32; labels; feat rows = per_obj[label]                 (:128-130, :353-362)
// Random choices (views, point indices) are the caller's inputs, so the step is reproducible.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

// grid (chunks, n_samples)
__global__ void __launch_bounds__(kThreads) views_any_kernel(const uint8_t* __restrict__ vis_mask, const int64_t* __restrict__ mask_off,
                                                             const int64_t* __restrict__ point_off,
                                                             const int32_t* __restrict__ view_list,
                                                             const int64_t* __restrict__ view_list_off, uint8_t* __restrict__ keep) {
  const int s = blockIdx.y;
  const int64_t p0 = point_off[s], n = point_off[s + 1] - p0;
  const int64_t l0 = view_list_off[s], l1 = view_list_off[s + 1];
  const uint8_t* m = vis_mask + mask_off[s];
  for (int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x; j < n; j += (int64_t)gridDim.x * kThreads) {
    uint8_t any = (l1 == l0) ? 1 : 0;  // no view list: the full cloud is used (use_full_pc)
    for (int64_t k = l0; k < l1; ++k) any |= m[(int64_t)view_list[k] * n + j];
    keep[p0 + j] = any ? 1 : 0;
  }
}

__global__ void __launch_bounds__(kThreads) kept_list_kernel(const uint8_t* __restrict__ keep, const int64_t* __restrict__ new_index,
                                                             const int64_t* __restrict__ point_off, int n_samples, int64_t total,
                                                             int64_t* __restrict__ kept_idx) {
  for (int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x; j < total; j += (int64_t)gridDim.x * kThreads)
    if (keep[j]) kept_idx[new_index[j]] = j;  // global point index; ascending inside a sample
}

// rows[r] = global point index of output row r, or -1 (flagged) when an index is out of range
__global__ void __launch_bounds__(kThreads) select_rows_kernel(const int64_t* __restrict__ kept_idx, const int64_t* __restrict__ kept_off,
                                                               const int64_t* __restrict__ indices,
                                                               const int64_t* __restrict__ out_off, int64_t* __restrict__ rows,
                                                               int* __restrict__ error) {
  const int s = blockIdx.y;
  const int64_t o0 = out_off[s], m = out_off[s + 1] - o0;
  const int64_t k0 = kept_off[s], kn = kept_off[s + 1] - k0;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < m; i += (int64_t)gridDim.x * kThreads) {
    const int64_t want = indices[o0 + i];
    if (want < 0 || want >= kn) {
      rows[o0 + i] = -1;
      atomicExch(error, 1);
    } else {
      rows[o0 + i] = kept_idx[k0 + want];
    }
  }
}

// numpy's xyz.mean(0) on a C-contiguous (M,3) fp64 array adds the rows one after another per column
__global__ void column_mean_kernel(const double* __restrict__ xyz, const int64_t* __restrict__ rows, const int64_t* __restrict__ out_off,
                                   int n_samples, double* __restrict__ mean) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_samples * 3) return;
  const int s = t / 3, a = t - 3 * s;
  const int64_t o0 = out_off[s], m = out_off[s + 1] - o0;
  double acc = 0.0;
  for (int64_t i = 0; i < m; ++i) {
    const int64_t j = rows[o0 + i];
    acc += (j >= 0) ? xyz[3 * j + a] : 0.0;
  }
  mean[t] = acc / (double)m;  // m == 0 -> NaN, like numpy (with a warning there)
}

// one warp per output row
__global__ void __launch_bounds__(kThreads) gather_rows_kernel(const double* __restrict__ xyz, const double* __restrict__ rgb,
                                                               const int64_t* __restrict__ label, const float* __restrict__ per_obj,
                                                               const int64_t* __restrict__ obj_off, const int64_t* __restrict__ rows,
                                                               const int64_t* __restrict__ out_off, const double* __restrict__ mean,
                                                               int n_samples, int dim, int label_as_u8, float* __restrict__ out_xyz,
                                                               float* __restrict__ out_rgb, int32_t* __restrict__ out_label,
                                                               float* __restrict__ out_feat, int* __restrict__ error) {
  const int lane = threadIdx.x & 31;
  const int64_t total = out_off[n_samples];
  for (int64_t r = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); r < total; r += (int64_t)gridDim.x * (kThreads / 32)) {
    int s = 0;  // samples per call are few (a batch): linear search
    while (s + 1 < n_samples && r >= out_off[s + 1]) ++s;
    const int64_t j = rows[r];
    if (j < 0) continue;
    if (lane < 3) {
      out_xyz[3 * r + lane] = (float)(xyz[3 * j + lane] - mean[3 * s + lane]);  // xyz -= mean in fp64, then .float()
      if (rgb) out_rgb[3 * r + lane] = (float)rgb[3 * j + lane];
    }
    const long long l = label[j];
    if (lane == 0) out_label[r] = label_as_u8 ? (int32_t)(uint8_t)l : (int32_t)l;  // label.astype(np.uint8) after the filter (:349)
    const int64_t q0 = obj_off[s], nq = obj_off[s + 1] - q0;
    if (l < 0 || l >= nq) {  // numpy would raise IndexError (negative labels would wrap; not meaningful here)
      if (lane == 0) atomicExch(error, 2);
      continue;
    }
    const float* src = per_obj + (q0 + l) * dim;
    float* dst = out_feat + r * dim;
    if ((dim & 3) == 0 && ((((uintptr_t)src) | ((uintptr_t)dst)) & 15) == 0) {
      for (int c = lane; c < dim / 4; c += 32) reinterpret_cast<float4*>(dst)[c] = __ldg(reinterpret_cast<const float4*>(src) + c);
    } else {
      for (int c = lane; c < dim; c += 32) dst[c] = __ldg(src + c);
    }
  }
}

unsigned chunks_for(int64_t max_n, int n_samples) {
  int64_t c = dc::ceil_div<int64_t>(max_n, kThreads * 4);
  const int64_t cap = dc::ceil_div<int64_t>((int64_t)dc::sm_count() * 8, n_samples);
  if (c > cap) c = cap;
  return (unsigned)(c < 1 ? 1 : c);
}

}  // namespace

extern "C" {

int dc_sample_keep_flags(const uint8_t* vis_mask, const int64_t* mask_off, const int64_t* point_off, const int32_t* view_list,
                         const int64_t* view_list_off, int n_samples, int64_t max_points_per_sample, uint8_t* keep,
                         dc_stream_t stream) {
  DC_CHECK_ARG(point_off && view_list_off && keep, "dc_sample_keep_flags: null pointer argument");
  if (n_samples <= 0 || max_points_per_sample <= 0) return DC_OK;
  DC_CHECK_ARG(n_samples <= 65535, "dc_sample_keep_flags: at most 65535 samples per call");
  dim3 grid(chunks_for(max_points_per_sample, n_samples), (unsigned)n_samples);
  views_any_kernel<<<grid, kThreads, 0, dc::as_stream(stream)>>>(vis_mask, mask_off, point_off, view_list, view_list_off, keep);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_sample_gather(const double* xyz, const double* rgb, const int64_t* label, const float* per_obj, const int64_t* obj_off,
                     const uint8_t* keep, const int64_t* new_index, const int64_t* kept_off, const int64_t* point_off,
                     const int64_t* indices, const int64_t* out_off, int n_samples, int64_t total_points, int64_t total_rows,
                     int64_t max_rows_per_sample, int dim, int label_as_u8, float* out_xyz, float* out_rgb, int32_t* out_label,
                     float* out_feat, int64_t* rows, int64_t* kept_idx, double* mean, int* error, dc_stream_t stream) {
  DC_CHECK_ARG(xyz && label && per_obj && obj_off && keep && new_index && kept_off && point_off && indices && out_off && out_xyz &&
                   out_label && out_feat && rows && kept_idx && mean && error,
               "dc_sample_gather: null pointer argument");
  DC_CHECK_ARG(!rgb || out_rgb, "dc_sample_gather: out_rgb missing");
  DC_CHECK_ARG(dim > 0, "dc_sample_gather: bad feature width");
  if (n_samples <= 0) return DC_OK;
  DC_CHECK_ARG(n_samples <= 65535, "dc_sample_gather: at most 65535 samples per call");
  cudaStream_t st = dc::as_stream(stream);
  DC_CUDA(cudaMemsetAsync(error, 0, sizeof(int), st));
  if (total_points > 0) {
    int64_t b = dc::ceil_div<int64_t>(total_points, kThreads * 4);
    if (b > (int64_t)dc::sm_count() * 8) b = (int64_t)dc::sm_count() * 8;
    kept_list_kernel<<<(unsigned)b, kThreads, 0, st>>>(keep, new_index, point_off, n_samples, total_points, kept_idx);
  }
  if (total_rows > 0) {
    dim3 grid(chunks_for(max_rows_per_sample, n_samples), (unsigned)n_samples);
    select_rows_kernel<<<grid, kThreads, 0, st>>>(kept_idx, kept_off, indices, out_off, rows, error);
  }
  column_mean_kernel<<<dc::ceil_div(n_samples * 3, 64), 64, 0, st>>>(xyz, rows, out_off, n_samples, mean);
  if (total_rows > 0) {
    int64_t b = dc::ceil_div<int64_t>(total_rows, kThreads / 32);
    if (b > (int64_t)dc::sm_count() * 16) b = (int64_t)dc::sm_count() * 16;
    gather_rows_kernel<<<(unsigned)b, kThreads, 0, st>>>(xyz, rgb, label, per_obj, obj_off, rows, out_off, mean, n_samples, dim,
                                                        label_as_u8, out_xyz, out_rgb, out_label, out_feat, error);
  }
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // extern "C"
