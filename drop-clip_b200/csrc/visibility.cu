// (1)+(2) Fused camera projection, depth-tolerance visibility and instance-mask lookup.
//
// Reference: MultiviewFeatureFusion.get_visibility_mask utils/feature_fusion.py:81-125,
// transform_pointcloud_to_camera_frame utils/transforms.py:52-61, `seg[ys, xs]`
// tools/preprocess_data.py:395-401. The reference loops over views on the host and launches
// ~10 small torch kernels + 3 H2D copies per view; here one launch covers a whole ragged batch
// of scenes. The kernel is point-major: a thread keeps kPointsPerThread points in registers and
// walks all views of its scene, so the point cloud is read once and every (view, point) result
// is written once, coalesced along the point axis.
//
// Arithmetic is fp64 and reproduces the reference's BLAS calls bit for bit: np.dot / `@` on the
// OpenBLAS that numpy ships evaluate every output element as a k-ascending chain of fused
// multiply-adds (see oracle/visibility_ref.c), spelled out below with __dmul_rn / __fma_rn so
// that the compiler can neither contract nor re-associate anything.
//
// Roofline: HBM-bound by design - 24 B/point + (4 B depth sample + 1 or 8 B mask) per
// (point, view); the fp64 work is ~35 DFMA-class instructions per (point, view).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPointsPerThread = 4;
constexpr int kPointsPerBlock = kThreads * kPointsPerThread;

struct VisParams {
  const double* points;
  const int64_t* point_off;
  const int64_t* view_off;
  const float* depths;
  const float* inv_poses;
  const double* intrinsics;
  const int64_t* mask_off;
  int height, width;
  double threshold;
  void* mask;
  uint8_t* any_visible;
  const void* seg;
  int seg_dtype;
  int32_t* point_object;
};

__device__ __forceinline__ int load_seg(const void* seg, int dtype, int64_t idx) {
  if (dtype == DC_U8) return (int)__ldg(reinterpret_cast<const uint8_t*>(seg) + idx);
  if (dtype == DC_I32) return __ldg(reinterpret_cast<const int32_t*>(seg) + idx);
  return (int)__ldg(reinterpret_cast<const long long*>(seg) + idx);
}

template <typename MaskT>
__global__ void __launch_bounds__(kThreads) project_visibility_kernel(VisParams p) {
  extern __shared__ double s_cam[];  // [n_views][12] inverse pose rows (fp64) then [9] intrinsics
  const int scene = blockIdx.y;
  const int64_t p0 = p.point_off[scene];
  const int64_t n_pts = p.point_off[scene + 1] - p0;
  const int64_t tile0 = (int64_t)blockIdx.x * kPointsPerBlock;
  if (tile0 >= n_pts) return;
  const int64_t v0 = p.view_off[scene];
  const int n_views = (int)(p.view_off[scene + 1] - v0);

  for (int i = threadIdx.x; i < n_views * 12; i += kThreads) {
    const int v = i / 12, e = i - v * 12;
    s_cam[i] = (double)__ldg(p.inv_poses + (v0 + v) * 16 + e);  // fp32 -> fp64 like np.dot's upcast
  }
  double* s_K = s_cam + n_views * 12;
  if (threadIdx.x < 9) s_K[threadIdx.x] = __ldg(p.intrinsics + (int64_t)scene * 9 + threadIdx.x);
  __syncthreads();

  double px[kPointsPerThread], py[kPointsPerThread], pz[kPointsPerThread];
  bool valid[kPointsPerThread];
  bool any[kPointsPerThread];
#pragma unroll
  for (int k = 0; k < kPointsPerThread; ++k) {
    const int64_t i = tile0 + k * kThreads + threadIdx.x;
    valid[k] = i < n_pts;
    any[k] = false;
    const int64_t j = valid[k] ? p0 + i : p0;
    px[k] = __ldg(p.points + 3 * j);
    py[k] = __ldg(p.points + 3 * j + 1);
    pz[k] = __ldg(p.points + 3 * j + 2);
  }

  const double K0 = s_K[0], K1 = s_K[1], K2 = s_K[2], K3 = s_K[3], K4 = s_K[4], K5 = s_K[5], K6 = s_K[6],
               K7 = s_K[7], K8 = s_K[8];
  const double w_lim = (double)p.width, h_lim = (double)p.height;
  const int64_t hw = (int64_t)p.height * p.width;
  MaskT* mask_scene = reinterpret_cast<MaskT*>(p.mask) + p.mask_off[scene];
  int32_t* pobj_scene = p.point_object ? p.point_object + p.mask_off[scene] : nullptr;

  for (int v = 0; v < n_views; ++v) {
    const double* m = s_cam + v * 12;
    const double m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5], m6 = m[6], m7 = m[7],
                 m8 = m[8], m9 = m[9], m10 = m[10], m11 = m[11];
    const float* depth = p.depths + (v0 + v) * hw;
    bool vis[kPointsPerThread];
    int64_t pix[kPointsPerThread];
    double qz[kPointsPerThread];
    bool inside[kPointsPerThread];
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k) {
      // dgemm(inv_pose, [p;1]): acc = a0*b0; acc = fma(a1,b1,acc); ...; last term is a3*1
      double cx = __dadd_rn(m3, __fma_rn(m2, pz[k], __fma_rn(m1, py[k], __dmul_rn(m0, px[k]))));
      double cy = __dadd_rn(m7, __fma_rn(m6, pz[k], __fma_rn(m5, py[k], __dmul_rn(m4, px[k]))));
      double cz = __dadd_rn(m11, __fma_rn(m10, pz[k], __fma_rn(m9, py[k], __dmul_rn(m8, px[k]))));
      cy = -cy;
      cz = -cz;
      // K @ c (structural zeros of K are kept so that non-finite inputs behave identically)
      const double qx = __fma_rn(K2, cz, __fma_rn(K1, cy, __dmul_rn(K0, cx)));
      const double qy = __fma_rn(K5, cz, __fma_rn(K4, cy, __dmul_rn(K3, cx)));
      qz[k] = __fma_rn(K8, cz, __fma_rn(K7, cy, __dmul_rn(K6, cx)));
      int pu = 0, pv = 0;
      bool in = true;
      if (qz[k] != 0.0) {
        const double uq = __ddiv_rn(qx, qz[k]);
        const double vq = __ddiv_rn(qy, qz[k]);
        // trunc-toward-zero into int64 then 0 <= . < limit  <=>  -1 < q < limit ; NaN/inf fail
        in = (uq > -1.0) && (uq < w_lim) && (vq > -1.0) && (vq < h_lim);
        if (in) {
          pu = (int)uq;
          pv = (int)vq;
        }
      }
      inside[k] = in && valid[k];
      pix[k] = (int64_t)pv * p.width + pu;
    }
    float sensor[kPointsPerThread];
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k) sensor[k] = inside[k] ? __ldg(depth + pix[k]) : 0.f;
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k) {
      vis[k] = inside[k] && (fabs((double)sensor[k] - qz[k]) <= p.threshold);
      any[k] |= vis[k];
    }
    const int64_t row = (int64_t)v * n_pts + tile0 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k) {
      if (valid[k]) {
        mask_scene[row + k * kThreads] = (MaskT)vis[k];
        if (pobj_scene)
          pobj_scene[row + k * kThreads] = vis[k] ? load_seg(p.seg, p.seg_dtype, (v0 + v) * hw + pix[k]) : -1;
      }
    }
  }
  if (p.any_visible) {
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k)
      if (valid[k]) p.any_visible[p0 + tile0 + k * kThreads + threadIdx.x] = any[k] ? 1 : 0;
  }
}

}  // namespace

extern "C" int dc_project_visibility(const double* points, const int64_t* point_off, const int64_t* view_off,
                                     const float* depths, const float* inv_poses, const double* intrinsics,
                                     const int64_t* mask_off, int n_scenes, int64_t max_points_per_scene,
                                     int max_views_per_scene, int height, int width, double threshold,
                                     void* mask, int mask_elem_size, uint8_t* any_visible, const void* seg,
                                     int seg_dtype, int32_t* point_object, dc_stream_t stream) {
  DC_CHECK_ARG(points && point_off && view_off && depths && inv_poses && intrinsics && mask_off && mask,
               "dc_project_visibility: null pointer argument");
  DC_CHECK_ARG(mask_elem_size == 1 || mask_elem_size == 8, "dc_project_visibility: mask_elem_size must be 1 or 8");
  DC_CHECK_ARG(height > 0 && width > 0 && (int64_t)height * width < (1ll << 31), "dc_project_visibility: bad image size");
  DC_CHECK_ARG(!point_object || seg, "dc_project_visibility: point_object needs seg");
  DC_CHECK_ARG(!seg || seg_dtype == DC_U8 || seg_dtype == DC_I32 || seg_dtype == DC_I64,
               "dc_project_visibility: seg dtype must be u8, i32 or i64");
  if (n_scenes <= 0 || max_points_per_scene <= 0) return DC_OK;
  DC_CHECK_ARG(n_scenes <= 65535, "dc_project_visibility: at most 65535 scenes per call");
  const size_t smem = ((size_t)max_views_per_scene * 12 + 9) * sizeof(double);
  DC_CHECK_ARG(smem <= 200 * 1024, "dc_project_visibility: too many views per scene (%d)", max_views_per_scene);
  VisParams p{points, point_off, view_off, depths, inv_poses, intrinsics, mask_off, height, width, threshold,
              mask, any_visible, seg, seg_dtype, point_object};
  dim3 grid((unsigned)dc::ceil_div<int64_t>(max_points_per_scene, kPointsPerBlock), (unsigned)n_scenes);
  if (mask_elem_size == 1) {
    if (smem > 48 * 1024)
      DC_CUDA(cudaFuncSetAttribute(project_visibility_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    project_visibility_kernel<uint8_t><<<grid, kThreads, smem, dc::as_stream(stream)>>>(p);
  } else {
    if (smem > 48 * 1024)
      DC_CUDA(cudaFuncSetAttribute(project_visibility_kernel<long long>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    project_visibility_kernel<long long><<<grid, kThreads, smem, dc::as_stream(stream)>>>(p);
  }
  DC_LAUNCH_CHECK();
  return DC_OK;
}
