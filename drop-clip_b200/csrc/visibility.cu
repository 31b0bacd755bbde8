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
// Measured (profiles/r01_*): the kernel is NOT HBM-bound - the depth gathers hit L2 (~85 %) and
// the fp64 pipe (2 cycles per warp instruction) plus issue slots set the pace. Three exact
// rewrites therefore cut fp64 work without changing a single result bit (each is argued where it
// is used): (a) the y/z sign flip is folded into pre-negated matrix rows, (b) structural zeros of
// a pinhole K are skipped for finite operands, (c) the two IEEE divisions share one refined
// reciprocal, whose quotients are only trusted when they are provably far from every integer;
// otherwise the thread falls back to the literal evaluation.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kPointsPerThread = 2;
constexpr int kPointsPerBlock = kThreads * kPointsPerThread;
constexpr double kBig = 1e100;  // operands below this magnitude cannot overflow anywhere in the chain

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

__device__ __forceinline__ double rcp_refined(double x) {
  // rcp.approx.ftz.f64 carries ~20 mantissa bits; two Newton steps take the relative error to a
  // few 2^-53. Only used for |x| in (1e-200, 1e200), where neither x nor 1/x is subnormal.
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double e = __fma_rn(-x, r, 1.0);
  r = __fma_rn(r, e, r);
  e = __fma_rn(-x, r, 1.0);
  r = __fma_rn(r, e, r);
  return r;
}

// Literal evaluation of one (point, view): the operation sequence of the reference
// (oracle/visibility_ref.c), every structural zero of K multiplied out so that non-finite operands
// propagate exactly as in numpy. `m` holds the inverse pose with rows 1 and 2 pre-negated, see (a).
__device__ __noinline__ bool literal_pixel(const double* __restrict__ m, const double* __restrict__ K, double x, double y,
                                           double z, int width, int height, int& pix, double& qz_out) {
  const double cx = __dadd_rn(m[3], __fma_rn(m[2], z, __fma_rn(m[1], y, __dmul_rn(m[0], x))));
  const double cy = __dadd_rn(m[7], __fma_rn(m[6], z, __fma_rn(m[5], y, __dmul_rn(m[4], x))));
  const double cz = __dadd_rn(m[11], __fma_rn(m[10], z, __fma_rn(m[9], y, __dmul_rn(m[8], x))));
  const double qx = __fma_rn(K[2], cz, __fma_rn(K[1], cy, __dmul_rn(K[0], cx)));
  const double qy = __fma_rn(K[5], cz, __fma_rn(K[4], cy, __dmul_rn(K[3], cx)));
  const double qz = __fma_rn(K[8], cz, __fma_rn(K[7], cy, __dmul_rn(K[6], cx)));
  qz_out = qz;
  int pu = 0, pv = 0;
  bool in = true;
  if (qz != 0.0) {
    const double uq = __ddiv_rn(qx, qz);
    const double vq = __ddiv_rn(qy, qz);
    // numpy truncates toward zero into int64, then tests 0 <= . < limit  <=>  -1 < q < limit; NaN/inf fail
    in = (uq > -1.0) && (uq < (double)width) && (vq > -1.0) && (vq < (double)height);
    if (in) {
      pu = (int)uq;
      pv = (int)vq;
    }
  }
  pix = pv * width + pu;
  return in;
}

// Classifies an approximate quotient q (relative error <= 2^-50) against [0, limit) under
// truncation toward zero. Returns 0 = surely outside, 1 = surely inside (pixel in `out`),
// 2 = too close to an integer (or too large) to decide without the exact quotient.
__device__ __forceinline__ int classify(double q, int limit, int& out) {
  const int i = __double2int_rz(q);  // saturating
  const double d = q - (double)i;    // in (-1, 1) unless saturated
  // |q| <= 1e6 bounds the absolute error by 1e6 * 2^-50 < 1e-9; |d| in (1e-8, 1 - 1e-8) then means no
  // integer lies between q and the correctly rounded exact quotient, so both truncate to i.
  if (!(fabs(fabs(d) - 0.5) < 0.5 - 1e-8) || !(fabs(q) <= 1.0e6)) return 2;
  out = i;
  return ((unsigned)i < (unsigned)limit) ? 1 : 0;
}

template <typename MaskT>
__global__ void __launch_bounds__(kThreads, 4) project_visibility_kernel(VisParams p) {
  extern __shared__ double s_cam[];  // [n_views][12] inverse pose rows (fp64), rows 1,2 negated; then [9] K
  __shared__ int s_ok;
  const int scene = blockIdx.y;
  const int64_t p0 = p.point_off[scene];
  const int64_t n_pts = p.point_off[scene + 1] - p0;
  const int64_t tile0 = (int64_t)blockIdx.x * kPointsPerBlock;
  if (tile0 >= n_pts) return;
  const int64_t v0 = p.view_off[scene];
  const int n_views = (int)(p.view_off[scene + 1] - v0);

  if (threadIdx.x == 0) s_ok = 1;
  __syncthreads();
  bool mine_ok = true;
  for (int i = threadIdx.x; i < n_views * 12; i += kThreads) {
    const int v = i / 12, e = i - v * 12;
    const double val = (double)__ldg(p.inv_poses + (v0 + v) * 16 + e);  // fp32 -> fp64 like np.dot's upcast
    mine_ok &= fabs(val) < kBig;
    // (a) round-to-nearest is sign-symmetric, so evaluating the chain with rows 1 and 2 negated gives
    //     exactly the negated camera-frame y and z (only the sign of an exact zero can differ, which
    //     no later step observes).
    s_cam[i] = (e >= 4) ? -val : val;
  }
  double* s_K = s_cam + n_views * 12;
  if (threadIdx.x < 9) {
    const double kv = __ldg(p.intrinsics + (int64_t)scene * 9 + threadIdx.x);
    s_K[threadIdx.x] = kv;
    mine_ok &= fabs(kv) < kBig;
  }
  if (!mine_ok) s_ok = 0;
  __syncthreads();
  const double K0 = s_K[0], K2 = s_K[2], K4 = s_K[4], K5 = s_K[5];
  // (b) pinhole structure K = [[fx,0,cx],[0,fy,cy],[0,0,1]], every matrix entry finite and < 1e100
  const bool pinhole = s_ok && s_K[1] == 0.0 && s_K[3] == 0.0 && s_K[6] == 0.0 && s_K[7] == 0.0 && s_K[8] == 1.0;

  double px[kPointsPerThread], py[kPointsPerThread], pz[kPointsPerThread];
  bool valid[kPointsPerThread], fast[kPointsPerThread], any[kPointsPerThread];
#pragma unroll
  for (int k = 0; k < kPointsPerThread; ++k) {
    const int64_t i = tile0 + k * kThreads + threadIdx.x;
    valid[k] = i < n_pts;
    any[k] = false;
    const int64_t j = valid[k] ? p0 + i : p0;
    px[k] = __ldg(p.points + 3 * j);
    py[k] = __ldg(p.points + 3 * j + 1);
    pz[k] = __ldg(p.points + 3 * j + 2);
    fast[k] = pinhole && fabs(px[k]) < kBig && fabs(py[k]) < kBig && fabs(pz[k]) < kBig;  // NaN -> false
  }

  const int64_t hw = (int64_t)p.height * p.width;
  MaskT* mask_scene = reinterpret_cast<MaskT*>(p.mask) + p.mask_off[scene];
  int32_t* pobj_scene = p.point_object ? p.point_object + p.mask_off[scene] : nullptr;

  for (int v = 0; v < n_views; ++v) {
    const double* m = s_cam + v * 12;
    const float* depth = p.depths + (v0 + v) * hw;
    bool inside[kPointsPerThread];
    int pix[kPointsPerThread];
    double qz[kPointsPerThread];
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k) {
      bool literal = !fast[k];
      inside[k] = false;
      pix[k] = 0;
      qz[k] = 0.0;
      if (fast[k]) {
        // dgemm(inv_pose, [p;1]): acc = a0*b0; acc = fma(a1,b1,acc); acc = fma(a2,b2,acc); acc += a3*1
        const double cx = __dadd_rn(m[3], __fma_rn(m[2], pz[k], __fma_rn(m[1], py[k], __dmul_rn(m[0], px[k]))));
        const double cy = __dadd_rn(m[7], __fma_rn(m[6], pz[k], __fma_rn(m[5], py[k], __dmul_rn(m[4], px[k]))));
        const double cz = __dadd_rn(m[11], __fma_rn(m[10], pz[k], __fma_rn(m[9], py[k], __dmul_rn(m[8], px[k]))));
        // (b) all operands are finite here, so fma(0, cy, acc) == acc, fma(fy, cy, 0*cx) == fy*cy and
        //     fma(1, cz, 0) == cz up to the sign of an exact zero.
        const double qx = __fma_rn(K2, cz, __dmul_rn(K0, cx));
        const double qy = __fma_rn(K5, cz, __dmul_rn(K4, cy));
        qz[k] = cz;
        const double aqz = fabs(cz);
        if (cz == 0.0) {
          inside[k] = true;  // the reference skips the division: pixel (0,0)
        } else if (aqz > 1e-200 && aqz < 1e200) {
          // (c) one refined reciprocal serves both quotients; see classify() for when it is trusted
          const double r = rcp_refined(cz);
          int pu = 0, pv = 0;
          const int su = classify(__dmul_rn(qx, r), p.width, pu);
          const int sv = classify(__dmul_rn(qy, r), p.height, pv);
          if (su == 1 && sv == 1) {
            inside[k] = true;
            pix[k] = pv * p.width + pu;
          } else if (su != 0 && sv != 0) {
            literal = true;  // undecided in some axis and not surely outside in the other
          }
        } else {
          literal = true;
        }
      }
      if (literal) inside[k] = literal_pixel(m, s_K, px[k], py[k], pz[k], p.width, p.height, pix[k], qz[k]);
      inside[k] = inside[k] && valid[k];
    }
    float sensor[kPointsPerThread];
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k) sensor[k] = inside[k] ? __ldg(depth + pix[k]) : 0.f;
    const int64_t row = (int64_t)v * n_pts + tile0 + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k) {
      // torch: abs(fp32 depth promoted to fp64 - z') <= threshold, compared in fp64
      const bool vis = inside[k] && (fabs((double)sensor[k] - qz[k]) <= p.threshold);
      any[k] |= vis;
      if (valid[k]) {
        mask_scene[row + k * kThreads] = (MaskT)vis;
        if (pobj_scene)
          pobj_scene[row + k * kThreads] = vis ? load_seg(p.seg, p.seg_dtype, (v0 + v) * hw + pix[k]) : -1;
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
