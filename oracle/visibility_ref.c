/* ORACLE (test infrastructure, never the product path).
 *
 * Plain-C, machine-independent restatement of the integer/fp64 part of DROP-CLIP's fusion:
 *   oracle_visibility   utils/feature_fusion.py:88-123  (+ utils/transforms.py:52-61)
 *   oracle_seg_counts   utils/feature_fusion.py:307,320 (np.unique(seg), (seg == obj).sum())
 *   oracle_quantize     data/dataset_blender.py:406-414 (floor(xyz / size) -> int32, ME semantics)
 *
 * Why C and not numpy: the reference computes the two small matrix products with BLAS dgemm.
 * On the OpenBLAS that numpy ships (0.3.30, Haswell/SkylakeX kernels) every output element is a
 * k-ascending chain of fused multiply-adds,  acc = a0*b0; acc = fma(a1,b1,acc); ...  (measured:
 * 0 mismatches in 4e5 elements against np.dot, tests/test_oracle_golden.py re-checks it). numpy
 * has no fma primitive, so the chain is spelled out here with fma() and compiled with
 * -ffp-contract=off; the CUDA kernel spells out the same chain with __fma_rn.
 *
 * Parity status: PINNED - tests/golden/vis_*.npz hold masks produced by the unmodified
 * reference; this file reproduces them bit for bit.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* numpy's float64 -> int64 cast on x86-64 is cvttsd2si: truncate toward zero, and NaN / inf /
 * out-of-range all give INT64_MIN ("integer indefinite"). */
static int64_t trunc_like_numpy(double x) {
  if (!(x > -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
  return (int64_t)x;
}

/* One view. inv_pose: 16 doubles (row-major inverse camera->world matrix, already inverted by np.linalg.inv in
 * the pose's own dtype and widened - np.dot promotes an fp32 inverse to fp64, utils/transforms.py:54-58),
 * K: 9 doubles row-major. Outputs may be NULL. */
void oracle_visibility(const double* pts, int64_t n, const float* depth, int64_t height, int64_t width,
                       const double* inv_pose, const double* K, double threshold,
                       int64_t* mask, int64_t* pix, double* zdepth) {
  double m[12];
  for (int i = 0; i < 12; ++i) m[i] = inv_pose[i];
  for (int64_t i = 0; i < n; ++i) {
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    double c[3];
    for (int r = 0; r < 3; ++r) {
      double acc = m[4 * r] * x;
      acc = fma(m[4 * r + 1], y, acc);
      acc = fma(m[4 * r + 2], z, acc);
      acc = fma(m[4 * r + 3], 1.0, acc);
      c[r] = acc;
    }
    c[1] = -c[1];
    c[2] = -c[2];
    double p[3];
    for (int r = 0; r < 3; ++r) {
      double acc = K[3 * r] * c[0];
      acc = fma(K[3 * r + 1], c[1], acc);
      acc = fma(K[3 * r + 2], c[2], acc);
      p[r] = acc;
    }
    int64_t u = 0, v = 0;
    if (p[2] != 0) {
      u = trunc_like_numpy(p[0] / p[2]);
      v = trunc_like_numpy(p[1] / p[2]);
    }
    int64_t vis = 0;
    if (u >= 0 && v >= 0 && u < width && v < height) {
      const double sensor = (double)depth[v * width + u];
      vis = fabs(sensor - p[2]) <= threshold;
    }
    if (mask) mask[i] = vis;
    if (pix) { pix[2 * i] = u; pix[2 * i + 1] = v; }
    if (zdepth) zdepth[i] = p[2];
  }
}

/* counts[b] = number of pixels with id b for b in [0,nbins); returns the number of pixels
 * whose id falls outside [0,nbins). */
int64_t oracle_seg_counts(const int64_t* seg, int64_t npix, int64_t nbins, int64_t* counts) {
  int64_t outside = 0;
  memset(counts, 0, sizeof(int64_t) * (size_t)nbins);
  for (int64_t i = 0; i < npix; ++i) {
    if (seg[i] >= 0 && seg[i] < nbins) counts[seg[i]]++; else outside++;
  }
  return outside;
}

/* floor(xyz / size) in fp32 (true division), converted to int32 like torch's .int(). */
void oracle_quantize(const float* xyz, int64_t n3, float size, int32_t* out) {
  for (int64_t i = 0; i < n3; ++i) out[i] = (int32_t)floorf(xyz[i] / size);
}

/* transform_pointcloud_to_world_frame / _to_camera_frame utils/transforms.py:43-61: rows 0..2 of np.dot(M, [p;1])
 * with M (4x4, 16 doubles row-major: an fp32 matrix is promoted by np.dot) - the same k-ascending FMA chain. */
void oracle_transform(const double* pts, int64_t n, const double* M, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    const double x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    for (int r = 0; r < 3; ++r) {
      double acc = M[4 * r] * x;
      acc = fma(M[4 * r + 1], y, acc);
      acc = fma(M[4 * r + 2], z, acc);
      out[3 * i + r] = fma(M[4 * r + 3], 1.0, acc);
    }
  }
}
