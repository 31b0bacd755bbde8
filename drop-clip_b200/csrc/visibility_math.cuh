// Exact fp64 projection arithmetic shared by the visibility kernels (see visibility.cu for the
// derivation of the shortcuts (a)-(c) and why none of them changes a result bit).
#pragma once
#include "common.cuh"

namespace dc {
namespace vis {

constexpr double kBig = 1e100;  // operands below this magnitude cannot overflow anywhere in the chain

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
static __device__ __noinline__ bool literal_pixel(const double* __restrict__ m, const double* __restrict__ K, double x, double y,
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


// One (point, view) evaluation. `m` = inverse pose rows with rows 1,2 pre-negated, K = full 3x3
// (s_K) plus the four pinhole entries in registers. Returns "inside the image" and sets pix / qz.
__device__ __forceinline__ bool project_point(const double* __restrict__ m, const double* __restrict__ K, double K0, double K2,
                                              double K4, double K5, bool fast, double x, double y, double z, int width,
                                              int height, int& pix, double& qz) {
  bool literal = !fast;
  bool inside = false;
  pix = 0;
  qz = 0.0;
  if (fast) {
    // dgemm(inv_pose, [p;1]): acc = a0*b0; acc = fma(a1,b1,acc); acc = fma(a2,b2,acc); acc += a3*1
    const double cx = __dadd_rn(m[3], __fma_rn(m[2], z, __fma_rn(m[1], y, __dmul_rn(m[0], x))));
    const double cy = __dadd_rn(m[7], __fma_rn(m[6], z, __fma_rn(m[5], y, __dmul_rn(m[4], x))));
    const double cz = __dadd_rn(m[11], __fma_rn(m[10], z, __fma_rn(m[9], y, __dmul_rn(m[8], x))));
    // (b) all operands are finite here, so fma(0, cy, acc) == acc, fma(fy, cy, 0*cx) == fy*cy and
    //     fma(1, cz, 0) == cz up to the sign of an exact zero.
    const double qx = __fma_rn(K2, cz, __dmul_rn(K0, cx));
    const double qy = __fma_rn(K5, cz, __dmul_rn(K4, cy));
    qz = cz;
    const double aqz = fabs(cz);
    if (cz == 0.0) {
      inside = true;  // the reference skips the division: pixel (0,0)
    } else if (aqz > 1e-200 && aqz < 1e200) {
      // (c) one refined reciprocal serves both quotients; see classify() for when it is trusted
      const double r = rcp_refined(cz);
      int pu = 0, pv = 0;
      const int su = classify(__dmul_rn(qx, r), width, pu);
      const int sv = classify(__dmul_rn(qy, r), height, pv);
      if (su == 1 && sv == 1) {
        inside = true;
        pix = pv * width + pu;
      } else if (su != 0 && sv != 0) {
        literal = true;  // undecided in some axis and not surely outside in the other
      }
    } else {
      literal = true;
    }
  }
  if (literal) inside = literal_pixel(m, K, x, y, z, width, height, pix, qz);
  return inside;
}

// Loads the per-scene camera constants into shared memory: [n_views][12] inverse pose rows
// (fp32 -> fp64, rows 1 and 2 negated) followed by the 9 intrinsics. Returns true when every
// entry is finite and below kBig and K has the pinhole structure (shortcut (b) allowed).
__device__ __forceinline__ bool load_cameras(double* s_cam, int* s_ok, const double* __restrict__ inv_poses, int64_t v0, int n_views,
                                             const double* __restrict__ intrinsics, int scene) {
  if (threadIdx.x == 0) *s_ok = 1;
  __syncthreads();
  bool mine_ok = true;
  for (int i = threadIdx.x; i < n_views * 12; i += blockDim.x) {
    const int v = i / 12, e = i - v * 12;
    const double val = __ldg(inv_poses + (v0 + v) * 16 + e);  // fp64 on the device: an fp32 inverse upcasts exactly like np.dot does
    mine_ok &= fabs(val) < kBig;
    // (a) round-to-nearest is sign-symmetric, so evaluating the chain with rows 1 and 2 negated gives
    //     exactly the negated camera-frame y and z (only the sign of an exact zero can differ, which
    //     no later step observes).
    s_cam[i] = (e >= 4) ? -val : val;
  }
  double* s_K = s_cam + n_views * 12;
  if (threadIdx.x < 9) {
    const double kv = __ldg(intrinsics + (int64_t)scene * 9 + threadIdx.x);
    s_K[threadIdx.x] = kv;
    mine_ok &= fabs(kv) < kBig;
  }
  if (!mine_ok) *s_ok = 0;
  __syncthreads();
  return *s_ok && s_K[1] == 0.0 && s_K[3] == 0.0 && s_K[6] == 0.0 && s_K[7] == 0.0 && s_K[8] == 1.0;
}

}  // namespace vis
}  // namespace dc
