// Pixel-level fusion (the `use_obj_prior=0` ablation path).
//
// Reference: MultiviewFeatureFusion.aggregate_features utils/feature_fusion.py:138-250 and the final
// division of fuse_points :266-268. Per view the reference bicubically upsamples the (ph,pw,C)
// patch map to (H,W,C) (943.7 MB at 480x640x768 fp32), L2-normalises every pixel, multiplies the
// whole map with the query matrix, builds a per-pixel similarity metric, and finally gathers the
// visible projected pixels. Only those gathered pixels matter, so this kernel evaluates the 16
// bicubic taps directly at each visible (point, view), in registers, and never materialises the
// map. Rows are accumulated over the views in view order (same order as `sum_features[mask] += feat3d`)
// and written once. The query similarity of an interpolated feature comes from a per-view table of
// (patch cell . query) dots interpolated with the same weights (patch_query_dots_kernel).
//
// Bicubic weights follow ATen's upsample_bicubic2d (align_corners=False, A=-0.75, source index
// scale*(dst+0.5)-0.5 un-clamped, taps clamped to the border).
#include <algorithm>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxPerLane = 32;  // channels per lane: dim <= 1024

struct PixParams {
  const double* points;
  const int64_t* point_off;
  const int64_t* view_off;
  const double* inv_poses;
  const double* intrinsics;
  const int64_t* mask_off;
  const uint8_t* visible;
  const void* seg;
  int seg_dtype;
  const float* patch;
  int ph, pw, dim;
  const float* queries;
  const int64_t* query_off;
  int sim_kernel, norm_feat;
  int height, width;
  const int64_t* perm;
  float* out_sum;
  float* out_weight;
  double* dots;  // [total_views][ph * pw][q_stride] fp64: patch row . query, see patch_query_dots_kernel
  int q_stride;
  int normalize;  // divide the sums by sum_v weight (similarity) or by the number of views that see the point
};

__device__ __forceinline__ float cubic1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

__device__ __forceinline__ void cubic_taps(int dst, float scale, int in_size, int (&idx)[4], float (&w)[4]) {
  const float A = -0.75f;
  const float src = scale * ((float)dst + 0.5f) - 0.5f;
  const float fl = floorf(src);
  const float t = src - fl;
  const int i0 = (int)fl;
  w[0] = cubic2(t + 1.f, A);
  w[1] = cubic1(t, A);
  w[2] = cubic1(1.f - t, A);
  w[3] = cubic2(2.f - t, A);
#pragma unroll
  for (int k = 0; k < 4; ++k) idx[k] = min(max(i0 - 1 + k, 0), in_size - 1);
}

// The same taps with fp64 weights (ATen's formulas evaluated in double precision, as F.interpolate does for fp64
// input): used for the similarity chain only, see "similarity weights in fp64" below.
__device__ __forceinline__ double cubic1d(double x, double A) { return ((A + 2.0) * x - (A + 3.0)) * x * x + 1.0; }
__device__ __forceinline__ double cubic2d(double x, double A) { return ((A * x - 5.0 * A) * x + 8.0 * A) * x - 4.0 * A; }
__device__ __forceinline__ int cubic_weights64(int dst, int in_size, int out_size, double (&w)[4]) {
  const double A = -0.75;
  const double src = ((double)in_size / (double)out_size) * ((double)dst + 0.5) - 0.5;
  const double fl = floor(src);
  const double t = src - fl;
  w[0] = cubic2d(t + 1.0, A);
  w[1] = cubic1d(t, A);
  w[2] = cubic1d(1.0 - t, A);
  w[3] = cubic2d(2.0 - t, A);
  return (int)fl;  // first tap is at index i0 - 1 (clamped by the caller)
}
// Taps of one axis with EXACT weights: the fp64 coefficients (for the similarity chain) and their fp32 roundings
// (for the fp32 feature fold). ATen's own fp32 evaluation of `scale * (dst + 0.5) - 0.5` carries an absolute error
// of ~1e-6 in the source coordinate (and depends on how its compiler contracted the expression into FMAs), which
// moves interpolated features by ~2e-7 - already 1e-3 of the smallest magnitudes the parity metric resolves
// (profiles/r02_reference_fp32_noise.md). Evaluating the coordinate in fp64 puts this kernel next to the exact
// value of the reference's formula rather than next to one particular build's rounding of it.
__device__ __forceinline__ void cubic_taps_exact(int dst, int in_size, int out_size, int (&idx)[4], float (&w)[4], double (&w64)[4]) {
  const int i0 = cubic_weights64(dst, in_size, out_size, w64);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    idx[k] = min(max(i0 - 1 + k, 0), in_size - 1);
    w[k] = (float)w64[k];
  }
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ int seg_at(const void* seg, int dtype, int64_t idx) {
  if (dtype == DC_U8) return (int)__ldg(reinterpret_cast<const uint8_t*>(seg) + idx);
  if (dtype == DC_I32) return __ldg(reinterpret_cast<const int32_t*>(seg) + idx);
  const long long v = __ldg(reinterpret_cast<const long long*>(seg) + idx);
  return (v < 0 || v > 0x7fffffff) ? -1 : (int)v;
}

// The similarity of an interpolated feature with a query is linear in the taps:
//   f . q = sum_t w_t (T_t . q),   f = sum_t w_t T_t  (16 bicubic taps T_t of the patch map)
// so the (patch cell, query) dot products are computed ONCE per view (ph * pw * Q dots, 16 M FMA at 24x32x21x768)
// instead of Q x C FMA and 64 KB of query reads per visible (point, view). The normalisation of `norm_feat`
// divides the similarity by |f| afterwards ((f / |f|) . q = (f . q) / |f|). Warp per patch cell, lanes over channels.
__global__ void __launch_bounds__(kThreads) patch_query_dots_kernel(PixParams p) {
  const int scene = blockIdx.y;
  const int64_t v0 = p.view_off[scene];
  const int n_views = (int)(p.view_off[scene + 1] - v0);
  const int n_q = (int)(p.query_off[scene + 1] - p.query_off[scene]);
  const float* q = p.queries + p.query_off[scene] * p.dim;
  const int lane = threadIdx.x & 31;
  const int64_t n_cells = (int64_t)p.ph * p.pw, n_rows = n_cells * n_views;
  for (int64_t r = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); r < n_rows; r += (int64_t)gridDim.x * kWarps) {
    const float* row = p.patch + (v0 * n_cells + r) * p.dim;
    double* out = p.dots + (v0 * n_cells + r) * p.q_stride;
    float t[kMaxPerLane];
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) t[k] = (k * 32 + lane < p.dim) ? __ldg(row + k * 32 + lane) : 0.f;
    for (int o = 0; o < n_q; o += 4) {  // four queries per round: their loads overlap (the kernel is latency-bound)
      double d[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (o + j < n_q) {
          const float* qo = q + (int64_t)(o + j) * p.dim;
#pragma unroll
          for (int k = 0; k < kMaxPerLane; ++k)
            if (k * 32 + lane < p.dim) d[j] = fma((double)t[k], (double)__ldg(qo + k * 32 + lane), d[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) d[j] = warp_sum_d(d[j]);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (lane == 0 && o + j < n_q) out[o + j] = d[j];
    }
  }
}

// Similarity weights in fp64. The weight of a visible (point, view) is clip(pos - max|mean(neg), 1e-6) of the
// similarities of its (normalised) interpolated feature with the scene's queries (calculate_sim,
// feature_fusion.py:65-73,182-196). On surfaces whose best two queries tie, pos - neg is a difference of two numbers
// ~1 that lands near the 1e-6 clip, and the fused feature of a point whose views all do that is a mean with weights
// of relative accuracy eps_fp32 / 1e-6: the reference's own fp32 result is then several percent away from the exact
// value of its formulas (profiles/r02_reference_fp32_noise.md: 0.7 % of the rows of a 480x640, V=8 scene are off by
// more than 1e-3). An fp32 evaluation here would add its own, different, noise on top, so the whole chain is exact
// instead: fp64 (patch . query) table, fp64 bicubic weights, pos - red(neg) in fp64, ONE division by |f| at the end
// ((pos - neg) / |f| = pos / |f| - neg / |f|, and max commutes with the positive scale), one rounding to fp32.
// What remains between this and the reference is the reference's rounding noise alone.
// interp_dot: lane l interpolates query o's dots (issued before the tap loads); finish_weight reduces over queries.
// Scenes with more than 32 queries take further rounds of interp_dot inside finish_weight.
__device__ __forceinline__ double interp_dot(const PixParams& p, const double* __restrict__ dots_view, const int (&iy)[4],
                                             const int (&ix)[4], const double (&wy)[4], const double (&wx)[4], int o, int n_q) {
  double mine = 0.0;
  if (o < n_q) {
#pragma unroll
    for (int ty = 0; ty < 4; ++ty) {
      const double* row = dots_view + ((int64_t)iy[ty] * p.pw) * p.q_stride + o;
      double r = __ldg(row + ix[0] * p.q_stride) * wx[0];
      r = fma(__ldg(row + ix[1] * p.q_stride), wx[1], r);
      r = fma(__ldg(row + ix[2] * p.q_stride), wx[2], r);
      r = fma(__ldg(row + ix[3] * p.q_stride), wx[3], r);
      mine = fma(r, wy[ty], mine);
    }
  }
  return mine;
}

template <typename More>
__device__ __forceinline__ float finish_weight(const PixParams& p, double first, float nrm, int id, int n_q, int lane, More&& more) {
  double pos = 0.0, red = (p.sim_kernel == DC_SIM_MAX) ? -INFINITY : 0.0;
  bool nan_seen = false;
  for (int o0 = 0; o0 < n_q; o0 += 32) {
    const int o = o0 + lane;
    const double mine = o0 == 0 ? first : more(o);
    if (id >= o0 && id < o0 + 32) pos = __shfl_sync(0xffffffffu, mine, id - o0);
    const bool is_neg = o < n_q && o != id;
    nan_seen |= __any_sync(0xffffffffu, is_neg && (mine != mine));
    if (p.sim_kernel == DC_SIM_MAX) red = fmax(red, warp_max_d(is_neg ? mine : -INFINITY));
    else red += warp_sum_d(is_neg ? mine : 0.0);
  }
  if (p.sim_kernel == DC_SIM_MEAN) red = red / (double)(n_q - 1);
  if (nan_seen) red = __longlong_as_double(0x7ff8000000000000ll);
  double w = pos - red;
  if (p.norm_feat) w = w / (double)nrm;  // nrm = 0 (an all-zero feature): 0 / 0 = NaN like the reference's f / |f|
  float weight = (float)w;
  if (weight == weight) weight = fmaxf(weight, 1e-6f);
  return weight;
}

// seg_at in two halves: the load (requested early) and the range check / narrowing (done where the id is needed), so
// that the scoreboard wait for this cold, scattered load sits behind the tap arithmetic.
__device__ __forceinline__ long long seg_raw(const void* seg, int dtype, int64_t idx) {
  long long v;
  if (dtype == DC_U8) v = (long long)__ldg(reinterpret_cast<const uint8_t*>(seg) + idx);
  else if (dtype == DC_I32) v = (long long)__ldg(reinterpret_cast<const int32_t*>(seg) + idx);
  else v = __ldg(reinterpret_cast<const long long*>(seg) + idx);
  return v;
}
__device__ __forceinline__ int seg_id(long long v) { return (v < 0 || v > 0x7fffffff) ? -1 : (int)v; }

__global__ void __launch_bounds__(kThreads) pixel_fuse_kernel(PixParams p) {
  extern __shared__ double s_cam[];  // [n_views][12] + [9]
  const int scene = blockIdx.y;
  const int64_t p0 = p.point_off[scene];
  const int64_t n_pts = p.point_off[scene + 1] - p0;
  const int64_t v0 = p.view_off[scene];
  const int n_views = (int)(p.view_off[scene + 1] - v0);
  for (int i = threadIdx.x; i < n_views * 12; i += kThreads) {
    const int v = i / 12, e = i - v * 12;
    s_cam[i] = __ldg(p.inv_poses + (v0 + v) * 16 + e);
  }
  double* s_K = s_cam + n_views * 12;
  if (threadIdx.x < 9) s_K[threadIdx.x] = __ldg(p.intrinsics + (int64_t)scene * 9 + threadIdx.x);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int per_lane = p.dim / 32;  // host guarantees dim % 32 == 0, dim <= 1024
  const int n_q = p.sim_kernel != DC_SIM_NONE ? (int)(p.query_off[scene + 1] - p.query_off[scene]) : 0;
  const int64_t hw = (int64_t)p.height * p.width;
  const uint8_t* vis_scene = p.visible + p.mask_off[scene];
  float* w_scene = p.out_weight ? p.out_weight + p.mask_off[scene] : nullptr;

  for (int64_t s_pos = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); s_pos < n_pts; s_pos += (int64_t)gridDim.x * kWarps) {
    const int64_t i = p.perm ? __ldg(p.perm + p0 + s_pos) : s_pos;  // spatially sorted processing order, original indexing
    const double x = __ldg(p.points + 3 * (p0 + i)), y = __ldg(p.points + 3 * (p0 + i) + 1),
                 z = __ldg(p.points + 3 * (p0 + i) + 2);
    float acc[kMaxPerLane];
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) acc[k] = 0.f;
    for (int v = 0; v < n_views; ++v) {
      const bool vis = vis_scene[(int64_t)v * n_pts + i] != 0;  // warp-uniform
      if (!vis) {
        if (w_scene && lane == 0) w_scene[(int64_t)v * n_pts + i] = 0.f;
        continue;
      }
      // same projection arithmetic as visibility.cu (the point is visible, so it is inside)
      const double* m = s_cam + v * 12;
      double cx = __dadd_rn(m[3], __fma_rn(m[2], z, __fma_rn(m[1], y, __dmul_rn(m[0], x))));
      double cy = __dadd_rn(m[7], __fma_rn(m[6], z, __fma_rn(m[5], y, __dmul_rn(m[4], x))));
      double cz = __dadd_rn(m[11], __fma_rn(m[10], z, __fma_rn(m[9], y, __dmul_rn(m[8], x))));
      cy = -cy;
      cz = -cz;
      const double qx = __fma_rn(s_K[2], cz, __fma_rn(s_K[1], cy, __dmul_rn(s_K[0], cx)));
      const double qy = __fma_rn(s_K[5], cz, __fma_rn(s_K[4], cy, __dmul_rn(s_K[3], cx)));
      const double qz = __fma_rn(s_K[8], cz, __fma_rn(s_K[7], cy, __dmul_rn(s_K[6], cx)));
      int pu = 0, pv = 0;
      if (qz != 0.0) {
        pu = (int)__ddiv_rn(qx, qz);
        pv = (int)__ddiv_rn(qy, qz);
      }
      int iy[4], ix[4];
      float wy[4], wx[4];
      double wy64[4], wx64[4];
      cubic_taps_exact(pv, p.ph, p.height, iy, wy, wy64);
      cubic_taps_exact(pu, p.pw, p.width, ix, wx, wx64);
      const float* pm = p.patch + (v0 + v) * (int64_t)p.ph * p.pw * p.dim;
      const double* dots_view = p.dots + (v0 + v) * (int64_t)p.ph * p.pw * p.q_stride;
      auto sim_dot = [&](int o) { return interp_dot(p, dots_view, iy, ix, wy64, wx64, o, n_q); };
      const double first_dot = p.sim_kernel != DC_SIM_NONE ? sim_dot(lane) : 0.0;
      float f[kMaxPerLane];
#pragma unroll
      for (int k = 0; k < kMaxPerLane; ++k) f[k] = 0.f;
      // ATen order: interpolate along x inside each of the 4 rows, then along y
#pragma unroll
      for (int ty = 0; ty < 4; ++ty) {
        const float* row = pm + (int64_t)iy[ty] * p.pw * p.dim;
        const float* t0 = row + (int64_t)ix[0] * p.dim;
        const float* t1 = row + (int64_t)ix[1] * p.dim;
        const float* t2 = row + (int64_t)ix[2] * p.dim;
        const float* t3 = row + (int64_t)ix[3] * p.dim;
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k) {
          if (k < per_lane) {
            const int c = k * 32 + lane;
            const float r = __ldg(t0 + c) * wx[0] + __ldg(t1 + c) * wx[1] + __ldg(t2 + c) * wx[2] + __ldg(t3 + c) * wx[3];
            f[k] = fmaf(r, wy[ty], f[k]);
          }
        }
      }
      float nrm = 1.f;
      if (p.norm_feat) {
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k)
          if (k < per_lane) ss = fmaf(f[k], f[k], ss);
        nrm = sqrtf(dc::warp_sum(ss));
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k)
          if (k < per_lane) f[k] = f[k] / nrm;
      }
      float weight = 1.f;
      if (p.sim_kernel != DC_SIM_NONE) {
        const int id = seg_at(p.seg, p.seg_dtype, (v0 + v) * hw + (int64_t)pv * p.width + pu);
        weight = 0.f;  // pixels whose id has no query keep metric 0 (quirk q13)
        if (id >= 0 && id < n_q)
          weight = finish_weight(p, first_dot, nrm, id, n_q, lane, sim_dot);
        if (w_scene && lane == 0) w_scene[(int64_t)v * n_pts + i] = weight;
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k)
          if (k < per_lane) acc[k] += f[k] * weight;  // feat2d[ys,xs] * metric, then +=  (two roundings)
      } else {
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k)
          if (k < per_lane) acc[k] += f[k];
      }
    }
    float* dst = p.out_sum + (p0 + i) * p.dim;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k)
      if (k < per_lane) dst[k * 32 + lane] = acc[k];
  }
}

// The fast path (dim = 128 * kChunks: CLIP widths 512 / 768 / 1024). A lane owns 4 consecutive channels of every
// 128-channel chunk (128-bit tap loads). A CTA owns kTilePts consecutive points of the (Morton-sorted) processing
// order and walks the views IN STEP (one barrier per view that sees any of them). Neighbouring points project into
// the same patch cells of a view, so while the CTA is on view v its warps keep hitting the same ~50 KB of taps in
// L1. A point-major version (one warp per point, all its views) let every warp drift through the views at its own
// pace: its ncu capture showed 43 % L1 hits, 10 GB of L2 reads per 8-view scene and 6.6 long-scoreboard stall cycles
// per issue; this kernel: 79 % L1 hits, 3.4 GB, 4.0 (profiles/r01_pixel_fuse.md).
// The per-point accumulators live in shared memory (kTilePts x dim fp32) so that the visible points of a view can
// be dealt to the warps round-robin whatever their position in the tile; a point is touched by one warp per view
// and the barrier orders the views, so every row is accumulated in view order exactly like `sum_features[mask] +=`.
constexpr int kTilePts = 32;

template <int kChunks>
__global__ void __launch_bounds__(kThreads, 2) pixel_fuse_tile_kernel(PixParams p) {
  extern __shared__ double s_cam[];  // [n_views][12] + [9], then the accumulators
  constexpr int kDim = 128 * kChunks;
  const int scene = blockIdx.y;
  const int64_t p0 = p.point_off[scene];
  const int64_t n_pts = p.point_off[scene + 1] - p0;
  const int64_t v0 = p.view_off[scene];
  const int n_views = (int)(p.view_off[scene + 1] - v0);
  const int64_t tile0 = (int64_t)blockIdx.x * kTilePts;
  if (tile0 >= n_pts) return;
  const int n_tile = (int)min((int64_t)kTilePts, n_pts - tile0);
  for (int i = threadIdx.x; i < n_views * 12; i += kThreads) {
    const int v = i / 12, e = i - v * 12;
    s_cam[i] = __ldg(p.inv_poses + (v0 + v) * 16 + e);
  }
  double* s_K = s_cam + n_views * 12;
  if (threadIdx.x < 9) s_K[threadIdx.x] = __ldg(p.intrinsics + (int64_t)scene * 9 + threadIdx.x);
  float4* s_acc = reinterpret_cast<float4*>(s_cam + (((size_t)n_views * 12 + 9 + 1) & ~(size_t)1));  // 16-byte aligned
  float* s_den = reinterpret_cast<float*>(s_acc + kTilePts * (kDim / 4));  // [kTilePts] denominators of fuse_points :266-268
  // per slot and view: the 8 exact (fp64) bicubic weights (y taps, x taps) and the first tap indices, written by the
  // lane that owns the slot, read by the warp that processes the pair
  double* s_tw = reinterpret_cast<double*>(s_den + kTilePts);  // [kTilePts][8]
  int* s_ti = reinterpret_cast<int*>(s_tw + kTilePts * 8);     // [kTilePts][2]
  for (int i = threadIdx.x; i < kTilePts * (kDim / 4); i += kThreads) s_acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (threadIdx.x < kTilePts) s_den[threadIdx.x] = 0.f;
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n_q = p.sim_kernel != DC_SIM_NONE ? (int)(p.query_off[scene + 1] - p.query_off[scene]) : 0;
  const int64_t hw = (int64_t)p.height * p.width;
  const uint8_t* vis_scene = p.visible + p.mask_off[scene];
  float* w_scene = p.out_weight ? p.out_weight + p.mask_off[scene] : nullptr;
  // lane l <-> tile slot l: original index of the point at sorted position tile0 + l
  const int64_t my_i = lane < n_tile ? (p.perm ? __ldg(p.perm + p0 + tile0 + lane) : tile0 + lane) : 0;
  const double px = __ldg(p.points + 3 * (p0 + my_i)), py = __ldg(p.points + 3 * (p0 + my_i) + 1), pz = __ldg(p.points + 3 * (p0 + my_i) + 2);

  for (int v = 0; v < n_views; ++v) {
    const bool vis = lane < n_tile && vis_scene[(int64_t)v * n_pts + my_i] != 0;
    const unsigned m = __ballot_sync(0xffffffffu, vis);  // identical in every warp of the CTA
    if (m == 0) continue;
    const int n_vis = __popc(m);
    const float4* pm = reinterpret_cast<const float4*>(p.patch + (v0 + v) * (int64_t)p.ph * p.pw * kDim);
    const double* dots_view = p.dots + (v0 + v) * (int64_t)p.ph * p.pw * p.q_stride;
    // lane-parallel projection: every lane projects ITS slot's point once per view (the same fp64 arithmetic as
    // visibility.cu; the point is visible, so it is inside); the pair loop below fetches pixels by shuffle
    const double* cam = s_cam + v * 12;
    int my_pu = 0, my_pv = 0;
    {
      double cx = __dadd_rn(cam[3], __fma_rn(cam[2], pz, __fma_rn(cam[1], py, __dmul_rn(cam[0], px))));
      double cy = __dadd_rn(cam[7], __fma_rn(cam[6], pz, __fma_rn(cam[5], py, __dmul_rn(cam[4], px))));
      double cz = __dadd_rn(cam[11], __fma_rn(cam[10], pz, __fma_rn(cam[9], py, __dmul_rn(cam[8], px))));
      cy = -cy;
      cz = -cz;
      const double qx = __fma_rn(s_K[2], cz, __fma_rn(s_K[1], cy, __dmul_rn(s_K[0], cx)));
      const double qy = __fma_rn(s_K[5], cz, __fma_rn(s_K[4], cy, __dmul_rn(s_K[3], cx)));
      const double qz = __fma_rn(s_K[8], cz, __fma_rn(s_K[7], cy, __dmul_rn(s_K[6], cx)));
      if (vis && qz != 0.0) {
        my_pu = (int)__ddiv_rn(qx, qz);
        my_pv = (int)__ddiv_rn(qy, qz);
      }
      // every warp computes and writes the same values (m, the pixels and hence these are identical in all warps);
      // the __syncthreads() that ends the previous visible view orders them against that view's readers
      if (vis) {
        double wy64[4], wx64[4];
        s_ti[2 * lane] = cubic_weights64(my_pv, p.ph, p.height, wy64);
        s_ti[2 * lane + 1] = cubic_weights64(my_pu, p.pw, p.width, wx64);
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
          s_tw[8 * lane + t4] = wy64[t4];
          s_tw[8 * lane + 4 + t4] = wx64[t4];
        }
      }
      __syncwarp();
    }
    for (int k = warp; k < n_vis; k += kWarps) {
      const int slot = __fns(m, 0, k + 1);
      const int64_t i = __shfl_sync(0xffffffffu, my_i, slot);
      const int pu = __shfl_sync(0xffffffffu, my_pu, slot), pv = __shfl_sync(0xffffffffu, my_pv, slot);
      int iy[4], ix[4];
      float wy[4], wx[4];
      {
        const int iy0 = s_ti[2 * slot], ix0 = s_ti[2 * slot + 1];
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
          iy[t4] = min(max(iy0 - 1 + t4, 0), p.ph - 1);
          ix[t4] = min(max(ix0 - 1 + t4, 0), p.pw - 1);
          wy[t4] = (float)s_tw[8 * slot + t4];
          wx[t4] = (float)s_tw[8 * slot + 4 + t4];
        }
      }
      // the instance id and the first round of query dots are requested before the taps: both are cold loads whose
      // latency then hides behind the 96 tap loads instead of sitting in front of the accumulator update
      long long raw_id = p.sim_kernel != DC_SIM_NONE ? seg_raw(p.seg, p.seg_dtype, (v0 + v) * hw + (int64_t)pv * p.width + pu) : -1;
      // the fp64 weights are re-read from shared memory inside the lambda (broadcast loads; nothing stays live)
      auto sim_dot = [&](int o) {
        double wy64[4], wx64[4];
#pragma unroll
        for (int t4 = 0; t4 < 4; ++t4) {
          wy64[t4] = s_tw[8 * slot + t4];
          wx64[t4] = s_tw[8 * slot + 4 + t4];
        }
        return interp_dot(p, dots_view, iy, ix, wy64, wx64, o, n_q);
      };
      const double first_dot = p.sim_kernel != DC_SIM_NONE ? sim_dot(lane) : 0.0;
      // Chunk-major: the 16 taps of one 128-channel chunk are requested back to back (16 independent 128-bit loads
      // in flight per lane) and folded in the ATen order (along x inside each row, then along y); finished chunks
      // occupy 4 registers each, so the loads of the next chunk have room without spilling.
      float4 f[kChunks];
      int row_off[4], col_off[4];
#pragma unroll
      for (int t4 = 0; t4 < 4; ++t4) {
        row_off[t4] = iy[t4] * p.pw * (kDim / 4) + lane;
        col_off[t4] = ix[t4] * (kDim / 4);
      }
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        float4 t[4][4];
#pragma unroll
        for (int ty = 0; ty < 4; ++ty)
#pragma unroll
          for (int tx = 0; tx < 4; ++tx) t[ty][tx] = __ldg(pm + (row_off[ty] + col_off[tx] + c * 32));
        // two-wide fp32 (fma.rn.f32x2): the same multiply-add chain per channel, half the instructions
        float2 lo = make_float2(0.f, 0.f), hi = lo;
#pragma unroll
        for (int ty = 0; ty < 4; ++ty) {
          const float2 w0 = make_float2(wx[0], wx[0]), w1 = make_float2(wx[1], wx[1]), w2 = make_float2(wx[2], wx[2]),
                       w3 = make_float2(wx[3], wx[3]), wv = make_float2(wy[ty], wy[ty]);
          float2 r = __fmul2_rn(make_float2(t[ty][0].x, t[ty][0].y), w0);
          r = __ffma2_rn(make_float2(t[ty][1].x, t[ty][1].y), w1, r);
          r = __ffma2_rn(make_float2(t[ty][2].x, t[ty][2].y), w2, r);
          r = __ffma2_rn(make_float2(t[ty][3].x, t[ty][3].y), w3, r);
          lo = __ffma2_rn(r, wv, lo);
          float2 q = __fmul2_rn(make_float2(t[ty][0].z, t[ty][0].w), w0);
          q = __ffma2_rn(make_float2(t[ty][1].z, t[ty][1].w), w1, q);
          q = __ffma2_rn(make_float2(t[ty][2].z, t[ty][2].w), w2, q);
          q = __ffma2_rn(make_float2(t[ty][3].z, t[ty][3].w), w3, q);
          hi = __ffma2_rn(q, wv, hi);
        }
        const float4 acc4 = make_float4(lo.x, lo.y, hi.x, hi.y);
        f[c] = acc4;
      }
      float nrm = 1.f;
      if (p.norm_feat) {
        float ss = 0.f;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) ss = fmaf(f[c].x, f[c].x, fmaf(f[c].y, f[c].y, fmaf(f[c].z, f[c].z, fmaf(f[c].w, f[c].w, ss))));
        nrm = sqrtf(dc::warp_sum(ss));
        // one correctly rounded reciprocal and 4 * kChunks multiplications instead of as many IEEE divisions (8-10
        // instructions each): the quotients differ from `feat2d /= norm` by at most 1 ulp, far inside the 1e-3 bar
        const float inv = __frcp_rn(nrm);
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          f[c].x = f[c].x * inv;
          f[c].y = f[c].y * inv;
          f[c].z = f[c].z * inv;
          f[c].w = f[c].w * inv;
        }
      }
      float4* acc = s_acc + slot * (kDim / 4) + lane;
      if (p.sim_kernel != DC_SIM_NONE) {
        asm volatile("" : "+l"(raw_id));  // keeps the narrowing (and the wait for the load) down here
        const int id = seg_id(raw_id);
        float weight = 0.f;  // pixels whose id has no query keep metric 0 (quirk q13)
        if (id >= 0 && id < n_q)
          weight = finish_weight(p, first_dot, nrm, id, n_q, lane, sim_dot);
        if (w_scene && lane == 0) w_scene[(int64_t)v * n_pts + i] = weight;
        if (lane == 0) s_den[slot] += weight;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {  // feat2d[ys,xs] * metric, then +=  (two roundings)
          float4 a = acc[c * 32];
          a.x += f[c].x * weight;
          a.y += f[c].y * weight;
          a.z += f[c].z * weight;
          a.w += f[c].w * weight;
          acc[c * 32] = a;
        }
      } else {
        if (lane == 0) s_den[slot] += 1.f;
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          float4 a = acc[c * 32];
          a.x += f[c].x;
          a.y += f[c].y;
          a.z += f[c].z;
          a.w += f[c].w;
          acc[c * 32] = a;
        }
      }
    }
    __syncthreads();  // the next view may hand a point to another warp
  }
  for (int slot = warp; slot < n_tile; slot += kWarps) {
    const int64_t i = __shfl_sync(0xffffffffu, my_i, slot);
    float4* dst = reinterpret_cast<float4*>(p.out_sum + (p0 + i) * kDim) + lane;
    const float den = s_den[slot];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      float4 a = s_acc[slot * (kDim / 4) + c * 32 + lane];
      if (p.normalize) a = make_float4(a.x / den, a.y / den, a.z / den, a.w / den);
      dst[c * 32] = a;
    }
  }
}

// similarity_mask entries of invisible (point, view) pairs are 0 (:238 writes visible pairs only into a zero tensor).
// The tile kernel writes the visible ones; this pass fills the rest with coalesced stores (the point-major kernels
// write them from inside their view loop, which in Morton order is one scattered 4-byte store per pair).
__global__ void __launch_bounds__(kThreads) zero_invisible_weights_kernel(const uint8_t* __restrict__ visible,
                                                                          const int64_t* __restrict__ mask_off,
                                                                          float* __restrict__ weight) {
  const int64_t begin = mask_off[blockIdx.y], end = mask_off[blockIdx.y + 1];
  for (int64_t j = begin + (int64_t)blockIdx.x * kThreads + threadIdx.x; j < end; j += (int64_t)gridDim.x * kThreads)
    if (!visible[j]) weight[j] = 0.f;
}

// feat[j,:] /= denom[j], denom = sum_v weight[v,j] (similarity) or sum_v visible[v,j]
__global__ void __launch_bounds__(kThreads) pixel_normalize_kernel(float* __restrict__ sums, const int64_t* __restrict__ point_off,
                                                                   const int64_t* __restrict__ view_off,
                                                                   const int64_t* __restrict__ mask_off,
                                                                   const uint8_t* __restrict__ visible,
                                                                   const float* __restrict__ weight, int dim) {
  const int scene = blockIdx.y;
  const int64_t p0 = point_off[scene];
  const int64_t n_pts = point_off[scene + 1] - p0;
  const int n_views = (int)(view_off[scene + 1] - view_off[scene]);
  const int lane = threadIdx.x & 31;
  for (int64_t i = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); i < n_pts; i += (int64_t)gridDim.x * kWarps) {
    float denom = 0.f;
    if (weight) {
      const float* w = weight + mask_off[scene];
      for (int v = 0; v < n_views; ++v) denom += w[(int64_t)v * n_pts + i];
    } else {
      const uint8_t* m = visible + mask_off[scene];
      int c = 0;
      for (int v = 0; v < n_views; ++v) c += m[(int64_t)v * n_pts + i] != 0;
      denom = (float)c;
    }
    float* row = sums + (p0 + i) * dim;
    for (int c = lane; c < dim; c += 32) row[c] = row[c] / denom;
  }
}

// generate_view_clip (data/dataset_blender.py:132-171): every point of a cloud is projected into ONE view
// (fp64 pose - the json world_matrix is inverted in fp64 there -, truncation toward zero, pixel (0,0) when the
// projected z is 0), the pixel is CLIPPED into the image instead of tested (:158-159), and the bicubically
// upsampled patch feature of that pixel is gathered (:151-156,165). No visibility, no weights. The reference
// materialises the (h, w, C) map per view; here the 16 taps are evaluated per point. Warp per (point, view);
// lanes stride over channels in float4 when the rows allow it.
template <bool kVec>
__global__ void __launch_bounds__(kThreads) view_clip_gather_kernel(const double* __restrict__ points, int64_t n_pts,
                                                                    const double* __restrict__ inv_poses,
                                                                    const double* __restrict__ intrinsics,
                                                                    const float* __restrict__ patch, int ph, int pw, int dim,
                                                                    int height, int width, float* __restrict__ out) {
  const int view = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const double* m = inv_poses + (int64_t)view * 16;
  const double* K = intrinsics;
  const float scale_y = (float)ph / (float)height, scale_x = (float)pw / (float)width;
  const float* pm = patch + (int64_t)view * ph * pw * dim;
  for (int64_t i = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); i < n_pts; i += (int64_t)gridDim.x * kWarps) {
    const double x = __ldg(points + 3 * i), y = __ldg(points + 3 * i + 1), z = __ldg(points + 3 * i + 2);
    const double cx = __dadd_rn(__ldg(m + 3), __fma_rn(__ldg(m + 2), z, __fma_rn(__ldg(m + 1), y, __dmul_rn(__ldg(m), x))));
    const double cy = -__dadd_rn(__ldg(m + 7), __fma_rn(__ldg(m + 6), z, __fma_rn(__ldg(m + 5), y, __dmul_rn(__ldg(m + 4), x))));
    const double cz = -__dadd_rn(__ldg(m + 11), __fma_rn(__ldg(m + 10), z, __fma_rn(__ldg(m + 9), y, __dmul_rn(__ldg(m + 8), x))));
    const double qx = __fma_rn(__ldg(K + 2), cz, __fma_rn(__ldg(K + 1), cy, __dmul_rn(__ldg(K), cx)));
    const double qy = __fma_rn(__ldg(K + 5), cz, __fma_rn(__ldg(K + 4), cy, __dmul_rn(__ldg(K + 3), cx)));
    const double qz = __fma_rn(__ldg(K + 8), cz, __fma_rn(__ldg(K + 7), cy, __dmul_rn(__ldg(K + 6), cx)));
    int pu = 0, pv = 0;
    if (qz != 0.0) {
      // fp64 -> int64 assignment: truncation; NaN, +-inf and anything beyond int64 become INT64_MIN, which np.clip turns into 0
      const double uq = __ddiv_rn(qx, qz), vq = __ddiv_rn(qy, qz);
      if (fabs(uq) < 9.2e18) pu = (int)max(0ll, min((long long)(width - 1), __double2ll_rz(uq)));
      if (fabs(vq) < 9.2e18) pv = (int)max(0ll, min((long long)(height - 1), __double2ll_rz(vq)));
    }
    int iy[4], ix[4];
    float wy[4], wx[4];
    cubic_taps(pv, scale_y, ph, iy, wy);
    cubic_taps(pu, scale_x, pw, ix, wx);
    float* row_out = out + ((int64_t)view * n_pts + i) * dim;
    if (kVec) {
      for (int c = lane; c < dim / 4; c += 32) {
        float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int ty = 0; ty < 4; ++ty) {
          const float* row = pm + (int64_t)iy[ty] * pw * dim;
          const float4 a = __ldg(reinterpret_cast<const float4*>(row + (int64_t)ix[0] * dim) + c);
          const float4 b = __ldg(reinterpret_cast<const float4*>(row + (int64_t)ix[1] * dim) + c);
          const float4 d = __ldg(reinterpret_cast<const float4*>(row + (int64_t)ix[2] * dim) + c);
          const float4 e = __ldg(reinterpret_cast<const float4*>(row + (int64_t)ix[3] * dim) + c);
          f.x = fmaf(a.x * wx[0] + b.x * wx[1] + d.x * wx[2] + e.x * wx[3], wy[ty], f.x);
          f.y = fmaf(a.y * wx[0] + b.y * wx[1] + d.y * wx[2] + e.y * wx[3], wy[ty], f.y);
          f.z = fmaf(a.z * wx[0] + b.z * wx[1] + d.z * wx[2] + e.z * wx[3], wy[ty], f.z);
          f.w = fmaf(a.w * wx[0] + b.w * wx[1] + d.w * wx[2] + e.w * wx[3], wy[ty], f.w);
        }
        __stcs(reinterpret_cast<float4*>(row_out) + c, f);  // written once, never re-read here
      }
    } else {
      for (int c = lane; c < dim; c += 32) {
        float f = 0.f;
#pragma unroll
        for (int ty = 0; ty < 4; ++ty) {
          const float* row = pm + (int64_t)iy[ty] * pw * dim + c;
          const float r = __ldg(row + (int64_t)ix[0] * dim) * wx[0] + __ldg(row + (int64_t)ix[1] * dim) * wx[1] +
                          __ldg(row + (int64_t)ix[2] * dim) * wx[2] + __ldg(row + (int64_t)ix[3] * dim) * wx[3];
          f = fmaf(r, wy[ty], f);
        }
        row_out[c] = f;
      }
    }
  }
}

unsigned blocks_for(int64_t max_points, int n_scenes) {
  int64_t want = dc::ceil_div<int64_t>(max_points, kWarps);
  int64_t cap = dc::ceil_div<int64_t>((int64_t)dc::sm_count() * 16, n_scenes);
  if (want > cap) want = cap;
  return (unsigned)(want < 1 ? 1 : want);
}

int query_stride(int max_queries) { return (max_queries + 3) & ~3; }  // whole 32-byte sectors per (cell) row of doubles

}  // namespace

extern "C" {

size_t dc_pixel_fuse_workspace(int64_t total_views, int patch_h, int patch_w, int max_queries_per_scene) {
  if (total_views <= 0 || patch_h <= 0 || patch_w <= 0 || max_queries_per_scene <= 0) return 0;
  return (size_t)total_views * patch_h * patch_w * query_stride(max_queries_per_scene) * sizeof(double);
}

int dc_pixel_fuse(const double* points, const int64_t* point_off, const int64_t* view_off, const double* inv_poses,
                  const double* intrinsics, const int64_t* mask_off, const uint8_t* visible, const void* seg, int seg_dtype,
                  const float* patch_feats, int patch_h, int patch_w, int dim, const float* queries,
                  const int64_t* query_off, int sim_kernel, int norm_feat, int n_scenes, int64_t max_points_per_scene,
                  int max_views_per_scene, int height, int width, const int64_t* perm, float* out_sum, float* out_weight,
                  int normalize, int64_t total_views, int max_queries_per_scene, void* workspace, size_t workspace_bytes,
                  dc_stream_t stream) {
  DC_CHECK_ARG(points && point_off && view_off && inv_poses && intrinsics && mask_off && visible && patch_feats && out_sum,
               "dc_pixel_fuse: null pointer argument");
  DC_CHECK_ARG(sim_kernel >= DC_SIM_NONE && sim_kernel <= DC_SIM_MEAN, "dc_pixel_fuse: Please set method in [mean, max]");
  DC_CHECK_ARG(sim_kernel == DC_SIM_NONE || (queries && query_off && seg), "dc_pixel_fuse: similarity needs queries and seg");
  DC_CHECK_ARG(dim > 0 && dim % 32 == 0 && dim <= 32 * kMaxPerLane, "dc_pixel_fuse: dim must be a multiple of 32, <= %d",
               32 * kMaxPerLane);
  DC_CHECK_ARG(patch_h > 0 && patch_w > 0 && height > 0 && width > 0, "dc_pixel_fuse: bad sizes");
  if (n_scenes <= 0 || max_points_per_scene <= 0) return DC_OK;
  DC_CHECK_ARG(n_scenes <= 65535, "dc_pixel_fuse: at most 65535 scenes per call");
  const size_t smem = ((size_t)max_views_per_scene * 12 + 9) * sizeof(double);
  DC_CHECK_ARG(smem <= 48 * 1024, "dc_pixel_fuse: too many views per scene (%d)", max_views_per_scene);
  PixParams p{points, point_off, view_off, inv_poses, intrinsics, mask_off, visible, seg, seg_dtype, patch_feats, patch_h,
              patch_w, dim, queries, query_off, sim_kernel, norm_feat, height, width, perm, out_sum, out_weight, nullptr, 0, normalize};
  dim3 grid(blocks_for(max_points_per_scene, n_scenes), (unsigned)n_scenes);
  const bool aligned = (((uintptr_t)patch_feats | (uintptr_t)out_sum) & 15) == 0;
  cudaStream_t st = dc::as_stream(stream);
  if (sim_kernel != DC_SIM_NONE) {
    DC_CHECK_ARG(total_views >= 0 && max_queries_per_scene >= 0, "dc_pixel_fuse: bad extents");
    const size_t need = dc_pixel_fuse_workspace(total_views, patch_h, patch_w, max_queries_per_scene);
    DC_CHECK_ARG(workspace && workspace_bytes >= need, "dc_pixel_fuse: workspace too small (%zu < %zu bytes)", workspace_bytes, need);
    DC_CHECK_ARG(((uintptr_t)workspace & 7) == 0, "dc_pixel_fuse: workspace must be 8-byte aligned");
    p.dots = static_cast<double*>(workspace);
    p.q_stride = query_stride(max_queries_per_scene);
    if (total_views > 0 && max_queries_per_scene > 0) {
      const int64_t rows = (int64_t)max_views_per_scene * patch_h * patch_w;
      dim3 dgrid(blocks_for(rows, n_scenes), (unsigned)n_scenes);
      patch_query_dots_kernel<<<dgrid, kThreads, 0, st>>>(p);
      DC_LAUNCH_CHECK();
    }
  }
  if (aligned && (dim == 768 || dim == 512 || dim == 1024)) {  // CLIP ViT-L/14, ViT-B, ViT-H widths
    const size_t cam_doubles = ((size_t)max_views_per_scene * 12 + 9 + 1) & ~(size_t)1;
    const size_t tsmem = cam_doubles * sizeof(double) + (size_t)kTilePts * dim * sizeof(float) + kTilePts * sizeof(float) +
                         kTilePts * (8 * sizeof(double) + 2 * sizeof(int));
    dim3 tgrid((unsigned)dc::ceil_div<int64_t>(max_points_per_scene, kTilePts), (unsigned)n_scenes);
    DC_CHECK_ARG(tsmem <= 200 * 1024, "dc_pixel_fuse: too many views per scene (%d)", max_views_per_scene);
    // the attribute belongs to the (function, device) pair, so it is set on every call for the device in use
    const void* fn = dim == 768 ? (const void*)pixel_fuse_tile_kernel<6>
                   : dim == 512 ? (const void*)pixel_fuse_tile_kernel<4> : (const void*)pixel_fuse_tile_kernel<8>;
    DC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    if (dim == 768) pixel_fuse_tile_kernel<6><<<tgrid, kThreads, tsmem, st>>>(p);
    else if (dim == 512) pixel_fuse_tile_kernel<4><<<tgrid, kThreads, tsmem, st>>>(p);
    else pixel_fuse_tile_kernel<8><<<tgrid, kThreads, tsmem, st>>>(p);
    if (out_weight) {
      DC_LAUNCH_CHECK();
      dim3 zgrid((unsigned)std::min<int64_t>(dc::ceil_div<int64_t>(max_points_per_scene * max_views_per_scene, kThreads * 8), 4096),
                 (unsigned)n_scenes);
      zero_invisible_weights_kernel<<<zgrid, kThreads, 0, st>>>(visible, mask_off, out_weight);
    }
  }
  else {
    pixel_fuse_kernel<<<grid, kThreads, smem, st>>>(p);
    if (normalize) {
      DC_LAUNCH_CHECK();
      DC_CHECK_ARG(sim_kernel == DC_SIM_NONE || out_weight, "dc_pixel_fuse: normalize needs out_weight on this path");
      pixel_normalize_kernel<<<grid, kThreads, 0, st>>>(out_sum, point_off, view_off, mask_off, visible,
                                                        sim_kernel != DC_SIM_NONE ? out_weight : nullptr, dim);
    }
  }
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_pixel_normalize(float* sums, const int64_t* point_off, const int64_t* view_off, const int64_t* mask_off,
                       const uint8_t* visible, const float* weight, int n_scenes, int64_t max_points_per_scene, int dim,
                       dc_stream_t stream) {
  DC_CHECK_ARG(sums && point_off && view_off && mask_off && (visible || weight), "dc_pixel_normalize: null pointer argument");
  if (n_scenes <= 0 || max_points_per_scene <= 0) return DC_OK;
  dim3 grid(blocks_for(max_points_per_scene, n_scenes), (unsigned)n_scenes);
  pixel_normalize_kernel<<<grid, kThreads, 0, dc::as_stream(stream)>>>(sums, point_off, view_off, mask_off, visible, weight, dim);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_view_clip_gather(const double* points, int64_t n_points, const double* inv_poses, const double* intrinsics,
                        const float* patch_feats, int n_views, int patch_h, int patch_w, int dim, int height, int width,
                        float* out, dc_stream_t stream) {
  DC_CHECK_ARG(points && inv_poses && intrinsics && patch_feats && out, "dc_view_clip_gather: null pointer argument");
  DC_CHECK_ARG(patch_h > 0 && patch_w > 0 && dim > 0 && height > 0 && width > 0, "dc_view_clip_gather: bad sizes");
  if (n_points <= 0 || n_views <= 0) return DC_OK;
  DC_CHECK_ARG(n_views <= 65535, "dc_view_clip_gather: at most 65535 views per call");
  dim3 grid(blocks_for(n_points, n_views), (unsigned)n_views);
  const bool vec = dim % 4 == 0 && (((uintptr_t)patch_feats | (uintptr_t)out) & 15) == 0;
  cudaStream_t st = dc::as_stream(stream);
  if (vec) view_clip_gather_kernel<true><<<grid, kThreads, 0, st>>>(points, n_points, inv_poses, intrinsics, patch_feats, patch_h,
                                                                    patch_w, dim, height, width, out);
  else view_clip_gather_kernel<false><<<grid, kThreads, 0, st>>>(points, n_points, inv_poses, intrinsics, patch_feats, patch_h,
                                                                  patch_w, dim, height, width, out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // extern "C"
