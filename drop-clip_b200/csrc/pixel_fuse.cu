// Pixel-level fusion (the `use_obj_prior=0` ablation path).
//
// Reference: MultiviewFeatureFusion.aggregate_features utils/feature_fusion.py:138-250 and the final
// division of fuse_points :266-268. Per view the reference bicubically upsamples the (ph,pw,C)
// patch map to (H,W,C) (943.7 MB at 480x640x768 fp32), L2-normalises every pixel, multiplies the
// whole map with the query matrix, builds a per-pixel similarity metric, and finally gathers the
// visible projected pixels. Only those gathered pixels matter, so this kernel evaluates the 16
// bicubic taps directly at each visible (point, view), in registers, and never materialises the
// map. It is point-major: a warp owns a point, walks its views in order (same accumulation
// order as `sum_features[mask] += feat3d`), and writes the point's row once.
//
// Bicubic weights follow ATen's upsample_bicubic2d (align_corners=False, A=-0.75, source index
// scale*(dst+0.5)-0.5 un-clamped, taps clamped to the border).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxPerLane = 32;  // channels per lane: dim <= 1024

struct PixParams {
  const double* points;
  const int64_t* point_off;
  const int64_t* view_off;
  const float* inv_poses;
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

__device__ __forceinline__ int seg_at(const void* seg, int dtype, int64_t idx) {
  if (dtype == DC_U8) return (int)__ldg(reinterpret_cast<const uint8_t*>(seg) + idx);
  if (dtype == DC_I32) return __ldg(reinterpret_cast<const int32_t*>(seg) + idx);
  const long long v = __ldg(reinterpret_cast<const long long*>(seg) + idx);
  return (v < 0 || v > 0x7fffffff) ? -1 : (int)v;
}

__global__ void __launch_bounds__(kThreads) pixel_fuse_kernel(PixParams p) {
  extern __shared__ double s_cam[];  // [n_views][12] + [9]
  const int scene = blockIdx.y;
  const int64_t p0 = p.point_off[scene];
  const int64_t n_pts = p.point_off[scene + 1] - p0;
  const int64_t v0 = p.view_off[scene];
  const int n_views = (int)(p.view_off[scene + 1] - v0);
  for (int i = threadIdx.x; i < n_views * 12; i += kThreads) {
    const int v = i / 12, e = i - v * 12;
    s_cam[i] = (double)__ldg(p.inv_poses + (v0 + v) * 16 + e);
  }
  double* s_K = s_cam + n_views * 12;
  if (threadIdx.x < 9) s_K[threadIdx.x] = __ldg(p.intrinsics + (int64_t)scene * 9 + threadIdx.x);
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int per_lane = p.dim / 32;  // host guarantees dim % 32 == 0, dim <= 1024
  const int n_q = p.sim_kernel != DC_SIM_NONE ? (int)(p.query_off[scene + 1] - p.query_off[scene]) : 0;
  const float* q = p.sim_kernel != DC_SIM_NONE ? p.queries + p.query_off[scene] * p.dim : nullptr;
  const float scale_y = (float)p.ph / (float)p.height, scale_x = (float)p.pw / (float)p.width;
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
      cubic_taps(pv, scale_y, p.ph, iy, wy);
      cubic_taps(pu, scale_x, p.pw, ix, wx);
      const float* pm = p.patch + (v0 + v) * (int64_t)p.ph * p.pw * p.dim;
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
      if (p.norm_feat) {
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k)
          if (k < per_lane) ss = fmaf(f[k], f[k], ss);
        const float nrm = sqrtf(dc::warp_sum(ss));
#pragma unroll
        for (int k = 0; k < kMaxPerLane; ++k)
          if (k < per_lane) f[k] = f[k] / nrm;
      }
      float weight = 1.f;
      if (p.sim_kernel != DC_SIM_NONE) {
        const int id = seg_at(p.seg, p.seg_dtype, (v0 + v) * hw + (int64_t)pv * p.width + pu);
        weight = 0.f;  // pixels whose id has no query keep metric 0 (quirk q13)
        if (id >= 0 && id < n_q) {
          float pos = 0.f, red = (p.sim_kernel == DC_SIM_MAX) ? -INFINITY : 0.f;
          bool nan_seen = false;
          for (int o = 0; o < n_q; ++o) {
            const float* qo = q + (int64_t)o * p.dim;
            float d = 0.f;
#pragma unroll
            for (int k = 0; k < kMaxPerLane; ++k)
              if (k < per_lane) d = fmaf(f[k], __ldg(qo + k * 32 + lane), d);
            d = dc::warp_sum(d);
            if (o == id) pos = d;
            else {
              nan_seen |= (d != d);
              if (p.sim_kernel == DC_SIM_MAX) red = fmaxf(red, d);
              else red += d;
            }
          }
          if (p.sim_kernel == DC_SIM_MEAN) red = red / (float)(n_q - 1);
          if (nan_seen) red = __int_as_float(0x7fc00000);
          weight = pos - red;
          if (weight == weight) weight = fmaxf(weight, 1e-6f);
        }
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

// Same arithmetic with 128-bit accesses: a lane owns 4 consecutive channels of every 128-channel chunk
// (dim = 128 * kChunks; CLIP ViT-L/14: 6 chunks), which cuts the load instructions per visible
// (point, view) from 16 * dim / 32 to 16 * dim / 128, and the Q per-query warp reductions (5 shuffles
// each) are replaced by one transposed reduction of up to 32 partial dot products (31 shuffles),
// after which lane o holds query o's similarity. The round-1 profile had this kernel at ~1 % of the
// fp32 peak, bound by the latency of 900 dependent 4-byte loads per lane and pair.
template <int kChunks>
__global__ void __launch_bounds__(kThreads) pixel_fuse_vec_kernel(PixParams p) {
  extern __shared__ double s_cam[];  // [n_views][12] + [9]
  const int scene = blockIdx.y;
  const int64_t p0 = p.point_off[scene];
  const int64_t n_pts = p.point_off[scene + 1] - p0;
  const int64_t v0 = p.view_off[scene];
  const int n_views = (int)(p.view_off[scene + 1] - v0);
  for (int i = threadIdx.x; i < n_views * 12; i += kThreads) {
    const int v = i / 12, e = i - v * 12;
    s_cam[i] = (double)__ldg(p.inv_poses + (v0 + v) * 16 + e);
  }
  double* s_K = s_cam + n_views * 12;
  if (threadIdx.x < 9) s_K[threadIdx.x] = __ldg(p.intrinsics + (int64_t)scene * 9 + threadIdx.x);
  __syncthreads();

  constexpr int kDim = 128 * kChunks;
  const int lane = threadIdx.x & 31;
  const int n_q = p.sim_kernel != DC_SIM_NONE ? (int)(p.query_off[scene + 1] - p.query_off[scene]) : 0;
  const float4* q4 = p.sim_kernel != DC_SIM_NONE ? reinterpret_cast<const float4*>(p.queries + p.query_off[scene] * kDim) : nullptr;
  const float scale_y = (float)p.ph / (float)p.height, scale_x = (float)p.pw / (float)p.width;
  const int64_t hw = (int64_t)p.height * p.width;
  const uint8_t* vis_scene = p.visible + p.mask_off[scene];
  float* w_scene = p.out_weight ? p.out_weight + p.mask_off[scene] : nullptr;

  for (int64_t s_pos = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5); s_pos < n_pts; s_pos += (int64_t)gridDim.x * kWarps) {
    const int64_t i = p.perm ? __ldg(p.perm + p0 + s_pos) : s_pos;  // spatially sorted processing order, original indexing
    const double x = __ldg(p.points + 3 * (p0 + i)), y = __ldg(p.points + 3 * (p0 + i) + 1),
                 z = __ldg(p.points + 3 * (p0 + i) + 2);
    float4 acc[kChunks];
#pragma unroll
    for (int k = 0; k < kChunks; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int v = 0; v < n_views; ++v) {
      const bool vis = vis_scene[(int64_t)v * n_pts + i] != 0;  // warp-uniform
      if (!vis) {
        if (w_scene && lane == 0) w_scene[(int64_t)v * n_pts + i] = 0.f;
        continue;
      }
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
      cubic_taps(pv, scale_y, p.ph, iy, wy);
      cubic_taps(pu, scale_x, p.pw, ix, wx);
      const float4* pm = reinterpret_cast<const float4*>(p.patch + (v0 + v) * (int64_t)p.ph * p.pw * kDim);
      float4 f[kChunks];
#pragma unroll
      for (int k = 0; k < kChunks; ++k) f[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      // ATen order: interpolate along x inside each of the 4 rows, then along y
#pragma unroll
      for (int ty = 0; ty < 4; ++ty) {
        const float4* row = pm + (int64_t)iy[ty] * p.pw * (kDim / 4);
        const float4* t0 = row + (int64_t)ix[0] * (kDim / 4) + lane;
        const float4* t1 = row + (int64_t)ix[1] * (kDim / 4) + lane;
        const float4* t2 = row + (int64_t)ix[2] * (kDim / 4) + lane;
        const float4* t3 = row + (int64_t)ix[3] * (kDim / 4) + lane;
#pragma unroll
        for (int k = 0; k < kChunks; ++k) {
          const float4 a = __ldg(t0 + k * 32), b = __ldg(t1 + k * 32), c = __ldg(t2 + k * 32), d = __ldg(t3 + k * 32);
          f[k].x = fmaf(a.x * wx[0] + b.x * wx[1] + c.x * wx[2] + d.x * wx[3], wy[ty], f[k].x);
          f[k].y = fmaf(a.y * wx[0] + b.y * wx[1] + c.y * wx[2] + d.y * wx[3], wy[ty], f[k].y);
          f[k].z = fmaf(a.z * wx[0] + b.z * wx[1] + c.z * wx[2] + d.z * wx[3], wy[ty], f[k].z);
          f[k].w = fmaf(a.w * wx[0] + b.w * wx[1] + c.w * wx[2] + d.w * wx[3], wy[ty], f[k].w);
        }
      }
      if (p.norm_feat) {
        float ss = 0.f;
#pragma unroll
        for (int k = 0; k < kChunks; ++k) ss = fmaf(f[k].x, f[k].x, fmaf(f[k].y, f[k].y, fmaf(f[k].z, f[k].z, fmaf(f[k].w, f[k].w, ss))));
        const float nrm = sqrtf(dc::warp_sum(ss));
#pragma unroll
        for (int k = 0; k < kChunks; ++k) {
          f[k].x = f[k].x / nrm;
          f[k].y = f[k].y / nrm;
          f[k].z = f[k].z / nrm;
          f[k].w = f[k].w / nrm;
        }
      }
      float weight = 1.f;
      if (p.sim_kernel != DC_SIM_NONE) {
        const int id = seg_at(p.seg, p.seg_dtype, (v0 + v) * hw + (int64_t)pv * p.width + pu);
        weight = 0.f;  // pixels whose id has no query keep metric 0 (quirk q13)
        if (id >= 0 && id < n_q) {
          float pos = 0.f, red = (p.sim_kernel == DC_SIM_MAX) ? -INFINITY : 0.f;
          bool nan_seen = false;
          for (int o0 = 0; o0 < n_q; o0 += 32) {
            float d[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              d[j] = 0.f;
              if (o0 + j < n_q) {  // warp-uniform
                const float4* qo = q4 + (int64_t)(o0 + j) * (kDim / 4) + lane;
#pragma unroll
                for (int k = 0; k < kChunks; ++k) {
                  const float4 t = __ldg(qo + k * 32);
                  d[j] = fmaf(f[k].x, t.x, fmaf(f[k].y, t.y, fmaf(f[k].z, t.z, fmaf(f[k].w, t.w, d[j]))));
                }
              }
            }
            // transposed reduction: afterwards lane l holds the full dot product of query o0 + l
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
#pragma unroll
              for (int j = 0; j < s; ++j) {
                const bool upper = (lane & s) != 0;
                const float send = upper ? d[j] : d[j + s];
                const float keep = upper ? d[j + s] : d[j];
                d[j] = keep + __shfl_xor_sync(0xffffffffu, send, s);
              }
            }
            const float mine = d[0];
            const int o = o0 + lane;
            if (id >= o0 && id < o0 + 32) pos = __shfl_sync(0xffffffffu, mine, id - o0);
            const bool is_neg = o < n_q && o != id;
            nan_seen |= __any_sync(0xffffffffu, is_neg && (mine != mine));
            if (p.sim_kernel == DC_SIM_MAX) red = fmaxf(red, dc::warp_max(is_neg ? mine : -INFINITY));
            else red += dc::warp_sum(is_neg ? mine : 0.f);
          }
          if (p.sim_kernel == DC_SIM_MEAN) red = red / (float)(n_q - 1);
          if (nan_seen) red = __int_as_float(0x7fc00000);
          weight = pos - red;
          if (weight == weight) weight = fmaxf(weight, 1e-6f);
        }
        if (w_scene && lane == 0) w_scene[(int64_t)v * n_pts + i] = weight;
#pragma unroll
        for (int k = 0; k < kChunks; ++k) {  // feat2d[ys,xs] * metric, then +=  (two roundings)
          acc[k].x += f[k].x * weight;
          acc[k].y += f[k].y * weight;
          acc[k].z += f[k].z * weight;
          acc[k].w += f[k].w * weight;
        }
      } else {
#pragma unroll
        for (int k = 0; k < kChunks; ++k) {
          acc[k].x += f[k].x;
          acc[k].y += f[k].y;
          acc[k].z += f[k].z;
          acc[k].w += f[k].w;
        }
      }
    }
    float4* dst = reinterpret_cast<float4*>(p.out_sum + (p0 + i) * kDim) + lane;
#pragma unroll
    for (int k = 0; k < kChunks; ++k) dst[k * 32] = acc[k];
  }
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

unsigned blocks_for(int64_t max_points, int n_scenes) {
  int64_t want = dc::ceil_div<int64_t>(max_points, kWarps);
  int64_t cap = dc::ceil_div<int64_t>((int64_t)dc::sm_count() * 16, n_scenes);
  if (want > cap) want = cap;
  return (unsigned)(want < 1 ? 1 : want);
}

}  // namespace

extern "C" {

int dc_pixel_fuse(const double* points, const int64_t* point_off, const int64_t* view_off, const float* inv_poses,
                  const double* intrinsics, const int64_t* mask_off, const uint8_t* visible, const void* seg, int seg_dtype,
                  const float* patch_feats, int patch_h, int patch_w, int dim, const float* queries,
                  const int64_t* query_off, int sim_kernel, int norm_feat, int n_scenes, int64_t max_points_per_scene,
                  int max_views_per_scene, int height, int width, const int64_t* perm, float* out_sum, float* out_weight,
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
              patch_w, dim, queries, query_off, sim_kernel, norm_feat, height, width, perm, out_sum, out_weight};
  dim3 grid(blocks_for(max_points_per_scene, n_scenes), (unsigned)n_scenes);
  const bool aligned = (((uintptr_t)patch_feats | (uintptr_t)queries | (uintptr_t)out_sum) & 15) == 0;
  cudaStream_t st = dc::as_stream(stream);
  if (aligned && dim == 768) pixel_fuse_vec_kernel<6><<<grid, kThreads, smem, st>>>(p);       // CLIP ViT-L/14
  else if (aligned && dim == 512) pixel_fuse_vec_kernel<4><<<grid, kThreads, smem, st>>>(p);  // CLIP ViT-B
  else if (aligned && dim == 1024) pixel_fuse_vec_kernel<8><<<grid, kThreads, smem, st>>>(p);
  else pixel_fuse_kernel<<<grid, kThreads, smem, st>>>(p);
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

}  // extern "C"
