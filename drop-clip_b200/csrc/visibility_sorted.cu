// (1) Projection + visibility with spatially sorted points.
//
// Why: the reference's point cloud arrives in Open3D hash order, so the 32 lanes of a warp gather
// 32 unrelated depth pixels = 32 distinct 128-B lines per load. ncu on the direct kernel
// (profiles/r01_ncu_project_visibility_v1_raw.csv) shows DRAM at 6 %, L2 hit rate 84 % and the
// time set by L1TEX line replays (~2 cycles per line). Sorting the points of a scene by a coarse
// Morton cell makes neighbouring lanes project to neighbouring pixels (~8 lines per load).
//
// Pipeline (all kernels scene-batched, results bit-identical to the direct kernel):
//   1. bbox per scene (fp32 is enough: cells only steer locality, never results)
//   2. counting sort by 15-bit Morton cell: histogram -> per-scene scan -> scatter (perm, rank)
//   3. visibility in sorted order; results bit-packed per point (one 32-bit word per 32 views,
//      plane-major so that lanes store contiguous words)
//   4. unpack: a thread per original point reads its words through `rank` and writes the
//      (V,N) mask rows coalesced - or, fused with the removal of never-visible points, writes the
//      compacted (V,N') mask directly (the form fuse_obj_prior returns, utils/feature_fusion.py
//      :277-281), so the full mask never touches HBM.
#include "visibility_math.cuh"

namespace {

using namespace dc::vis;

constexpr int kThreads = 256;
constexpr int kPointsPerThread = 2;
constexpr int kPointsPerBlock = kThreads * kPointsPerThread;
constexpr int kCellBits = 5;                    // per axis
constexpr int kCells = 1 << (3 * kCellBits);    // 32768 bins per scene

struct SortedWs {
  float* bbox;        // [n_scenes][6] min xyz, max xyz
  int* counts;        // [n_scenes][kCells]
  int64_t* perm;      // [total_points] scene-local point index at each sorted position
  size_t total;
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

SortedWs carve(void* ws, int64_t total_points, int n_scenes) {
  size_t off = 0;
  auto take = [&](size_t b) { size_t o = off; off = align_up(off + b, 256); return o; };
  const size_t o_bbox = take(sizeof(float) * 6 * (size_t)n_scenes);
  const size_t o_counts = take(sizeof(int) * (size_t)kCells * (size_t)n_scenes);
  const size_t o_perm = take(sizeof(int64_t) * (size_t)(total_points > 0 ? total_points : 1));
  SortedWs w{};
  w.total = off;
  if (ws) {
    uint8_t* b = reinterpret_cast<uint8_t*>(ws);
    w.bbox = reinterpret_cast<float*>(b + o_bbox);
    w.counts = reinterpret_cast<int*>(b + o_counts);
    w.perm = reinterpret_cast<int64_t*>(b + o_perm);
  }
  return w;
}

__device__ __forceinline__ void atomic_min_f(float* a, float v) {
  if (v >= 0) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* a, float v) {
  if (v >= 0) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned*>(a), __float_as_uint(v));
}

__global__ void init_bbox_kernel(float* bbox, int n_scenes) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_scenes * 6) bbox[i] = (i % 6 < 3) ? INFINITY : -INFINITY;
}

__device__ __forceinline__ float finite_or(double v, float alt) {
  const float f = (float)v;
  return (fabsf(f) < 1e30f) ? f : alt;  // NaN / inf / huge coordinates do not stretch the grid
}

__global__ void __launch_bounds__(kThreads) bbox_kernel(const double* __restrict__ points, const int64_t* __restrict__ point_off,
                                                        float* __restrict__ bbox) {
  const int scene = blockIdx.y;
  const int64_t p0 = point_off[scene], n = point_off[scene + 1] - p0;
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const double v = __ldg(points + 3 * (p0 + i) + a);
      lo[a] = fminf(lo[a], finite_or(v, INFINITY));
      hi[a] = fmaxf(hi[a], finite_or(v, -INFINITY));
    }
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    lo[a] = dc::warp_min(lo[a]);
    hi[a] = dc::warp_max(hi[a]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (lo[a] != INFINITY) atomic_min_f(bbox + scene * 6 + a, lo[a] + 0.0f);
      if (hi[a] != -INFINITY) atomic_max_f(bbox + scene * 6 + 3 + a, hi[a] + 0.0f);
    }
  }
}

__device__ __forceinline__ unsigned spread3(unsigned v) {  // 5 bits -> every third bit
  v &= 0x1f;
  v = (v | (v << 8)) & 0x100f;
  v = (v | (v << 4)) & 0x10c3;
  v = (v | (v << 2)) & 0x1249;
  return v;
}

__device__ __forceinline__ int cell_of(const double* __restrict__ points, int64_t j, const float* __restrict__ bb) {
  const float ext = fmaxf(fmaxf(bb[3] - bb[0], bb[4] - bb[1]), fmaxf(bb[5] - bb[2], 1e-20f));
  const float inv = (float)(1 << kCellBits) / ext;  // isotropic cells
  unsigned c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float f = (finite_or(__ldg(points + 3 * j + a), bb[a]) - bb[a]) * inv;
    int q = (int)f;
    q = q < 0 ? 0 : (q > (1 << kCellBits) - 1 ? (1 << kCellBits) - 1 : q);
    c[a] = (unsigned)q;
  }
  return (int)(spread3(c[0]) | (spread3(c[1]) << 1) | (spread3(c[2]) << 2));
}

__global__ void __launch_bounds__(kThreads) cell_count_kernel(const double* __restrict__ points, const int64_t* __restrict__ point_off,
                                                              const float* __restrict__ bbox, int* __restrict__ counts) {
  const int scene = blockIdx.y;
  const int64_t p0 = point_off[scene], n = point_off[scene + 1] - p0;
  const float* bb = bbox + scene * 6;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads)
    atomicAdd(counts + (int64_t)scene * kCells + cell_of(points, p0 + i, bb), 1);
}

// one CTA per scene: exclusive scan of its kCells counters, in place
__global__ void __launch_bounds__(1024) cell_scan_kernel(int* __restrict__ counts) {
  __shared__ int s[1024];
  int* c = counts + (int64_t)blockIdx.x * kCells;
  constexpr int per = kCells / 1024;
  int loc[per], sum = 0;
#pragma unroll
  for (int k = 0; k < per; ++k) {
    loc[k] = c[threadIdx.x * per + k];
    sum += loc[k];
  }
  s[threadIdx.x] = sum;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int add = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
    __syncthreads();
    s[threadIdx.x] += add;
    __syncthreads();
  }
  int run = s[threadIdx.x] - sum;
#pragma unroll
  for (int k = 0; k < per; ++k) {
    c[threadIdx.x * per + k] = run;
    run += loc[k];
  }
}

__global__ void __launch_bounds__(kThreads) cell_scatter_kernel(const double* __restrict__ points, const int64_t* __restrict__ point_off,
                                                                const float* __restrict__ bbox, int* __restrict__ counts,
                                                                int64_t* __restrict__ perm, int64_t* __restrict__ rank) {
  const int scene = blockIdx.y;
  const int64_t p0 = point_off[scene], n = point_off[scene + 1] - p0;
  const float* bb = bbox + scene * 6;
  for (int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (int64_t)gridDim.x * kThreads) {
    const int pos = atomicAdd(counts + (int64_t)scene * kCells + cell_of(points, p0 + i, bb), 1);
    perm[p0 + pos] = i;   // order inside a cell depends on scheduling; results do not (they are un-permuted)
    rank[p0 + i] = pos;
  }
}

struct SortedParams {
  const double* points;
  const int64_t* point_off;
  const int64_t* view_off;
  const float* depths;
  const float* inv_poses;
  const double* intrinsics;
  const int64_t* perm;
  int64_t total_points;
  int height, width;
  double threshold;
  uint32_t* records;  // [n_words][total_points], indexed by sorted position
  uint8_t* any_visible;
};

__global__ void __launch_bounds__(kThreads, 4) visibility_sorted_kernel(SortedParams p) {
  extern __shared__ double s_cam[];
  __shared__ int s_ok;
  const int scene = blockIdx.y;
  const int64_t p0 = p.point_off[scene];
  const int64_t n_pts = p.point_off[scene + 1] - p0;
  const int64_t tile0 = (int64_t)blockIdx.x * kPointsPerBlock;
  if (tile0 >= n_pts) return;
  const int64_t v0 = p.view_off[scene];
  const int n_views = (int)(p.view_off[scene + 1] - v0);
  const bool pinhole = load_cameras(s_cam, &s_ok, p.inv_poses, v0, n_views, p.intrinsics, scene);
  const double* s_K = s_cam + n_views * 12;
  const double K0 = s_K[0], K2 = s_K[2], K4 = s_K[4], K5 = s_K[5];

  double px[kPointsPerThread], py[kPointsPerThread], pz[kPointsPerThread];
  bool valid[kPointsPerThread], fast[kPointsPerThread];
  int64_t orig[kPointsPerThread];
  uint32_t word[kPointsPerThread], any[kPointsPerThread];
#pragma unroll
  for (int k = 0; k < kPointsPerThread; ++k) {
    const int64_t s = tile0 + k * kThreads + threadIdx.x;
    valid[k] = s < n_pts;
    orig[k] = valid[k] ? p0 + p.perm[p0 + s] : p0;
    px[k] = __ldg(p.points + 3 * orig[k]);
    py[k] = __ldg(p.points + 3 * orig[k] + 1);
    pz[k] = __ldg(p.points + 3 * orig[k] + 2);
    fast[k] = pinhole && fabs(px[k]) < kBig && fabs(py[k]) < kBig && fabs(pz[k]) < kBig;
    word[k] = 0;
    any[k] = 0;
  }
  const int64_t hw = (int64_t)p.height * p.width;
  for (int v = 0; v < n_views; ++v) {
    const double* m = s_cam + v * 12;
    const float* depth = p.depths + (v0 + v) * hw;
    bool inside[kPointsPerThread];
    int pix[kPointsPerThread];
    double qz[kPointsPerThread];
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k)
      inside[k] = project_point(m, s_K, K0, K2, K4, K5, fast[k], px[k], py[k], pz[k], p.width, p.height, pix[k], qz[k]) && valid[k];
    float sensor[kPointsPerThread];
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k) sensor[k] = inside[k] ? __ldg(depth + pix[k]) : 0.f;
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k) {
      const bool vis = inside[k] && (fabs((double)sensor[k] - qz[k]) <= p.threshold);
      word[k] |= (vis ? 1u : 0u) << (v & 31);
    }
    if ((v & 31) == 31 || v == n_views - 1) {
      const int64_t plane = (int64_t)(v >> 5) * p.total_points + p0 + tile0 + threadIdx.x;
#pragma unroll
      for (int k = 0; k < kPointsPerThread; ++k) {
        if (valid[k]) p.records[plane + k * kThreads] = word[k];
        any[k] |= word[k];
        word[k] = 0;
      }
    }
  }
  if (p.any_visible) {
#pragma unroll
    for (int k = 0; k < kPointsPerThread; ++k)
      if (valid[k]) p.any_visible[orig[k]] = any[k] ? 1 : 0;
  }
}

// thread per original point; TOut mask element. kCompact: write only kept points at their rank.
template <typename TOut, bool kCompact>
__global__ void __launch_bounds__(kThreads) unpack_kernel(const uint32_t* __restrict__ records, const int64_t* __restrict__ rank,
                                                          const int64_t* __restrict__ point_off, const int64_t* __restrict__ view_off,
                                                          const int64_t* __restrict__ mask_off, int64_t total_points,
                                                          const uint8_t* __restrict__ any_visible,
                                                          const int64_t* __restrict__ new_index, const int64_t* __restrict__ kept_off,
                                                          TOut* __restrict__ out) {
  const int scene = blockIdx.y;
  const int64_t p0 = point_off[scene];
  const int64_t n = point_off[scene + 1] - p0;
  const int64_t i = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  if (kCompact && !any_visible[p0 + i]) return;
  const int n_v = (int)(view_off[scene + 1] - view_off[scene]);
  const int64_t s = p0 + rank[p0 + i];
  int64_t stride = n, col = i;
  if (kCompact) {
    const int64_t kept0 = kept_off[scene];
    stride = kept_off[scene + 1] - kept0;
    col = new_index[p0 + i] - kept0;
  }
  TOut* dst = out + mask_off[scene] + col;
  for (int w = 0; w * 32 < n_v; ++w) {
    const uint32_t bits = __ldg(records + (int64_t)w * total_points + s);
    const int lim = min(32, n_v - w * 32);
#pragma unroll 8
    for (int b = 0; b < lim; ++b) dst[(int64_t)(w * 32 + b) * stride] = (TOut)((bits >> b) & 1u);
  }
}

}  // namespace

extern "C" {

size_t dc_visibility_sorted_workspace(int64_t total_points, int n_scenes) {
  return carve(nullptr, total_points, n_scenes > 0 ? n_scenes : 1).total;
}

int dc_project_visibility_sorted(const double* points, const int64_t* point_off, const int64_t* view_off, const float* depths,
                                 const float* inv_poses, const double* intrinsics, int n_scenes, int64_t total_points,
                                 int64_t max_points_per_scene, int max_views_per_scene, int height, int width,
                                 double threshold, uint32_t* records, int64_t* rank, uint8_t* any_visible, void* workspace,
                                 size_t workspace_bytes, dc_stream_t stream) {
  DC_CHECK_ARG(points && point_off && view_off && depths && inv_poses && intrinsics && records && rank && workspace,
               "dc_project_visibility_sorted: null pointer argument");
  DC_CHECK_ARG(height > 0 && width > 0 && (int64_t)height * width < (1ll << 31), "dc_project_visibility_sorted: bad image size");
  if (n_scenes <= 0 || max_points_per_scene <= 0 || total_points <= 0) return DC_OK;
  DC_CHECK_ARG(n_scenes <= 65535, "dc_project_visibility_sorted: at most 65535 scenes per call");
  DC_CHECK_ARG(max_points_per_scene < (1ll << 31), "dc_project_visibility_sorted: at most 2^31 points per scene");
  const size_t smem = ((size_t)max_views_per_scene * 12 + 9) * sizeof(double);
  DC_CHECK_ARG(smem <= 48 * 1024, "dc_project_visibility_sorted: too many views per scene (%d)", max_views_per_scene);
  SortedWs w = carve(workspace, total_points, n_scenes);
  if (workspace_bytes < w.total)
    return dc::fail(DC_ERR_WORKSPACE, "dc_project_visibility_sorted: workspace %zu < %zu", workspace_bytes, w.total);
  cudaStream_t st = dc::as_stream(stream);
  init_bbox_kernel<<<dc::ceil_div(n_scenes * 6, 128), 128, 0, st>>>(w.bbox, n_scenes);
  DC_CUDA(cudaMemsetAsync(w.counts, 0, sizeof(int) * (size_t)kCells * n_scenes, st));
  int64_t chunks = dc::ceil_div<int64_t>(max_points_per_scene, kThreads * 4);
  const int64_t cap = dc::ceil_div<int64_t>((int64_t)dc::sm_count() * 8, n_scenes);
  if (chunks > cap) chunks = cap;
  if (chunks < 1) chunks = 1;
  dim3 g1((unsigned)chunks, (unsigned)n_scenes);
  bbox_kernel<<<g1, kThreads, 0, st>>>(points, point_off, w.bbox);
  cell_count_kernel<<<g1, kThreads, 0, st>>>(points, point_off, w.bbox, w.counts);
  cell_scan_kernel<<<(unsigned)n_scenes, 1024, 0, st>>>(w.counts);
  cell_scatter_kernel<<<g1, kThreads, 0, st>>>(points, point_off, w.bbox, w.counts, w.perm, rank);
  SortedParams p{points, point_off, view_off, depths, inv_poses, intrinsics, w.perm, total_points, height, width, threshold,
                 records, any_visible};
  dim3 grid((unsigned)dc::ceil_div<int64_t>(max_points_per_scene, kPointsPerBlock), (unsigned)n_scenes);
  visibility_sorted_kernel<<<grid, kThreads, smem, st>>>(p);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_unpack_visibility(const uint32_t* records, const int64_t* rank, const int64_t* point_off, const int64_t* view_off,
                         const int64_t* mask_off, int n_scenes, int64_t total_points, int64_t max_points_per_scene, void* mask,
                         int mask_elem_size, dc_stream_t stream) {
  DC_CHECK_ARG(records && rank && point_off && view_off && mask_off && mask, "dc_unpack_visibility: null pointer argument");
  DC_CHECK_ARG(mask_elem_size == 1 || mask_elem_size == 8, "dc_unpack_visibility: mask_elem_size must be 1 or 8");
  if (n_scenes <= 0 || max_points_per_scene <= 0) return DC_OK;
  dim3 grid((unsigned)dc::ceil_div<int64_t>(max_points_per_scene, kThreads), (unsigned)n_scenes);
  cudaStream_t st = dc::as_stream(stream);
  if (mask_elem_size == 1)
    unpack_kernel<uint8_t, false><<<grid, kThreads, 0, st>>>(records, rank, point_off, view_off, mask_off, total_points, nullptr,
                                                             nullptr, nullptr, (uint8_t*)mask);
  else
    unpack_kernel<long long, false><<<grid, kThreads, 0, st>>>(records, rank, point_off, view_off, mask_off, total_points, nullptr,
                                                               nullptr, nullptr, (long long*)mask);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_unpack_visibility_compact(const uint32_t* records, const int64_t* rank, const int64_t* point_off, const int64_t* view_off,
                                 const uint8_t* any_visible, const int64_t* new_index, const int64_t* kept_off,
                                 const int64_t* out_off, int n_scenes, int64_t total_points, int64_t max_points_per_scene,
                                 void* out, int out_elem_size, dc_stream_t stream) {
  DC_CHECK_ARG(records && rank && point_off && view_off && any_visible && new_index && kept_off && out_off && out,
               "dc_unpack_visibility_compact: null pointer argument");
  DC_CHECK_ARG(out_elem_size == 1 || out_elem_size == 8, "dc_unpack_visibility_compact: out_elem_size must be 1 or 8");
  if (n_scenes <= 0 || max_points_per_scene <= 0) return DC_OK;
  dim3 grid((unsigned)dc::ceil_div<int64_t>(max_points_per_scene, kThreads), (unsigned)n_scenes);
  cudaStream_t st = dc::as_stream(stream);
  if (out_elem_size == 1)
    unpack_kernel<uint8_t, true><<<grid, kThreads, 0, st>>>(records, rank, point_off, view_off, out_off, total_points, any_visible,
                                                            new_index, kept_off, (uint8_t*)out);
  else
    unpack_kernel<long long, true><<<grid, kThreads, 0, st>>>(records, rank, point_off, view_off, out_off, total_points, any_visible,
                                                              new_index, kept_off, (long long*)out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // extern "C"
