// (4) Object-level segmented weighted mean, scatter back to points, and the stream compaction
// of never-visible points.
//
// Reference: einsum("kvc,kv->kc") / sum_v w  utils/feature_fusion.py:333-335; reconstruct_per_obj_feat
// :127-136 (CPU np.argwhere per object) and feat[label] data/dataset_blender.py:128-130; the boolean
// compactions :277-281, :257-264.
#include "common.cuh"

namespace {

// ------------------------------------------------------------------ segmented weighted mean
// Segment = (scene, object); its members are at most one feature row per view. One CTA per
// segment: each of the 8 warps walks a strided subset of the views with its slice of the feature
// vector in registers (coalesced 128-bit row reads), the 8 partial vectors are combined through
// shared memory in a fixed order (deterministic), sum_v w by warp shuffle.
constexpr int kWmThreads = 256;
constexpr int kWmWarps = kWmThreads / 32;

template <typename T>
__device__ __forceinline__ void load4(const T* p, float (&o)[4]);
template <>
__device__ __forceinline__ void load4<float>(const float* p, float (&o)[4]) {
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
}
template <>
__device__ __forceinline__ void load4<__half>(const __half* p, float (&o)[4]) {
  const uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
  const __half2 a = *reinterpret_cast<const __half2*>(&raw.x), b = *reinterpret_cast<const __half2*>(&raw.y);
  o[0] = __low2float(a); o[1] = __high2float(a); o[2] = __low2float(b); o[3] = __high2float(b);
}

// dim <= 4 * 32 * kMaxChunks
constexpr int kMaxChunks = 8;  // up to dim = 1024

template <typename T>
__global__ void __launch_bounds__(kWmThreads) segmented_wmean_kernel(
    const T* __restrict__ feats, int dim, const int32_t* __restrict__ object_row, const float* __restrict__ weight_obj,
    const int64_t* __restrict__ view_off, const int64_t* __restrict__ query_off, const int64_t* __restrict__ wobj_off,
    float* __restrict__ fused) {
  extern __shared__ float s_part[];  // [kWmWarps][dim] then [kWmWarps] weight sums
  const int scene = blockIdx.y;
  const int obj = blockIdx.x;
  const int n_q = (int)(query_off[scene + 1] - query_off[scene]);
  if (obj >= n_q) return;
  const int n_v = (int)(view_off[scene + 1] - view_off[scene]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t* rows = object_row + wobj_off[scene] + (int64_t)obj * n_v;
  const float* w = weight_obj + wobj_off[scene] + (int64_t)obj * n_v;
  const int chunks = dim / 128;  // host guarantees dim % 128 == 0 and chunks <= kMaxChunks
  float acc[kMaxChunks][4];
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
  float wsum = 0.f;
  for (int v = warp; v < n_v; v += kWmWarps) {
    const int32_t r = __ldg(rows + v);
    const float wv = __ldg(w + v);
    wsum += wv;  // every lane holds the same partial
    if (r < 0) continue;  // absent view: the reference multiplies a zero row by weight 0
    const T* src = feats + (int64_t)r * dim;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      if (c < chunks) {
        float x[4];
        load4<T>(src + c * 128 + lane * 4, x);
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[c][i] = fmaf(wv, x[i], acc[c][i]);
      }
    }
  }
  float* s_w = s_part + kWmWarps * dim;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c)
    if (c < chunks)
      *reinterpret_cast<float4*>(s_part + warp * dim + c * 128 + lane * 4) = make_float4(acc[c][0], acc[c][1], acc[c][2], acc[c][3]);
  if (lane == 0) s_w[warp] = wsum;
  __syncthreads();
  float total_w = 0.f;
#pragma unroll
  for (int k = 0; k < kWmWarps; ++k) total_w += s_w[k];
  float* out = fused + (query_off[scene] + obj) * dim;
  for (int c = threadIdx.x; c < dim; c += kWmThreads) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kWmWarps; ++k) t += s_part[k * dim + c];
    out[c] = t / total_w;  // 0/0 = NaN for objects seen in no view (quirk q10)
  }
}

// Same arithmetic per channel (views in the same order inside a warp, the 8 partial vectors combined in the
// same order), organised for memory-level parallelism: 16-byte loads (8 fp16 or 4 fp32 per lane), the row
// indices and weights of kWmUnroll views fetched first and all their feature loads issued before the first
// accumulation. The plain kernel above chains two dependent global loads per view with a single 8-byte load
// per lane in flight (measured 2.6 TB/s = 39 % of HBM on the bench workload).
constexpr int kWmUnroll = 4;

template <typename T, int kChunks>  // dim = kChunks * 32 * (16 / sizeof(T))
__global__ void __launch_bounds__(kWmThreads) segmented_wmean_vec_kernel(
    const T* __restrict__ feats, const int32_t* __restrict__ object_row, const float* __restrict__ weight_obj,
    const int64_t* __restrict__ view_off, const int64_t* __restrict__ query_off, const int64_t* __restrict__ wobj_off,
    float* __restrict__ fused) {
  constexpr int kPer = 16 / (int)sizeof(T);
  constexpr int kDim = kChunks * 32 * kPer;
  extern __shared__ float s_part[];  // [kWmWarps][kDim] then [kWmWarps] weight sums
  const int scene = blockIdx.y;
  const int obj = blockIdx.x;
  const int n_q = (int)(query_off[scene + 1] - query_off[scene]);
  if (obj >= n_q) return;
  const int n_v = (int)(view_off[scene + 1] - view_off[scene]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int32_t* rows = object_row + wobj_off[scene] + (int64_t)obj * n_v;
  const float* w = weight_obj + wobj_off[scene] + (int64_t)obj * n_v;
  float acc[kChunks][kPer];
#pragma unroll
  for (int c = 0; c < kChunks; ++c)
#pragma unroll
    for (int i = 0; i < kPer; ++i) acc[c][i] = 0.f;
  float wsum = 0.f;
  for (int v0 = warp; v0 < n_v; v0 += kWmWarps * kWmUnroll) {
    int32_t r[kWmUnroll];
    float wv[kWmUnroll];
#pragma unroll
    for (int u = 0; u < kWmUnroll; ++u) {
      const int v = v0 + u * kWmWarps;
      r[u] = (v < n_v) ? __ldg(rows + v) : -1;
      wv[u] = (v < n_v) ? __ldg(w + v) : 0.f;
    }
    int4 raw[kWmUnroll][kChunks];
#pragma unroll
    for (int u = 0; u < kWmUnroll; ++u)
#pragma unroll
      for (int c = 0; c < kChunks; ++c)
        raw[u][c] = (r[u] >= 0) ? dc::ld_stream(reinterpret_cast<const int4*>(feats + (int64_t)r[u] * kDim) + c * 32 + lane)
                                : make_int4(0, 0, 0, 0);
#pragma unroll
    for (int u = 0; u < kWmUnroll; ++u) {
      wsum += wv[u];  // every lane holds the same partial; absent views carry weight 0 in the reference too
      if (r[u] < 0) continue;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        if (sizeof(T) == 2) {
          const __half2* h2 = reinterpret_cast<const __half2*>(&raw[u][c]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h2[j]);
            acc[c][2 * j] = fmaf(wv[u], f.x, acc[c][2 * j]);
            acc[c][2 * j + 1] = fmaf(wv[u], f.y, acc[c][2 * j + 1]);
          }
        } else {
          const float* f = reinterpret_cast<const float*>(&raw[u][c]);
#pragma unroll
          for (int j = 0; j < kPer; ++j) acc[c][j] = fmaf(wv[u], f[j], acc[c][j]);
        }
      }
    }
  }
  float* s_w = s_part + kWmWarps * kDim;
#pragma unroll
  for (int c = 0; c < kChunks; ++c)
#pragma unroll
    for (int i = 0; i < kPer; i += 4)
      *reinterpret_cast<float4*>(s_part + warp * kDim + (c * 32 + lane) * kPer + i) =
          make_float4(acc[c][i], acc[c][i + 1], acc[c][i + 2], acc[c][i + 3]);
  if (lane == 0) s_w[warp] = wsum;
  __syncthreads();
  float total_w = 0.f;
#pragma unroll
  for (int k = 0; k < kWmWarps; ++k) total_w += s_w[k];
  float* out = fused + (query_off[scene] + obj) * kDim;
  for (int c = threadIdx.x; c < kDim; c += kWmThreads) {
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kWmWarps; ++k) t += s_part[k * kDim + c];
    out[c] = t / total_w;  // 0/0 = NaN for objects seen in no view (quirk q10)
  }
}

// ------------------------------------------------------------------ scatter to points
// One warp per point row, 128-bit stores; the (Q x dim) source table stays in L1/L2.
__global__ void __launch_bounds__(256) scatter_to_points_kernel(
    const float* __restrict__ fused, const int64_t* __restrict__ query_off, const int64_t* __restrict__ labels,
    const int64_t* __restrict__ point_off, int dim, int skip_first, float* __restrict__ out) {
  const int scene = blockIdx.y;
  const int64_t p0 = point_off[scene];
  const int64_t n = point_off[scene + 1] - p0;
  const int n_q = (int)(query_off[scene + 1] - query_off[scene]);
  const int lane = threadIdx.x & 31;
  const int64_t warps_per_grid = (int64_t)gridDim.x * 8;
  const int vec = dim / 4;
  for (int64_t i = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5); i < n; i += warps_per_grid) {
    const int64_t lab = __ldg(labels + p0 + i);
    const bool hit = lab >= (skip_first ? 1 : 0) && lab < n_q;
    const float4* src = reinterpret_cast<const float4*>(fused + (query_off[scene] + (hit ? lab : 0)) * dim);
    int4* dst = reinterpret_cast<int4*>(out + (p0 + i) * dim);
    for (int c = lane; c < vec; c += 32) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (hit) v = __ldg(src + c);
      dc::st_stream(dst + c, *reinterpret_cast<int4*>(&v));
    }
  }
}

// ------------------------------------------------------------------ compaction
// Three-step exclusive scan of the 0/1 flags: per-block counts, scan of the block counts by one
// block, then per-element ranks.
constexpr int kScanThreads = 256;
constexpr int kScanItems = 16;  // flags per thread
constexpr int kScanBlock = kScanThreads * kScanItems;

__global__ void __launch_bounds__(kScanThreads) scan_count_kernel(const uint8_t* __restrict__ flags, int64_t n,
                                                                  int64_t* __restrict__ block_sums) {
  __shared__ int s_warp[kScanThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanItems;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) c += flags[base + k] != 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < kScanThreads / 32; ++w) t += s_warp[w];
    block_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) scan_block_sums_kernel(int64_t* __restrict__ block_sums, int64_t n_blocks) {
  // single CTA, serial over chunks of 1024 with a running carry (n_blocks is ~ n / 4096)
  __shared__ int64_t s[1024];
  int64_t carry = 0;
  for (int64_t c0 = 0; c0 < n_blocks; c0 += 1024) {
    const int64_t i = c0 + threadIdx.x;
    const int64_t v = i < n_blocks ? block_sums[i] : 0;
    s[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int64_t add = threadIdx.x >= o ? s[threadIdx.x - o] : 0;
      __syncthreads();
      s[threadIdx.x] += add;
      __syncthreads();
    }
    if (i < n_blocks) block_sums[i] = carry + s[threadIdx.x] - v;  // exclusive
    const int64_t chunk_total = s[1023];
    __syncthreads();
    carry += chunk_total;
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_rank_kernel(const uint8_t* __restrict__ flags, int64_t n,
                                                                 const int64_t* __restrict__ block_sums,
                                                                 int64_t* __restrict__ new_index) {
  __shared__ int s_warp[kScanThreads / 32];
  __shared__ int s_rank[kScanBlock];  // block-local exclusive ranks, staged so that the 8-byte stores are coalesced
  const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanItems;
  int f[kScanItems];
  int c = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    f[k] = (base + k < n) ? (flags[base + k] != 0) : 0;
    c += f[k];
  }
  // exclusive scan of c over the block
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int warp_base = 0;
  for (int w = 0; w < warp; ++w) warp_base += s_warp[w];
  int run = warp_base + (incl - c);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    s_rank[threadIdx.x * kScanItems + k] = run;
    run += f[k];
  }
  __syncthreads();
  const int64_t block_base = block_sums[blockIdx.x];
  const int64_t first = (int64_t)blockIdx.x * kScanBlock;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    const int j = k * kScanThreads + threadIdx.x;  // consecutive lanes -> consecutive elements
    if (first + j < n) new_index[first + j] = block_base + s_rank[j];
  }
}

__global__ void kept_offsets_kernel(const uint8_t* __restrict__ flags, const int64_t* __restrict__ new_index,
                                    const int64_t* __restrict__ point_off, int n_scenes, int64_t total,
                                    int64_t* __restrict__ kept_off) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n_scenes) return;
  const int64_t j = point_off[s];
  kept_off[s] = (j < total) ? new_index[j] : (total > 0 ? new_index[total - 1] + (flags[total - 1] != 0) : 0);
}

// out_off[s] = sum over earlier scenes of views * kept points: the layout of the compacted (V_s, N'_s) blocks
__global__ void mask_offsets_kernel(const int64_t* __restrict__ kept_off, const int64_t* __restrict__ view_off, int n_scenes,
                                    int64_t* __restrict__ out_off) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  int64_t run = 0;
  for (int s = 0; s < n_scenes; ++s) {
    out_off[s] = run;
    run += (view_off[s + 1] - view_off[s]) * (kept_off[s + 1] - kept_off[s]);
  }
  out_off[n_scenes] = run;
}

// rows of row_bytes: one thread per unit (4-byte words when the row width and both bases allow it, bytes otherwise, so
// uint8 colours / labels and fp16 points compact on the device too)
template <typename U>
__global__ void __launch_bounds__(256) compact_rows_kernel(const U* __restrict__ in, int units_per_row,
                                                           const uint8_t* __restrict__ flags,
                                                           const int64_t* __restrict__ new_index, int64_t n,
                                                           U* __restrict__ out) {
  const int64_t total = n * units_per_row;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = t / units_per_row;
    if (flags[j]) out[new_index[j] * units_per_row + (t - j * units_per_row)] = in[t];
  }
}

// Point-major: a thread owns one kept point and walks the views, so the flag and the rank are
// read once per point instead of once per (view, point); reads and writes stay coalesced along
// the point axis because ranks are monotone in the point index.
template <typename T, typename TOut = T>
__global__ void __launch_bounds__(256) compact_mask_kernel(const T* __restrict__ mask, const int64_t* __restrict__ mask_off,
                                                           const int64_t* __restrict__ point_off,
                                                           const int64_t* __restrict__ view_off,
                                                           const uint8_t* __restrict__ flags,
                                                           const int64_t* __restrict__ new_index,
                                                           const int64_t* __restrict__ kept_off,
                                                           const int64_t* __restrict__ out_off, TOut* __restrict__ out) {
  const int scene = blockIdx.y;
  const int64_t p0 = point_off[scene];
  const int64_t n = point_off[scene + 1] - p0;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flags[p0 + i]) return;
  const int n_v = (int)(view_off[scene + 1] - view_off[scene]);
  const int64_t kept0 = kept_off[scene];
  const int64_t n_kept = kept_off[scene + 1] - kept0;
  const T* src = mask + mask_off[scene] + i;
  TOut* dst = out + out_off[scene] + (new_index[p0 + i] - kept0);
#pragma unroll 4
  for (int v = 0; v < n_v; ++v) dst[(int64_t)v * n_kept] = (TOut)src[(int64_t)v * n];
}

}  // namespace

extern "C" {

int dc_segmented_wmean(const void* feats, int feat_dtype, int dim, const int32_t* object_row, const float* weight_obj,
                       const int64_t* view_off, const int64_t* query_off, const int64_t* wobj_off, int n_scenes,
                       int max_queries_per_scene, float* fused, dc_stream_t stream) {
  DC_CHECK_ARG(feats && object_row && weight_obj && view_off && query_off && wobj_off && fused,
               "dc_segmented_wmean: null pointer argument");
  DC_CHECK_ARG(feat_dtype == DC_F16 || feat_dtype == DC_F32, "dc_segmented_wmean: features must be fp16 or fp32");
  DC_CHECK_ARG(dim > 0 && dim % 128 == 0 && dim <= 128 * kMaxChunks, "dc_segmented_wmean: dim must be a multiple of 128, <= %d",
               128 * kMaxChunks);
  if (n_scenes <= 0 || max_queries_per_scene <= 0) return DC_OK;
  DC_CHECK_ARG(n_scenes <= 65535, "dc_segmented_wmean: at most 65535 scenes per call");
  dim3 grid((unsigned)max_queries_per_scene, (unsigned)n_scenes);
  const size_t smem = sizeof(float) * ((size_t)kWmWarps * dim + kWmWarps);
  cudaStream_t st = dc::as_stream(stream);
  if (dim == 768 && ((uintptr_t)feats & 15) == 0) {  // CLIP ViT-L/14 width: wide loads, unrolled over views
    if (feat_dtype == DC_F16)
      segmented_wmean_vec_kernel<__half, 3><<<grid, kWmThreads, smem, st>>>((const __half*)feats, object_row, weight_obj, view_off,
                                                                           query_off, wobj_off, fused);
    else
      segmented_wmean_vec_kernel<float, 6><<<grid, kWmThreads, smem, st>>>((const float*)feats, object_row, weight_obj, view_off,
                                                                          query_off, wobj_off, fused);
    DC_LAUNCH_CHECK();
    return DC_OK;
  }
  if (feat_dtype == DC_F16)
    segmented_wmean_kernel<__half><<<grid, kWmThreads, smem, st>>>((const __half*)feats, dim, object_row, weight_obj, view_off,
                                                                  query_off, wobj_off, fused);
  else
    segmented_wmean_kernel<float><<<grid, kWmThreads, smem, st>>>((const float*)feats, dim, object_row, weight_obj, view_off,
                                                                 query_off, wobj_off, fused);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_scatter_to_points(const float* fused, const int64_t* query_off, const int64_t* labels, const int64_t* point_off,
                         int n_scenes, int64_t max_points_per_scene, int dim, int skip_first, float* out,
                         dc_stream_t stream) {
  DC_CHECK_ARG(fused && query_off && labels && point_off && out, "dc_scatter_to_points: null pointer argument");
  DC_CHECK_ARG(dim > 0 && dim % 4 == 0, "dc_scatter_to_points: dim must be a multiple of 4");
  DC_CHECK_ARG(((uintptr_t)out & 15) == 0 && ((uintptr_t)fused & 15) == 0, "dc_scatter_to_points: 16-byte alignment required");
  if (n_scenes <= 0 || max_points_per_scene <= 0) return DC_OK;
  DC_CHECK_ARG(n_scenes <= 65535, "dc_scatter_to_points: at most 65535 scenes per call");
  int64_t want = dc::ceil_div<int64_t>(max_points_per_scene, 8 * 4);  // ~4 rows per warp
  int64_t cap = dc::ceil_div<int64_t>((int64_t)dc::sm_count() * 16, n_scenes);
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  dim3 grid((unsigned)want, (unsigned)n_scenes);
  scatter_to_points_kernel<<<grid, 256, 0, dc::as_stream(stream)>>>(fused, query_off, labels, point_off, dim, skip_first, out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

size_t dc_compact_workspace(int64_t total_points) {
  return sizeof(int64_t) * (size_t)(dc::ceil_div<int64_t>(total_points > 0 ? total_points : 1, kScanBlock) + 1);
}

int dc_compact_scan(const uint8_t* any_visible, int64_t total_points, const int64_t* point_off, int n_scenes,
                    int64_t* new_index, int64_t* kept_off, void* workspace, size_t workspace_bytes, dc_stream_t stream) {
  DC_CHECK_ARG(any_visible && point_off && new_index && kept_off && workspace, "dc_compact_scan: null pointer argument");
  if (workspace_bytes < dc_compact_workspace(total_points))
    return dc::fail(DC_ERR_WORKSPACE, "dc_compact_scan: workspace too small");
  cudaStream_t st = dc::as_stream(stream);
  if (total_points <= 0) {
    DC_CUDA(cudaMemsetAsync(kept_off, 0, sizeof(int64_t) * (size_t)(n_scenes + 1), st));
    return DC_OK;
  }
  int64_t* block_sums = reinterpret_cast<int64_t*>(workspace);
  const int64_t n_blocks = dc::ceil_div<int64_t>(total_points, kScanBlock);
  scan_count_kernel<<<(unsigned)n_blocks, kScanThreads, 0, st>>>(any_visible, total_points, block_sums);
  scan_block_sums_kernel<<<1, 1024, 0, st>>>(block_sums, n_blocks);
  scan_rank_kernel<<<(unsigned)n_blocks, kScanThreads, 0, st>>>(any_visible, total_points, block_sums, new_index);
  kept_offsets_kernel<<<dc::ceil_div(n_scenes + 1, 128), 128, 0, st>>>(any_visible, new_index, point_off, n_scenes, total_points, kept_off);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_compact_mask_offsets(const int64_t* kept_off, const int64_t* view_off, int n_scenes, int64_t* out_off, dc_stream_t stream) {
  DC_CHECK_ARG(kept_off && view_off && out_off && n_scenes >= 0, "dc_compact_mask_offsets: bad argument");
  mask_offsets_kernel<<<1, 32, 0, dc::as_stream(stream)>>>(kept_off, view_off, n_scenes, out_off);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_compact_rows(const void* in, int64_t row_bytes, const uint8_t* any_visible, const int64_t* new_index,
                    int64_t total_points, void* out, dc_stream_t stream) {
  DC_CHECK_ARG(in && any_visible && new_index && out, "dc_compact_rows: null pointer argument");
  DC_CHECK_ARG(row_bytes > 0 && row_bytes <= (1 << 30), "dc_compact_rows: row_bytes must be positive");
  if (total_points <= 0) return DC_OK;
  const bool words = row_bytes % 4 == 0 && (((uintptr_t)in | (uintptr_t)out) & 3) == 0;
  const int units = (int)(words ? row_bytes / 4 : row_bytes);
  const int64_t total = total_points * units;
  int64_t blocks = dc::ceil_div<int64_t>(total, 256);
  const int64_t cap = (int64_t)dc::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (words)
    compact_rows_kernel<uint32_t><<<(unsigned)blocks, 256, 0, dc::as_stream(stream)>>>((const uint32_t*)in, units, any_visible,
                                                                                      new_index, total_points, (uint32_t*)out);
  else
    compact_rows_kernel<uint8_t><<<(unsigned)blocks, 256, 0, dc::as_stream(stream)>>>((const uint8_t*)in, units, any_visible,
                                                                                     new_index, total_points, (uint8_t*)out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_compact_mask(const void* mask, int elem_size, const int64_t* mask_off, const int64_t* point_off,
                    const int64_t* view_off, const uint8_t* any_visible, const int64_t* new_index, const int64_t* kept_off,
                    const int64_t* out_off, int n_scenes, int64_t max_points_per_scene, int max_views_per_scene, void* out,
                    dc_stream_t stream) {
  DC_CHECK_ARG(mask && mask_off && point_off && view_off && any_visible && new_index && kept_off && out_off && out,
               "dc_compact_mask: null pointer argument");
  DC_CHECK_ARG(elem_size == 1 || elem_size == 4 || elem_size == 8 || elem_size == 18,
               "dc_compact_mask: elem_size must be 1, 4, 8 or 18 (uint8 in, int64 out)");
  if (n_scenes <= 0 || max_points_per_scene <= 0 || max_views_per_scene <= 0) return DC_OK;
  DC_CHECK_ARG(n_scenes <= 65535, "dc_compact_mask: at most 65535 scenes per call");
  (void)max_views_per_scene;
  dim3 grid((unsigned)dc::ceil_div<int64_t>(max_points_per_scene, 256), (unsigned)n_scenes);
  cudaStream_t st = dc::as_stream(stream);
  if (elem_size == 18)
    compact_mask_kernel<uint8_t, long long><<<grid, 256, 0, st>>>((const uint8_t*)mask, mask_off, point_off, view_off, any_visible, new_index, kept_off, out_off, (long long*)out);
  else if (elem_size == 1)
    compact_mask_kernel<uint8_t><<<grid, 256, 0, st>>>((const uint8_t*)mask, mask_off, point_off, view_off, any_visible, new_index, kept_off, out_off, (uint8_t*)out);
  else if (elem_size == 4)
    compact_mask_kernel<uint32_t><<<grid, 256, 0, st>>>((const uint32_t*)mask, mask_off, point_off, view_off, any_visible, new_index, kept_off, out_off, (uint32_t*)out);
  else
    compact_mask_kernel<unsigned long long><<<grid, 256, 0, st>>>((const unsigned long long*)mask, mask_off, point_off, view_off, any_visible, new_index, kept_off, out_off, (unsigned long long*)out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // extern "C"
