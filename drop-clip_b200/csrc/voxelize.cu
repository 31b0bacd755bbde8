// (5) Voxelisation to MinkowskiEngine-style sparse coordinates for a batch of samples.
//
// Reference: ME.utils.sparse_quantize(coordinates, features, labels, ignore_label, return_index,
// return_inverse, quantization_size) as called at data/dataset_blender.py:406-414 and
// data/dataset.py:164-172 (third-party, un-vendored; semantics restated in
// oracle/projections_ref.py::sparse_quantize_ref):
//   q = floor(xyz / size) (fp32 true division) -> int32; unique voxels with the index of their
//   first occurrence; inverse map; a voxel whose points disagree on the label gets ignore_label.
//
// ME inserts sequentially into a CPU hash map. The GPU formulation keeps the hash map but makes
// every step order-independent so the result is deterministic and equals the sequential one:
//   1. insert: 63-bit packed key, atomicCAS claim + linear probing inside the sample's own table
//      region; atomicMin keeps the smallest point index per voxel (= first occurrence); labels
//      are merged with a CAS that degrades to a CONFLICT marker;
//   2. a point is a "voxel head" iff it is its voxel's first occurrence; an exclusive scan of the
//      head flags yields the voxel rank in first-occurrence order (ME's canonical order);
//   3. emit coords / unique_map / labels at the heads, inverse_map at every point.
// A sort-based unique would need 96-bit keys and a second sort to restore first-occurrence order;
// the hash form reads every point twice and writes each output once (HBM-bound, ~60 B/point).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr unsigned long long kEmpty = ~0ull;
constexpr int kNoLabel = INT32_MIN;
constexpr int kConflict = INT32_MIN + 1;
constexpr int kCoordBias = 1 << 20;  // coordinates must lie in [-2^20, 2^20)

struct VoxWorkspace {
  unsigned long long* keys;  // [2 * total]
  int* first;                // [2 * total]  smallest sample-local point index of the voxel
  int* label;                // [2 * total]
  int* slot_of;              // [total]
  uint8_t* head;             // [total]
  int64_t* rank;             // [total]
  int64_t* block_sums;       // scan scratch
  int* error;                // [1]
};

__device__ __forceinline__ int sample_of(const int64_t* __restrict__ off, int n_samples, int64_t j) {
  int lo = 0, hi = n_samples - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= j) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

__global__ void __launch_bounds__(kThreads) vox_insert_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ sample_off,
                                                              int n_samples, int64_t total, float voxel_size,
                                                              const int32_t* __restrict__ labels, VoxWorkspace w) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= total) return;
  const int b = sample_of(sample_off, n_samples, j);
  const int64_t s0 = sample_off[b];
  const int64_t cap = 2 * (sample_off[b + 1] - s0);
  // fp32 true division then floor, like torch's `coordinates / quantization_size` followed by floor().int()
  const int qx = (int)floorf(__fdiv_rn(xyz[3 * j], voxel_size));
  const int qy = (int)floorf(__fdiv_rn(xyz[3 * j + 1], voxel_size));
  const int qz = (int)floorf(__fdiv_rn(xyz[3 * j + 2], voxel_size));
  if (qx < -kCoordBias || qx >= kCoordBias || qy < -kCoordBias || qy >= kCoordBias || qz < -kCoordBias || qz >= kCoordBias) {
    atomicExch(w.error, 1);
    w.slot_of[j] = -1;
    return;
  }
  const unsigned long long key = ((unsigned long long)(qx + kCoordBias) << 42) | ((unsigned long long)(qy + kCoordBias) << 21) |
                                 (unsigned long long)(qz + kCoordBias);
  int64_t slot = (int64_t)(mix64(key) % (uint64_t)cap);
  unsigned long long* keys = w.keys + 2 * s0;
  while (true) {
    const unsigned long long old = atomicCAS(keys + slot, kEmpty, key);
    if (old == kEmpty || old == key) break;
    if (++slot == cap) slot = 0;
  }
  const int64_t g = 2 * s0 + slot;
  atomicMin(w.first + g, (int)(j - s0));
  if (labels) {
    const int lab = labels[j];
    const int old = atomicCAS(w.label + g, kNoLabel, lab);
    if (old != kNoLabel && old != lab) atomicExch(w.label + g, kConflict);
  }
  w.slot_of[j] = (int)slot;
}

__global__ void __launch_bounds__(kThreads) vox_heads_kernel(const int64_t* __restrict__ sample_off, int n_samples,
                                                             int64_t total, VoxWorkspace w) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= total) return;
  const int b = sample_of(sample_off, n_samples, j);
  const int64_t s0 = sample_off[b];
  const int slot = w.slot_of[j];
  w.head[j] = (slot >= 0 && w.first[2 * s0 + slot] == (int)(j - s0)) ? 1 : 0;
}

__global__ void __launch_bounds__(kThreads) vox_emit_kernel(const float* __restrict__ xyz, const int64_t* __restrict__ sample_off,
                                                            int n_samples, int64_t total, float voxel_size, int has_labels,
                                                            int32_t ignore_label, VoxWorkspace w, int32_t* __restrict__ coords,
                                                            int64_t* __restrict__ unique_map, int64_t* __restrict__ inverse_map,
                                                            int32_t* __restrict__ voxel_labels) {
  const int64_t j = (int64_t)blockIdx.x * kThreads + threadIdx.x;
  if (j >= total) return;
  const int b = sample_of(sample_off, n_samples, j);
  const int64_t s0 = sample_off[b];
  const int slot = w.slot_of[j];
  if (slot < 0) { inverse_map[j] = -1; return; }
  const int64_t g = 2 * s0 + slot;
  const int64_t base = w.rank[s0];  // heads before this sample
  const int64_t head_j = s0 + w.first[g];
  const int64_t vox = w.rank[head_j] - base;
  inverse_map[j] = vox;
  if (head_j == j) {
    const int64_t o = s0 + vox;
    coords[3 * o] = (int)floorf(__fdiv_rn(xyz[3 * j], voxel_size));
    coords[3 * o + 1] = (int)floorf(__fdiv_rn(xyz[3 * j + 1], voxel_size));
    coords[3 * o + 2] = (int)floorf(__fdiv_rn(xyz[3 * j + 2], voxel_size));
    unique_map[o] = j - s0;
    if (has_labels) {
      const int lab = w.label[g];
      voxel_labels[o] = (lab == kConflict) ? ignore_label : lab;
    }
  }
}

__global__ void vox_offsets_kernel(const int64_t* __restrict__ sample_off, int n_samples, int64_t total,
                                   const uint8_t* __restrict__ head, const int64_t* __restrict__ rank,
                                   const int* __restrict__ error, int64_t* __restrict__ voxel_off) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > n_samples) return;
  if (*error) {  // a coordinate fell outside the packable range: poison the result so the host can raise
    voxel_off[s] = -1;
    return;
  }
  const int64_t j = sample_off[s];
  voxel_off[s] = (j < total) ? rank[j] : (total > 0 ? rank[total - 1] + head[total - 1] : 0);
}

// one warp per output row
__global__ void __launch_bounds__(kThreads) voxel_gather_kernel(const uint32_t* __restrict__ in, int words, const int64_t* __restrict__ sample_off,
                                                                const int64_t* __restrict__ voxel_off, const int64_t* __restrict__ unique_map,
                                                                int n_samples, uint32_t* __restrict__ out) {
  const int64_t n_out = voxel_off[n_samples];
  const int lane = threadIdx.x & 31;
  for (int64_t o = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); o < n_out; o += (int64_t)gridDim.x * (kThreads / 32)) {
    const int b = sample_of(voxel_off, n_samples, o);
    const int64_t k = o - voxel_off[b];
    const int64_t src = sample_off[b] + unique_map[sample_off[b] + k];
    for (int c = lane; c < words; c += 32) out[o * words + c] = in[src * words + c];
  }
}

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// shared with fusion_ops.cu's scan (re-declared here to stay self-contained)
constexpr int kScanItems = 16;
constexpr int kScanBlock = kThreads * kScanItems;

__global__ void __launch_bounds__(kThreads) vscan_count(const uint8_t* __restrict__ f, int64_t n, int64_t* __restrict__ sums) {
  __shared__ int s_warp[kThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanItems;
  int c = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) c += f[base + k];
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < kThreads / 32; ++k) t += s_warp[k];
    sums[blockIdx.x] = t;
  }
}
__global__ void vscan_sums(int64_t* sums, int64_t n_blocks) {  // n_blocks is tiny (total / 4096): one thread
  if (threadIdx.x || blockIdx.x) return;
  int64_t run = 0;
  for (int64_t i = 0; i < n_blocks; ++i) { const int64_t v = sums[i]; sums[i] = run; run += v; }
}
__global__ void __launch_bounds__(kThreads) vscan_rank(const uint8_t* __restrict__ f, int64_t n, const int64_t* __restrict__ sums,
                                                       int64_t* __restrict__ rank) {
  __shared__ int s_warp[kThreads / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanBlock + (int64_t)threadIdx.x * kScanItems;
  int fl[kScanItems], c = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { fl[k] = (base + k < n) ? f[base + k] : 0; c += fl[k]; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  int wb = 0;
  for (int k = 0; k < warp; ++k) wb += s_warp[k];
  int64_t run = sums[blockIdx.x] + wb + (incl - c);
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) { if (base + k < n) rank[base + k] = run; run += fl[k]; }
}

VoxWorkspace carve(void* ws, int64_t total, size_t* bytes) {
  size_t off = 0;
  auto take = [&](size_t n) { size_t o = off; off = align_up(off + n, 256); return o; };
  const size_t t = (size_t)(total > 0 ? total : 1);
  const size_t o_keys = take(16 * t), o_first = take(8 * t), o_label = take(8 * t), o_slot = take(4 * t), o_head = take(t),
               o_rank = take(8 * t), o_sums = take(8 * (t / kScanBlock + 2)), o_err = take(4);
  if (bytes) *bytes = off;
  VoxWorkspace w{};
  if (ws) {
    uint8_t* b = reinterpret_cast<uint8_t*>(ws);
    w.keys = reinterpret_cast<unsigned long long*>(b + o_keys);
    w.first = reinterpret_cast<int*>(b + o_first);
    w.label = reinterpret_cast<int*>(b + o_label);
    w.slot_of = reinterpret_cast<int*>(b + o_slot);
    w.head = b + o_head;
    w.rank = reinterpret_cast<int64_t*>(b + o_rank);
    w.block_sums = reinterpret_cast<int64_t*>(b + o_sums);
    w.error = reinterpret_cast<int*>(b + o_err);
  }
  return w;
}

__global__ void fill_i32(int* p, int64_t n, int v) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

extern "C" {

size_t dc_voxelize_workspace(int64_t total_points) {
  size_t bytes = 0;
  carve(nullptr, total_points, &bytes);
  return bytes;
}

int dc_voxelize(const float* xyz, const int64_t* sample_off, int n_samples, int64_t total_points, float voxel_size,
                const int32_t* labels, int32_t ignore_label, int32_t* coords, int64_t* unique_map, int64_t* inverse_map,
                int32_t* voxel_labels, int64_t* voxel_off, void* workspace, size_t workspace_bytes, dc_stream_t stream) {
  DC_CHECK_ARG(xyz && sample_off && coords && unique_map && inverse_map && voxel_off && workspace,
               "dc_voxelize: null pointer argument");
  DC_CHECK_ARG(!labels || voxel_labels, "dc_voxelize: labels need voxel_labels");
  DC_CHECK_ARG(voxel_size > 0.f, "dc_voxelize: voxel_size must be positive");
  DC_CHECK_ARG(n_samples >= 1, "dc_voxelize: need at least one sample");
  DC_CHECK_ARG(total_points < (1ll << 30), "dc_voxelize: too many points for one call");
  size_t need = 0;
  VoxWorkspace w = carve(workspace, total_points, &need);
  if (workspace_bytes < need) return dc::fail(DC_ERR_WORKSPACE, "dc_voxelize: workspace %zu < %zu", workspace_bytes, need);
  cudaStream_t st = dc::as_stream(stream);
  if (total_points <= 0) {
    DC_CUDA(cudaMemsetAsync(voxel_off, 0, sizeof(int64_t) * (size_t)(n_samples + 1), st));
    return DC_OK;
  }
  const size_t t = (size_t)total_points;
  DC_CUDA(cudaMemsetAsync(w.keys, 0xFF, 16 * t, st));
  DC_CUDA(cudaMemsetAsync(w.error, 0, 4, st));
  int64_t fill_blocks = dc::ceil_div<int64_t>(2 * total_points, 256);
  if (fill_blocks > (int64_t)dc::sm_count() * 8) fill_blocks = (int64_t)dc::sm_count() * 8;
  const unsigned fill_grid = (unsigned)fill_blocks;
  fill_i32<<<fill_grid, 256, 0, st>>>(w.first, 2 * total_points, INT32_MAX);
  if (labels) fill_i32<<<fill_grid, 256, 0, st>>>(w.label, 2 * total_points, kNoLabel);
  const unsigned grid = (unsigned)dc::ceil_div<int64_t>(total_points, kThreads);
  vox_insert_kernel<<<grid, kThreads, 0, st>>>(xyz, sample_off, n_samples, total_points, voxel_size, labels, w);
  vox_heads_kernel<<<grid, kThreads, 0, st>>>(sample_off, n_samples, total_points, w);
  const int64_t n_blocks = dc::ceil_div<int64_t>(total_points, kScanBlock);
  vscan_count<<<(unsigned)n_blocks, kThreads, 0, st>>>(w.head, total_points, w.block_sums);
  vscan_sums<<<1, 32, 0, st>>>(w.block_sums, n_blocks);
  vscan_rank<<<(unsigned)n_blocks, kThreads, 0, st>>>(w.head, total_points, w.block_sums, w.rank);
  vox_emit_kernel<<<grid, kThreads, 0, st>>>(xyz, sample_off, n_samples, total_points, voxel_size, labels != nullptr,
                                             ignore_label, w, coords, unique_map, inverse_map, voxel_labels);
  vox_offsets_kernel<<<dc::ceil_div(n_samples + 1, 128), 128, 0, st>>>(sample_off, n_samples, total_points, w.head, w.rank,
                                                                      w.error, voxel_off);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

int dc_voxel_gather(const void* in, int64_t row_bytes, const int64_t* sample_off, const int64_t* voxel_off,
                    const int64_t* unique_map, int n_samples, int64_t total_points, void* out, dc_stream_t stream) {
  DC_CHECK_ARG(in && sample_off && voxel_off && unique_map && out, "dc_voxel_gather: null pointer argument");
  DC_CHECK_ARG(row_bytes > 0 && row_bytes % 4 == 0, "dc_voxel_gather: row_bytes must be a positive multiple of 4");
  if (total_points <= 0) return DC_OK;
  int64_t blocks = dc::ceil_div<int64_t>(total_points, kThreads / 32);
  const int64_t cap = (int64_t)dc::sm_count() * 16;
  if (blocks > cap) blocks = cap;
  voxel_gather_kernel<<<(unsigned)blocks, kThreads, 0, dc::as_stream(stream)>>>((const uint32_t*)in, (int)(row_bytes / 4), sample_off,
                                                                             voxel_off, unique_map, n_samples, (uint32_t*)out);
  DC_LAUNCH_CHECK();
  return DC_OK;
}

}  // extern "C"
