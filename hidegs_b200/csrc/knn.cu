// knn.cu — mean squared distance to the 3 nearest neighbours of every point.
//
// Replaces simple-knn's distCUDA2 -> SimpleKNN::knn (submodules/simple-knn/spatial.cu:15-26,
// simple_knn.cu:186-222, kernels :70-183).  The reference sorts the points along a 30-bit Morton curve,
// takes the 6 curve neighbours of a point to bound its 3-NN radius and then scans EVERY box of 1024 sorted
// points whose bounding box intersects that radius; the answer is the exact 3-NN (self excluded, duplicates
// counted) distance triple, independent of the traversal order.
//
// Design (B200): same exact search, restructured so that a point does not test all N/1024 boxes:
//   * bounding box, Morton codes, sort (CUB, library) and a gather into a sorted float4 array (x, y, z, source
//     index) so that the scan reads contiguous 16-byte records;
//   * a two-level box hierarchy over the sorted array: leaves of 128 points, inner boxes of 32 leaves;
//   * one thread per sorted point (warps hold curve neighbours, so they prune the same boxes).
// The squared distance uses the reference's expression and contraction, so results are bit-identical.
// Nothing synchronises with the host (the reference does two blocking cudaMemcpy + cudaMalloc/cudaFree).
#include "common.cuh"
#include "../../include/hidegs_geometry.h"

#include <cfloat>
#include <cub/device/device_radix_sort.cuh>

namespace hg {

namespace {

constexpr int kLeaf = 128;   // sorted points per leaf box
constexpr int kFan = 32;     // leaves per inner box
constexpr int kBBoxBlocks = 148 * 2;

struct Box { float lo[3], hi[3]; };

__device__ __forceinline__ void box_init(Box& b) {
#pragma unroll
  for (int i = 0; i < 3; ++i) { b.lo[i] = FLT_MAX; b.hi[i] = -FLT_MAX; }
}
__device__ __forceinline__ void box_add(Box& b, float x, float y, float z) {
  b.lo[0] = fminf(b.lo[0], x); b.hi[0] = fmaxf(b.hi[0], x);
  b.lo[1] = fminf(b.lo[1], y); b.hi[1] = fmaxf(b.hi[1], y);
  b.lo[2] = fminf(b.lo[2], z); b.hi[2] = fmaxf(b.hi[2], z);
}
__device__ __forceinline__ void box_merge_warp(Box& b) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      b.lo[i] = fminf(b.lo[i], __shfl_xor_sync(0xffffffffu, b.lo[i], o));
      b.hi[i] = fmaxf(b.hi[i], __shfl_xor_sync(0xffffffffu, b.hi[i], o));
    }
  }
}

// CTA-wide box merge (blockDim.x <= 1024); result valid in thread 0.
__device__ __forceinline__ void box_merge_cta(Box& b, Box* sm) {
  box_merge_warp(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) sm[warp] = b;
  __syncthreads();
  if (warp == 0) {
    const int nw = (blockDim.x + 31) >> 5;
    if (lane < nw) b = sm[lane];
    else box_init(b);
    box_merge_warp(b);
  }
}

__global__ void __launch_bounds__(256)
bbox_partial_kernel(const float* __restrict__ pts, const int64_t N, Box* __restrict__ partial) {
  __shared__ Box sm[32];
  Box b;
  box_init(b);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x)
    box_add(b, pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]);
  box_merge_cta(b, sm);
  if (threadIdx.x == 0) partial[blockIdx.x] = b;
}

__global__ void __launch_bounds__(kBBoxBlocks <= 512 ? 512 : 1024)
bbox_final_kernel(const Box* __restrict__ partial, int n, Box* __restrict__ out) {
  __shared__ Box sm[32];
  Box b;
  box_init(b);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const Box o = partial[i];
    box_add(b, o.lo[0], o.lo[1], o.lo[2]);
    box_add(b, o.hi[0], o.hi[1], o.hi[2]);
  }
  box_merge_cta(b, sm);
  if (threadIdx.x == 0) *out = b;
}

__device__ __forceinline__ uint32_t spread10(uint32_t x) {
  x = (x | (x << 16)) & 0x030000FFu;
  x = (x | (x << 8)) & 0x0300F00Fu;
  x = (x | (x << 4)) & 0x030C30C3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}

__global__ void __launch_bounds__(256)
morton_kernel(const float* __restrict__ pts, const int64_t N, const Box* __restrict__ bbox,
              uint32_t* __restrict__ codes, uint32_t* __restrict__ ids) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const Box b = *bbox;
  uint32_t c[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float ext = b.hi[a] - b.lo[a];
    float t = ext > 0.f ? (pts[3 * i + a] - b.lo[a]) / ext : 0.f;
    t = fminf(fmaxf(t, 0.f), 1.f);
    c[a] = spread10((uint32_t)(t * 1023.0f));
  }
  codes[i] = c[0] | (c[1] << 1) | (c[2] << 2);
  ids[i] = (uint32_t)i;
}

__global__ void __launch_bounds__(256)
gather_kernel(const float* __restrict__ pts, const uint32_t* __restrict__ ids_sorted, const int64_t N,
              float4* __restrict__ sorted) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const uint32_t s = ids_sorted[i];
  sorted[i] = make_float4(pts[3 * (size_t)s], pts[3 * (size_t)s + 1], pts[3 * (size_t)s + 2], __uint_as_float(s));
}

// leaves: one CTA (kLeaf threads) per leaf
__global__ void __launch_bounds__(kLeaf)
leaf_box_kernel(const float4* __restrict__ sorted, const int64_t N, Box* __restrict__ leaves) {
  __shared__ Box sm[32];
  const int64_t i = (int64_t)blockIdx.x * kLeaf + threadIdx.x;
  Box b;
  box_init(b);
  if (i < N) {
    const float4 p = sorted[i];
    box_add(b, p.x, p.y, p.z);
  }
  box_merge_cta(b, sm);
  if (threadIdx.x == 0) leaves[blockIdx.x] = b;
}

// inner boxes: one warp per inner box
__global__ void __launch_bounds__(256)
inner_box_kernel(const Box* __restrict__ leaves, const int n_leaves, Box* __restrict__ inner, const int n_inner) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_inner) return;
  Box b;
  box_init(b);
  const int l = w * kFan + lane;
  if (l < n_leaves) b = leaves[l];
  box_merge_warp(b);
  if (lane == 0) inner[w] = b;
}

// distBoxPoint (simple_knn.cu:121-131)
__device__ __forceinline__ float box_dist2(const Box& b, float x, float y, float z) {
  float dx = 0.f, dy = 0.f, dz = 0.f;
  if (x < b.lo[0] || x > b.hi[0]) dx = fminf(fabsf(x - b.lo[0]), fabsf(x - b.hi[0]));
  if (y < b.lo[1] || y > b.hi[1]) dy = fminf(fabsf(y - b.lo[1]), fabsf(y - b.hi[1]));
  if (z < b.lo[2] || z > b.hi[2]) dz = fminf(fabsf(z - b.lo[2]), fabsf(z - b.hi[2]));
  return dx * dx + dy * dy + dz * dz;
}

// updateKBest<3> (simple_knn.cu:133-148) with the reference's compiled distance expression
// d.x*d.x + d.y*d.y + d.z*d.z  ->  fma(dz, dz, fma(dx, dx, dy*dy))  (SASS of the sm_100 build: FMUL y, FFMA x, FFMA z).
__device__ __forceinline__ void update3(float rx, float ry, float rz, const float4 q, float (&best)[3]) {
  const float dx = __fsub_rn(q.x, rx), dy = __fsub_rn(q.y, ry), dz = __fsub_rn(q.z, rz);
  float dist = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    if (best[j] > dist) {
      const float t = best[j];
      best[j] = dist;
      dist = t;
    }
  }
}

__global__ void __launch_bounds__(128)
knn3_kernel(const float4* __restrict__ sorted, const int64_t N, const Box* __restrict__ leaves, const int n_leaves,
            const Box* __restrict__ inner, const int n_inner, float* __restrict__ out) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  const float4 me = sorted[idx];
  float best[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
  // curve neighbours bound the search radius (simple_knn.cu:159-164)
  const int64_t lo = idx - 3 > 0 ? idx - 3 : 0, hi = idx + 3 < N - 1 ? idx + 3 : N - 1;
  for (int64_t i = lo; i <= hi; ++i) {
    if (i == idx) continue;
    update3(me.x, me.y, me.z, sorted[i], best);
  }
  const float reject = best[2];
  best[0] = best[1] = best[2] = FLT_MAX;
  for (int s = 0; s < n_inner; ++s) {
    const float ds = box_dist2(inner[s], me.x, me.y, me.z);
    if (ds > reject || ds > best[2]) continue;
    const int l1 = min(n_leaves, (s + 1) * kFan);
    for (int l = s * kFan; l < l1; ++l) {
      const float dl = box_dist2(leaves[l], me.x, me.y, me.z);
      if (dl > reject || dl > best[2]) continue;
      const int64_t i1 = min(N, (int64_t)(l + 1) * kLeaf);
      for (int64_t i = (int64_t)l * kLeaf; i < i1; ++i) {
        if (i == idx) continue;
        update3(me.x, me.y, me.z, sorted[i], best);
      }
    }
  }
  out[__float_as_uint(me.w)] = (best[0] + best[1] + best[2]) / 3.0f;
}

struct Carve {
  char* p;
  explicit Carve(void* base) : p((char*)(((uintptr_t)base + 255) / 256 * 256)) {}
  template <typename T>
  T* take(size_t n) {
    T* r = (T*)p;
    p += (n * sizeof(T) + 255) / 256 * 256;
    return r;
  }
};

size_t knn_sort_bytes(int64_t N) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (int)N, 0, 30);
  return bytes;
}

}  // namespace
}  // namespace hg

using namespace hg;

extern "C" {

size_t hg_dist2_knn3_workspace_bytes(int64_t N) {
  if (N <= 0) return 256;
  const size_t n = (size_t)N;
  const size_t n_leaves = (n + kLeaf - 1) / kLeaf, n_inner = (n_leaves + kFan - 1) / kFan;
  return 4 * (n * 4 + 256) + (n * 16 + 256) + (n_leaves + n_inner + kBBoxBlocks + 1) * sizeof(Box) + 4 * 256 +
         knn_sort_bytes(N) + 1024;
}

int hg_dist2_knn3(const float* points, int64_t N, float* out, void* ws, void* st_) {
  if (N < 0 || N > 0x7fffffff || (N > 0 && (!points || !out || !ws))) {
    set_error("hg_dist2_knn3: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  if (N == 0) return HG_OK;
  cudaStream_t st = (cudaStream_t)st_;
  const int n_leaves = (int)((N + kLeaf - 1) / kLeaf), n_inner = (n_leaves + kFan - 1) / kFan;
  Carve cv(ws);
  uint32_t* codes = cv.take<uint32_t>(N);
  uint32_t* codes_sorted = cv.take<uint32_t>(N);
  uint32_t* ids = cv.take<uint32_t>(N);
  uint32_t* ids_sorted = cv.take<uint32_t>(N);
  float4* sorted = cv.take<float4>(N);
  Box* leaves = cv.take<Box>(n_leaves);
  Box* inner = cv.take<Box>(n_inner);
  Box* partial = cv.take<Box>(kBBoxBlocks);
  Box* bbox = cv.take<Box>(1);
  size_t sort_bytes = knn_sort_bytes(N);
  char* sort_tmp = cv.take<char>(sort_bytes);

  const unsigned nb = (unsigned)((N + 255) / 256);
  bbox_partial_kernel<<<kBBoxBlocks, 256, 0, st>>>(points, N, partial);
  HG_POST_LAUNCH(false, st, "knn_bbox");
  bbox_final_kernel<<<1, 512, 0, st>>>(partial, kBBoxBlocks, bbox);
  HG_POST_LAUNCH(false, st, "knn_bbox_final");
  morton_kernel<<<nb, 256, 0, st>>>(points, N, bbox, codes, ids);
  HG_POST_LAUNCH(false, st, "knn_morton");
  HG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(sort_tmp, sort_bytes, codes, codes_sorted, ids, ids_sorted, (int)N, 0, 30,
                                              st));
  count_launch(5);
  gather_kernel<<<nb, 256, 0, st>>>(points, ids_sorted, N, sorted);
  HG_POST_LAUNCH(false, st, "knn_gather");
  leaf_box_kernel<<<n_leaves, kLeaf, 0, st>>>(sorted, N, leaves);
  HG_POST_LAUNCH(false, st, "knn_leaf_boxes");
  inner_box_kernel<<<(n_inner * 32 + 255) / 256, 256, 0, st>>>(leaves, n_leaves, inner, n_inner);
  HG_POST_LAUNCH(false, st, "knn_inner_boxes");
  knn3_kernel<<<(unsigned)((N + 127) / 128), 128, 0, st>>>(sorted, N, leaves, n_leaves, inner, n_inner, out);
  HG_POST_LAUNCH(false, st, "knn3");
  return HG_OK;
}

}  // extern "C"
