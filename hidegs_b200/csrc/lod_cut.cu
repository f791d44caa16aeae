// lod_cut.cu — hierarchy LOD cut: which nodes of the Gaussian hierarchy are rendered for a camera and a target
// granularity, and the interpolation weight of every rendered node towards its parent.
//
// Replaces Switching::expandToSize (submodules/gaussianhierarchy/runtime_switching.cu:496-527: markNodesForSize
// :402-431, cub InclusiveSum, putRenderIndices :55-80, blocking cudaMemcpy of the count, three thrust::device_vector
// allocations per call) and Switching::getTsIndexed / computeTsIndexed (:433-494), bound as
// gaussian_hierarchy._C.expand_to_size / get_interpolation_weights (ext.cpp:19-20, torch/torch_interface.cpp:77-119).
//
// Same kernels' arithmetic (projected size = box.minn.w / distance(viewpoint, box), FLT_MAX inside the box) on a
// caller-provided workspace; node + parent boxes are read as float4 pairs; the count travels through the pinned
// slot pattern (one event wait instead of a device-wide blocking copy).
#include "common.cuh"
#include "../../include/hidegs_geometry.h"

#include <cfloat>
#include <cub/device/device_scan.cuh>

namespace hg {

namespace {

struct Node {  // types.h:47-56
  int depth, parent, start, count_leafs, count_merged, start_children, count_children;
};

// computeSizeGPU / inboxCUDA / pointboxdistCUDA (runtime_switching.cu:109-143)
__device__ __forceinline__ float projected_size(const float4* __restrict__ boxes, int node, float vx, float vy,
                                                float vz) {
  const float4 lo = __ldg(boxes + 2 * (size_t)node), hi = __ldg(boxes + 2 * (size_t)node + 1);
  const bool inside = vx >= lo.x && vx <= hi.x && vy >= lo.y && vy <= hi.y && vz >= lo.z && vz <= hi.z;
  if (inside) return FLT_MAX;
  const float cx = fmaxf(lo.x, fminf(hi.x, vx)), cy = fmaxf(lo.y, fminf(hi.y, vy)), cz = fmaxf(lo.z, fminf(hi.z, vz));
  const float dx = __fsub_rn(vx, cx), dy = __fsub_rn(vy, cy), dz = __fsub_rn(vz, cz);
  // diff.x*diff.x + diff.y*diff.y + diff.z*diff.z as nvcc contracts it: fma(dz, dz, fma(dx, dx, dy*dy))
  const float d2 = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
  return __fdiv_rn(lo.w, __fsqrt_rn(d2));
}

__global__ void __launch_bounds__(256)
mark_nodes_kernel(const Node* __restrict__ nodes, const float4* __restrict__ boxes, const int N,
                  const float* __restrict__ viewpoint, const float target, int* __restrict__ counts,
                  int* __restrict__ markers) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  const float vx = __ldg(viewpoint), vy = __ldg(viewpoint + 1), vz = __ldg(viewpoint + 2);
  const Node node = nodes[idx];
  const float size = projected_size(boxes, idx, vx, vy, vz);
  int count = 0;
  if (size >= target) {
    count = node.count_leafs;
  } else if (node.parent != -1) {
    if (projected_size(boxes, node.parent, vx, vy, vz) >= target) {
      count = node.count_leafs;
      if (node.depth != 0) count += node.count_merged;
    }
  }
  if (count != 0 && markers) markers[idx] = 1;
  counts[idx] = count;
}

__global__ void __launch_bounds__(256)
put_indices_kernel(const Node* __restrict__ nodes, const int N, const int* __restrict__ counts,
                   const int* __restrict__ offsets, const int capacity, int* __restrict__ render_indices,
                   int* __restrict__ parent_indices, int* __restrict__ node_of_index) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= N) return;
  const int count = counts[idx];
  if (count == 0) return;
  const Node node = nodes[idx];
  const int offset = idx == 0 ? 0 : offsets[idx - 1];
  const int parent_gaussian = node.parent != -1 ? nodes[node.parent].start : -1;
  for (int i = 0; i < count && offset + i < capacity; ++i) {
    render_indices[offset + i] = node.start + i;
    if (parent_indices) parent_indices[offset + i] = parent_gaussian;
    if (node_of_index) node_of_index[offset + i] = idx;
  }
}

__global__ void __launch_bounds__(256)
interpolation_weights_kernel(const Node* __restrict__ nodes, const float4* __restrict__ boxes, const int n,
                             const int* __restrict__ indices, const float vx, const float vy, const float vz,
                             const float target, float* __restrict__ ts, int* __restrict__ kids) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const int node_id = indices[idx];
  const Node node = nodes[node_id];
  float t;
  if (node.parent == -1) {
    t = 1.0f;
  } else {
    const float parentsize = projected_size(boxes, node.parent, vx, vy, vz);
    if (parentsize > __fmul_rn(2.0f, target)) {
      t = 1.0f;
    } else {
      const float size = projected_size(boxes, node_id, vx, vy, vz);
      const float start = fmaxf(__fmul_rn(0.5f, parentsize), size);
      const float diff = __fsub_rn(parentsize, start);
      if (diff <= 0) {
        t = 1.0f;
      } else {
        const float tdiff = fmaxf(0.0f, __fsub_rn(target, start));
        t = fmaxf(__fsub_rn(1.0f, __fdiv_rn(tdiff, diff)), 0.0f);
      }
    }
  }
  ts[idx] = t;
  kids[idx] = node.parent == -1 ? 1 : nodes[node.parent].count_children;
}

size_t lod_scan_bytes(int N) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, (int*)nullptr, (int*)nullptr, N);
  return bytes;
}

}  // namespace
}  // namespace hg

using namespace hg;

extern "C" {

size_t hg_expand_to_size_workspace_bytes(int32_t N) {
  if (N <= 0) return 256;
  return 2 * ((size_t)N * 4 + 256) + lod_scan_bytes(N) + 1024;
}

int hg_expand_to_size(const int32_t* nodes, const float* boxes, int32_t N, float target_size, const float* viewpoint,
                      int32_t capacity, int32_t* render_indices, int32_t* parent_indices,
                      int32_t* nodes_for_render_indices, int32_t* node_markers, void* ws, int32_t* out_count,
                      void* st_) {
  if (N < 0 || capacity < 0 || !out_count ||
      (N > 0 && (!nodes || !boxes || !viewpoint || !render_indices || !ws)) || ((uintptr_t)boxes & 15) != 0) {
    set_error("hg_expand_to_size: bad argument (boxes must be 16-byte aligned)");
    return HG_ERR_INVALID_ARG;
  }
  *out_count = 0;
  if (N == 0) return HG_OK;
  cudaStream_t st = (cudaStream_t)st_;
  char* p = (char*)(((uintptr_t)ws + 255) / 256 * 256);
  int* counts = (int*)p;
  p += ((size_t)N * 4 + 255) / 256 * 256;
  int* offsets = (int*)p;
  p += ((size_t)N * 4 + 255) / 256 * 256;
  size_t scan_bytes = lod_scan_bytes(N);
  const int blocks = (N + 255) / 256;
  mark_nodes_kernel<<<blocks, 256, 0, st>>>((const Node*)nodes, (const float4*)boxes, N, viewpoint, target_size, counts,
                                            node_markers);
  HG_POST_LAUNCH(false, st, "lod_mark_nodes");
  HG_CUDA_TRY(cub::DeviceScan::InclusiveSum(p, scan_bytes, counts, offsets, N, st));
  count_launch(2);
  put_indices_kernel<<<blocks, 256, 0, st>>>((const Node*)nodes, N, counts, offsets, capacity, render_indices,
                                             parent_indices, nodes_for_render_indices);
  HG_POST_LAUNCH(false, st, "lod_put_indices");
  int total = 0;
  HG_CUDA_TRY(cudaMemcpyAsync(&total, offsets + N - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
  HG_CUDA_TRY(cudaStreamSynchronize(st));
  *out_count = total;
  if (total > capacity) {
    set_error("hg_expand_to_size: the cut holds %d Gaussians but the index buffers hold %d", total, capacity);
    return HG_ERR_INVALID_ARG;
  }
  return HG_OK;
}

int hg_interpolation_weights(const int32_t* indices, int32_t n, float target_size, const int32_t* nodes,
                             const float* boxes, float vx, float vy, float vz, float* ts, int32_t* kids, void* st_) {
  if (n < 0 || (n > 0 && (!indices || !nodes || !boxes || !ts || !kids)) || ((uintptr_t)boxes & 15) != 0) {
    set_error("hg_interpolation_weights: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  if (n == 0) return HG_OK;
  cudaStream_t st = (cudaStream_t)st_;
  interpolation_weights_kernel<<<(n + 255) / 256, 256, 0, st>>>((const Node*)nodes, (const float4*)boxes, n, indices, vx,
                                                                vy, vz, target_size, ts, kids);
  HG_POST_LAUNCH(false, st, "lod_interpolation_weights");
  return HG_OK;
}

}  // extern "C"
