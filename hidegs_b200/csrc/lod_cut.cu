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


// ---- parent interpolation of a hierarchy cut (gaussian_renderer/__init__.py:278-318, the `interp_python` branch of
// render_post): row e < E of every output = t x[child] + (1 - t) x[parent] (the parent's quaternion negated when it
// points away from the child's), rows E .. E+S-1 = the model's last S rows (skybox).  One gather pass instead of
// ~25 PyTorch gathers / lerps / cats; each product and the sum are rounded separately, as the op chain rounds them.
// 16 threads per entry: the SH row (3 M floats) is strided over the 16 lanes (64-byte segments), lanes 0..10 also
// carry the 11 scalar attributes.
struct InterpSrc {
  const float* means; const float* scales; const float* rots; const float* opacity; const float* shs;
};
struct InterpDst {
  float* means; float* scales; float* rots; float* opacity; float* shs;
};

__device__ __forceinline__ float lerp_sep(float t, float ti, float a, float b) {
  return __fadd_rn(__fmul_rn(t, a), __fmul_rn(ti, b));
}

__device__ __forceinline__ float quat_sign(const float* __restrict__ rots, long long c, long long p) {
  const float4 rc = __ldg((const float4*)(rots + 4 * c)), rp = __ldg((const float4*)(rots + 4 * p));
  const float dot = __fmaf_rn(rc.w, rp.w, __fmaf_rn(rc.z, rp.z, __fmaf_rn(rc.y, rp.y, __fmul_rn(rc.x, rp.x))));
  return dot < 0.f ? -1.f : 1.f;
}

__global__ void __launch_bounds__(256)
hier_interp_fwd_kernel(const InterpSrc src, const long long N, const int M3, const int* __restrict__ render_indices,
                       const int* __restrict__ parent_indices, const float* __restrict__ ts, const long long E,
                       const long long S, const InterpDst dst) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long e = gid >> 4;
  const int sub = (int)(gid & 15);
  if (e >= E + S) return;
  if (e >= E) {  // skybox tail: plain copy of row N - S + (e - E)
    const long long r = N - S + (e - E);
    for (int k = sub; k < M3; k += 16) dst.shs[e * M3 + k] = __ldg(src.shs + r * M3 + k);
    if (sub < 3) dst.means[e * 3 + sub] = __ldg(src.means + r * 3 + sub);
    else if (sub < 6) dst.scales[e * 3 + sub - 3] = __ldg(src.scales + r * 3 + sub - 3);
    else if (sub < 10) dst.rots[e * 4 + sub - 6] = __ldg(src.rots + r * 4 + sub - 6);
    else if (sub == 10) dst.opacity[e] = __ldg(src.opacity + r);
    return;
  }
  long long c = render_indices[e], p = parent_indices[e];
  if (c < 0) c += N;  // torch index semantics of the reference's gathers
  if (p < 0) p += N;
  const float t = __ldg(ts + e), ti = __fsub_rn(1.0f, t);
  for (int k = sub; k < M3; k += 16)
    dst.shs[e * M3 + k] = lerp_sep(t, ti, __ldg(src.shs + c * M3 + k), __ldg(src.shs + p * M3 + k));
  if (sub < 3) {
    dst.means[e * 3 + sub] = lerp_sep(t, ti, __ldg(src.means + c * 3 + sub), __ldg(src.means + p * 3 + sub));
  } else if (sub < 6) {
    const int k = sub - 3;
    dst.scales[e * 3 + k] = lerp_sep(t, ti, __ldg(src.scales + c * 3 + k), __ldg(src.scales + p * 3 + k));
  } else if (sub < 10) {
    const int k = sub - 6;
    const float sg = quat_sign(src.rots, c, p);
    dst.rots[e * 4 + k] = lerp_sep(t, ti, __ldg(src.rots + c * 4 + k), __fmul_rn(sg, __ldg(src.rots + p * 4 + k)));
  } else if (sub == 10) {
    dst.opacity[e] = lerp_sep(t, ti, __ldg(src.opacity + c), __ldg(src.opacity + p));
  }
}

// Backward: d x[child] += t g, d x[parent] += (1 - t) g (sign-flipped for the quaternion), skybox rows += g.  The
// destination rows are pre-zeroed by the caller; a parent is shared by its children, so every write is a RED.
__global__ void __launch_bounds__(256)
hier_interp_bwd_kernel(const float* __restrict__ rots, const long long N, const int M3,
                       const int* __restrict__ render_indices, const int* __restrict__ parent_indices,
                       const float* __restrict__ ts, const long long E, const long long S, const InterpSrc g,
                       const InterpDst d) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long e = gid >> 4;
  const int sub = (int)(gid & 15);
  if (e >= E + S) return;
  long long c, p;
  float t, ti;
  if (e >= E) {
    c = p = N - S + (e - E);
    t = 1.f;
    ti = 0.f;
  } else {
    c = render_indices[e];
    p = parent_indices[e];
    if (c < 0) c += N;
    if (p < 0) p += N;
    t = __ldg(ts + e);
    ti = __fsub_rn(1.0f, t);
  }
  const bool par = e < E;
  if (g.shs && d.shs)
    for (int k = sub; k < M3; k += 16) {
      const float v = __ldg(g.shs + e * M3 + k);
      atomicAdd(d.shs + c * M3 + k, t * v);
      if (par) atomicAdd(d.shs + p * M3 + k, ti * v);
    }
  if (sub < 3) {
    if (g.means && d.means) {
      const float v = __ldg(g.means + e * 3 + sub);
      atomicAdd(d.means + c * 3 + sub, t * v);
      if (par) atomicAdd(d.means + p * 3 + sub, ti * v);
    }
  } else if (sub < 6) {
    if (g.scales && d.scales) {
      const int k = sub - 3;
      const float v = __ldg(g.scales + e * 3 + k);
      atomicAdd(d.scales + c * 3 + k, t * v);
      if (par) atomicAdd(d.scales + p * 3 + k, ti * v);
    }
  } else if (sub < 10) {
    if (g.rots && d.rots) {
      const int k = sub - 6;
      const float v = __ldg(g.rots + e * 4 + k);
      atomicAdd(d.rots + c * 4 + k, t * v);
      if (par) atomicAdd(d.rots + p * 4 + k, quat_sign(rots, c, p) * ti * v);
    }
  } else if (sub == 10) {
    if (g.opacity && d.opacity) {
      const float v = __ldg(g.opacity + e);
      atomicAdd(d.opacity + c, t * v);
      if (par) atomicAdd(d.opacity + p, ti * v);
    }
  }
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

static int interp_args_ok(const char* what, int64_t N, int32_t M, const void* ri, const void* pi, const void* ts, int64_t E,
                          int64_t S) {
  if (N < 0 || M < 0 || E < 0 || S < 0 || S > N || (E > 0 && (!ri || !pi || !ts)) || (E + S) * 16 > (int64_t)INT32_MAX * 256) {
    set_error("%s: bad argument", what);
    return HG_ERR_INVALID_ARG;
  }
  return HG_OK;
}

int hg_hier_interpolate(const float* means3D, const float* scales, const float* rotations, const float* opacity,
                        const float* shs, int64_t N, int32_t M, const int32_t* render_indices,
                        const int32_t* parent_indices, const float* ts, int64_t E, int64_t S, float* out_means3D,
                        float* out_scales, float* out_rotations, float* out_opacity, float* out_shs, void* st_) {
  int rc = interp_args_ok("hg_hier_interpolate", N, M, render_indices, parent_indices, ts, E, S);
  if (rc != HG_OK) return rc;
  if (E + S == 0) return HG_OK;
  if (!means3D || !scales || !rotations || !opacity || (M > 0 && !shs) || !out_means3D || !out_scales ||
      !out_rotations || !out_opacity || (M > 0 && !out_shs) || ((uintptr_t)rotations & 15) != 0) {
    set_error("hg_hier_interpolate: null tensor (rotations must be 16-byte aligned)");
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)st_;
  const InterpSrc src{means3D, scales, rotations, opacity, shs};
  const InterpDst dst{out_means3D, out_scales, out_rotations, out_opacity, out_shs};
  const long long threads = (E + S) * 16;
  hier_interp_fwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(src, N, 3 * M, render_indices, parent_indices,
                                                                           ts, E, S, dst);
  HG_POST_LAUNCH(false, st, "hier_interp_fwd");
  return HG_OK;
}

int hg_hier_interpolate_backward(const float* rotations, int64_t N, int32_t M, const int32_t* render_indices,
                                 const int32_t* parent_indices, const float* ts, int64_t E, int64_t S,
                                 const float* g_means3D, const float* g_scales, const float* g_rotations,
                                 const float* g_opacity, const float* g_shs, float* d_means3D, float* d_scales,
                                 float* d_rotations, float* d_opacity, float* d_shs, void* st_) {
  int rc = interp_args_ok("hg_hier_interpolate_backward", N, M, render_indices, parent_indices, ts, E, S);
  if (rc != HG_OK) return rc;
  if (E + S == 0) return HG_OK;
  if (!rotations || ((uintptr_t)rotations & 15) != 0) {
    set_error("hg_hier_interpolate_backward: rotations missing or not 16-byte aligned");
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)st_;
  const InterpSrc g{g_means3D, g_scales, g_rotations, g_opacity, g_shs};
  const InterpDst d{d_means3D, d_scales, d_rotations, d_opacity, d_shs};
  const long long threads = (E + S) * 16;
  hier_interp_bwd_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(rotations, N, 3 * M, render_indices,
                                                                           parent_indices, ts, E, S, g, d);
  HG_POST_LAUNCH(false, st, "hier_interp_bwd");
  return HG_OK;
}

}  // extern "C"
