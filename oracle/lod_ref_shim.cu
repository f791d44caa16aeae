// oracle/lod_ref_shim.cu — C entry points over the UNMODIFIED reference Switching class (TEST INFRASTRUCTURE).
// Compiled together with /root/reference/submodules/gaussianhierarchy/runtime_switching.cu (from where it lies) into
// oracle/_ref/ref_lod_switching.so by oracle/Makefile; the reference's torch binding (torch/torch_interface.cpp) pulls
// in Eigen-based loaders that are irrelevant here, so the two run-time entry points are exposed through this shim.
#include <cuda_runtime.h>
#include "runtime_switching.h"

extern "C" {

// torch_interface.cpp:77-98 ExpandToSize
int ref_expand_to_size(int N, float size, int* nodes, float* boxes, float* viewpoint, float zx, float zy, float zz,
                       int* render_indices, int* parent_indices, int* nodes_for_render_indices) {
  return Switching::expandToSize(N, size, nodes, boxes, viewpoint, zx, zy, zz, render_indices, nullptr, parent_indices,
                                 nodes_for_render_indices);
}

// torch_interface.cpp:100-119 GetTsIndexed
void ref_get_ts_indexed(int N, int* indices, float size, int* nodes, float* boxes, float vx, float vy, float vz, float zx,
                        float zy, float zz, float* ts, int* kids) {
  Switching::getTsIndexed(N, indices, size, nodes, boxes, vx, vy, vz, zx, zy, zz, ts, kids, 0);
  cudaDeviceSynchronize();
}

}
