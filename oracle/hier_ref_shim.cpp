// oracle/hier_ref_shim.cpp — C entry points over the UNMODIFIED reference HierarchyLoader / HierarchyWriter / Traversal
// (TEST INFRASTRUCTURE).  Compiled together with hierarchy_loader.cpp, hierarchy_writer.cpp and traversal.cpp from
// where they lie under /root/reference/submodules/gaussianhierarchy (Eigen from its dependencies/ directory) into
// oracle/_ref/ref_hier_io.so by oracle/Makefile.  Mirrors torch/torch_interface.cpp:18-83 without the torch types.
#include <cstring>
#include <vector>

#include "hierarchy_loader.h"
#include "hierarchy_writer.h"
#include "traversal.h"

extern "C" {

// Two-call protocol: first with null outputs to learn P and N, then with buffers.
int ref_hier_load(const char* filename, int* P, int* N, float* pos, float* shs, float* alphas, float* scales, float* rot,
                  int* nodes, float* boxes) {
  try {
    HierarchyLoader loader;
    std::vector<Eigen::Vector3f> vpos, vscales;
    std::vector<SHs> vshs;
    std::vector<float> valphas;
    std::vector<Eigen::Vector4f> vrot;
    std::vector<Node> vnodes;
    std::vector<Box> vboxes;
    loader.load(filename, vpos, vshs, valphas, vscales, vrot, vnodes, vboxes);
    *P = (int)vpos.size();
    *N = (int)vnodes.size();
    if (pos) {
      memcpy(pos, vpos.data(), vpos.size() * 12);
      memcpy(shs, vshs.data(), vshs.size() * 192);
      memcpy(alphas, valphas.data(), valphas.size() * 4);
      memcpy(scales, vscales.data(), vscales.size() * 12);
      memcpy(rot, vrot.data(), vrot.size() * 16);
      memcpy(nodes, vnodes.data(), vnodes.size() * sizeof(Node));
      memcpy(boxes, vboxes.data(), vboxes.size() * sizeof(Box));
    }
    return 0;
  } catch (...) {
    return 1;
  }
}

int ref_hier_write(const char* filename, int P, int N, float* pos, float* shs, float* opacities, float* log_scales,
                   float* rotations, int* nodes, float* boxes, int compressed) {
  try {
    HierarchyWriter writer;
    writer.write(filename, P, N, (Eigen::Vector3f*)pos, (SHs*)shs, opacities, (Eigen::Vector3f*)log_scales,
                 (Eigen::Vector4f*)rotations, (Node*)nodes, (Box*)boxes, compressed != 0);
    return 0;
  } catch (...) {
    return 1;
  }
}

int ref_expand_to_target(int* nodes, int target, int* out, int capacity) {
  std::vector<int> idx = Traversal::expandToTarget((Node*)nodes, target);
  const int n = (int)idx.size();
  if (out) memcpy(out, idx.data(), sizeof(int) * (size_t)(n < capacity ? n : capacity));
  return n;
}

}
