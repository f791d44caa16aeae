// oracle/hier_ref_shim.cpp — C entry points over the UNMODIFIED reference HierarchyLoader / HierarchyWriter / Traversal
// and the point-cloud Loader (TEST INFRASTRUCTURE).  Compiled together with hierarchy_loader.cpp, hierarchy_writer.cpp,
// loader.cpp and traversal.cpp from
// where they lie under /root/reference/submodules/gaussianhierarchy (Eigen from its dependencies/ directory) into
// oracle/_ref/ref_hier_io.so by oracle/Makefile.  Mirrors torch/torch_interface.cpp:18-83 without the torch types.
#include <cstring>
#include <vector>

#include "hierarchy_loader.h"
#include "hierarchy_writer.h"
#include "loader.h"
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

// The reference's own C++ readers of what GaussianModel.save_ply / save_pt write (loader.cpp:76-160: the hierarchy
// builder consumes point_cloud.ply / point_cloud.bin through them): an independent consumer of the files
// hidegs_b200/ply_io.py writes.  Outputs per Gaussian: position 3, shs 48 (the loader's own order), opacity (after its
// sigmoid), scale 3 (after exp), rotation 4 (normalised), covariance 6.  Two-call protocol as above.
static int ref_copy_out(const std::vector<Gaussian>& g, int* count, float* pos, float* shs, float* opacity, float* scale,
                        float* rot, float* cov) {
  *count = (int)g.size();
  if (!pos) return 0;
  for (size_t i = 0; i < g.size(); ++i) {
    memcpy(pos + 3 * i, g[i].position.data(), 12);
    memcpy(shs + 48 * i, g[i].shs.data(), 192);
    opacity[i] = g[i].opacity;
    memcpy(scale + 3 * i, g[i].scale.data(), 12);
    memcpy(rot + 4 * i, g[i].rotation.data(), 16);
    memcpy(cov + 6 * i, g[i].covariance.data(), 24);
  }
  return 0;
}

int ref_load_ply(const char* filename, int skybox, int* count, float* pos, float* shs, float* opacity, float* scale,
                 float* rot, float* cov) {
  try {
    std::vector<Gaussian> g;
    Loader::loadPly(filename, g, skybox);
    return ref_copy_out(g, count, pos, shs, opacity, scale, rot, cov);
  } catch (...) {
    return 1;
  }
}

int ref_load_bin(const char* filename, int skybox, int* count, float* pos, float* shs, float* opacity, float* scale,
                 float* rot, float* cov) {
  try {
    std::vector<Gaussian> g;
    Loader::loadBin(filename, g, skybox);
    return ref_copy_out(g, count, pos, shs, opacity, scale, rot, cov);
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
