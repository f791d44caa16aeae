// hier_io.cu — the .hier hierarchy file format (both variants) and the static depth cut (include/hidegs_hierarchy.h).
//
// Replaces HierarchyLoader::load / HierarchyWriter::write / Traversal::expandToTarget of the reference
// (submodules/gaussianhierarchy/hierarchy_loader.cpp:26-128, hierarchy_writer.cpp:27-118, traversal.cpp:14-38).
//
// Design: the file is a handful of contiguous sections, so it is read / written with ONE sequential pass over a raw
// image; the half <-> float and HalfNode <-> Node conversions (element-by-element host loops with per-section
// std::vectors in the reference) are kernels over that image when the tensors live in HBM (a compressed 6M-Gaussian
// scene is ~700 MB: disk -> pinned host -> HBM -> fp32 tensors without a host-side widening pass), and a plain
// single-pass host loop for the API-compatible CPU-tensor path.
#include "common.cuh"

#include <cuda_fp16.h>

#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/hidegs_hierarchy.h"

namespace hg {
namespace {

struct HalfNodeRec {  // types.h:58-64
  int32_t parent, start, start_children;
  int16_t dccc[4];  // depth, count_children, count_leafs, count_merged
};
static_assert(sizeof(HalfNodeRec) == 20, "half node record is 20 bytes");

void fill_layout(int64_t P, int64_t N, bool compressed, hg_hier_layout* L) {
  const int64_t e = compressed ? 2 : 4;  // bytes per non-position scalar
  L->P = P;
  L->N = N;
  L->compressed = compressed ? 1 : 0;
  int64_t off = 4;
  L->pos = off;      off += P * 12;
  L->rot = off;      off += P * 4 * e;
  L->scale = off;    off += P * 3 * e;
  L->opacity = off;  off += P * e;
  L->sh = off;       off += P * 48 * e;
  off += 4;  // node count
  L->nodes = off;    off += N * (compressed ? (int64_t)sizeof(HalfNodeRec) : 28);
  L->boxes = off;    off += N * 8 * e;
  L->file_bytes = off;
}

// host half conversion lives in hier_half.cpp (F16C when the CPU has it, scalar otherwise)
}  // namespace
void half_to_float_array(const uint16_t* src, float* dst, size_t n);
void float_to_half_array(const float* src, uint16_t* dst, size_t n);
namespace {

struct File {
  FILE* f = nullptr;
  ~File() {
    if (f) fclose(f);
  }
};

int probe(const char* filename, hg_hier_layout* out, File* keep_open) {
  File local;
  File& fl = keep_open ? *keep_open : local;
  fl.f = filename ? fopen(filename, "rb") : nullptr;
  if (!fl.f) {
    set_error("File not found!");  // hierarchy_loader.cpp:38
    return HG_ERR_INVALID_ARG;
  }
  int32_t P = 0, N = 0;
  if (fread(&P, 4, 1, fl.f) != 1) {
    set_error("%s: truncated hierarchy file (no Gaussian count)", filename);
    return HG_ERR_INVALID_ARG;
  }
  const bool compressed = P < 0;
  const int64_t allP = compressed ? -(int64_t)P : P;
  hg_hier_layout L;
  fill_layout(allP, 0, compressed, &L);
  if (fseek(fl.f, (long)(L.nodes - 4), SEEK_SET) != 0 || fread(&N, 4, 1, fl.f) != 1 || N < 0) {
    set_error("%s: truncated hierarchy file (no node count)", filename);
    return HG_ERR_INVALID_ARG;
  }
  fill_layout(allP, N, compressed, out);
  return HG_OK;
}

bool read_exact(FILE* f, int64_t off, void* dst, int64_t bytes) {
  if (bytes == 0) return true;
  if (fseek(f, (long)off, SEEK_SET) != 0) return false;
  return (int64_t)fread(dst, 1, (size_t)bytes, f) == bytes;
}

// ---- device kernels ----
__global__ void widen_half_kernel(const __half* __restrict__ src, float* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __half2float(src[i]);
}
__global__ void narrow_half_kernel(const float* __restrict__ src, __half* __restrict__ dst, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __float2half_rn(src[i]);
}
// The half sections start at 2-byte aligned offsets only, so bytes are fetched as 16-bit words.
__global__ void widen_nodes_kernel(const uint16_t* __restrict__ src, int32_t* __restrict__ nodes, int64_t N) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const uint16_t* r = src + i * 10;
    const int32_t parent = (int32_t)((uint32_t)r[0] | ((uint32_t)r[1] << 16));
    const int32_t start = (int32_t)((uint32_t)r[2] | ((uint32_t)r[3] << 16));
    const int32_t start_children = (int32_t)((uint32_t)r[4] | ((uint32_t)r[5] << 16));
    int32_t* o = nodes + i * 7;
    o[0] = (int16_t)r[6];  // depth
    o[1] = parent;
    o[2] = start;
    o[3] = (int16_t)r[8];  // count_leafs
    o[4] = (int16_t)r[9];  // count_merged
    o[5] = start_children;
    o[6] = (int16_t)r[7];  // count_children
  }
}
__global__ void narrow_nodes_kernel(const int32_t* __restrict__ nodes, uint16_t* __restrict__ dst, int64_t N,
                                    int32_t* __restrict__ overflow) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int32_t* n = nodes + i * 7;
    if (n[0] > 32000 || n[6] > 32000 || n[3] > 32000 || n[4] > 32000) atomicExch(overflow, 1);
    uint16_t* r = dst + i * 10;
    r[0] = (uint16_t)((uint32_t)n[1] & 0xffffu);  r[1] = (uint16_t)((uint32_t)n[1] >> 16);
    r[2] = (uint16_t)((uint32_t)n[2] & 0xffffu);  r[3] = (uint16_t)((uint32_t)n[2] >> 16);
    r[4] = (uint16_t)((uint32_t)n[5] & 0xffffu);  r[5] = (uint16_t)((uint32_t)n[5] >> 16);
    r[6] = (uint16_t)(int16_t)n[0];
    r[7] = (uint16_t)(int16_t)n[6];
    r[8] = (uint16_t)(int16_t)n[3];
    r[9] = (uint16_t)(int16_t)n[4];
  }
}

int grid_for(int64_t n) {
  const int64_t b = (n + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

}  // namespace
}  // namespace hg

extern "C" {

int hg_hier_layout_for(int64_t P, int64_t N, int32_t compressed, hg_hier_layout* out) {
  if (!out || P < 0 || N < 0 || P > 0x7fffffff || N > 0x7fffffff) {
    hg::set_error("hg_hier_layout_for: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  hg::fill_layout(P, N, compressed != 0, out);
  return HG_OK;
}

int hg_hier_probe(const char* filename, hg_hier_layout* out) {
  if (!out) {
    hg::set_error("hg_hier_probe: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  return hg::probe(filename, out, nullptr);
}

int hg_hier_read_raw(const char* filename, void* raw, int64_t capacity) {
  hg::File fl;
  hg_hier_layout L;
  const int rc = hg::probe(filename, &L, &fl);
  if (rc != HG_OK) return rc;
  if (!raw || capacity < L.file_bytes) {
    hg::set_error("hg_hier_read_raw: buffer too small (%lld < %lld)", (long long)capacity, (long long)L.file_bytes);
    return HG_ERR_INVALID_ARG;
  }
  if (!hg::read_exact(fl.f, 0, raw, L.file_bytes)) {
    hg::set_error("%s: truncated hierarchy file (%lld bytes expected)", filename, (long long)L.file_bytes);
    return HG_ERR_INVALID_ARG;
  }
  return HG_OK;
}

int hg_hier_load(const char* filename, float* pos, float* shs, float* alphas, float* scales, float* rot,
                 int32_t* nodes, float* boxes) {
  hg::File fl;
  hg_hier_layout L;
  const int rc = hg::probe(filename, &L, &fl);
  if (rc != HG_OK) return rc;
  const int64_t P = L.P, N = L.N;
  if ((P && (!pos || !shs || !alphas || !scales || !rot)) || (N && (!nodes || !boxes))) {
    hg::set_error("hg_hier_load: output arrays are mandatory");
    return HG_ERR_INVALID_ARG;
  }
  bool ok = hg::read_exact(fl.f, L.pos, pos, P * 12);
  if (!L.compressed) {
    ok = ok && hg::read_exact(fl.f, L.rot, rot, P * 16) && hg::read_exact(fl.f, L.scale, scales, P * 12) &&
         hg::read_exact(fl.f, L.opacity, alphas, P * 4) && hg::read_exact(fl.f, L.sh, shs, P * 192) &&
         hg::read_exact(fl.f, L.nodes, nodes, N * 28) && hg::read_exact(fl.f, L.boxes, boxes, N * 32);
  } else {
    // one staging buffer, section by section, widened in place order (hierarchy_loader.cpp:87-126)
    std::vector<uint16_t> tmp;
    auto widen = [&](int64_t off, float* dst, int64_t count) {
      tmp.resize((size_t)count);
      if (!hg::read_exact(fl.f, off, tmp.data(), count * 2)) return false;
      hg::half_to_float_array(tmp.data(), dst, (size_t)count);
      return true;
    };
    ok = ok && widen(L.rot, rot, P * 4) && widen(L.scale, scales, P * 3) && widen(L.opacity, alphas, P) &&
         widen(L.sh, shs, P * 48) && widen(L.boxes, boxes, N * 8);
    if (ok) {
      std::vector<hg::HalfNodeRec> hn((size_t)N);
      ok = hg::read_exact(fl.f, L.nodes, hn.data(), N * (int64_t)sizeof(hg::HalfNodeRec));
      for (int64_t i = 0; ok && i < N; ++i) {
        const hg::HalfNodeRec& h = hn[(size_t)i];
        int32_t* o = nodes + i * 7;
        o[0] = h.dccc[0];
        o[1] = h.parent;
        o[2] = h.start;
        o[3] = h.dccc[2];
        o[4] = h.dccc[3];
        o[5] = h.start_children;
        o[6] = h.dccc[1];
      }
    }
  }
  if (!ok) {
    hg::set_error("%s: truncated hierarchy file (%lld bytes expected)", filename, (long long)L.file_bytes);
    return HG_ERR_INVALID_ARG;
  }
  return HG_OK;
}

int hg_hier_write(const char* filename, int64_t P, int64_t N, const float* pos, const float* shs,
                  const float* opacities, const float* log_scales, const float* rotations, const int32_t* nodes,
                  const float* boxes, int32_t compressed) {
  hg_hier_layout L;
  if (hg_hier_layout_for(P, N, compressed, &L) != HG_OK) return HG_ERR_INVALID_ARG;
  if ((P && (!pos || !shs || !opacities || !log_scales || !rotations)) || (N && (!nodes || !boxes))) {
    hg::set_error("hg_hier_write: input arrays are mandatory");
    return HG_ERR_INVALID_ARG;
  }
  if (compressed) {
    for (int64_t i = 0; i < N; ++i) {
      const int32_t* n = nodes + i * 7;
      if (n[0] > 32000 || n[6] > 32000 || n[3] > 32000 || n[4] > 32000) {
        hg::set_error("Would lose information!");  // hierarchy_writer.cpp:97
        return HG_ERR_INVALID_ARG;
      }
    }
  }
  hg::File fl;
  fl.f = filename ? fopen(filename, "wb") : nullptr;
  if (!fl.f) {
    hg::set_error("File not created!");  // hierarchy_writer.cpp:41
    return HG_ERR_INVALID_ARG;
  }
  bool ok = true;
  auto put = [&](const void* p, int64_t bytes) {
    if (bytes) ok = ok && (int64_t)fwrite(p, 1, (size_t)bytes, fl.f) == bytes;
  };
  const int32_t headP = compressed ? -(int32_t)P : (int32_t)P, headN = (int32_t)N;
  put(&headP, 4);
  put(pos, P * 12);
  if (!compressed) {
    put(rotations, P * 16);
    put(log_scales, P * 12);
    put(opacities, P * 4);
    put(shs, P * 192);
    put(&headN, 4);
    put(nodes, N * 28);
    put(boxes, N * 32);
  } else {
    std::vector<uint16_t> tmp;
    auto narrow = [&](const float* src, int64_t count) {
      tmp.resize((size_t)count);
      hg::float_to_half_array(src, tmp.data(), (size_t)count);
      put(tmp.data(), count * 2);
    };
    narrow(rotations, P * 4);
    narrow(log_scales, P * 3);
    narrow(opacities, P);
    narrow(shs, P * 48);
    put(&headN, 4);
    std::vector<hg::HalfNodeRec> hn((size_t)N);
    for (int64_t i = 0; i < N; ++i) {
      const int32_t* n = nodes + i * 7;
      hg::HalfNodeRec& h = hn[(size_t)i];
      h.parent = n[1];
      h.start = n[2];
      h.start_children = n[5];
      h.dccc[0] = (int16_t)n[0];
      h.dccc[1] = (int16_t)n[6];
      h.dccc[2] = (int16_t)n[3];
      h.dccc[3] = (int16_t)n[4];
    }
    put(hn.data(), N * (int64_t)sizeof(hg::HalfNodeRec));
    narrow(boxes, N * 8);
  }
  if (!ok) {
    hg::set_error("%s: short write", filename);
    return HG_ERR_INVALID_ARG;
  }
  return HG_OK;
}

int hg_hier_decode_device(const void* raw_dev, const hg_hier_layout* L, float* pos, float* shs, float* alphas,
                          float* scales, float* rot, int32_t* nodes, float* boxes, void* stream_) {
  if (!raw_dev || !L || L->P < 0 || L->N < 0) {
    hg::set_error("hg_hier_decode_device: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  const char* raw = (const char*)raw_dev;
  const int64_t P = L->P, N = L->N;
  if (P) HG_CUDA_TRY(cudaMemcpyAsync(pos, raw + L->pos, P * 12, cudaMemcpyDeviceToDevice, stream));
  if (!L->compressed) {
    if (P) {
      HG_CUDA_TRY(cudaMemcpyAsync(rot, raw + L->rot, P * 16, cudaMemcpyDeviceToDevice, stream));
      HG_CUDA_TRY(cudaMemcpyAsync(scales, raw + L->scale, P * 12, cudaMemcpyDeviceToDevice, stream));
      HG_CUDA_TRY(cudaMemcpyAsync(alphas, raw + L->opacity, P * 4, cudaMemcpyDeviceToDevice, stream));
      HG_CUDA_TRY(cudaMemcpyAsync(shs, raw + L->sh, P * 192, cudaMemcpyDeviceToDevice, stream));
    }
    if (N) {
      HG_CUDA_TRY(cudaMemcpyAsync(nodes, raw + L->nodes, N * 28, cudaMemcpyDeviceToDevice, stream));
      HG_CUDA_TRY(cudaMemcpyAsync(boxes, raw + L->boxes, N * 32, cudaMemcpyDeviceToDevice, stream));
    }
    return HG_OK;
  }
  auto widen = [&](int64_t off, float* dst, int64_t n) {
    if (n) hg::widen_half_kernel<<<hg::grid_for(n), 256, 0, stream>>>((const __half*)(raw + off), dst, n);
  };
  widen(L->rot, rot, P * 4);
  widen(L->scale, scales, P * 3);
  widen(L->opacity, alphas, P);
  widen(L->sh, shs, P * 48);
  widen(L->boxes, boxes, N * 8);
  if (N) hg::widen_nodes_kernel<<<hg::grid_for(N), 256, 0, stream>>>((const uint16_t*)(raw + L->nodes), nodes, N);
  hg::count_launch(5);
  HG_POST_LAUNCH(false, stream, "hier_decode");
  return HG_OK;
}

int hg_hier_encode_device(void* raw_dev, const hg_hier_layout* L, const float* pos, const float* shs,
                          const float* opacities, const float* log_scales, const float* rotations,
                          const int32_t* nodes, const float* boxes, int32_t* overflow_flag, void* stream_) {
  if (!raw_dev || !L || L->P < 0 || L->N < 0 || !overflow_flag) {
    hg::set_error("hg_hier_encode_device: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t stream = (cudaStream_t)stream_;
  char* raw = (char*)raw_dev;
  const int64_t P = L->P, N = L->N;
  const int32_t headP = L->compressed ? -(int32_t)P : (int32_t)P, headN = (int32_t)N;
  HG_CUDA_TRY(cudaMemsetAsync(overflow_flag, 0, 4, stream));
  HG_CUDA_TRY(cudaMemcpyAsync(raw, &headP, 4, cudaMemcpyHostToDevice, stream));
  HG_CUDA_TRY(cudaMemcpyAsync(raw + L->nodes - 4, &headN, 4, cudaMemcpyHostToDevice, stream));
  if (P) HG_CUDA_TRY(cudaMemcpyAsync(raw + L->pos, pos, P * 12, cudaMemcpyDeviceToDevice, stream));
  if (!L->compressed) {
    if (P) {
      HG_CUDA_TRY(cudaMemcpyAsync(raw + L->rot, rotations, P * 16, cudaMemcpyDeviceToDevice, stream));
      HG_CUDA_TRY(cudaMemcpyAsync(raw + L->scale, log_scales, P * 12, cudaMemcpyDeviceToDevice, stream));
      HG_CUDA_TRY(cudaMemcpyAsync(raw + L->opacity, opacities, P * 4, cudaMemcpyDeviceToDevice, stream));
      HG_CUDA_TRY(cudaMemcpyAsync(raw + L->sh, shs, P * 192, cudaMemcpyDeviceToDevice, stream));
    }
    if (N) {
      HG_CUDA_TRY(cudaMemcpyAsync(raw + L->nodes, nodes, N * 28, cudaMemcpyDeviceToDevice, stream));
      HG_CUDA_TRY(cudaMemcpyAsync(raw + L->boxes, boxes, N * 32, cudaMemcpyDeviceToDevice, stream));
    }
    return HG_OK;
  }
  auto narrow = [&](int64_t off, const float* src, int64_t n) {
    if (n) hg::narrow_half_kernel<<<hg::grid_for(n), 256, 0, stream>>>(src, (__half*)(raw + off), n);
  };
  narrow(L->rot, rotations, P * 4);
  narrow(L->scale, log_scales, P * 3);
  narrow(L->opacity, opacities, P);
  narrow(L->sh, shs, P * 48);
  narrow(L->boxes, boxes, N * 8);
  if (N)
    hg::narrow_nodes_kernel<<<hg::grid_for(N), 256, 0, stream>>>(nodes, (uint16_t*)(raw + L->nodes), N, overflow_flag);
  hg::count_launch(5);
  HG_POST_LAUNCH(false, stream, "hier_encode");
  return HG_OK;
}

int64_t hg_expand_to_target(const int32_t* nodes, int64_t N, int32_t target, int32_t* out, int64_t capacity) {
  if (!nodes || N <= 0 || capacity < 0 || (capacity && !out)) {
    hg::set_error("hg_expand_to_target: bad argument");
    return -1;
  }
  // explicit stack, children pushed in reverse so that they are visited in the reference's recursion order
  std::vector<int32_t> stack;
  stack.push_back(0);
  int64_t count = 0;
  auto emit = [&](int32_t v) {
    if (count < capacity) out[count] = v;
    ++count;
  };
  while (!stack.empty()) {
    const int32_t id = stack.back();
    stack.pop_back();
    if (id < 0 || id >= N) {
      hg::set_error("hg_expand_to_target: node index %d out of range", id);
      return -1;
    }
    const int32_t* n = nodes + (int64_t)id * 7;
    const int32_t depth = n[0], start = n[2], leafs = n[3], merged = n[4], first_child = n[5], children = n[6];
    for (int32_t i = 0; i < leafs; ++i) emit(start + i);
    if (depth <= target) {
      for (int32_t i = 0; i < merged; ++i) emit(start + leafs + i);
    } else {
      for (int32_t i = children - 1; i >= 0; --i) stack.push_back(first_child + i);
    }
  }
  return count;
}

}  // extern "C"
