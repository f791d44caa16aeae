// binning.cu — tile-instance generation, key sort and per-tile ranges.
//
// Replaces, in order, cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:321),
// duplicateWithKeys (:70-115), cub::DeviceRadixSort::SortPairs on the 64-bit
// tile|depth key (:357-362, bits [0, 32+getHigherMsb(tiles))) and
// identifyTileRanges (:120-142) with its memset (:364).
//
// Results are bit-identical by construction: the same keys are generated in
// the same emission order (ascending slot, y-major / x-minor tiles) and the
// sort is a stable LSD radix sort over the same bit range, so ties keep
// emission order exactly as with the reference.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace hg {

namespace {

// getHigherMsb (rasterizer_impl.cu:35-50): note getHigherMsb(16) == 5.
inline uint32_t higher_msb(uint32_t n) {
  uint32_t msb = sizeof(n) * 4;
  uint32_t step = msb;
  while (step > 1) {
    step /= 2;
    if (n >> msb) msb += step;
    else msb -= step;
  }
  if (n >> msb) msb++;
  return msb;
}

__global__ void __launch_bounds__(256)
emit_keys_kernel(const int P, const float* __restrict__ depths,
                 const uint32_t* __restrict__ offsets, const uint2* __restrict__ rects,
                 const int* __restrict__ radii, const uint32_t grid_x,
                 uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= P) return;
  if (radii[idx] <= 0) return;
  uint32_t off = (idx == 0) ? 0u : offsets[idx - 1];
  const uint2 r = rects[idx];
  const uint32_t minx = r.x & 0xffffu, miny = r.x >> 16;
  const uint32_t maxx = r.y & 0xffffu, maxy = r.y >> 16;
  const uint64_t dbits = (uint64_t)__float_as_uint(depths[idx]);
  for (uint32_t y = miny; y < maxy; ++y) {
    for (uint32_t x = minx; x < maxx; ++x) {
      keys[off] = ((uint64_t)(y * grid_x + x) << 32) | dbits;
      vals[off] = (uint32_t)idx;
      ++off;
    }
  }
}

__global__ void __launch_bounds__(256)
tile_ranges_kernel(const int L, const uint64_t* __restrict__ keys, uint2* __restrict__ ranges) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L) return;
  const uint32_t cur = (uint32_t)(keys[idx] >> 32);
  if (idx == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = (uint32_t)(keys[idx - 1] >> 32);
    if (cur != prev) {
      ranges[prev].y = idx;
      ranges[cur].x = idx;
    }
    // As in the reference the end marker sits inside the else branch: a list
    // with a single instance (L == 1) leaves its tile range at (0, 0).
    if (idx == L - 1) ranges[cur].y = L;
  }
}

}  // namespace

size_t scan_temp_bytes(int P) {
  size_t bytes = 0;
  cub::DeviceScan::InclusiveSum(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, P);
  return bytes;
}

size_t sort_temp_bytes(int64_t R) {
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint64_t*)nullptr, (uint64_t*)nullptr,
                                  (uint32_t*)nullptr, (uint32_t*)nullptr, (int)R);
  return bytes;
}

int launch_scan(const GeomState& g, int P, size_t temp_bytes, cudaStream_t stream, bool debug) {
  HG_CUDA_TRY(cub::DeviceScan::InclusiveSum(g.scan_temp, temp_bytes, g.tiles_touched,
                                            g.point_offsets, P, stream));
  HG_POST_LAUNCH(debug, stream, "scan");
  return HG_OK;
}

int launch_binning(const hg_raster_inputs& in, const GeomState& g, const BinState& b,
                   const ImageState& img, const int* radii, int R, dim3 grid, size_t sort_bytes,
                   cudaStream_t stream) {
  emit_keys_kernel<<<(in.P + 255) / 256, 256, 0, stream>>>(in.P, g.depths, g.point_offsets,
                                                           g.rects, radii, grid.x,
                                                           b.keys_unsorted, b.vals_unsorted);
  HG_POST_LAUNCH(in.debug, stream, "emit_keys");

  const int bit = (int)higher_msb(grid.x * grid.y);
  HG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(b.sort_temp, sort_bytes, b.keys_unsorted, b.keys,
                                              b.vals_unsorted, b.vals, R, 0, 32 + bit, stream));
  count_launch(2 * ((32 + bit + 7) / 8));

  HG_CUDA_TRY(cudaMemsetAsync(img.ranges, 0, (size_t)grid.x * grid.y * sizeof(uint2), stream));
  tile_ranges_kernel<<<(R + 255) / 256, 256, 0, stream>>>(R, b.keys, img.ranges);
  HG_POST_LAUNCH(in.debug, stream, "tile_ranges");
  return HG_OK;
}

}  // namespace hg
