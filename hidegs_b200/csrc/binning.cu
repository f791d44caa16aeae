// binning.cu — tile-instance generation, sort and per-tile ranges.
//
// Replaces, in order, cub::DeviceScan::InclusiveSum (rasterizer_impl.cu:321), duplicateWithKeys (:70-115),
// cub::DeviceRadixSort::SortPairs on the 64-bit tile|depth key (:357-362, bits [0, 32+getHigherMsb(tiles))) and
// identifyTileRanges (:120-142) with its memset (:364).
//
// The reference sorts R tile instances (R ~ 5.4 M at config 2) on 45..47-bit keys: six 8-bit radix passes over
// 12-byte pairs.  The same order is produced here with far less traffic by sorting in two stages:
//   1. the P slots are sorted ONCE by depth (32-bit keys, stable; culled slots carry 0xFFFFFFFF and end up last);
//      this runs while the host waits for num_rendered;
//   2. tile instances are emitted in that depth order (y-major / x-minor inside a splat, as duplicateWithKeys does)
//      with the tile id as their only key and sorted STABLY on getHigherMsb(tiles) bits (13 at 1080p: two passes
//      over 8-byte pairs).
// A stable LSD radix sort orders by (tile, depth, emission order); emission order among equal (tile, depth) is the
// ascending slot index in both schemes (stage 1 is stable over the slot index; a splat appears at most once per
// tile), so point_list and the tile ranges are bit-identical to the reference's, ties included.  The 64-bit keys
// themselves are only reconstructed on request (hg_raster_debug_keys) for the parity tests.
#include "common.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

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

// Reference emission (duplicateWithKeys, rasterizer_impl.cu:70-115): 64-bit keys in ascending slot order.  Used
// only by hg_raster_debug_keys.
__global__ void __launch_bounds__(256)
emit_keys_reference_kernel(const int P, const float* __restrict__ depths,
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
      if (keys) keys[off] = ((uint64_t)(y * grid_x + x) << 32) | dbits;
      if (vals) vals[off] = (uint32_t)idx;
      ++off;
    }
  }
}

__global__ void __launch_bounds__(256)
rebuild_sorted_keys_kernel(const int R, const uint32_t* __restrict__ tiles_sorted,
                           const uint32_t* __restrict__ point_list, const float* __restrict__ depths,
                           uint64_t* __restrict__ keys) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R) return;
  keys[i] = ((uint64_t)tiles_sorted[i] << 32) | (uint64_t)__float_as_uint(depths[point_list[i]]);
}

// tiles_touched in depth order (input of the second scan)
struct TilesInDepthOrder {
  const uint32_t* tiles_touched;
  __host__ __device__ __forceinline__ uint32_t operator()(const uint32_t& slot) const { return tiles_touched[slot]; }
};

// Stage 2 emission: the i-th nearest splat writes its tile ids and its slot index behind those of all nearer
// splats.  Culled slots have no tiles and sit at the end of the order.
__global__ void __launch_bounds__(256)
emit_instances_kernel(const int P, const uint32_t* __restrict__ depth_order,
                      const uint32_t* __restrict__ offsets_sorted, const uint32_t* __restrict__ tiles_touched,
                      const uint2* __restrict__ rects, const uint32_t grid_x, uint32_t* __restrict__ tile_ids,
                      uint32_t* __restrict__ vals) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const uint32_t idx = depth_order[i];
  if (tiles_touched[idx] == 0) return;
  uint32_t off = (i == 0) ? 0u : offsets_sorted[i - 1];
  const uint2 r = rects[idx];
  const uint32_t minx = r.x & 0xffffu, miny = r.x >> 16;
  const uint32_t maxx = r.y & 0xffffu, maxy = r.y >> 16;
  for (uint32_t y = miny; y < maxy; ++y) {
    for (uint32_t x = minx; x < maxx; ++x) {
      tile_ids[off] = y * grid_x + x;
      vals[off] = idx;
      ++off;
    }
  }
}

__global__ void __launch_bounds__(256)
tile_ranges_kernel(const int L, const uint32_t* __restrict__ tiles, uint2* __restrict__ ranges) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= L) return;
  const uint32_t cur = tiles[idx];
  if (idx == 0) {
    ranges[cur].x = 0;
  } else {
    const uint32_t prev = tiles[idx - 1];
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
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, (uint32_t*)nullptr, (int)R);
  return bytes;
}

size_t depth_sort_temp_bytes(int P) {
  size_t a = 0, b = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, a, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr,
                                  (uint32_t*)nullptr, P);
  cub::TransformInputIterator<uint32_t, TilesInDepthOrder, const uint32_t*> it(nullptr, TilesInDepthOrder{nullptr});
  cub::DeviceScan::InclusiveSum(nullptr, b, it, (uint32_t*)nullptr, P);
  return a > b ? a : b;  // the depth-order scan reuses the depth sort's temp storage
}

int launch_scan(const GeomState& g, int P, size_t temp_bytes, cudaStream_t stream, bool debug) {
  HG_CUDA_TRY(cub::DeviceScan::InclusiveSum(g.scan_temp, temp_bytes, g.tiles_touched,
                                            g.point_offsets, P, stream));
  HG_POST_LAUNCH(debug, stream, "scan");
  return HG_OK;
}

int launch_depth_sort(const GeomState& g, int P, size_t temp_bytes, cudaStream_t stream, bool debug) {
  HG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(g.depth_sort_temp, temp_bytes, reinterpret_cast<const uint32_t*>(g.depths),
                                              g.depth_sorted, g.slot_ids, g.depth_order, P, 0, 32, stream));
  count_launch(4);
  HG_POST_LAUNCH(debug, stream, "depth_sort");
  return HG_OK;
}

int launch_binning(const hg_raster_inputs& in, const GeomState& g, const BinState& b,
                   const ImageState& img, const int* radii, int R, dim3 grid, size_t sort_bytes,
                   size_t dtemp, cudaStream_t stream) {
  (void)radii;
  cub::TransformInputIterator<uint32_t, TilesInDepthOrder, const uint32_t*> tiles_sorted(g.depth_order,
                                                                                        TilesInDepthOrder{g.tiles_touched});
  HG_CUDA_TRY(cub::DeviceScan::InclusiveSum(g.depth_sort_temp, dtemp, tiles_sorted, g.offsets_sorted, in.P, stream));
  HG_POST_LAUNCH(in.debug, stream, "scan_depth_order");
  emit_instances_kernel<<<(in.P + 255) / 256, 256, 0, stream>>>(in.P, g.depth_order, g.offsets_sorted, g.tiles_touched,
                                                                g.rects, grid.x, b.keys_unsorted, b.vals_unsorted);
  HG_POST_LAUNCH(in.debug, stream, "emit_instances");

  const int bit = (int)higher_msb(grid.x * grid.y);
  HG_CUDA_TRY(cub::DeviceRadixSort::SortPairs(b.sort_temp, sort_bytes, b.keys_unsorted, b.keys,
                                              b.vals_unsorted, b.vals, R, 0, bit, stream));
  count_launch(1 + (bit + 7) / 8);

  HG_CUDA_TRY(cudaMemsetAsync(img.ranges, 0, (size_t)grid.x * grid.y * sizeof(uint2), stream));
  tile_ranges_kernel<<<(R + 255) / 256, 256, 0, stream>>>(R, b.keys, img.ranges);
  HG_POST_LAUNCH(in.debug, stream, "tile_ranges");
  return HG_OK;
}

int launch_debug_keys(int P, const GeomState& g, const BinState& b, const int* radii, int R,
                      dim3 grid, uint64_t* keys_unsorted, uint32_t* vals_unsorted, uint64_t* keys_sorted,
                      cudaStream_t stream) {
  if (keys_unsorted || vals_unsorted) {
    emit_keys_reference_kernel<<<(P + 255) / 256, 256, 0, stream>>>(P, g.depths, g.point_offsets, g.rects, radii,
                                                                       grid.x, keys_unsorted, vals_unsorted);
    HG_POST_LAUNCH(true, stream, "emit_keys_reference");
  }
  if (keys_sorted) {
    rebuild_sorted_keys_kernel<<<(R + 255) / 256, 256, 0, stream>>>(R, b.keys, b.vals, g.depths, keys_sorted);
    HG_POST_LAUNCH(true, stream, "rebuild_sorted_keys");
  }
  return HG_OK;
}

}  // namespace hg
