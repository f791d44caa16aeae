// tile_instances.cuh — warp-cooperative enumeration of tile instances.
//
// duplicateWithKeys (rasterizer_impl.cu:70-115) walks a splat's tile rectangle in ONE thread: a splat that covers
// 2000 tiles keeps a lane busy for 2000 iterations while the other 31 idle.  Here the instances of a warp's 32 splats
// are numbered 0..total-1 (inclusive scan of the per-lane counts) and handed out 32 at a time, one per lane, whatever
// the rectangle sizes: lane l of step j0 finds the owner of instance j0 + l by a 5-step binary search over the scan
// (shuffles) and turns the local index into a tile id (y-major / x-minor, as the reference).
#pragma once
#include <stdint.h>

namespace hg {

struct WarpInstances {
  uint32_t incl, cnt, total, origin, width;
  int lane;
  // cnt: instances of this lane's splat (0 = none); origin: minx | miny << 16; width: maxx - minx
  __device__ __forceinline__ WarpInstances(uint32_t cnt_, uint32_t origin_, uint32_t width_, int lane_)
      : cnt(cnt_), origin(origin_), width(width_), lane(lane_) {
    incl = cnt_;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane_ >= o) incl += v;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
  }
  // Instance j0 + lane: returns whether it exists; owner = lane whose splat it belongs to, tile = its tile id.
  // Must be called by all 32 lanes.
  __device__ __forceinline__ bool at(uint32_t j0, uint32_t grid_x, int& owner, uint32_t& tile) const {
    const uint32_t j = j0 + (uint32_t)lane;
    int lo = 0;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      const uint32_t v = __shfl_sync(0xffffffffu, incl, lo + s - 1);
      if (v <= j) lo += s;
    }
    lo = lo > 31 ? 31 : lo;
    const uint32_t o_incl = __shfl_sync(0xffffffffu, incl, lo);
    const uint32_t o_cnt = __shfl_sync(0xffffffffu, cnt, lo);
    const uint32_t o_org = __shfl_sync(0xffffffffu, origin, lo);
    const uint32_t o_w = __shfl_sync(0xffffffffu, width, lo);
    owner = lo;
    const bool on = j < total;
    tile = 0;
    if (on) {
      const uint32_t k = j - (o_incl - o_cnt);
      const uint32_t ty = k / o_w;
      const uint32_t tx = k - ty * o_w;
      tile = ((o_org >> 16) + ty) * grid_x + (o_org & 0xffffu) + tx;
    }
    return on;
  }
};

}  // namespace hg
