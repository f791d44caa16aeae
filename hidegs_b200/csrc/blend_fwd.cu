// blend_fwd.cu — per-tile front-to-back alpha blending.
//
// Replaces renderCUDA<3,5> forward of the reference
// (cuda_rasterizer/forward.cu:440-610, launched at :639).
//
// Design (B200):
//  * one 256-thread CTA per 16x16 tile; each WARP owns an 8x4 pixel sub-tile so
//    that all decisions that let the reference skip work per pixel can be taken
//    per warp here;
//  * the tile's sorted list is consumed in batches of 256 entries.  Each thread
//    gathers ONE 64-byte splat record (4 x LDG.128, L2-resident: 64 MB at 1M
//    Gaussians) into registers while the previous batch is being blended
//    (register double buffering), then publishes it to shared memory as SoA;
//  * per warp, 32 entries at a time are tested lane-parallel against the warp's
//    sub-tile with an EXACT conservative bound (minimum of the Gaussian's
//    quadratic form over the 8x4 rectangle vs ln(255*opacity)); only entries
//    that can reach alpha >= 1/255 somewhere in the sub-tile are evaluated.
//    Skipped entries are exactly those the reference would `continue` on for
//    every pixel of the sub-tile, so results are unchanged;
//  * the per-pixel arithmetic of kept entries (power, exp, alpha, T) uses
//    explicit round-to-nearest intrinsics in the reference's compiled order so
//    that alpha thresholds, early termination, n_contrib and out_observe are
//    bit-identical;
//  * out_observe is counted with one warp ballot per entry into a shared
//    counter and flushed with one global atomic per (tile, entry);
//  * early termination: lane (pixel) -> warp (ballot) -> CTA (__syncthreads_or).
#include "common.cuh"

namespace hg {

namespace {

constexpr int kBatch = HG_BLOCK_SIZE;  // 256 entries per staging round

struct Prefetch {
  int id;
  float4 r0, r1, r2, r3;
  float it, ifrac;
};

// Conservative keep/cull threshold for one entry: an upper bound on the value
// q = -power below which alpha can reach 1/255.  +inf = never cull, -1 = always.
__device__ __forceinline__ float cull_tau(float a, float b, float c, float o, bool interp) {
  if (o < 0.00392156862f) return -1.0f;  // alpha <= o < 1/255 for every pixel
  const float det = a * c - b * b;
  if (interp || !(det > 0.0f) || !(a > 0.0f) || !(c > 0.0f)) return __int_as_float(0x7f800000);
  return __logf(255.0f * o) * 1.001f + 2e-3f;
}

// Minimum of q(d) = 0.5*(a dx^2 + c dy^2) + b dx dy over the pixel rectangle
// [x0,x1]x[y0,y1] for a Gaussian centred at (mx,my); returns true if the entry
// may contribute inside the rectangle.
__device__ __forceinline__ bool may_touch(float mx, float my, float a, float b, float c,
                                          float tau, float x0, float x1, float y0, float y1) {
  const float dx = fminf(fmaxf(mx, x0), x1) - mx;  // offset to the nearest point, 0 if inside
  const float dy = fminf(fmaxf(my, y0), y1) - my;
  if (!(tau < __int_as_float(0x7f800000))) return true;
  if (tau < 0.0f) return false;
  // Candidate on the vertical edge through dx (free dy) and on the horizontal
  // edge through dy (free dx).
  const float dy1 = fminf(fmaxf(__fdividef(-b * dx, c), y0 - my), y1 - my);
  const float dx2 = fminf(fmaxf(__fdividef(-b * dy, a), x0 - mx), x1 - mx);
  const float s1 = 0.5f * (a * dx * dx + c * dy1 * dy1);
  const float q1 = s1 + b * dx * dy1 - 1e-5f * s1;
  const float s2 = 0.5f * (a * dx2 * dx2 + c * dy * dy);
  const float q2 = s2 + b * dx2 * dy - 1e-5f * s2;
  float q = (dx != 0.0f) ? q1 : q2;
  if (dx != 0.0f && dy != 0.0f) q = fminf(q1, q2);
  if (dx == 0.0f && dy == 0.0f) q = 0.0f;
  return !(q > tau);
}

template <bool GEO, bool DEPTH, bool INTERP>
__global__ void __launch_bounds__(HG_BLOCK_SIZE)
blend_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ records, const float* __restrict__ ts,
                 const int* __restrict__ kids, const int W, const int H, const float focal_x,
                 const float focal_y, const float cx, const float cy,
                 const float* __restrict__ bg_color, float* __restrict__ final_T,
                 uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                 float* __restrict__ out_invdepth, int* __restrict__ out_observe,
                 float* __restrict__ out_all_map, float* __restrict__ out_plane_depth) {
  __shared__ float4 s_a[kBatch];  // x, y, conic.a, conic.b
  __shared__ float4 s_b[kBatch];  // conic.c, opacity, tau, -
  __shared__ float4 s_c[kBatch];  // r, g, b, 1/depth
  __shared__ float4 s_d[GEO ? kBatch : 1];  // all_map 0..3
  __shared__ float s_e[GEO ? kBatch : 1];   // all_map 4
  __shared__ float2 s_i[INTERP ? kBatch : 1];
  __shared__ int s_id[kBatch];
  __shared__ int s_obs[kBatch];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = blockIdx.y * gridDim.x + blockIdx.x;
  // Warp -> 8x4 sub-tile, lane -> pixel.
  const int wx0 = blockIdx.x * HG_BLOCK_X + (warp & 1) * 8;
  const int wy0 = blockIdx.y * HG_BLOCK_Y + (warp >> 1) * 4;
  const int pxi = wx0 + (lane & 7), pyi = wy0 + (lane >> 3);
  const bool inside = pxi < W && pyi < H;
  const float pixx = (float)pxi, pixy = (float)pyi;
  const float fx0 = (float)wx0, fx1 = (float)(wx0 + 7), fy0 = (float)wy0, fy1 = (float)(wy0 + 3);

  const uint2 range = ranges[tile];
  const int n = (int)(range.y - range.x);
  const int nb = (n + kBatch - 1) / kBatch;

  bool done = !inside;
  float T = 1.0f;
  uint32_t last_contributor = 0;
  float C0 = 0.f, C1 = 0.f, C2 = 0.f, Dinv = 0.f;
  float A0 = 0.f, A1 = 0.f, A2 = 0.f, A3 = 0.f, A4 = 0.f;

  Prefetch pf;
  auto prefetch = [&](int b) {
    const int i = b * kBatch + tid;
    if (i < n) {
      pf.id = (int)__ldg(point_list + range.x + i);
      const float4* r = records + 4 * (size_t)pf.id;
      pf.r0 = __ldg(r);
      pf.r1 = __ldg(r + 1);
      pf.r2 = __ldg(r + 2);
      pf.r3 = __ldg(r + 3);
      if (INTERP) {
        pf.it = __ldg(ts + pf.id);
        pf.ifrac = __frcp_rn((float)__ldg(kids + pf.id));
      }
    }
  };
  if (nb > 0) prefetch(0);

  bool pending_flush = false;
  for (int b = 0; b < nb; ++b) {
    const int any_active = __syncthreads_or(!done);
    if (pending_flush) {
      const int c = s_obs[tid];
      if (c) atomicAdd(out_observe + s_id[tid], c);
      pending_flush = false;
    }
    if (!any_active) break;

    const int cnt = min(kBatch, n - b * kBatch);
    if (tid < cnt) {
      const float a = pf.r0.z, bb = pf.r0.w, c = pf.r1.x, o = pf.r1.y;
      s_a[tid] = pf.r0;
      s_b[tid] = make_float4(c, o, cull_tau(a, bb, c, o, INTERP), 0.f);
      s_c[tid] = make_float4(pf.r1.z, pf.r1.w, pf.r2.x, pf.r2.y);
      if (GEO) {
        s_d[tid] = make_float4(pf.r2.z, pf.r2.w, pf.r3.x, pf.r3.y);
        s_e[tid] = pf.r3.z;
      }
      if (INTERP) s_i[tid] = make_float2(pf.it, pf.ifrac);
      s_id[tid] = pf.id;
    }
    s_obs[tid] = 0;
    __syncthreads();
    if (b + 1 < nb) prefetch(b + 1);
    pending_flush = true;

    if (__ballot_sync(0xffffffffu, !done) == 0) continue;  // whole warp finished
    const uint32_t base = (uint32_t)(b * kBatch);
    for (int c0 = 0; c0 < cnt; c0 += 32) {
      const int j = c0 + lane;
      bool keep = false;
      if (j < cnt) {
        const float4 ea = s_a[j];
        const float4 eb = s_b[j];
        keep = may_touch(ea.x, ea.y, ea.z, ea.w, eb.x, eb.z, fx0, fx1, fy0, fy1);
      }
      uint32_t mask = __ballot_sync(0xffffffffu, keep);
      while (mask) {
        const int k = c0 + __ffs(mask) - 1;
        mask &= mask - 1;
        bool observed = false;
        if (!done) {
          const float4 ea = s_a[k];
          const float2 eb = *reinterpret_cast<const float2*>(&s_b[k]);
          const float dx = __fsub_rn(ea.x, pixx), dy = __fsub_rn(ea.y, pixy);
          // power = -0.5f*(a dx dx + c dy dy) - b dx dy, as compiled (forward.cu:536).
          const float quad = __fmaf_rn(dx, __fmul_rn(dx, ea.z), __fmul_rn(dy, __fmul_rn(dy, eb.x)));
          const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, ea.w)));
          if (!(power > 0.0f)) {
            float alpha = fminf(0.99f, __fmul_rn(eb.y, expf(power)));
            if (INTERP) {
              const float2 it = s_i[k];
              const float kidsqrt = __fsub_rn(1.0f, __powf(__fsub_rn(1.0f, alpha), it.y));
              alpha = __fmaf_rn(alpha, it.x, __fmul_rn(__fsub_rn(1.0f, it.x), kidsqrt));
            }
            if (!(alpha < 1.0f / 255.0f)) {
              const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
              if (test_T < 0.0001f) {
                done = true;
              } else {
                const float4 ec = s_c[k];
                C0 = __fmaf_rn(T, __fmul_rn(alpha, ec.x), C0);
                C1 = __fmaf_rn(T, __fmul_rn(alpha, ec.y), C1);
                C2 = __fmaf_rn(T, __fmul_rn(alpha, ec.z), C2);
                if (DEPTH) Dinv = __fmaf_rn(T, __fmul_rn(alpha, ec.w), Dinv);
                if (GEO) {
                  const float4 ed = s_d[k];
                  const float ee = s_e[k];
                  A0 = __fmaf_rn(T, __fmul_rn(alpha, ed.x), A0);
                  A1 = __fmaf_rn(T, __fmul_rn(alpha, ed.y), A1);
                  A2 = __fmaf_rn(T, __fmul_rn(alpha, ed.z), A2);
                  A3 = __fmaf_rn(T, __fmul_rn(alpha, ed.w), A3);
                  A4 = __fmaf_rn(T, __fmul_rn(alpha, ee), A4);
                }
                observed = T > 0.5f;
                T = test_T;
                last_contributor = base + (uint32_t)k + 1u;
              }
            }
          }
        }
        const uint32_t om = __ballot_sync(0xffffffffu, observed);
        if (om && lane == 0) atomicAdd(&s_obs[k], __popc(om));
      }
      if (__ballot_sync(0xffffffffu, !done) == 0) break;
    }
  }
  if (pending_flush) {
    __syncthreads();
    const int c = s_obs[tid];
    if (c) atomicAdd(out_observe + s_id[tid], c);
  }

  if (inside) {
    const size_t HW = (size_t)H * W;
    const size_t pix = (size_t)pyi * W + pxi;
    final_T[pix] = T;
    n_contrib[pix] = last_contributor;
    out_color[pix] = __fmaf_rn(T, __ldg(bg_color), C0);
    out_color[HW + pix] = __fmaf_rn(T, __ldg(bg_color + 1), C1);
    out_color[2 * HW + pix] = __fmaf_rn(T, __ldg(bg_color + 2), C2);
    if (DEPTH) out_invdepth[pix] = Dinv;
    if (GEO) {
      out_all_map[pix] = A0;
      out_all_map[HW + pix] = A1;
      out_all_map[2 * HW + pix] = A2;
      out_all_map[3 * HW + pix] = A3;
      out_all_map[4 * HW + pix] = A4;
      // plane depth (forward.cu:474,607): float ray, double add/div.
      const float rayx = __fdiv_rn(__fsub_rn(pixx, cx), focal_x);
      const float rayy = __fdiv_rn(__fsub_rn(pixy, cy), focal_y);
      const float den = __fadd_rn(A2, __fmaf_rn(rayx, A0, __fmul_rn(rayy, A1)));
      out_plane_depth[pix] = (float)__ddiv_rn((double)A4, -__dadd_rn((double)den, 1.0e-8));
    } else {
      out_all_map[pix] = 0.f;
      out_all_map[HW + pix] = 0.f;
      out_all_map[2 * HW + pix] = 0.f;
      out_all_map[3 * HW + pix] = 0.f;
      out_all_map[4 * HW + pix] = 0.f;
      out_plane_depth[pix] = 0.f;
    }
  }
}

template <bool GEO, bool DEPTH>
int dispatch(bool interp, dim3 grid, cudaStream_t stream, const uint2* ranges,
             const uint32_t* point_list, const float4* records, const float* ts, const int* kids,
             int W, int H, float fx, float fy, const float* bg, float* final_T,
             uint32_t* n_contrib, float* out_color, float* out_invdepth, int* out_observe,
             float* out_all_map, float* out_plane_depth) {
  const float cx = float(W * 0.5f), cy = float(H * 0.5f);
  if (interp)
    blend_fwd_kernel<GEO, DEPTH, true><<<grid, HG_BLOCK_SIZE, 0, stream>>>(
        ranges, point_list, records, ts, kids, W, H, fx, fy, cx, cy, bg, final_T, n_contrib,
        out_color, out_invdepth, out_observe, out_all_map, out_plane_depth);
  else
    blend_fwd_kernel<GEO, DEPTH, false><<<grid, HG_BLOCK_SIZE, 0, stream>>>(
        ranges, point_list, records, ts, kids, W, H, fx, fy, cx, cy, bg, final_T, n_contrib,
        out_color, out_invdepth, out_observe, out_all_map, out_plane_depth);
  return 0;
}

}  // namespace

int launch_blend_fwd(const hg_raster_inputs& in, const GeomState& g, const BinState& b,
                     const ImageState& img, dim3 grid, float focal_x, float focal_y,
                     float* out_color, float* out_invdepth, int* out_observe, float* out_all_map,
                     float* out_plane_depth, bool empty_scene, cudaStream_t stream) {
  const size_t HW = (size_t)in.W * in.H;
  if (empty_scene) {
    // The reference returns before rendering when nothing is visible
    // (rasterizer_impl.cu:332-333): every output keeps its zero fill.
    HG_CUDA_TRY(cudaMemsetAsync(out_color, 0, 3 * HW * sizeof(float), stream));
    if (out_invdepth) HG_CUDA_TRY(cudaMemsetAsync(out_invdepth, 0, HW * sizeof(float), stream));
    HG_CUDA_TRY(cudaMemsetAsync(out_all_map, 0, 5 * HW * sizeof(float), stream));
    HG_CUDA_TRY(cudaMemsetAsync(out_plane_depth, 0, HW * sizeof(float), stream));
    HG_CUDA_TRY(cudaMemsetAsync(img.final_T, 0, HW * sizeof(float), stream));
    HG_CUDA_TRY(cudaMemsetAsync(img.n_contrib, 0, HW * sizeof(uint32_t), stream));
    return HG_OK;
  }
  const bool interp = in.ts != nullptr && in.kids != nullptr;
  const bool geo = in.render_geo != 0;
  const bool depth = out_invdepth != nullptr;
#define HG_ARGS                                                                              \
  interp, grid, stream, img.ranges, b.vals, g.records, in.ts, in.kids, in.W, in.H, focal_x, \
      focal_y, in.background, img.final_T, img.n_contrib, out_color, out_invdepth,           \
      out_observe, out_all_map, out_plane_depth
  if (geo && depth) dispatch<true, true>(HG_ARGS);
  else if (geo) dispatch<true, false>(HG_ARGS);
  else if (depth) dispatch<false, true>(HG_ARGS);
  else dispatch<false, false>(HG_ARGS);
#undef HG_ARGS
  HG_POST_LAUNCH(in.debug, stream, "blend_fwd");
  return HG_OK;
}

}  // namespace hg
