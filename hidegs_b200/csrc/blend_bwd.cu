// blend_bwd.cu — per-tile back-to-front gradient pass of the alpha blend.
//
// Replaces renderCUDA<3,5> backward of the reference
// (cuda_rasterizer/backward.cu:499-772, launched at :883).
//
// Design (B200):
//  * same tiling as the forward (256-thread CTA per 16x16 tile, warp = 8x4
//    sub-tile, register-double-buffered 64-byte record gathers, lane-parallel
//    exact sub-tile culling), traversing the list from the tile's LAST
//    contributor (block max of n_contrib) instead of the end of the range;
//  * the nine blended channels (rgb, 5 geometry channels, inverse depth) share
//    one recurrence: dL/dalpha only needs sum_ch (c_ch - accum_ch) * dL/dch, so
//    each pixel carries ONE scalar accumulator of g = <features, dL/dpixel>
//    instead of nine, which halves the FP32 work per contributing pair;
//  * the reference issues 15 same-address float atomics per (pixel, Gaussian)
//    pair.  Here the 15 per-lane partials are reduced across the warp with a
//    recursive-halving butterfly (16 SHFL + 16 FADD instead of 75 + 75), and
//    the 16 lanes that end up owning one component each issue ONE coalesced
//    RED.ADD.F32 into the Gaussian's 64-byte accumulator row: one L2 atomic
//    transaction per (warp, contributing entry).
#include "common.cuh"

namespace hg {

namespace {

constexpr int kBatch = HG_BLOCK_SIZE;

struct Prefetch {
  int id;
  float4 r0, r1, r2, r3;
  float it, ifrac;
};

__device__ __forceinline__ float cull_tau(float a, float b, float c, float o, bool interp) {
  if (o < 0.00392156862f) return -1.0f;
  const float det = a * c - b * b;
  if (interp || !(det > 0.0f) || !(a > 0.0f) || !(c > 0.0f)) return __int_as_float(0x7f800000);
  return __logf(255.0f * o) * 1.001f + 2e-3f;
}

__device__ __forceinline__ bool may_touch(float mx, float my, float a, float b, float c,
                                          float tau, float x0, float x1, float y0, float y1) {
  const float dx = fminf(fmaxf(mx, x0), x1) - mx;
  const float dy = fminf(fmaxf(my, y0), y1) - my;
  if (!(tau < __int_as_float(0x7f800000))) return true;
  if (tau < 0.0f) return false;
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

// Sum v[0..15] over the 32 lanes.  On return lane L holds in v[0] the total of
// component comp(L) = 8*b4 + 4*b3 + 2*b2 + b1 (b_i = bit i of L); lanes with
// bit 0 clear are the owners.
__device__ __forceinline__ void warp_reduce16(float (&v)[16], int lane) {
  {
    const bool hi = lane & 16;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float send = hi ? v[i] : v[i + 8];
      const float keep = hi ? v[i + 8] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool hi = lane & 8;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float send = hi ? v[i] : v[i + 4];
      const float keep = hi ? v[i + 4] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool hi = lane & 4;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float send = hi ? v[i] : v[i + 2];
      const float keep = hi ? v[i + 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
  }
  {
    const bool hi = lane & 2;
    const float send = hi ? v[0] : v[1];
    const float keep = hi ? v[1] : v[0];
    v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  }
  v[0] += __shfl_xor_sync(0xffffffffu, v[0], 1);
}

template <bool GEO, bool DEPTH, bool INTERP>
__global__ void __launch_bounds__(HG_BLOCK_SIZE)
blend_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                 const float4* __restrict__ records, const float* __restrict__ ts,
                 const int* __restrict__ kids, const int W, const int H, const float fx,
                 const float fy, const float* __restrict__ bg_color,
                 const float* __restrict__ all_map_pixels, const float* __restrict__ final_Ts,
                 const uint32_t* __restrict__ n_contrib, const float* __restrict__ dL_dpixels,
                 const float* __restrict__ dL_dout_all_maps,
                 const float* __restrict__ dL_dout_plane_depths,
                 const float* __restrict__ dL_invdepths, float* __restrict__ accum) {
  __shared__ float4 s_a[kBatch];
  __shared__ float4 s_b[kBatch];
  __shared__ float4 s_c[kBatch];
  __shared__ float4 s_d[GEO ? kBatch : 1];
  __shared__ float s_e[GEO ? kBatch : 1];
  __shared__ float2 s_i[INTERP ? kBatch : 1];
  __shared__ int s_id[kBatch];
  __shared__ int s_max[HG_BLOCK_SIZE / 32];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = blockIdx.y * gridDim.x + blockIdx.x;
  const int wx0 = blockIdx.x * HG_BLOCK_X + (warp & 1) * 8;
  const int wy0 = blockIdx.y * HG_BLOCK_Y + (warp >> 1) * 4;
  const int pxi = wx0 + (lane & 7), pyi = wy0 + (lane >> 3);
  const bool inside = pxi < W && pyi < H;
  const float pixx = (float)pxi, pixy = (float)pyi;
  const float fx0 = (float)wx0, fx1 = (float)(wx0 + 7), fy0 = (float)wy0, fy1 = (float)(wy0 + 3);
  const size_t HW = (size_t)H * W;
  const size_t pix = (size_t)pyi * W + pxi;

  const uint2 range = ranges[tile];
  const int n = (int)(range.y - range.x);

  const float T_final = inside ? final_Ts[pix] : 0.f;
  const int last_contributor = inside ? min((int)n_contrib[pix], n) : 0;

  // Per-pixel upstream gradients w[0..2]=rgb, w[3]=invdepth, w[4..8]=all_map.
  float w[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) w[i] = 0.f;
  float bg_dot = 0.f;
  if (inside) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      w[c] = dL_dpixels[c * HW + pix];
      bg_dot += __ldg(bg_color + c) * w[c];
    }
    if (DEPTH) w[3] = dL_invdepths[pix];
    if (GEO) {
#pragma unroll
      for (int c = 0; c < 5; ++c) w[4 + c] = dL_dout_all_maps[c * HW + pix];
      // Fold dL/dplane_depth into the geometry channels (backward.cu:583-592).
      const float rayx = (float)(((double)pixx - W * 0.5) / (double)fx);
      const float rayy = (float)(((double)pixy - H * 0.5) / (double)fy);
      const float nx = all_map_pixels[pix], ny = all_map_pixels[HW + pix],
                  nz = all_map_pixels[2 * HW + pix];
      const float dist = all_map_pixels[4 * HW + pix];
      const float tmp = (float)((double)(nx * rayx + ny * rayy + nz) + 1.0e-8);
      const float dpd = dL_dout_plane_depths[pix];
      w[8] += (-dpd / tmp);
      w[4] += dpd * (dist / (tmp * tmp) * rayx);
      w[5] += dpd * (dist / (tmp * tmp) * rayy);
      w[6] += dpd * (dist / (tmp * tmp));
    }
  }

  // Tile-wide last contributor: nothing behind it received any weight.
  int wmax = last_contributor;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
  if (lane == 0) s_max[warp] = wmax;
  __syncthreads();
  int n_eff = 0;
#pragma unroll
  for (int i = 0; i < HG_BLOCK_SIZE / 32; ++i) n_eff = max(n_eff, s_max[i]);
  const int nb = (n_eff + kBatch - 1) / kBatch;

  float T = T_final;
  float acc_g = 0.f, last_g = 0.f, last_alpha = 0.f;
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

  Prefetch pf;
  auto prefetch = [&](int b) {
    const int q = n_eff - 1 - (b * kBatch + tid);  // position in the tile's list
    if (q >= 0) {
      pf.id = (int)__ldg(point_list + range.x + q);
      const float4* r = records + 4 * (size_t)pf.id;
      pf.r0 = __ldg(r);
      pf.r1 = __ldg(r + 1);
      pf.r2 = __ldg(r + 2);
      pf.r3 = __ldg(r + 3);
      if (INTERP) {
        pf.it = __ldg(ts + pf.id);
        pf.ifrac = 1.0f / (float)__ldg(kids + pf.id);
      }
    }
  };
  if (nb > 0) prefetch(0);

  for (int b = 0; b < nb; ++b) {
    __syncthreads();
    const int cnt = min(kBatch, n_eff - b * kBatch);
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
    __syncthreads();
    if (b + 1 < nb) prefetch(b + 1);

    // Position of slot k of this batch: q = n_eff - 1 - (b*kBatch + k).
    const int q_first = n_eff - 1 - b * kBatch;
    if (q_first - (cnt - 1) >= wmax) continue;  // the whole batch lies behind this warp
    for (int c0 = 0; c0 < cnt; c0 += 32) {
      if (q_first - (c0 + 31) >= wmax) continue;
      const int j = c0 + lane;
      bool keep = false;
      if (j < cnt && q_first - j < wmax) {
        const float4 ea = s_a[j];
        const float4 eb = s_b[j];
        keep = may_touch(ea.x, ea.y, ea.z, ea.w, eb.x, eb.z, fx0, fx1, fy0, fy1);
      }
      uint32_t mask = __ballot_sync(0xffffffffu, keep);
      while (mask) {
        const int k = c0 + __ffs(mask) - 1;
        mask &= mask - 1;
        const int q = q_first - k;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0.f;
        bool contributed = false;
        if (q < last_contributor) {
          const float4 ea = s_a[k];
          const float2 eb = *reinterpret_cast<const float2*>(&s_b[k]);
          const float dx = __fsub_rn(ea.x, pixx), dy = __fsub_rn(ea.y, pixy);
          const float quad = __fmaf_rn(dx, __fmul_rn(dx, ea.z), __fmul_rn(dy, __fmul_rn(dy, eb.x)));
          const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, ea.w)));
          if (!(power > 0.0f)) {
            const float G = expf(power);
            const float test_alpha = __fmul_rn(eb.y, G);
            const bool nullalpha = test_alpha > 0.99f;
            const float my_alpha = fminf(0.99f, test_alpha);
            float alpha = my_alpha;
            float opac_mult = 1.0f;
            if (INTERP) {
              const float2 it = s_i[k];
              const float kidsqrt = 1.0f - powf(1.0f - my_alpha, it.y);
              alpha = it.x * my_alpha + (1.0f - it.x) * kidsqrt;
              opac_mult = it.x - powf(1.0f - my_alpha, it.y - 1.0f) * (it.x - 1.0f) * it.y;
            }
            if (!(alpha < 1.0f / 255.0f)) {
              contributed = true;
              const float rinv = 1.0f / (1.0f - alpha);
              T = T * rinv;
              const float weight = alpha * T;
              const float4 ec = s_c[k];
              float g = ec.x * w[0] + ec.y * w[1] + ec.z * w[2];
              v[0] = weight * w[0];
              v[1] = weight * w[1];
              v[2] = weight * w[2];
              if (DEPTH) {
                g += ec.w * w[3];
                v[3] = weight * w[3];
              }
              if (GEO) {
                const float4 ed = s_d[k];
                const float ee = s_e[k];
                g += ed.x * w[4] + ed.y * w[5] + ed.z * w[6] + ed.w * w[7] + ee * w[8];
                v[4] = weight * w[4];
                v[5] = weight * w[5];
                v[6] = weight * w[6];
                v[7] = weight * w[7];
                v[8] = weight * w[8];
              }
              acc_g = last_alpha * last_g + (1.0f - last_alpha) * acc_g;
              last_g = g;
              last_alpha = alpha;
              float dL_dalpha = (g - acc_g) * T;
              dL_dalpha += (-T_final * rinv) * bg_dot;
              if (nullalpha) dL_dalpha = 0.f;
              const float dL_dG = eb.y * dL_dalpha;
              const float gdx = G * dx, gdy = G * dy;
              const float dG_ddelx = -gdx * ea.z - gdy * ea.w;
              const float dG_ddely = -gdy * eb.x - gdx * ea.w;
              v[9] = dL_dG * dG_ddelx * ddelx_dx;
              v[10] = dL_dG * dG_ddely * ddely_dy;
              v[11] = -0.5f * gdx * dx * dL_dG;
              v[12] = -0.5f * gdx * dy * dL_dG;
              v[13] = -0.5f * gdy * dy * dL_dG;
              v[14] = opac_mult * G * dL_dalpha;
            }
          }
        }
        if (__ballot_sync(0xffffffffu, contributed) == 0) continue;
        warp_reduce16(v, lane);
        if ((lane & 1) == 0) {
          const int comp = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 +
                           ((lane >> 1) & 1);
          if (comp < 15) atomicAdd(accum + (size_t)s_id[k] * HG_ACC_FLOATS + comp, v[0]);
        }
      }
    }
  }
}

}  // namespace

int launch_blend_bwd(const hg_raster_inputs& in, const GeomState& g, const BinState& b,
                     const ImageState& img, dim3 grid, float focal_x, float focal_y,
                     const float* all_map_pixels, const float* dL_dpix,
                     const float* dL_dout_all_map, const float* dL_dout_plane_depth,
                     const float* dL_dout_invdepth, float* accum, cudaStream_t stream) {
  const bool interp = in.ts != nullptr && in.kids != nullptr;
  const bool geo = in.render_geo != 0;
  const bool depth = dL_dout_invdepth != nullptr;
#define HG_LAUNCH(G_, D_, I_)                                                                 \
  blend_bwd_kernel<G_, D_, I_><<<grid, HG_BLOCK_SIZE, 0, stream>>>(                           \
      img.ranges, b.vals, g.records, in.ts, in.kids, in.W, in.H, focal_x, focal_y,            \
      in.background, all_map_pixels, img.final_T, img.n_contrib, dL_dpix, dL_dout_all_map,    \
      dL_dout_plane_depth, dL_dout_invdepth, accum)
  if (interp) {
    if (geo && depth) HG_LAUNCH(true, true, true);
    else if (geo) HG_LAUNCH(true, false, true);
    else if (depth) HG_LAUNCH(false, true, true);
    else HG_LAUNCH(false, false, true);
  } else {
    if (geo && depth) HG_LAUNCH(true, true, false);
    else if (geo) HG_LAUNCH(true, false, false);
    else if (depth) HG_LAUNCH(false, true, false);
    else HG_LAUNCH(false, false, false);
  }
#undef HG_LAUNCH
  HG_POST_LAUNCH(in.debug, stream, "blend_bwd");
  return HG_OK;
}

}  // namespace hg
