// blend_fwd.cu — per-tile front-to-back alpha blending.
//
// Replaces renderCUDA<3,5> forward of the reference
// (cuda_rasterizer/forward.cu:440-610, launched at :639).
//
// Design (B200), blend_fwd2_kernel:
//  * one 128-thread CTA per 16x16 tile; each WARP owns an 8x8 pixel sub-tile, two pixels per lane (rows r, r+4), so
//    that all decisions that let the reference skip work per pixel can be taken per warp and half here;
//  * the tile's sorted list is consumed in batches of 128 entries.  Each thread gathers ONE 64-byte splat record
//    (4 x LDG.128, L2-resident: 64 MB at 1M Gaussians) into registers while the previous batch is being blended
//    (register double buffering), then publishes it to shared memory as one 80-byte staging record
//    (blend_common.cuh): one address per entry in the blend loop, every field at an immediate offset;
//  * per warp and half, 32 entries at a time are tested lane-parallel against the 8x4 rectangle with an EXACT
//    conservative bound (minimum of the Gaussian's quadratic form over the rectangle vs ln(255*opacity)); only entries
//    that can reach alpha >= 1/255 somewhere in it are evaluated.  Skipped entries are exactly those the reference
//    would `continue` on for every pixel of the rectangle, so results are unchanged;
//  * the per-pixel arithmetic of kept entries (power, exp, alpha, T) uses explicit round-to-nearest intrinsics in the
//    reference's compiled order so that alpha thresholds, early termination, n_contrib and out_observe are
//    bit-identical; when an entry reaches both halves the two pixels of a lane run through a select-predicated,
//    branch-free twin of the per-pixel code in ONE basic block (their dependency chains interleave);
//  * out_observe is counted with one warp ballot per entry into a shared counter and flushed with one global atomic
//    per (tile, entry);
//  * early termination: lane (pixel) -> warp (ballot) -> CTA (__syncthreads_or).
// Measured and dropped: reading the staged entry through a shared-space address kept in a register (volatile ld.shared:
// the compiler otherwise re-derives the shared window with S2UR + ULEA per entry) and a single-lane red.shared for the
// observe count: 8 fewer instructions per (warp, entry), 0.512 vs 0.507 ms — the ordered volatile loads cost more.
// Removed after measurement: the one-pixel-per-lane first version (0.66 vs 0.55 ms at config 2) and a variant that
// staged every batch with one cp.async.bulk per thread into an mbarrier-released ring (SASS UBLKCP.S.G +
// SYNCS.ARRIVE.TRANS64, dump kept in profiles/r02_sass_excerpts.md): 0.674 vs 0.661 ms for the register-double-buffered
// gather of the same build — the gathers hit L2 and are hidden already.
#include "blend_common.cuh"

#include <cstdlib>

namespace hg {

namespace {

// Warp per 8x8 sub-tile, TWO pixels per lane (rows r and r+4), 128-thread CTA per tile.  The loop control, the broadcast
// reads of the staged entry and the observe bookkeeping are shared by the two pixels; each half of the sub-tile has its
// own exact cull test.
constexpr int kThreadsB = 128;
constexpr int kBatchB = kThreadsB;

struct FwdPixel {
  float T, C0, C1, C2, Dinv, A0, A1, A2, A3, A4;
  uint32_t last_contributor;
};

// The blend state of a lane's two pixels, kept as register PAIRS (.x: pixel A in rows 0..3 of the sub-tile, .y: pixel B
// four rows below) so that the evaluation of an entry that reaches both halves runs on the packed FP32 instructions.
struct FwdPair {
  float2 T, C0, C1, C2, Dinv, A0, A1, A2, A3, A4;
  uint32_t lastA, lastB;
  bool doneA, doneB;
};

template <int HALF>
__device__ __forceinline__ float& half_of(float2& v) { return HALF ? v.y : v.x; }

template <int HALF>
__device__ __forceinline__ FwdPixel pixel_of(const FwdPair& s) {
  return HALF ? FwdPixel{s.T.y, s.C0.y, s.C1.y, s.C2.y, s.Dinv.y, s.A0.y, s.A1.y, s.A2.y, s.A3.y, s.A4.y, s.lastB}
              : FwdPixel{s.T.x, s.C0.x, s.C1.x, s.C2.x, s.Dinv.x, s.A0.x, s.A1.x, s.A2.x, s.A3.x, s.A4.x, s.lastA};
}

// An entry that reaches BOTH halves: the per-pixel code of fwd_pair (below) for the lane's two pixels at once, branch
// free (a pixel that does not contribute adds fma(0, feature, C) = C) and on FADD2 / FMUL2 / FFMA2: the same operations
// in the same order with the same roundings per half — alpha thresholds, early termination, n_contrib and out_observe
// keep their bits — in roughly half the issue slots (15 packed + 10 scalar FP32 instructions for the pair up to alpha
// instead of 44, 9 FFMA2 instead of 18 FFMA for the nine blended channels).  dx is shared by the two pixels (same column);
// -(dy * (dx * b)) is formed as dy * (dx * -b) (round-to-nearest is sign-symmetric) because the packed FMA has no
// operand negation.
template <bool GEO, bool DEPTH, bool INTERP>
__device__ __forceinline__ void fwd_pair_flat2(FwdPair& s, const float4* __restrict__ e, const float4 ea, const float2 eb,
                                               float pixx, float2 neg_pixy, uint32_t index1, bool& obsA, bool& obsB) {
  const float dx = __fsub_rn(ea.x, pixx);
  const float dxa = __fmul_rn(dx, ea.z);
  const float ndxb = __fmul_rn(dx, -ea.w);
  const float2 dy = __fadd2_rn(bc2(ea.y), neg_pixy);
  const float2 quad = __ffma2_rn(bc2(dx), bc2(dxa), __fmul2_rn(dy, __fmul2_rn(dy, bc2(eb.x))));
  const float2 power = __ffma2_rn(quad, bc2(-0.5f), __fmul2_rn(dy, bc2(ndxb)));
  float2 alpha = __fmul2_rn(bc2(eb.y), expf_pair(power));
  alpha.x = fminf(0.99f, alpha.x);
  alpha.y = fminf(0.99f, alpha.y);
  if (INTERP) {
    const float4 e4 = e[4];
    const float kA = __fsub_rn(1.0f, __powf(__fsub_rn(1.0f, alpha.x), e4.z));
    const float kB = __fsub_rn(1.0f, __powf(__fsub_rn(1.0f, alpha.y), e4.z));
    const float om = __fsub_rn(1.0f, e4.y);
    alpha.x = __fmaf_rn(alpha.x, e4.y, __fmul_rn(om, kA));
    alpha.y = __fmaf_rn(alpha.y, e4.y, __fmul_rn(om, kB));
  }
  const bool okA = !s.doneA && !(power.x > 0.0f) && !(alpha.x < 1.0f / 255.0f);
  const bool okB = !s.doneB && !(power.y > 0.0f) && !(alpha.y < 1.0f / 255.0f);
  const float2 test_T = __fmul2_rn(s.T, __ffma2_rn(alpha, bc2(-1.0f), bc2(1.0f)));  // T * (1 - alpha)
  const bool termA = okA && test_T.x < 0.0001f, termB = okB && test_T.y < 0.0001f;
  const bool contribA = okA && !termA, contribB = okB && !termB;
  s.doneA = s.doneA || termA;
  s.doneB = s.doneB || termB;
  float2 wgt = __fmul2_rn(alpha, s.T);
  wgt.x = contribA ? wgt.x : 0.0f;
  wgt.y = contribB ? wgt.y : 0.0f;
  const float4 ec = e[2];
  s.C0 = __ffma2_rn(wgt, bc2(ec.x), s.C0);
  s.C1 = __ffma2_rn(wgt, bc2(ec.y), s.C1);
  s.C2 = __ffma2_rn(wgt, bc2(ec.z), s.C2);
  if (DEPTH) s.Dinv = __ffma2_rn(wgt, bc2(ec.w), s.Dinv);
  if (GEO) {
    const float4 ed = e[3];
    const float ee = e[4].x;
    s.A0 = __ffma2_rn(wgt, bc2(ed.x), s.A0);
    s.A1 = __ffma2_rn(wgt, bc2(ed.y), s.A1);
    s.A2 = __ffma2_rn(wgt, bc2(ed.z), s.A2);
    s.A3 = __ffma2_rn(wgt, bc2(ed.w), s.A3);
    s.A4 = __ffma2_rn(wgt, bc2(ee), s.A4);
  }
  obsA = contribA && s.T.x > 0.5f;
  obsB = contribB && s.T.y > 0.5f;
  s.T.x = contribA ? test_T.x : s.T.x;
  s.T.y = contribB ? test_T.y : s.T.y;
  s.lastA = contribA ? index1 : s.lastA;
  s.lastB = contribB ? index1 : s.lastB;
}

// An entry that reaches one half only: the reference's per-pixel code with its early-outs.
template <bool GEO, bool DEPTH, bool INTERP, int HALF>
__device__ __forceinline__ bool fwd_pair(FwdPair& s, const float4* __restrict__ e, const float4 ea, const float2 eb,
                                         float pixx, float2 neg_pixy, uint32_t index1) {
  bool observed = false;
  bool& done = HALF ? s.doneB : s.doneA;
  if (!done) {
    const float dx = __fsub_rn(ea.x, pixx), dy = __fadd_rn(ea.y, HALF ? neg_pixy.y : neg_pixy.x);  // y - pixy
    const float quad = __fmaf_rn(dx, __fmul_rn(dx, ea.z), __fmul_rn(dy, __fmul_rn(dy, eb.x)));
    const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, ea.w)));
    if (!(power > 0.0f)) {
      float alpha = fminf(0.99f, __fmul_rn(eb.y, expf(power)));
      if (INTERP) {
        const float4 e4 = e[4];
        const float kidsqrt = __fsub_rn(1.0f, __powf(__fsub_rn(1.0f, alpha), e4.z));
        alpha = __fmaf_rn(alpha, e4.y, __fmul_rn(__fsub_rn(1.0f, e4.y), kidsqrt));
      }
      if (!(alpha < 1.0f / 255.0f)) {
        float& T = half_of<HALF>(s.T);
        const float test_T = __fmul_rn(T, __fsub_rn(1.0f, alpha));
        if (test_T < 0.0001f) {
          done = true;
        } else {
          const float wgt = __fmul_rn(alpha, T);
          const float4 ec = e[2];
          half_of<HALF>(s.C0) = __fmaf_rn(wgt, ec.x, half_of<HALF>(s.C0));
          half_of<HALF>(s.C1) = __fmaf_rn(wgt, ec.y, half_of<HALF>(s.C1));
          half_of<HALF>(s.C2) = __fmaf_rn(wgt, ec.z, half_of<HALF>(s.C2));
          if (DEPTH) half_of<HALF>(s.Dinv) = __fmaf_rn(wgt, ec.w, half_of<HALF>(s.Dinv));
          if (GEO) {
            const float4 ed = e[3];
            const float ee = e[4].x;
            half_of<HALF>(s.A0) = __fmaf_rn(wgt, ed.x, half_of<HALF>(s.A0));
            half_of<HALF>(s.A1) = __fmaf_rn(wgt, ed.y, half_of<HALF>(s.A1));
            half_of<HALF>(s.A2) = __fmaf_rn(wgt, ed.z, half_of<HALF>(s.A2));
            half_of<HALF>(s.A3) = __fmaf_rn(wgt, ed.w, half_of<HALF>(s.A3));
            half_of<HALF>(s.A4) = __fmaf_rn(wgt, ee, half_of<HALF>(s.A4));
          }
          observed = T > 0.5f;
          T = test_T;
          (HALF ? s.lastB : s.lastA) = index1;
        }
      }
    }
  }
  return observed;
}

template <bool GEO, bool DEPTH>
__device__ __forceinline__ void write_pixel(const FwdPixel& s, bool inside, size_t pix, size_t HW, float pixx, float pixy,
                                            float cx, float cy, float focal_x, float focal_y,
                                            const float* __restrict__ bg_color, float* __restrict__ final_T,
                                            uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                                            float* __restrict__ out_invdepth, float* __restrict__ out_all_map,
                                            float* __restrict__ out_plane_depth) {
  if (!inside) return;
  final_T[pix] = s.T;
  n_contrib[pix] = s.last_contributor;
  out_color[pix] = __fmaf_rn(s.T, __ldg(bg_color), s.C0);
  out_color[HW + pix] = __fmaf_rn(s.T, __ldg(bg_color + 1), s.C1);
  out_color[2 * HW + pix] = __fmaf_rn(s.T, __ldg(bg_color + 2), s.C2);
  if (DEPTH) out_invdepth[pix] = s.Dinv;
  if (GEO) {
    out_all_map[pix] = s.A0;
    out_all_map[HW + pix] = s.A1;
    out_all_map[2 * HW + pix] = s.A2;
    out_all_map[3 * HW + pix] = s.A3;
    out_all_map[4 * HW + pix] = s.A4;
    // plane depth (forward.cu:474,607): float ray, double add/div.
    const float rayx = __fdiv_rn(__fsub_rn(pixx, cx), focal_x);
    const float rayy = __fdiv_rn(__fsub_rn(pixy, cy), focal_y);
    const float den = __fadd_rn(s.A2, __fmaf_rn(rayx, s.A0, __fmul_rn(rayy, s.A1)));
    out_plane_depth[pix] = (float)__ddiv_rn((double)s.A4, -__dadd_rn((double)den, 1.0e-8));
  } else {
    out_all_map[pix] = 0.f;
    out_all_map[HW + pix] = 0.f;
    out_all_map[2 * HW + pix] = 0.f;
    out_all_map[3 * HW + pix] = 0.f;
    out_all_map[4 * HW + pix] = 0.f;
    out_plane_depth[pix] = 0.f;
  }
}

template <bool GEO, bool DEPTH, bool INTERP>
__global__ void __launch_bounds__(kThreadsB)
blend_fwd2_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                  const float4* __restrict__ records, const float* __restrict__ ts,
                  const int* __restrict__ kids, const int W, const int H, const float focal_x,
                  const float focal_y, const float cx, const float cy,
                  const float* __restrict__ bg_color, float* __restrict__ final_T,
                  uint32_t* __restrict__ n_contrib, float* __restrict__ out_color,
                  float* __restrict__ out_invdepth, int* __restrict__ out_observe,
                  float* __restrict__ out_all_map, float* __restrict__ out_plane_depth) {
  __shared__ float4 s_rec[kBatchB * kRecQuads];
  __shared__ int s_obs[kBatchB];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = blockIdx.y * gridDim.x + blockIdx.x;
  const int wx0 = blockIdx.x * HG_BLOCK_X + (warp & 1) * 8;
  const int wy0 = blockIdx.y * HG_BLOCK_Y + (warp >> 1) * 8;
  const int pxi = wx0 + (lane & 7), pyA = wy0 + (lane >> 3), pyB = pyA + 4;
  const bool insideA = pxi < W && pyA < H, insideB = pxi < W && pyB < H;
  const float pixx = (float)pxi, pixyA = (float)pyA, pixyB = (float)pyB;
  const float fx0 = (float)wx0, fx1 = (float)(wx0 + 7);
  // (the packed two-half test of the backward, may_touch2, costs the forward 8 registers = one resident CTA: 0.521 vs 0.507 ms)
  const float fyA0 = (float)wy0, fyA1 = (float)(wy0 + 3), fyB0 = (float)(wy0 + 4), fyB1 = (float)(wy0 + 7);

  const uint2 range = ranges[tile];
  const int n = (int)(range.y - range.x);
  const int nb = (n + kBatchB - 1) / kBatchB;

  const float2 z2 = make_float2(0.f, 0.f);
  FwdPair S{make_float2(1.0f, 1.0f), z2, z2, z2, z2, z2, z2, z2, z2, z2, 0u, 0u, !insideA, !insideB};
  const float2 neg_pixy = make_float2(-pixyA, -pixyB);

  Prefetch pf;
  auto prefetch = [&](int b) {
    const int i = b * kBatchB + tid;
    if (i < n) gather_record<INTERP>(pf, point_list, records, ts, kids, range.x + i);
  };
  if (nb > 0) prefetch(0);

  bool pending_flush = false;
  for (int b = 0; b < nb; ++b) {
    const int any_active = __syncthreads_or(!(S.doneA && S.doneB));
    if (pending_flush) {
      const int c = s_obs[tid];
      if (c) atomicAdd(out_observe + __float_as_int(s_rec[kRecQuads * tid + 1].w), c);
      pending_flush = false;
    }
    if (!any_active) break;

    const int cnt = min(kBatchB, n - b * kBatchB);
    if (tid < cnt) stage_record<GEO, INTERP>(s_rec, tid, pf);
    s_obs[tid] = 0;
    __syncthreads();
    if (b + 1 < nb) prefetch(b + 1);
    pending_flush = true;

    uint32_t liveA = __ballot_sync(0xffffffffu, !S.doneA), liveB = __ballot_sync(0xffffffffu, !S.doneB);
    if ((liveA | liveB) == 0) continue;  // whole warp finished
    const uint32_t base = (uint32_t)(b * kBatchB);
    for (int c0 = 0; c0 < cnt; c0 += 32) {
      const int j = c0 + lane;
      bool keepA = false, keepB = false;
      if (j < cnt) {
        const float4 ea = s_rec[kRecQuads * j];
        const float4 eb = s_rec[kRecQuads * j + 1];
        keepA = liveA != 0 && may_touch(ea.x, ea.y, ea.z, ea.w, eb.x, eb.z, fx0, fx1, fyA0, fyA1);
        keepB = liveB != 0 && may_touch(ea.x, ea.y, ea.z, ea.w, eb.x, eb.z, fx0, fx1, fyB0, fyB1);
      }
      const uint32_t maskA = __ballot_sync(0xffffffffu, keepA), maskB = __ballot_sync(0xffffffffu, keepB);
      uint32_t mask = maskA | maskB;
      while (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        const int k = c0 + bit;
        const float4* e = s_rec + kRecQuads * k;
        const float4 ea = e[0];
        const float2 eb = *reinterpret_cast<const float2*>(e + 1);
        const uint32_t index1 = base + (uint32_t)k + 1u;
        bool obsA = false, obsB = false;
        const bool hasA = (maskA >> bit) & 1u, hasB = (maskB >> bit) & 1u;
        if (hasA && hasB) {
          fwd_pair_flat2<GEO, DEPTH, INTERP>(S, e, ea, eb, pixx, neg_pixy, index1, obsA, obsB);
        } else if (hasA) {
          obsA = fwd_pair<GEO, DEPTH, INTERP, 0>(S, e, ea, eb, pixx, neg_pixy, index1);
        } else {
          obsB = fwd_pair<GEO, DEPTH, INTERP, 1>(S, e, ea, eb, pixx, neg_pixy, index1);
        }
        const uint32_t oa = __ballot_sync(0xffffffffu, obsA), ob = __ballot_sync(0xffffffffu, obsB);
        if ((oa | ob) && lane == 0) atomicAdd(&s_obs[k], __popc(oa) + __popc(ob));
      }
      liveA = __ballot_sync(0xffffffffu, !S.doneA);
      liveB = __ballot_sync(0xffffffffu, !S.doneB);
      if ((liveA | liveB) == 0) break;
    }
  }
  if (pending_flush) {
    __syncthreads();
    const int c = s_obs[tid];
    if (c) atomicAdd(out_observe + __float_as_int(s_rec[kRecQuads * tid + 1].w), c);
  }

  const size_t HW = (size_t)H * W;
  write_pixel<GEO, DEPTH>(pixel_of<0>(S), insideA, (size_t)pyA * W + pxi, HW, pixx, pixyA, cx, cy, focal_x, focal_y, bg_color, final_T,
                          n_contrib, out_color, out_invdepth, out_all_map, out_plane_depth);
  write_pixel<GEO, DEPTH>(pixel_of<1>(S), insideB, (size_t)pyB * W + pxi, HW, pixx, pixyB, cx, cy, focal_x, focal_y, bg_color, final_T,
                          n_contrib, out_color, out_invdepth, out_all_map, out_plane_depth);
}

template <bool GEO, bool DEPTH>
int dispatch(bool interp, dim3 grid, cudaStream_t stream, const uint2* ranges,
             const uint32_t* point_list, const float4* records, const float* ts, const int* kids,
             int W, int H, float fx, float fy, const float* bg, float* final_T,
             uint32_t* n_contrib, float* out_color, float* out_invdepth, int* out_observe,
             float* out_all_map, float* out_plane_depth) {
  const float cx = float(W * 0.5f), cy = float(H * 0.5f);
  if (interp)
    blend_fwd2_kernel<GEO, DEPTH, true><<<grid, kThreadsB, 0, stream>>>(
        ranges, point_list, records, ts, kids, W, H, fx, fy, cx, cy, bg, final_T, n_contrib,
        out_color, out_invdepth, out_observe, out_all_map, out_plane_depth);
  else
    blend_fwd2_kernel<GEO, DEPTH, false><<<grid, kThreadsB, 0, stream>>>(
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
