// blend_bwd.cu — per-tile back-to-front gradient pass of the alpha blend.
//
// Replaces renderCUDA<3,5> backward of the reference
// (cuda_rasterizer/backward.cu:499-772, launched at :883).
//
// One kernel, blend_bwd3_kernel: 128-thread CTA per 16x16 tile, a warp per 8x8 sub-tile, two pixels per lane; every
// warp walks the tile's list on its own (no CTA barrier in the loop) back to front from ITS last contributor, with
// register-double-buffered gathers of the 64-byte splat records and a lane-parallel exact sub-tile cull.  The nine
// blended channels (rgb, inverse depth, 5 geometry channels) share ONE scalar recurrence: dL/dalpha only needs
// sum_ch (c_ch - accum_ch) * dL/dch, so each pixel carries one accumulator of g = <features, dL/dpixel> instead of
// nine.  The reduction of the per-pair gradients over the warp's 64 pixels runs on the tensor cores (3xTF32 mma.sync
// against per-pixel constant tiles, see below).  The reference issues 15 same-address float atomics per (pixel,
// Gaussian) pair; here one 64-byte accumulator row per Gaussian receives a few vector REDs per (warp, 16 entries).
// (Rounds 1-2 also carried a one-pixel-per-lane kernel and a two-pixel kernel with a 16-value shuffle butterfly per
// entry, 1.6 and 1.34 ms at config 2 against 1.09 ms; both were removed once this kernel covered the hierarchy
// interpolation as well.)
#include "blend_common.cuh"

#include <cstdlib>
#include <mutex>

namespace hg {

namespace {

// Two pixels per lane (rows r and r + 4 of the warp's 8x8 sub-tile), 128-thread CTA per 16x16 tile.
constexpr int kThreadsB = 128;


struct PixelState {
  float w[9];
  float bgT, T, acc_g, last_g, last_alpha;
  int last_contributor;
};

// The recurrence state of a lane's two pixels as register PAIRS (.x: pixel A, .y: pixel B four rows below), so that an
// entry that reaches both halves is evaluated with the packed FP32 instructions (blend_common.cuh).
struct PairState {
  float2 w[9];
  float2 bgT, T, acc_g, last_g, last_alpha;
  int lastA, lastB;
};

template <int HALF>
__device__ __forceinline__ float& half_of(float2& v) { return HALF ? v.y : v.x; }
template <int HALF>
__device__ __forceinline__ float half_of(const float2& v) { return HALF ? v.y : v.x; }

template <bool GEO, bool DEPTH>
__device__ __forceinline__ void load_pixel_state(PixelState& s, bool inside, size_t pix, size_t HW, int W, int H,
                                                 float pixx, float pixy, float fx, float fy, int n,
                                                 const float* __restrict__ bg_color,
                                                 const float* __restrict__ all_map_pixels,
                                                 const float* __restrict__ final_Ts,
                                                 const uint32_t* __restrict__ n_contrib,
                                                 const float* __restrict__ dL_dpixels,
                                                 const float* __restrict__ dL_dout_all_maps,
                                                 const float* __restrict__ dL_dout_plane_depths,
                                                 const float* __restrict__ dL_invdepths) {
#pragma unroll
  for (int i = 0; i < 9; ++i) s.w[i] = 0.f;
  const float T_final = inside ? final_Ts[pix] : 0.f;
  s.last_contributor = inside ? min((int)n_contrib[pix], n) : 0;
  float bg_dot = 0.f;
  if (inside) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s.w[c] = dL_dpixels[c * HW + pix];
      bg_dot += __ldg(bg_color + c) * s.w[c];
    }
    if (DEPTH) s.w[3] = dL_invdepths[pix];
    if (GEO) {
#pragma unroll
      for (int c = 0; c < 5; ++c) s.w[4 + c] = dL_dout_all_maps[c * HW + pix];
      // Fold dL/dplane_depth into the geometry channels (backward.cu:583-592).
      const float rayx = (float)(((double)pixx - W * 0.5) / (double)fx);
      const float rayy = (float)(((double)pixy - H * 0.5) / (double)fy);
      const float nx = all_map_pixels[pix], ny = all_map_pixels[HW + pix], nz = all_map_pixels[2 * HW + pix];
      const float dist = all_map_pixels[4 * HW + pix];
      const float tmp = (float)((double)(nx * rayx + ny * rayy + nz) + 1.0e-8);
      const float dpd = dL_dout_plane_depths[pix];
      s.w[8] += (-dpd / tmp);
      s.w[4] += dpd * (dist / (tmp * tmp) * rayx);
      s.w[5] += dpd * (dist / (tmp * tmp) * rayy);
      s.w[6] += dpd * (dist / (tmp * tmp));
    }
  }
  s.bgT = -T_final * bg_dot;
  s.T = T_final;
  s.acc_g = s.last_g = s.last_alpha = 0.f;
}

// ------------------------------------------------------------------------------------------------------------------
// The cross-pixel reduction as tensor-core matrix products.
//
// Every gradient a (pixel, entry) pair contributes is a per-pair scalar times a per-PIXEL constant:
//   dL/d(rgb, 1/z, all_map)[entry] = sum_pix  wgt(pix, entry) * dL/dpixel[ch](pix)           wgt = alpha * T
//   dL/d(mean2D, conic, opacity)   = polynomials in the six moments  sum_pix p(pix, entry) * {1, xi, eta, xi^2, xi eta,
//                                    eta^2}  with p = alpha * dL/dalpha and (xi, eta) the pixel's offset from the sub-tile
//                                    centre (expand dx = (x_e - cx) - xi in dG/dx, dx^2, dx dy, dy^2).
// So for a group of 16 entries the whole reduction over the warp's 64 pixels is  D[16 x n] = A[16 x 64] * B[64 x n]:
// the lanes park (wgt, p) of their two pixels in a per-warp shared-memory tile, and once 16 entries are collected the
// warp runs mma.sync.m16n8k8 (TF32 inputs, FP32 accumulate) against constant B tiles built once per CTA — the nine
// upstream-gradient channels of the warp's pixels and the six moments.  FP32 accuracy is kept by splitting A and the
// gradient channels into TF32 hi + lo parts (three products, the lo*lo term ~2^-22 is dropped); the moment tile is
// exact in TF32 (|xi|, |eta| <= 3.5, products <= 12.25).  This takes the 16-value shuffle butterfly (74 issue slots) and
// the 15 per-pixel partial products (~50) of a shuffle-based reduction off the FP32 issue port: per (warp, entry) it costs one
// STS.128 + ~15 slots of the amortised flush.  The moments are turned into the reference's gradients per entry by the
// lane that owns the row, and the 16-float accumulator row receives vector REDs (red.global.add.v2/v4.f32).
// With the hierarchy interpolation (INTERP) the opacity gradient needs a third per-pair scalar; only its plain sum is
// needed, so it takes one shuffle butterfly per (warp, entry) instead of a third A tile.
constexpr int kGroupC = 16;               // entries per flush (the M of the MMA)
constexpr int kWarpsC = kThreadsB / 32;

constexpr int kRecQuadsC = 4;             // staged entry: 64 B (x y a b | c o am4 id | r g b 1/z | am0..3)
constexpr int kAStride = 68;              // floats per A row: 64 pixels + 4 pad (conflict-free ldmatrix rows, STS.64)

template <bool INTERP>
struct alignas(16) BwdMmaSmem {
  float4 rec[kWarpsC][32 * kRecQuadsC];   // per-WARP staging of the 32 entries a round looks at (see the main loop)
  // wgt, row-major [entry][k]; k = 2 lane + {0: pixel A, 1: pixel B}.  The four pad floats of a row (columns 64..67, never
  // read by ldmatrix) hold the row's first meta quad (x, y, conic a, b):
  alignas(16) float a_w[kWarpsC][kGroupC * kAStride];
  // p, same layout; pad = the second meta quad (conic c, opacity, sum of the third per-pair scalar [INTERP], slot id).
  // One address per parked row reaches all four stores with immediate offsets.
  float a_p[kWarpsC][kGroupC * kAStride];
  float2 b_ch0[kWarpsC][8][32];           // channels 0..7 of the warp's pixels, fragment order: (b0, b1) per lane
  float2 b_ch8[kWarpsC][8][4];            // channel 8 (column 0 of its n-tile: the lanes with g == 0)
  float2 zero2;                           // what the lanes with g != 0 read instead
  float2 b_mom[8][32];                    // the six moments, fragment order (same for every warp)
  // hierarchy interpolation only (last, so that the layout above is the same in both variants):
  float2 tf[INTERP ? kWarpsC : 1][INTERP ? 32 : 1];        // (t, 1 / kids) of the staged entries
  uint32_t info[kWarpsC][32];             // per staged entry: list position | kHasA | kHasB (which halves it can reach)
};
constexpr uint32_t kHasA = 1u << 30, kHasB = 1u << 31;
constexpr uint32_t kARowBytes = kAStride * 4;
constexpr uint32_t kAPOffset = kWarpsC * kGroupC * kAStride * 4;  // from a row of a_w to the same row of a_p
constexpr uint32_t kMetaOffset = 64 * 4;                          // from a row's start to its pad quad

__device__ __forceinline__ void sts_v2(uint32_t addr, float a, float b) {
  asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void sts_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// A-operand quad (a0, a1, a2, a3) of mma.m16n8k8 for one k-step straight from a row-major fp32 tile: each of the four
// 8x8 "b16" matrices of ldmatrix is 8 rows x 16 bytes = 8 rows x 4 fp32, and lane l receives word (l / 4, l % 4) —
// exactly the fragment layout.  Lane l supplies the row address of matrix l / 8: rows (l % 8) + 8 ((l / 8) & 1),
// columns 4 (l / 16) .. +3 of the k-step.
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const float* smem_row_ptr) {
  const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem_row_ptr);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}

// v = hi + lo with hi the TF32 truncation of v (one LOP3; cvt.rna.tf32 is emulated with five instructions on sm_100a)
// and lo = v - hi exact; the MMA ignores the bits of lo below TF32, so |error| <= 2^-20 |v|.
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
  // (opaque to the optimiser: it would otherwise feed the MMA the unmasked register through an extra MOV)
  asm("and.b32 %0, %1, 0xffffe000;" : "=r"(hi) : "r"(__float_as_uint(v)));
  lo = __float_as_uint(v - __uint_as_float(hi));
}

// The same for two values at once: the two truncations are LOP3s, the two exact remainders ONE packed FADD2.
__device__ __forceinline__ void split_tf32_pair(float v0, float v1, uint32_t& hi0, uint32_t& lo0, uint32_t& hi1,
                                                uint32_t& lo1) {
  asm("and.b32 %0, %1, 0xffffe000;" : "=r"(hi0) : "r"(__float_as_uint(v0)));
  asm("and.b32 %0, %1, 0xffffe000;" : "=r"(hi1) : "r"(__float_as_uint(v1)));
  const float2 lo = __fadd2_rn(make_float2(v0, v1), make_float2(-__uint_as_float(hi0), -__uint_as_float(hi1)));
  lo0 = __float_as_uint(lo.x);
  lo1 = __float_as_uint(lo.y);
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// One (pixel, entry) pair: the recurrence, returning the two per-pair scalars
// wgt = alpha * T_before and p = alpha * dL/dalpha (= opacity * G * dL/dalpha: the factor every geometric gradient
// carries; the opacity gradient is sum p / opacity).  A pair that does not contribute is folded in as alpha = 0, which
// leaves the recurrence untouched without any state select: T / (1 - 0) = T, and the next entry's
// acc = 0 * g + 1 * acc_new re-derives the same accumulator.
// With the hierarchy interpolation (INTERP; backward.cu:659-672, 757-765) the blended alpha is
// t a + (1 - t) (1 - (1 - a)^(1/kids)) of a = min(0.99, opacity G); the geometric gradients keep the factor
// opacity * G * dL/dalpha (the reference does not chain them through the interpolation), and the opacity gradient needs
// a THIRD per-pair scalar p3 = opac_mult(a) * p, returned through `p3` and summed over the warp's pixels outside.
// ACC selects how (1 - a)^(1/kids) is evaluated: false = ex2(y lg2 x) as the FORWARD does (forward.cu:550 uses __powf),
// true = the accurate pow of the reference's backward (~35 instructions each, twice per pair).  The fast evaluation
// reports through `near` when the interpolated alpha lies within 2e-6 of the 1/255 cut — ten times its error — so that
// the caller can redo the rare pair whose classification could differ from the reference's.
template <bool GEO, bool DEPTH, bool INTERP, bool ACC, int HALF>
__device__ __forceinline__ bool pixel_pair_c(PairState& ps, const float4 ea, const float4 eb, const float4 ec,
                                             const float4 ed, const float2 tf, float pixx, float2 neg_pixy, int q,
                                             float& wgt, float& p, float& p3, bool& near) {
  struct {  // this half's view of the pair state
    float &T, &acc_g, &last_g, &last_alpha;
    const float bgT;
    const int last_contributor;
    float w[9];
  } s{half_of<HALF>(ps.T), half_of<HALF>(ps.acc_g), half_of<HALF>(ps.last_g), half_of<HALF>(ps.last_alpha),
      half_of<HALF>(ps.bgT), HALF ? ps.lastB : ps.lastA,
      {half_of<HALF>(ps.w[0]), half_of<HALF>(ps.w[1]), half_of<HALF>(ps.w[2]), half_of<HALF>(ps.w[3]),
       half_of<HALF>(ps.w[4]), half_of<HALF>(ps.w[5]), half_of<HALF>(ps.w[6]), half_of<HALF>(ps.w[7]),
       half_of<HALF>(ps.w[8])}};
  const float dx = __fsub_rn(ea.x, pixx), dy = __fadd_rn(ea.y, half_of<HALF>(neg_pixy));  // y - pixy
  const float quad = __fmaf_rn(dx, __fmul_rn(dx, ea.z), __fmul_rn(dy, __fmul_rn(dy, eb.x)));
  const float power = __fmaf_rn(quad, -0.5f, -__fmul_rn(dy, __fmul_rn(dx, ea.w)));
  // expf, not a bare ex2.approx: the pair must be classified (alpha >= 1/255, clamp at 0.99) exactly as the forward
  // classified it.  (Measured: the one-MUFU variant is 3 % faster and flips one borderline pair in 3 M gradients.)
  const float test_alpha = __fmul_rn(eb.y, expf(power));
  float blend_alpha = fminf(0.99f, test_alpha);
  float opac_mult = 1.0f;
  if (INTERP) {
    const float my_alpha = blend_alpha;
    if (ACC) {
      const float kidsqrt = 1.0f - powf(1.0f - my_alpha, tf.y);
      blend_alpha = tf.x * my_alpha + (1.0f - tf.x) * kidsqrt;
      opac_mult = tf.x - powf(1.0f - my_alpha, tf.y - 1.0f) * (tf.x - 1.0f) * tf.y;
    } else {
      const float lg = __log2f(1.0f - my_alpha);  // both powers from one logarithm
      const float kidsqrt = 1.0f - exp2f(tf.y * lg);
      blend_alpha = tf.x * my_alpha + (1.0f - tf.x) * kidsqrt;
      opac_mult = tf.x - exp2f((tf.y - 1.0f) * lg) * (tf.x - 1.0f) * tf.y;
      near = near || fabsf(blend_alpha - 1.0f / 255.0f) < 2e-6f;
    }
  }
  const bool valid = (q < s.last_contributor) && !(power > 0.0f) && !(blend_alpha < 1.0f / 255.0f);
  const float alpha = valid ? blend_alpha : 0.f;
  float rinv;  // 1 - alpha is in [0.01, 1]: the bare MUFU.RCP, without __fdividef's denormal-range scaling (same bits)
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv) : "f"(1.0f - alpha));
  const float Tn = s.T * rinv;
  wgt = alpha * Tn;
  float g = ec.x * s.w[0] + ec.y * s.w[1] + ec.z * s.w[2];
  if (DEPTH) g += ec.w * s.w[3];
  if (GEO) {
    const float ee = eb.z;
    g += ed.x * s.w[4] + ed.y * s.w[5] + ed.z * s.w[6] + ed.w * s.w[7] + ee * s.w[8];
  }
  const float acc_new = s.last_alpha * s.last_g + (1.0f - s.last_alpha) * s.acc_g;
  const float dL_dalpha = (g - acc_new) * Tn + s.bgT * rinv;
  s.T = Tn;
  s.acc_g = acc_new;
  s.last_g = g;
  s.last_alpha = alpha;
  // the clamp has no gradient (backward.cu:680-682).  Without interpolation alpha == opacity * G here.
  p = test_alpha > 0.99f ? 0.f : (INTERP ? (valid ? test_alpha : 0.f) : alpha) * dL_dalpha;
  p3 = INTERP ? opac_mult * p : 0.f;
  return valid;
}

// Both pixels of the lane at once (an entry that reaches both halves; no hierarchy interpolation): the code above on
// FADD2 / FMUL2 / FFMA2.  The classification (power, expf, the 0.99 clamp and the 1/255 cut) keeps the forward's
// operations, order and roundings per half (see fwd_pair_flat2 in blend_fwd.cu); the gradient arithmetic behind it is
// the same expressions with explicit contractions.
template <bool GEO, bool DEPTH>
__device__ __forceinline__ bool pixel_pair_c2(PairState& s, const float4 ea, const float4 eb, const float4 ec,
                                              const float4 ed, float pixx, float2 neg_pixy, int q, float2& wgt,
                                              float2& p) {
  const float dx = __fsub_rn(ea.x, pixx);
  const float dxa = __fmul_rn(dx, ea.z);
  const float ndxb = __fmul_rn(dx, -ea.w);
  const float2 dy = __fadd2_rn(bc2(ea.y), neg_pixy);
  const float2 quad = __ffma2_rn(bc2(dx), bc2(dxa), __fmul2_rn(dy, __fmul2_rn(dy, bc2(eb.x))));
  const float2 power = __ffma2_rn(quad, bc2(-0.5f), __fmul2_rn(dy, bc2(ndxb)));
  const float2 test_alpha = __fmul2_rn(bc2(eb.y), expf_pair(power));
  const float bA = fminf(0.99f, test_alpha.x), bB = fminf(0.99f, test_alpha.y);
  const bool validA = (q < s.lastA) && !(power.x > 0.0f) && !(bA < 1.0f / 255.0f);
  const bool validB = (q < s.lastB) && !(power.y > 0.0f) && !(bB < 1.0f / 255.0f);
  const float2 alpha = make_float2(validA ? bA : 0.f, validB ? bB : 0.f);
  const float2 om = __ffma2_rn(alpha, bc2(-1.0f), bc2(1.0f));
  float2 rinv;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv.x) : "f"(om.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rinv.y) : "f"(om.y));
  const float2 Tn = __fmul2_rn(s.T, rinv);
  wgt = __fmul2_rn(alpha, Tn);
  float2 g = __ffma2_rn(bc2(ec.z), s.w[2], __ffma2_rn(bc2(ec.y), s.w[1], __fmul2_rn(bc2(ec.x), s.w[0])));
  if (DEPTH) g = __ffma2_rn(bc2(ec.w), s.w[3], g);
  if (GEO) {
    g = __ffma2_rn(bc2(ed.x), s.w[4], g);
    g = __ffma2_rn(bc2(ed.y), s.w[5], g);
    g = __ffma2_rn(bc2(ed.z), s.w[6], g);
    g = __ffma2_rn(bc2(ed.w), s.w[7], g);
    g = __ffma2_rn(bc2(eb.z), s.w[8], g);
  }
  const float2 om_last = __ffma2_rn(s.last_alpha, bc2(-1.0f), bc2(1.0f));
  const float2 acc_new = __ffma2_rn(s.last_alpha, s.last_g, __fmul2_rn(om_last, s.acc_g));
  const float2 diff = __fadd2_rn(g, make_float2(-acc_new.x, -acc_new.y));
  const float2 dL_dalpha = __ffma2_rn(diff, Tn, __fmul2_rn(s.bgT, rinv));
  s.T = Tn;
  s.acc_g = acc_new;
  s.last_g = g;
  s.last_alpha = alpha;
  p = __fmul2_rn(alpha, dL_dalpha);
  p.x = test_alpha.x > 0.99f ? 0.f : p.x;
  p.y = test_alpha.y > 0.99f ? 0.f : p.y;
  return validA || validB;
}

template <bool GEO, bool DEPTH, bool INTERP>
// (Measured: 3 CTAs / SM with 148 registers — no rematerialisation pressure, 12 warps — 1.123 ms against 1.014 ms at
// 4 CTAs / 125 registers: the kernel wants warps more than registers.)
__global__ void __launch_bounds__(kThreadsB, INTERP ? 3 : 4)
blend_bwd3_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                  const float4* __restrict__ records, const float* __restrict__ ts, const int* __restrict__ kids,
                  const int W, const int H, const float fx, const float fy,
                  const float* __restrict__ bg_color, const float* __restrict__ all_map_pixels,
                  const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                  const float* __restrict__ dL_dpixels, const float* __restrict__ dL_dout_all_maps,
                  const float* __restrict__ dL_dout_plane_depths, const float* __restrict__ dL_invdepths,
                  float* __restrict__ accum) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BwdMmaSmem<INTERP>& sm = *reinterpret_cast<BwdMmaSmem<INTERP>*>(smem_raw);

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = blockIdx.y * gridDim.x + blockIdx.x;
  const int wx0 = blockIdx.x * HG_BLOCK_X + (warp & 1) * 8;
  const int wy0 = blockIdx.y * HG_BLOCK_Y + (warp >> 1) * 8;
  const int pxi = wx0 + (lane & 7), pyA = wy0 + (lane >> 3), pyB = pyA + 4;
  const bool insideA = pxi < W && pyA < H, insideB = pxi < W && pyB < H;
  const float pixx = (float)pxi, pixyA = (float)pyA, pixyB = (float)pyB;
  const float fx0 = (float)wx0, fx1 = (float)(wx0 + 7);
  const float2 fy0 = make_float2((float)wy0, (float)(wy0 + 4)), fy1 = make_float2((float)(wy0 + 3), (float)(wy0 + 7));
  const float cx = (float)wx0 + 3.5f, cy = (float)wy0 + 3.5f;  // moments are taken about the sub-tile centre
  const size_t HW = (size_t)H * W;

  const uint2 range = ranges[tile];
  const int n = (int)(range.y - range.x);

  PairState S;
  PixelState A, B;
  load_pixel_state<GEO, DEPTH>(A, insideA, (size_t)pyA * W + pxi, HW, W, H, pixx, pixyA, fx, fy, n, bg_color,
                               all_map_pixels, final_Ts, n_contrib, dL_dpixels, dL_dout_all_maps, dL_dout_plane_depths,
                               dL_invdepths);
  load_pixel_state<GEO, DEPTH>(B, insideB, (size_t)pyB * W + pxi, HW, W, H, pixx, pixyB, fx, fy, n, bg_color,
                               all_map_pixels, final_Ts, n_contrib, dL_dpixels, dL_dout_all_maps, dL_dout_plane_depths,
                               dL_invdepths);

#pragma unroll
  for (int c = 0; c < 9; ++c) S.w[c] = make_float2(A.w[c], B.w[c]);
  S.bgT = make_float2(A.bgT, B.bgT);
  S.T = make_float2(A.T, B.T);
  S.acc_g = S.last_g = S.last_alpha = make_float2(0.f, 0.f);
  S.lastA = A.last_contributor;
  S.lastB = B.last_contributor;
  const float2 neg_pixy = make_float2(-pixyA, -pixyB);
  // last contributor of each half and of the tile; the slot id of this lane's first list entry is requested right away
  // so that the list read (and, behind it, the record gather) overlaps the construction of the B tiles
  int wmaxA = A.last_contributor, wmaxB = B.last_contributor;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wmaxA = max(wmaxA, __shfl_xor_sync(0xffffffffu, wmaxA, o));
    wmaxB = max(wmaxB, __shfl_xor_sync(0xffffffffu, wmaxB, o));
  }
  const int wmax = max(wmaxA, wmaxB);
  int first_id = 0;
  if (wmax - 1 - lane >= 0) first_id = (int)__ldg(point_list + range.x + (wmax - 1 - lane));

  // ---- constant B tiles.  K index of a pixel: k = 2 lane + {0: pixel A, 1: pixel B} of the lane that owns it; k-step
  // s covers k = 8 s .. 8 s + 7, and fragment lane (g, t) holds b0 = B[8 s + t][g], b1 = B[8 s + t + 4][g].
  {
    const int ks = lane >> 2, jA = 2 * (lane & 3);  // this lane's pixel A is column jA of k-step ks, pixel B column jA + 1
    float* const ch0 = reinterpret_cast<float*>(&sm.b_ch0[warp][ks][0]);
    float* const ch8 = reinterpret_cast<float*>(&sm.b_ch8[warp][ks][0]);
    // column j of a k-step -> fragment lane t = j & 3, component (j >> 2): float index 2 (4 g + t) + (j >> 2)
    const int fa = 2 * (jA & 3) + (jA >> 2), fb = 2 * ((jA + 1) & 3) + ((jA + 1) >> 2);
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      ch0[8 * c + fa] = A.w[c];
      ch0[8 * c + fb] = B.w[c];
    }
    if (GEO) {
      ch8[fa] = A.w[8];
      ch8[fb] = B.w[8];
      if (tid == 0) sm.zero2 = make_float2(0.f, 0.f);
    }
    // moments of fragment lane (g, t) at k-step s: the pixels at k = 8 s + t and 8 s + t + 4
    for (int i = tid; i < 8 * 32; i += kThreadsB) {
      const int s_ = i >> 5, l_ = i & 31, g_ = l_ >> 2, t_ = l_ & 3;
      float m[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int k = 8 * s_ + t_ + 4 * h, owner = k >> 1;
        const float xi = (float)(owner & 7) - 3.5f, eta = (float)((owner >> 3) + 4 * (k & 1)) - 3.5f;
        switch (g_) {
          case 0: m[h] = 1.f; break;
          case 1: m[h] = xi; break;
          case 2: m[h] = eta; break;
          case 3: m[h] = xi * xi; break;
          case 4: m[h] = xi * eta; break;
          case 5: m[h] = eta * eta; break;
          default: m[h] = 0.f; break;
        }
      }
      sm.b_mom[s_][l_] = make_float2(m[0], m[1]);
    }
  }

  __syncthreads();  // the B tiles are complete; from here on the warps of the CTA never synchronise with each other
  const float ddelx_dx = 0.5f * W, ddely_dy = 0.5f * H;

  float* const a_w = sm.a_w[warp];
  float* const a_p = sm.a_p[warp];
  const int fg = lane >> 2, ft = lane & 3;                      // fragment coordinates of this lane
  // ldmatrix row address this lane supplies: matrix lane / 8 -> rows (lane % 8) + 8 (matrix & 1), columns 4 (matrix >> 1)
  const int a_ld = ((lane & 7) + 8 * ((lane >> 3) & 1)) * kAStride + 4 * (lane >> 4);  // + 8 * k-step
  const float2* const b8_src = fg == 0 ? &sm.b_ch8[warp][0][ft] : &sm.zero2;  // + 4 * k-step for g == 0
  const int b8_step = fg == 0 ? 4 : 0;
  int rows = 0;  // entries parked in the A tile
  // shared-space address of this lane's (pixel A, pixel B) pair in row 0 of a_w.  Opaque to the optimiser: under the
  // 128-register cap it otherwise re-derives the address from S2R (thread id, shared window) for every parked entry.
  uint32_t park_base;
  asm volatile("mov.u32 %0, %1;" : "=r"(park_base) : "r"((uint32_t)__cvta_generic_to_shared(a_w + 2 * lane)));

  // Reduce the parked rows over the warp's 64 pixels and add them to the accumulator rows.
  auto flush = [&]() {
    __syncwarp();
    // two accumulator chains per tile (hi*hi | the two cross terms) halve the depth of the dependent MMA sequence
    float dc0[4] = {0.f, 0.f, 0.f, 0.f}, dc8[4] = {0.f, 0.f, 0.f, 0.f}, dm[4] = {0.f, 0.f, 0.f, 0.f};
    float dc0x[4] = {0.f, 0.f, 0.f, 0.f}, dc8x[4] = {0.f, 0.f, 0.f, 0.f}, dmx[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int s_ = 0; s_ < 8; ++s_) {
      uint32_t wq[4], pq[4];
      ldmatrix_x4(wq, a_w + a_ld + 8 * s_);
      ldmatrix_x4(pq, a_p + a_ld + 8 * s_);
      uint32_t w0h, w0l, w1h, w1l, w2h, w2l, w3h, w3l, p0h, p0l, p1h, p1l, p2h, p2l, p3h, p3l;
      split_tf32_pair(__uint_as_float(wq[0]), __uint_as_float(wq[1]), w0h, w0l, w1h, w1l);
      split_tf32_pair(__uint_as_float(wq[2]), __uint_as_float(wq[3]), w2h, w2l, w3h, w3l);
      split_tf32_pair(__uint_as_float(pq[0]), __uint_as_float(pq[1]), p0h, p0l, p1h, p1l);
      split_tf32_pair(__uint_as_float(pq[2]), __uint_as_float(pq[3]), p2h, p2l, p3h, p3l);
      const float2 bq = sm.b_ch0[warp][s_][lane];
      uint32_t b0h, b0l, b1h, b1l;
      split_tf32_pair(bq.x, bq.y, b0h, b0l, b1h, b1l);
      mma_tf32(dc0, w0h, w1h, w2h, w3h, b0h, b1h);
      mma_tf32(dc0x, w0l, w1l, w2l, w3l, b0h, b1h);
      mma_tf32(dc0x, w0h, w1h, w2h, w3h, b0l, b1l);
      if (GEO) {
        const float2 b8 = b8_src[s_ * b8_step];
        split_tf32_pair(b8.x, b8.y, b0h, b0l, b1h, b1l);
        mma_tf32(dc8, w0h, w1h, w2h, w3h, b0h, b1h);
        mma_tf32(dc8x, w0l, w1l, w2l, w3l, b0h, b1h);
        mma_tf32(dc8x, w0h, w1h, w2h, w3h, b0l, b1l);
      }
      const float2 bm = sm.b_mom[s_][lane];
      mma_tf32(dm, p0h, p1h, p2h, p3h, __float_as_uint(bm.x), __float_as_uint(bm.y));
      mma_tf32(dmx, p0l, p1l, p2l, p3l, __float_as_uint(bm.x), __float_as_uint(bm.y));
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dc0[i] += dc0x[i];
      dc8[i] += dc8x[i];
      dm[i] += dmx[i];
    }
    // D fragments: lane (g, t) holds columns 2t, 2t+1 of rows g (d[0], d[1]) and g + 8 (d[2], d[3]).
    // Moment columns: 0 S1, 1 Sx, 2 Sy, 3 Sxx, 4 Sxy, 5 Syy -> collect all six in the lane with t == 0.
    float Sy[2], Sxx[2], Sxy[2], Syy[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      Sy[h] = __shfl_down_sync(0xffffffffu, dm[2 * h], 1);
      Sxx[h] = __shfl_down_sync(0xffffffffu, dm[2 * h + 1], 1);
      Sxy[h] = __shfl_down_sync(0xffffffffu, dm[2 * h], 2);
      Syy[h] = __shfl_down_sync(0xffffffffu, dm[2 * h + 1], 2);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int row = fg + 8 * h;
      if (row < rows) {
        const float4 m1 = *reinterpret_cast<const float4*>(a_p + row * kAStride + 64);
        float* const arow = accum + (size_t)__float_as_int(m1.w) * HG_ACC_FLOATS;
        red_add_v2(arow + 2 * ft, dc0[2 * h], dc0[2 * h + 1]);
        if (ft == 0) {
          const float4 m0 = *reinterpret_cast<const float4*>(a_w + row * kAStride + 64);
          const float S1 = dm[2 * h], Sx = dm[2 * h + 1];
          const float u = m0.x - cx, v = m0.y - cy, o = m1.y;
          const float Dx = u * S1 - Sx, Dy = v * S1 - Sy[h];                 // sum p dx, sum p dy
          const float Qxx = u * (Dx - Sx) + Sxx[h];                          // sum p dx^2
          const float Qxy = u * Dy - v * Sx + Sxy[h];                        // sum p dx dy
          const float Qyy = v * (Dy - Sy[h]) + Syy[h];                       // sum p dy^2
          const float g9 = -ddelx_dx * (m0.z * Dx + m0.w * Dy);
          const float g10 = -ddely_dy * (m1.x * Dy + m0.w * Dx);
          red_add_v4(arow + 8, dc8[2 * h], g9, g10, -0.5f * Qxx);
          red_add_v2(arow + 12, -0.5f * Qxy, -0.5f * Qyy);
          atomicAdd(arow + 14, __fdividef(INTERP ? m1.z : S1, o));
        }
      }
    }
    rows = 0;
    __syncwarp();
  };

  // ---- main loop, per warp and barrier-free.  Each warp walks the tile's list itself, back to front from ITS last
  // contributor, 32 entries per round: lane j gathers entry j's 64-byte record (register double-buffered, the records
  // are L2 resident), tests it against the warp's two 8x4 halves from registers, and only the entries that can reach
  // the sub-tile are parked in the warp's private shared-memory slots for the broadcast reads of the blend loop.  The
  // four warps of a tile re-read the same records from L2 (4x the gather traffic of a CTA-wide staging, ~1.9 GB per
  // view, far from the L2 limit) and in exchange never wait for each other: no __syncthreads, no idle tail per batch.
  float4* const w_rec = sm.rec[warp];
  uint32_t* const w_info = sm.info[warp];
  const uint32_t lanes_below = (1u << lane) - 1u;
  Prefetch pf;
  auto prefetch = [&](int base) {
    const int q = base - lane;
    if (q >= 0) gather_record<INTERP>(pf, point_list, records, ts, kids, range.x + q);
  };
  if (wmax - 1 - lane >= 0) gather_record_id<INTERP>(pf, first_id, records, ts, kids);
  // (Measured and dropped: fetching the list's slot ids one round further ahead than the records, so that the record
  // loads never wait for the list read — 1.036 vs 1.022 ms, one more live register.)

  for (int base = wmax - 1; base >= 0; base -= 32) {
    const int q_mine = base - lane;
    bool keepA = false, keepB = false;
    const float a = pf.r0.z, bb = pf.r0.w, c = pf.r1.x, o = pf.r1.y;
    if (q_mine >= 0) {
      const float tau = cull_tau(a, bb, c, o, INTERP);
      may_touch2(pf.r0.x, pf.r0.y, a, bb, c, tau, fx0, fx1, fy0, fy1, keepA, keepB);
      keepA = keepA && q_mine < wmaxA;
      keepB = keepB && q_mine < wmaxB;
    }
    const uint32_t maskA = __ballot_sync(0xffffffffu, keepA), maskB = __ballot_sync(0xffffffffu, keepB);
    const uint32_t mask = maskA | maskB;
    if (keepA || keepB) {  // survivors are parked densely, in list order: the blend loop below is a plain counted loop
      const int slot = __popc(mask & lanes_below);
      float4* d = w_rec + kRecQuadsC * slot;
      d[0] = pf.r0;
      d[1] = make_float4(c, o, pf.r3.z, __int_as_float(pf.id));
      d[2] = make_float4(pf.r1.z, pf.r1.w, pf.r2.x, pf.r2.y);
      if (GEO) d[3] = make_float4(pf.r2.z, pf.r2.w, pf.r3.x, pf.r3.y);
      if (INTERP) sm.tf[warp][slot] = make_float2(pf.it, pf.ifrac);
      w_info[slot] = (uint32_t)q_mine | (keepA ? kHasA : 0u) | (keepB ? kHasB : 0u);
    }
    __syncwarp();
    if (base >= 32) prefetch(base - 32);
    // (Measured and dropped after the packed-FP32 rewrite: every survivor through the packed pair path — a half it cannot
    // reach entering with alpha = 0 — with test_alpha of entry i + 1 computed next to the recurrence of entry i: 1.012 vs
    // 0.976 ms; the entries that reach one half only are too many to pay a pair evaluation each.)
    // (Measured and dropped: software-pipelining this loop by one entry — next entry's first two quads fetched while the
    // current one is evaluated — 1.27 vs 1.09 ms: eight more live registers at the 128-register cap cost more than the
    // short-scoreboard stalls they remove.)
    const int cnt = __popc(mask);
    for (int i = 0; i < cnt; ++i) {
      const float4* e = w_rec + kRecQuadsC * i;
      const uint32_t info = w_info[i];
      const float4 ea = e[0];
      const float4 eb = e[1];
      const float4 ec = e[2];      float4 ed = make_float4(0.f, 0.f, 0.f, 0.f);
      if (GEO) ed = e[3];
      float2 tf = make_float2(1.f, 1.f);
      if (INTERP) tf = sm.tf[warp][i];
      const int q = (int)(info & (kHasA - 1u));
      const bool hasA = (info & kHasA) != 0u, hasB = (info & kHasB) != 0u;
      float wA = 0.f, pA = 0.f, wB = 0.f, pB = 0.f, p3A = 0.f, p3B = 0.f;
      bool any, near = false;
      // state before this entry (only read again on the rare redo below)
      const float2 S_T = S.T, S_acc = S.acc_g, S_lg = S.last_g, S_la = S.last_alpha;
      if (hasA && hasB) {  // one basic block: the two pixels' dependency chains interleave
        if (!INTERP) {
          float2 w2, p2;
          any = pixel_pair_c2<GEO, DEPTH>(S, ea, eb, ec, ed, pixx, neg_pixy, q, w2, p2);
          wA = w2.x; wB = w2.y; pA = p2.x; pB = p2.y;
        } else {
          any = pixel_pair_c<GEO, DEPTH, INTERP, false, 0>(S, ea, eb, ec, ed, tf, pixx, neg_pixy, q, wA, pA, p3A, near);
          any |= pixel_pair_c<GEO, DEPTH, INTERP, false, 1>(S, ea, eb, ec, ed, tf, pixx, neg_pixy, q, wB, pB, p3B, near);
        }
      } else if (hasA) {
        any = pixel_pair_c<GEO, DEPTH, INTERP, false, 0>(S, ea, eb, ec, ed, tf, pixx, neg_pixy, q, wA, pA, p3A, near);
      } else {
        any = pixel_pair_c<GEO, DEPTH, INTERP, false, 1>(S, ea, eb, ec, ed, tf, pixx, neg_pixy, q, wB, pB, p3B, near);
      }
      if (INTERP && __any_sync(0xffffffffu, near)) {
        // a pair within the error of the fast pow of the 1/255 cut: evaluate this entry again, for the whole warp, the
        // way the reference's backward does
        S.T = S_T; S.acc_g = S_acc; S.last_g = S_lg; S.last_alpha = S_la;
        wA = pA = wB = pB = p3A = p3B = 0.f;
        any = false;
        if (hasA) any = pixel_pair_c<GEO, DEPTH, INTERP, true, 0>(S, ea, eb, ec, ed, tf, pixx, neg_pixy, q, wA, pA, p3A, near);
        if (hasB) any |= pixel_pair_c<GEO, DEPTH, INTERP, true, 1>(S, ea, eb, ec, ed, tf, pixx, neg_pixy, q, wB, pB, p3B, near);
      }
      if (__ballot_sync(0xffffffffu, any) == 0) continue;
      float s3 = eb.z;
      if (INTERP) {  // the third scalar only needs its plain sum over the 64 pixels: one butterfly, no third A tile
        s3 = p3A + p3B;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s3 += __shfl_xor_sync(0xffffffffu, s3, o);
      }
      const uint32_t row_addr = park_base + rows * kARowBytes;
      sts_v2(row_addr, wA, wB);
      sts_v2(row_addr + kAPOffset, pA, pB);
      if (lane == 0) {
        sts_v4(row_addr + kMetaOffset, ea.x, ea.y, ea.z, ea.w);
        sts_v4(row_addr + kAPOffset + kMetaOffset, eb.x, eb.y, s3, eb.w);
      }
      if (++rows == kGroupC) flush();
    }
    __syncwarp();  // every lane is done with this round's slots
  }
  if (rows > 0) flush();
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
  // the opt-in for > 48 KB of dynamic shared memory is a per-device function attribute: set once per device
    static std::mutex attr_mu;
    static bool attr_done[64] = {}, attr_ok[64] = {};
    int device = 0;
    HG_CUDA_TRY(cudaGetDevice(&device));
    if (device < 0 || device >= 64) {
      set_error("blend_bwd: device index %d out of range", device);
      return HG_ERR_CUDA;
    }
    {
      std::lock_guard<std::mutex> lk(attr_mu);
      if (!attr_done[device]) {
        bool ok = true;
        const int b0 = (int)sizeof(BwdMmaSmem<false>), b1 = (int)sizeof(BwdMmaSmem<true>);
        const auto attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
        ok &= cudaFuncSetAttribute(blend_bwd3_kernel<true, true, false>, attr, b0) == cudaSuccess;
        ok &= cudaFuncSetAttribute(blend_bwd3_kernel<true, false, false>, attr, b0) == cudaSuccess;
        ok &= cudaFuncSetAttribute(blend_bwd3_kernel<false, true, false>, attr, b0) == cudaSuccess;
        ok &= cudaFuncSetAttribute(blend_bwd3_kernel<false, false, false>, attr, b0) == cudaSuccess;
        ok &= cudaFuncSetAttribute(blend_bwd3_kernel<true, true, true>, attr, b1) == cudaSuccess;
        ok &= cudaFuncSetAttribute(blend_bwd3_kernel<true, false, true>, attr, b1) == cudaSuccess;
        ok &= cudaFuncSetAttribute(blend_bwd3_kernel<false, true, true>, attr, b1) == cudaSuccess;
        ok &= cudaFuncSetAttribute(blend_bwd3_kernel<false, false, true>, attr, b1) == cudaSuccess;
        attr_ok[device] = ok;
        attr_done[device] = true;
      }
    }
    if (!attr_ok[device]) {
      set_error("blend_bwd: could not reserve %zu bytes of shared memory", sizeof(BwdMmaSmem<true>));
      return HG_ERR_CUDA;
    }
#define HG_LAUNCH3(G_, D_, I_)                                                                            \
  blend_bwd3_kernel<G_, D_, I_><<<grid, kThreadsB, sizeof(BwdMmaSmem<I_>), stream>>>(                     \
      img.ranges, b.vals, g.records, in.ts, in.kids, in.W, in.H, focal_x, focal_y, in.background,         \
      all_map_pixels, img.final_T, img.n_contrib, dL_dpix, dL_dout_all_map, dL_dout_plane_depth,          \
      dL_dout_invdepth, accum)
    if (interp) {
      if (geo && depth) HG_LAUNCH3(true, true, true);
      else if (geo) HG_LAUNCH3(true, false, true);
      else if (depth) HG_LAUNCH3(false, true, true);
      else HG_LAUNCH3(false, false, true);
    } else {
      if (geo && depth) HG_LAUNCH3(true, true, false);
      else if (geo) HG_LAUNCH3(true, false, false);
      else if (depth) HG_LAUNCH3(false, true, false);
      else HG_LAUNCH3(false, false, false);
    }
#undef HG_LAUNCH3
  HG_POST_LAUNCH(in.debug, stream, "blend_bwd");
  return HG_OK;
}

}  // namespace hg
