// losses.cu — image / patch losses of the HiDeGS training step as fused sm_100a kernels.
//
// Replaces the PyTorch op chains of utils/loss_utils.py (l1_loss :18-19, l2_loss :21-22, ssim :24-64,
// get_img_grad_weight :66-78, lncc :80-115) and compute_scale_regularization
// (scripts/frequency_regularization.py:1403-1444).  All of them are HBM-bound streaming passes: every
// image is read once (plus a 5-pixel halo for SSIM), reductions are two-stage and deterministic
// (per-block partials in fixed order, final sum in double), nothing synchronises with the host.
#include "common.cuh"
#include "reduce.cuh"
#include "../../include/hidegs_losses.h"

#include <cmath>

namespace hg {

namespace {

constexpr int kRedBlocks = 148 * 4;  // a multiple of the SM count
constexpr int kRedThreads = 256;

// Sum over the CTA (any block shape); `tid` is the linear thread index.  Valid in thread 0.
__device__ __forceinline__ float block_sum(float v, float* smem, int tid, int nthreads) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int lane = tid & 31, warp = tid >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  const int nw = (nthreads + 31) >> 5;
  v = (tid < nw) ? smem[tid] : 0.f;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  }
  __syncthreads();
  return v;
}

// ---------------------------------------------------------------- l1 / l2
template <bool L2>
__global__ void __launch_bounds__(kRedThreads)
pixel_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                  float* __restrict__ grad, double* __restrict__ partial) {
  __shared__ float red[32];
  float acc = 0.f;
  const float inv_n = 1.0f / (float)n;
  const int64_t n4 = n / 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 x = __ldg((const float4*)a + i), y = __ldg((const float4*)b + i);
    const float d[4] = {x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w};
    float4 g;
    float* gp = &g.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (L2) {
        acc += d[k] * d[k];
        gp[k] = 2.0f * d[k] * inv_n;
      } else {
        acc += fabsf(d[k]);
        gp[k] = (d[k] > 0.f ? 1.f : (d[k] < 0.f ? -1.f : 0.f)) * inv_n;
      }
    }
    if (grad) ((float4*)grad)[i] = g;
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = a[i] - b[i];
    acc += L2 ? d * d : fabsf(d);
    if (grad) grad[i] = L2 ? 2.0f * d * inv_n : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * inv_n;
  }
  const float s = block_sum(acc, red, threadIdx.x, kRedThreads);
  if (threadIdx.x == 0) partial[blockIdx.x] = (double)s;
}

__global__ void finalize_mean_kernel(const double* __restrict__ partial, int n_partial, double denom,
                                     float* __restrict__ out) {
  __shared__ double sm[32];
  const double s = cta_sum_strided(partial, n_partial, 1, 0, sm);
  if (threadIdx.x == 0) out[0] = (float)(s / denom);
}

template <bool L2>
int run_pixel_loss(const float* a, const float* b, int64_t n, float* out, float* grad, void* ws,
                   cudaStream_t st) {
  if (!a || !b || !out || !ws || n <= 0) {
    set_error("pixel loss: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  const bool vec_ok = (((uintptr_t)a | (uintptr_t)b | (uintptr_t)grad) & 15) == 0;
  if (!vec_ok) {
    set_error("pixel loss: pointers must be 16-byte aligned");
    return HG_ERR_INVALID_ARG;
  }
  double* partial = (double*)ws;
  pixel_loss_kernel<L2><<<kRedBlocks, kRedThreads, 0, st>>>(a, b, n, grad, partial);
  HG_POST_LAUNCH(false, st, "pixel_loss");
  finalize_mean_kernel<<<1, 256, 0, st>>>(partial, kRedBlocks, (double)n, out);
  HG_POST_LAUNCH(false, st, "finalize_mean");
  return HG_OK;
}

// d (g * mean |a - b|) / da  (or the squared difference); grad_b = -grad_a.  The upstream scalar is read from the device.
template <bool L2>
__global__ void __launch_bounds__(256)
pixel_loss_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, const float* __restrict__ gscale,
                      float* __restrict__ grad_a, float* __restrict__ grad_b) {
  const float s = (gscale ? __ldg(gscale) : 1.0f) / (float)n;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float d = __ldg(a + i) - __ldg(b + i);
    const float g = L2 ? 2.0f * d * s : (d > 0.f ? s : (d < 0.f ? -s : 0.f));
    if (grad_a) grad_a[i] = g;
    if (grad_b) grad_b[i] = -g;
  }
}

// total = clamp(lambda_freq * freq + lambda_scale * scale * [count > 0], 0, 1) of frequency_regularization_pyramid_scale
// (:1636-1660) and its two partial derivatives, in one launch: out = [total, d total / d freq, d total / d scale]
__global__ void freq_total_kernel(const float* __restrict__ freq, const float* __restrict__ scale,
                                  const float* __restrict__ count, float lambda_freq, float lambda_scale,
                                  float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float nonempty = (count && count[0] > 0.f) ? 1.f : 0.f;
  float t = 0.f;
  if (freq) t = lambda_freq * freq[0];
  if (scale) {
    const float term = lambda_scale * scale[0] * nonempty;
    t = freq ? t + term : term;
  }
  const float gate = (t >= 0.f && t <= 1.0f) ? 1.f : 0.f;  // d clamp(t, 0, 1) / dt
  out[0] = fminf(fmaxf(t, 0.f), 1.0f);
  out[1] = freq ? gate * lambda_freq : 0.f;
  out[2] = scale ? gate * lambda_scale * nonempty : 0.f;
}

// ---------------------------------------------------------------- composed image gradient of the training loss
__global__ void __launch_bounds__(256)
training_image_grad_kernel(const float* __restrict__ color, const float* __restrict__ gt, const float* g_ssim,
                           const float* g_freq, int64_t n, float w_l1_over_n, float w_ssim,
                           const float* __restrict__ w_freq, float* out) {
  const float wf = w_freq ? __ldg(w_freq) : 0.f;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float c = color[i];
    const float d = fminf(fmaxf(c, 0.f), 1.f) - gt[i];
    float g = (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f)) * w_l1_over_n;
    if (g_ssim) g = fmaf(w_ssim, g_ssim[i], g);
    if (g_freq) g = fmaf(wf, g_freq[i], g);
    out[i] = (c >= 0.f && c <= 1.f) ? g : 0.f;  // d clamp(c, 0, 1) / dc
  }
}

// ---------------------------------------------------------------- value of the composed training loss
// (1 - l) L1 + l (1 - SSIM) + regulariser total + normal term, the operations of the Python expression in its order
// (separately rounded), in one launch instead of six scalar torch kernels per view.
__global__ void training_loss_value_kernel(const float* __restrict__ l1, const float* __restrict__ ssim,
                                           const float* __restrict__ extra0, const float* __restrict__ extra1,
                                           float lambda_dssim, float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float v = __fadd_rn(__fmul_rn(1.0f - lambda_dssim, l1[0]), __fmul_rn(lambda_dssim, __fsub_rn(1.0f, ssim[0])));
  if (extra0) v = __fadd_rn(v, extra0[0]);
  if (extra1) v = __fadd_rn(v, extra1[0]);
  out[0] = v;
}

// ---------------------------------------------------------------- SSIM
// Separable 11-tap Gaussian (sigma 1.5) in shared memory: 32x32 output tile per 256-thread CTA, 42x42 input tile
// (halo 5, zero padded as F.conv2d(padding=5)).  Both passes are register blocked — a thread produces 4 consecutive
// outputs from a 14-value sliding window — so a tap costs 0.3 shared-memory loads instead of one.
// The 11 window weights travel as a by-value kernel argument (constant bank of the launch): no per-device symbol to
// initialise, nothing on the default stream.
struct Gauss11 { float w[11]; };
constexpr int kTile = 32, kHalo = 5, kIn = kTile + 2 * kHalo;  // 42
constexpr int kSsimThreads = 256;

// horizontal pass over NQ quantities derived from the NS staged planes; work item = (row, group of 4 columns)
template <int NS, int NQ, typename F>
__device__ __forceinline__ void conv_rows4(const float (*src)[kIn][kIn + 1], float (*dst)[kIn][kTile + 1], int tid,
                                           const Gauss11& gw, F make) {
  for (int w = tid; w < kIn * (kTile / 4); w += kSsimThreads) {
    const int r = w / (kTile / 4), c0 = (w - r * (kTile / 4)) * 4;
    float acc[NQ][4];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[q][j] = 0.f;
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      float in[NS];
#pragma unroll
      for (int p = 0; p < NS; ++p) in[p] = src[p][r][c0 + k];
      float val[NQ];
      make(in, val);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int tap = k - j;  // output c0+j uses inputs c0+j .. c0+j+10
        if (tap >= 0 && tap < 11) {
#pragma unroll
          for (int q = 0; q < NQ; ++q) acc[q][j] += gw.w[tap] * val[q];
        }
      }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int j = 0; j < 4; ++j) dst[q][r][c0 + j] = acc[q][j];
  }
}

// vertical pass: thread (tx, ty) produces rows 4 ty .. 4 ty + 3 of column tx
template <int NQ>
__device__ __forceinline__ void conv_cols4(const float (*h)[kIn][kTile + 1], int tx, int ty, const Gauss11& gw,
                                           float (&out)[NQ][4]) {
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int j = 0; j < 4; ++j) out[q][j] = 0.f;
#pragma unroll
  for (int k = 0; k < 14; ++k) {
    float v[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) v[q] = h[q][4 * ty + k][tx];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int tap = k - j;
      if (tap >= 0 && tap < 11) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) out[q][j] += gw.w[tap] * v[q];
      }
    }
  }
}

template <int NS>
__device__ __forceinline__ void load_tile(float (*dst)[kIn][kIn + 1], const float* const (&planes)[NS], int H, int W, int x0,
                                          int y0, int tid) {
  for (int i = tid; i < kIn * kIn; i += kSsimThreads) {
    const int r = i / kIn, c = i - r * kIn;
    const int y = y0 + r - kHalo, x = x0 + c - kHalo;
    const bool in = y >= 0 && y < H && x >= 0 && x < W;  // zero padding
#pragma unroll
    for (int p = 0; p < NS; ++p) dst[p][r][c] = in ? __ldg(planes[p] + (size_t)y * W + x) : 0.f;
  }
}

__global__ void __launch_bounds__(kSsimThreads)
ssim_fwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2, int H, int W, const Gauss11 gw,
                float* __restrict__ maps, size_t map_stride, double* __restrict__ partial) {
  __shared__ float s[2][kIn][kIn + 1];
  __shared__ float h[5][kIn][kTile + 1];
  __shared__ float red[32];
  const int plane = blockIdx.z;
  const size_t base = (size_t)plane * H * W;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const int tid = threadIdx.x;
  const float* const planes[2] = {img1 + base, img2 + base};
  load_tile<2>(s, planes, H, W, x0, y0, tid);
  __syncthreads();
  conv_rows4<2, 5>(s, h, tid, gw, [](const float (&in)[2], float (&v)[5]) {
    v[0] = in[0]; v[1] = in[1]; v[2] = in[0] * in[0]; v[3] = in[1] * in[1]; v[4] = in[0] * in[1];
  });
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;
  float o[5][4];
  conv_cols4<5>(h, tx, ty, gw, o);
  float val = 0.f;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = x0 + tx, y = y0 + 4 * ty + j;
    if (x < W && y < H) {
      const float mu1 = o[0][j], mu2 = o[1][j], e11 = o[2][j], e22 = o[3][j], e12 = o[4][j];
      const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
      const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
      const float sg1 = e11 - mu1_sq, sg2 = e22 - mu2_sq, sg12 = e12 - mu12;
      const float A = mu1_sq + mu2_sq + C1, B = sg1 + sg2 + C2, C = 2.f * mu12 + C1, D = 2.f * sg12 + C2;
      // two reciprocals instead of seven divisions (A >= C1 > 0; B >= C2 up to rounding of the variances)
      const float iA = 1.0f / A, iB = 1.0f / B;
      const float iAB = iA * iB;
      const float ssim = (C * D) * iAB;
      val += ssim;
      if (maps) {
        const size_t p = base + (size_t)y * W + x;
        // d ssim / d mu1 (through mu1, mu1^2 and mu1 mu2), d ssim / d E[x^2], d ssim / d E[xy]
        maps[p] = 2.f * mu2 * (D - C) * iAB - 2.f * mu1 * ssim * (iA - iB);
        maps[map_stride + p] = -ssim * iB;
        maps[2 * map_stride + p] = (2.f * C) * iAB;
      }
    }
  }
  const float sum = block_sum(val, red, tid, kSsimThreads);
  if (tid == 0) partial[((size_t)plane * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = (double)sum;
}

__global__ void ssim_finalize_kernel(const double* __restrict__ partial, int per_item, double denom,
                                     float* __restrict__ out) {
  __shared__ double sm[32];
  const double s = cta_sum_strided(partial + (size_t)blockIdx.x * per_item, per_item, 1, 0, sm);
  if (threadIdx.x == 0) out[blockIdx.x] = (float)(s / denom);
}

__global__ void __launch_bounds__(kSsimThreads)
ssim_bwd_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                const float* __restrict__ maps, size_t map_stride, const float* __restrict__ gscale,
                int C, int H, int W, const Gauss11 gw, float* __restrict__ grad) {
  __shared__ float m[3][kIn][kIn + 1];
  __shared__ float h[3][kIn][kTile + 1];
  const int plane = blockIdx.z;
  const size_t base = (size_t)plane * H * W;
  const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
  const int tid = threadIdx.x;
  const float* const planes[3] = {maps + base, maps + map_stride + base, maps + 2 * map_stride + base};
  load_tile<3>(m, planes, H, W, x0, y0, tid);
  __syncthreads();
  conv_rows4<3, 3>(m, h, tid, gw, [](const float (&in)[3], float (&v)[3]) { v[0] = in[0]; v[1] = in[1]; v[2] = in[2]; });
  __syncthreads();
  const int tx = tid & 31, ty = tid >> 5;
  float o[3][4];
  conv_cols4<3>(h, tx, ty, gw, o);
  const float sc = __ldg(gscale + plane / C) / ((float)C * (float)H * (float)W);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int x = x0 + tx, y = y0 + 4 * ty + j;
    if (x < W && y < H) {
      const size_t p = base + (size_t)y * W + x;
      grad[p] = sc * (o[0][j] + 2.f * __ldg(img1 + p) * o[1][j] + __ldg(img2 + p) * o[2][j]);
    }
  }
}

// ---------------------------------------------------------------- SSIM, any odd window
// The tiled kernels above are specialised for the reference's default window (11 taps, the one every call site uses).
// ssim(img1, img2, window_size) with another odd size takes these two-pass kernels: a horizontal pass into a scratch
// of five (forward) or three (backward) planes, a vertical pass that finishes the map.  Same separable Gaussian
// (gaussian(window_size, 1.5), loss_utils.py:24-26), same zero padding of window_size / 2, same derivative maps.
constexpr int kMaxWindow = 63;
struct GaussN { float w[kMaxWindow]; int n; };

template <int NQ, bool SSIM_IN>
__global__ void __launch_bounds__(256)
window_rows_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
                   size_t in_stride, int H, int W, const GaussN gw, float* __restrict__ tmp, size_t plane_elems) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const size_t base = (size_t)blockIdx.z * H * W + (size_t)y * W;
  const int R = gw.n / 2;
  float acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.f;
  for (int k = 0; k < gw.n; ++k) {
    const int xx = x + k - R;
    if (xx < 0 || xx >= W) continue;  // zero padding
    const float w = gw.w[k];
    if (SSIM_IN) {
      const float p = __ldg(a + base + xx), q = __ldg(b + base + xx);
      acc[0] += w * p; acc[1] += w * q; acc[2] += w * (p * p); acc[3] += w * (q * q); acc[4] += w * (p * q);
    } else {
      acc[0] += w * __ldg(a + base + xx);
      acc[1] += w * __ldg(b + base + xx);
      acc[2] += w * __ldg(c + base + xx);
    }
  }
  (void)in_stride;
#pragma unroll
  for (int q = 0; q < NQ; ++q) tmp[q * plane_elems + base + x] = acc[q];
}

__global__ void __launch_bounds__(256)
window_cols_ssim_kernel(const float* __restrict__ tmp, size_t plane_elems, int H, int W, const GaussN gw,
                        float* __restrict__ maps, size_t map_stride, double* __restrict__ partial) {
  __shared__ float red[32];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const size_t base = (size_t)blockIdx.z * H * W;
  float val = 0.f;
  if (x < W && y < H) {
    const int R = gw.n / 2;
    float o[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < gw.n; ++k) {
      const int yy = y + k - R;
      if (yy < 0 || yy >= H) continue;
      const float w = gw.w[k];
#pragma unroll
      for (int q = 0; q < 5; ++q) o[q] += w * __ldg(tmp + q * plane_elems + base + (size_t)yy * W + x);
    }
    const float mu1 = o[0], mu2 = o[1], e11 = o[2], e22 = o[3], e12 = o[4];
    const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
    const float mu1_sq = mu1 * mu1, mu2_sq = mu2 * mu2, mu12 = mu1 * mu2;
    const float sg1 = e11 - mu1_sq, sg2 = e22 - mu2_sq, sg12 = e12 - mu12;
    const float A = mu1_sq + mu2_sq + C1, B = sg1 + sg2 + C2, C = 2.f * mu12 + C1, D = 2.f * sg12 + C2;
    const float iA = 1.0f / A, iB = 1.0f / B, iAB = iA * iB;
    const float ssim = (C * D) * iAB;
    val = ssim;
    if (maps) {
      const size_t p = base + (size_t)y * W + x;
      maps[p] = 2.f * mu2 * (D - C) * iAB - 2.f * mu1 * ssim * (iA - iB);
      maps[map_stride + p] = -ssim * iB;
      maps[2 * map_stride + p] = (2.f * C) * iAB;
    }
  }
  const float sum = block_sum(val, red, threadIdx.x, 256);
  if (threadIdx.x == 0) partial[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = (double)sum;
}

__global__ void __launch_bounds__(256)
window_cols_grad_kernel(const float* __restrict__ tmp, size_t plane_elems, const float* __restrict__ img1,
                        const float* __restrict__ img2, const float* __restrict__ gscale, int C, int H, int W,
                        const GaussN gw, float* __restrict__ grad) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const size_t base = (size_t)blockIdx.z * H * W;
  const int R = gw.n / 2;
  float o[3] = {0.f, 0.f, 0.f};
  for (int k = 0; k < gw.n; ++k) {
    const int yy = y + k - R;
    if (yy < 0 || yy >= H) continue;
    const float w = gw.w[k];
#pragma unroll
    for (int q = 0; q < 3; ++q) o[q] += w * __ldg(tmp + q * plane_elems + base + (size_t)yy * W + x);
  }
  const size_t p = base + (size_t)y * W + x;
  const float sc = __ldg(gscale + blockIdx.z / C) / ((float)C * (float)H * (float)W);
  grad[p] = sc * (o[0] + 2.f * __ldg(img1 + p) * o[1] + __ldg(img2 + p) * o[2]);
}

// ---------------------------------------------------------------- get_img_grad_weight
__global__ void __launch_bounds__(256)
grad_weight_raw_kernel(const float* __restrict__ img, int C, int H, int W, float* __restrict__ out,
                       float* __restrict__ pmin, float* __restrict__ pmax) {
  __shared__ float smin[8], smax[8];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  float v = 0.f;
  const bool interior = x >= 1 && x < W - 1 && y >= 1 && y < H - 1;
  if (interior) {
    float gx = 0.f, gy = 0.f;
    for (int c = 0; c < C; ++c) {
      const float* p = img + ((size_t)c * H + y) * W + x;
      gx += fabsf(__ldg(p + 1) - __ldg(p - 1));
      gy += fabsf(__ldg(p - W) - __ldg(p + W));
    }
    v = fmaxf(gx / (float)C, gy / (float)C);
    out[(size_t)y * W + x] = v;
  }
  float lo = interior ? v : __int_as_float(0x7f800000), hi = interior ? v : -__int_as_float(0x7f800000);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { lo = fminf(lo, smin[i]); hi = fmaxf(hi, smax[i]); }
    pmin[blockIdx.y * gridDim.x + blockIdx.x] = lo;
    pmax[blockIdx.y * gridDim.x + blockIdx.x] = hi;
  }
}

__global__ void minmax_finalize_kernel(const float* __restrict__ pmin, const float* __restrict__ pmax,
                                       int n, float* __restrict__ mm) {
  __shared__ float smin[256], smax[256];
  float lo = __int_as_float(0x7f800000), hi = -__int_as_float(0x7f800000);
  for (int i = threadIdx.x; i < n; i += blockDim.x) { lo = fminf(lo, pmin[i]); hi = fmaxf(hi, pmax[i]); }
  smin[threadIdx.x] = lo; smax[threadIdx.x] = hi;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      smin[threadIdx.x] = fminf(smin[threadIdx.x], smin[threadIdx.x + o]);
      smax[threadIdx.x] = fmaxf(smax[threadIdx.x], smax[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { mm[0] = smin[0]; mm[1] = smax[0]; }
}

__global__ void __launch_bounds__(256)
grad_weight_norm_kernel(int H, int W, const float* __restrict__ mm, float* __restrict__ out) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= W || y >= H) return;
  const bool interior = x >= 1 && x < W - 1 && y >= 1 && y < H - 1;
  const size_t p = (size_t)y * W + x;
  out[p] = interior ? (out[p] - mm[0]) / (mm[1] - mm[0]) : 1.0f;
}

// ---------------------------------------------------------------- lncc (one warp per patch row)
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <bool BWD>
__global__ void __launch_bounds__(256)
lncc_kernel(const float* __restrict__ ref, const float* __restrict__ nea, int bs, int tps,
            float* __restrict__ ncc, uint8_t* __restrict__ mask, const float* __restrict__ grad_ncc,
            float* __restrict__ grad_ref, float* __restrict__ grad_nea) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= bs) return;
  const float* r = ref + (size_t)row * tps;
  const float* n = nea + (size_t)row * tps;
  float sr = 0.f, sn = 0.f, srr = 0.f, snn = 0.f, srn = 0.f;
  for (int i = lane; i < tps; i += 32) {
    const float a = __ldg(r + i), b = __ldg(n + i);
    sr += a; sn += b; srr += a * a; snn += b * b; srn += a * b;
  }
  sr = warp_sum(sr); sn = warp_sum(sn); srr = warp_sum(srr); snn = warp_sum(snn); srn = warp_sum(srn);
  const float ra = sr / (float)tps, na = sn / (float)tps;
  const float cross = srn - na * sr, rv = srr - ra * sr, nv = snn - na * sn;
  const float den = rv * nv + 1e-8f;
  const float cc = cross * cross / den;
  const float raw = 1.f - cc;
  const float v = fminf(fmaxf(raw, 0.f), 2.f);
  if (!BWD) {
    if (lane == 0) { ncc[row] = v; mask[row] = v < 0.9f ? 1 : 0; }
    return;
  }
  // clamp passes the gradient where 0 <= raw <= 2 (torch.clamp); mean over a single column is the identity
  const float g = (raw >= 0.f && raw <= 2.f) ? -__ldg(grad_ncc + row) : 0.f;
  const float k1 = 2.f * cross / den, k2 = cross * cross / (den * den);
  for (int i = lane; i < tps; i += 32) {
    const float a = __ldg(r + i), b = __ldg(n + i);
    grad_ref[(size_t)row * tps + i] = g * (k1 * (b - na) - k2 * nv * 2.f * (a - ra));
    grad_nea[(size_t)row * tps + i] = g * (k1 * (a - ra) - k2 * rv * 2.f * (b - na));
  }
}

// ---------------------------------------------------------------- scale regularisation
struct ScaleRegCtl { double sum; double count; float coef; float pad; };

__global__ void __launch_bounds__(kRedThreads)
scale_reg_sum_kernel(const float* __restrict__ scaling, int64_t N, const int64_t* __restrict__ vis_idx,
                     const uint8_t* __restrict__ vis_mask, int64_t n_items, double* __restrict__ psum,
                     double* __restrict__ pcnt) {
  __shared__ float red[32];
  float acc = 0.f, cnt = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t g;
    if (vis_idx) {
      g = vis_idx[i];
      if (g < 0 || g >= N) continue;
    } else {
      if (!vis_mask[i]) continue;
      g = i;
    }
    const float m = fmaxf(fmaxf(scaling[3 * g], scaling[3 * g + 1]), scaling[3 * g + 2]);
    if (m > 0.01f) { acc += (m - 0.01f) * (m - 0.01f); cnt += 1.f; }
  }
  const float s = block_sum(acc, red, threadIdx.x, kRedThreads);
  const float c = block_sum(cnt, red, threadIdx.x, kRedThreads);
  if (threadIdx.x == 0) { psum[blockIdx.x] = (double)s; pcnt[blockIdx.x] = (double)c; }
}

__global__ void scale_reg_finalize_kernel(const double* __restrict__ psum, const double* __restrict__ pcnt,
                                          int n, float* __restrict__ out, ScaleRegCtl* __restrict__ ctl) {
  __shared__ double sm[32];
  const double s = cta_sum_strided(psum, n, 1, 0, sm);
  const double c = cta_sum_strided(pcnt, n, 1, 0, sm);
  if (threadIdx.x != 0) return;
  float loss = 0.f, coef = 0.f;
  if (c > 0.0) {
    const float raw = (float)(s / c);
    loss = fminf(fmaxf(raw, 0.f), 0.01f);
    coef = (raw >= 0.f && raw <= 0.01f) ? (float)(2.0 / c) : 0.f;  // saturated clamp: zero gradient
  }
  out[0] = loss;
  ctl->sum = s; ctl->count = c; ctl->coef = coef;
}

__global__ void __launch_bounds__(kRedThreads)
scale_reg_grad_kernel(const float* __restrict__ scaling, int64_t N, const int64_t* __restrict__ vis_idx,
                      const uint8_t* __restrict__ vis_mask, int64_t n_items,
                      const ScaleRegCtl* __restrict__ ctl, float* __restrict__ grad) {
  const float coef = ctl->coef;
  if (coef == 0.f) return;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_items;
       i += (int64_t)gridDim.x * blockDim.x) {
    int64_t g;
    if (vis_idx) {
      g = vis_idx[i];
      if (g < 0 || g >= N) continue;
    } else {
      if (!vis_mask[i]) continue;
      g = i;
    }
    const float a = scaling[3 * g], b = scaling[3 * g + 1], c = scaling[3 * g + 2];
    const float m = fmaxf(fmaxf(a, b), c);
    if (m > 0.01f) {
      const int arg = (a == m) ? 0 : ((b == m) ? 1 : 2);  // first maximum, as torch.max(dim)
      atomicAdd(grad + 3 * g + arg, coef * (m - 0.01f));
    }
  }
}

// gaussian(11, 1.5) of the reference (loss_utils.py:24-26): float32 values normalised by their float32 sum.
Gauss11 gauss_window() {
  Gauss11 g;
  float sum = 0.f;
  for (int x = 0; x < 11; ++x) {
    g.w[x] = (float)std::exp(-(double)((x - 5) * (x - 5)) / (2.0 * 1.5 * 1.5));
    sum += g.w[x];
  }
  for (int x = 0; x < 11; ++x) g.w[x] = g.w[x] / sum;
  return g;
}

// gaussian(n, 1.5), same evaluation order
bool gauss_window_n(int n, GaussN* g) {
  if (n < 1 || n > kMaxWindow || (n & 1) == 0) return false;
  g->n = n;
  float sum = 0.f;
  for (int x = 0; x < n; ++x) {
    g->w[x] = (float)std::exp(-(double)((x - n / 2) * (x - n / 2)) / (2.0 * 1.5 * 1.5));
    sum += g->w[x];
  }
  for (int x = 0; x < n; ++x) g->w[x] = g->w[x] / sum;
  for (int x = n; x < kMaxWindow; ++x) g->w[x] = 0.f;
  return true;
}

}  // namespace
}  // namespace hg

using namespace hg;

extern "C" {

size_t hg_reduce_workspace_bytes(int64_t) { return sizeof(double) * 2 * kRedBlocks + 256; }

int hg_l1_loss(const float* a, const float* b, int64_t n, float* out, float* grad_a, void* ws, void* st) {
  return run_pixel_loss<false>(a, b, n, out, grad_a, ws, (cudaStream_t)st);
}
int hg_l2_loss(const float* a, const float* b, int64_t n, float* out, float* grad_a, void* ws, void* st) {
  return run_pixel_loss<true>(a, b, n, out, grad_a, ws, (cudaStream_t)st);
}

int hg_pixel_loss_backward(const float* a, const float* b, int64_t n, int32_t squared, const float* gscale, float* grad_a,
                           float* grad_b, void* st_) {
  if (n < 0 || (n > 0 && (!a || !b || (!grad_a && !grad_b)))) {
    set_error("hg_pixel_loss_backward: NULL pointer");
    return HG_ERR_INVALID_ARG;
  }
  if (n == 0) return HG_OK;
  cudaStream_t st = (cudaStream_t)st_;
  const int64_t blocks = (n + 255) / 256;
  const unsigned grid = (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
  if (squared) pixel_loss_bwd_kernel<true><<<grid, 256, 0, st>>>(a, b, n, gscale, grad_a, grad_b);
  else pixel_loss_bwd_kernel<false><<<grid, 256, 0, st>>>(a, b, n, gscale, grad_a, grad_b);
  HG_POST_LAUNCH(false, st, "pixel_loss_bwd");
  return HG_OK;
}

int hg_freq_total(const float* freq_loss, const float* scale_loss, const float* hf_count, float lambda_freq,
                  float lambda_scale, float* out3, void* st_) {
  if (!out3 || (scale_loss && !hf_count)) {
    set_error("hg_freq_total: NULL pointer");
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)st_;
  freq_total_kernel<<<1, 32, 0, st>>>(freq_loss, scale_loss, hf_count, lambda_freq, lambda_scale, out3);
  HG_POST_LAUNCH(false, st, "freq_total");
  return HG_OK;
}

int hg_training_loss_value(const float* l1, const float* ssim, const float* extra0, const float* extra1,
                           float lambda_dssim, float* out, void* st_) {
  if (!l1 || !ssim || !out) {
    set_error("hg_training_loss_value: NULL pointer");
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)st_;
  training_loss_value_kernel<<<1, 32, 0, st>>>(l1, ssim, extra0, extra1, lambda_dssim, out);
  HG_POST_LAUNCH(false, st, "training_loss_value");
  return HG_OK;
}

int hg_training_image_grad(const float* color, const float* gt, const float* g_ssim, const float* g_freq, int64_t n,
                           float w_l1, float w_ssim, const float* w_freq, float* out, void* st_) {
  if (n < 0 || (n > 0 && (!color || !gt || !out))) {
    set_error("hg_training_image_grad: NULL pointer");
    return HG_ERR_INVALID_ARG;
  }
  if (n == 0) return HG_OK;
  cudaStream_t st = (cudaStream_t)st_;
  const int64_t blocks = (n + 255) / 256;
  training_image_grad_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, st>>>(
      color, gt, g_ssim, g_freq, n, w_l1 / (float)n, w_ssim, w_freq, out);
  HG_POST_LAUNCH(false, st, "training_image_grad");
  return HG_OK;
}

size_t hg_ssim_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W) {
  const size_t tiles = (size_t)((W + kTile - 1) / kTile) * ((H + kTile - 1) / kTile);
  return sizeof(double) * tiles * (size_t)B * C + 256;
}

int hg_ssim(const float* img1, const float* img2, int32_t B, int32_t C, int32_t H, int32_t W, float* out,
            float* maps, void* ws, void* st_) {
  if (!img1 || !img2 || !out || !ws || B <= 0 || C <= 0 || H <= 0 || W <= 0 || (size_t)B * C > 65535) {
    set_error("hg_ssim: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)st_;
  const dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, B * C);
  ssim_fwd_kernel<<<grid, kSsimThreads, 0, st>>>(img1, img2, H, W, gauss_window(), maps, (size_t)B * C * H * W, (double*)ws);
  HG_POST_LAUNCH(false, st, "ssim_fwd");
  ssim_finalize_kernel<<<B, 1024, 0, st>>>((const double*)ws, (int)(grid.x * grid.y * C), (double)C * H * W, out);
  HG_POST_LAUNCH(false, st, "ssim_finalize");
  return HG_OK;
}

int hg_ssim_backward(const float* img1, const float* img2, const float* maps, const float* gscale, int32_t B,
                     int32_t C, int32_t H, int32_t W, float* grad_img1, void* st_) {
  if (!img1 || !img2 || !maps || !gscale || !grad_img1 || B <= 0 || C <= 0 || H <= 0 || W <= 0) {
    set_error("hg_ssim_backward: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)st_;
  const dim3 grid((W + kTile - 1) / kTile, (H + kTile - 1) / kTile, B * C);
  ssim_bwd_kernel<<<grid, kSsimThreads, 0, st>>>(img1, img2, maps, (size_t)B * C * H * W, gscale, C, H, W,
                                                      gauss_window(), grad_img1);
  HG_POST_LAUNCH(false, st, "ssim_bwd");
  return HG_OK;
}

size_t hg_ssim_window_workspace_bytes(int32_t B, int32_t C, int32_t H, int32_t W) {
  const size_t blocks = (size_t)((W + 31) / 32) * ((H + 7) / 8);
  return sizeof(double) * blocks * (size_t)B * C + 256 + sizeof(float) * 5 * (size_t)B * C * H * W;
}

int hg_ssim_window(const float* img1, const float* img2, int32_t B, int32_t C, int32_t H, int32_t W, int32_t window_size,
                   float* out, float* maps, void* ws, void* st_) {
  GaussN gw;
  if (!img1 || !img2 || !out || !ws || B <= 0 || C <= 0 || H <= 0 || W <= 0 || (size_t)B * C > 65535 ||
      !gauss_window_n(window_size, &gw)) {
    set_error("hg_ssim_window: bad argument (window_size must be odd, 1..%d)", kMaxWindow);
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)st_;
  const dim3 grid((W + 31) / 32, (H + 7) / 8, B * C);
  const size_t n = (size_t)B * C * H * W;
  double* partial = (double*)ws;
  float* tmp = (float*)((char*)ws + align_up(sizeof(double) * grid.x * grid.y * grid.z, 256));
  window_rows_kernel<5, true><<<grid, 256, 0, st>>>(img1, img2, nullptr, 0, H, W, gw, tmp, n);
  HG_POST_LAUNCH(false, st, "ssim_window_rows");
  window_cols_ssim_kernel<<<grid, 256, 0, st>>>(tmp, n, H, W, gw, maps, n, partial);
  HG_POST_LAUNCH(false, st, "ssim_window_cols");
  ssim_finalize_kernel<<<B, 1024, 0, st>>>(partial, (int)(grid.x * grid.y * C), (double)C * H * W, out);
  HG_POST_LAUNCH(false, st, "ssim_finalize");
  return HG_OK;
}

int hg_ssim_window_backward(const float* img1, const float* img2, const float* maps, const float* gscale, int32_t B,
                            int32_t C, int32_t H, int32_t W, int32_t window_size, float* grad_img1, void* ws, void* st_) {
  GaussN gw;
  if (!img1 || !img2 || !maps || !gscale || !grad_img1 || !ws || B <= 0 || C <= 0 || H <= 0 || W <= 0 ||
      !gauss_window_n(window_size, &gw)) {
    set_error("hg_ssim_window_backward: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)st_;
  const dim3 grid((W + 31) / 32, (H + 7) / 8, B * C);
  const size_t n = (size_t)B * C * H * W;
  float* tmp = (float*)((char*)ws + align_up(sizeof(double) * grid.x * grid.y * grid.z, 256));
  window_rows_kernel<3, false><<<grid, 256, 0, st>>>(maps, maps + n, maps + 2 * n, n, H, W, gw, tmp, n);
  HG_POST_LAUNCH(false, st, "ssim_window_bwd_rows");
  window_cols_grad_kernel<<<grid, 256, 0, st>>>(tmp, n, img1, img2, gscale, C, H, W, gw, grad_img1);
  HG_POST_LAUNCH(false, st, "ssim_window_bwd_cols");
  return HG_OK;
}

size_t hg_img_grad_weight_workspace_bytes(int32_t H, int32_t W) {
  const size_t blocks = (size_t)((W + 31) / 32) * ((H + 7) / 8);
  return sizeof(float) * (2 * blocks + 2) + 256;
}

int hg_img_grad_weight(const float* img, int32_t C, int32_t H, int32_t W, float* out, void* ws, void* st_) {
  if (!img || !out || !ws || C <= 0 || H < 3 || W < 3) {
    set_error("hg_img_grad_weight: bad argument (needs H, W >= 3)");
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)st_;
  const dim3 grid((W + 31) / 32, (H + 7) / 8);
  const int blocks = grid.x * grid.y;
  float* pmin = (float*)ws;
  float* pmax = pmin + blocks;
  float* mm = pmax + blocks;
  grad_weight_raw_kernel<<<grid, 256, 0, st>>>(img, C, H, W, out, pmin, pmax);
  HG_POST_LAUNCH(false, st, "grad_weight_raw");
  minmax_finalize_kernel<<<1, 256, 0, st>>>(pmin, pmax, blocks, mm);
  HG_POST_LAUNCH(false, st, "minmax_finalize");
  grad_weight_norm_kernel<<<grid, 256, 0, st>>>(H, W, mm, out);
  HG_POST_LAUNCH(false, st, "grad_weight_norm");
  return HG_OK;
}

int hg_lncc(const float* ref, const float* nea, int32_t bs, int32_t tps, float* ncc, uint8_t* mask, void* st_) {
  if (!ref || !nea || !ncc || !mask || bs < 0 || tps <= 0) {
    set_error("hg_lncc: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  if (bs == 0) return HG_OK;
  cudaStream_t st = (cudaStream_t)st_;
  lncc_kernel<false><<<(bs + 7) / 8, 256, 0, st>>>(ref, nea, bs, tps, ncc, mask, nullptr, nullptr, nullptr);
  HG_POST_LAUNCH(false, st, "lncc");
  return HG_OK;
}

int hg_lncc_backward(const float* ref, const float* nea, const float* grad_ncc, int32_t bs, int32_t tps,
                     float* grad_ref, float* grad_nea, void* st_) {
  if (!ref || !nea || !grad_ncc || !grad_ref || !grad_nea || bs < 0 || tps <= 0) {
    set_error("hg_lncc_backward: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  if (bs == 0) return HG_OK;
  cudaStream_t st = (cudaStream_t)st_;
  lncc_kernel<true><<<(bs + 7) / 8, 256, 0, st>>>(ref, nea, bs, tps, nullptr, nullptr, grad_ncc, grad_ref, grad_nea);
  HG_POST_LAUNCH(false, st, "lncc_bwd");
  return HG_OK;
}

size_t hg_scale_reg_workspace_bytes(int64_t) { return sizeof(double) * 2 * kRedBlocks + sizeof(ScaleRegCtl) + 256; }

int hg_scale_reg(const float* scaling, int64_t N, const int64_t* vis_idx, const uint8_t* vis_mask, int64_t n_vis,
                 float* out, float* grad_scaling, void* ws, void* st_) {
  if (!scaling || !out || !ws || N < 0 || (!vis_idx && !vis_mask && n_vis != 0)) {
    set_error("hg_scale_reg: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  cudaStream_t st = (cudaStream_t)st_;
  const int64_t n_items = vis_mask ? N : (vis_idx ? n_vis : 0);  // an empty index list selects nothing
  double* psum = (double*)ws;
  double* pcnt = psum + kRedBlocks;
  ScaleRegCtl* ctl = (ScaleRegCtl*)(pcnt + kRedBlocks);
  scale_reg_sum_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(scaling, N, vis_idx, vis_mask, n_items, psum, pcnt);
  HG_POST_LAUNCH(false, st, "scale_reg_sum");
  scale_reg_finalize_kernel<<<1, 256, 0, st>>>(psum, pcnt, kRedBlocks, out, ctl);
  HG_POST_LAUNCH(false, st, "scale_reg_finalize");
  if (grad_scaling) {
    HG_CUDA_TRY(cudaMemsetAsync(grad_scaling, 0, sizeof(float) * 3 * (size_t)N, st));
    scale_reg_grad_kernel<<<kRedBlocks, kRedThreads, 0, st>>>(scaling, N, vis_idx, vis_mask, n_items, ctl, grad_scaling);
    HG_POST_LAUNCH(false, st, "scale_reg_grad");
  }
  return HG_OK;
}

}  // extern "C"
