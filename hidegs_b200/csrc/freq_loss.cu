// freq_loss.cu — multi-scale frequency regulariser: gray pyramid, Sobel/Laplacian loss, shared-memory
// mixed-radix FFT, spectral (log-magnitude / phase / band-energy) loss, its gradient back to the
// rendered image, and the high-frequency mask.
//
// Replaces TrueFrequencyPyramidRegularizer of scripts/frequency_regularization.py: build_pyramid
// (:1073-1082), _compute_spatial_frequency_loss (:1327-1360), compute_fft_features (:1084-1164),
// _compute_fft_frequency_loss (:1362-1401), compute_true_frequency_loss (:1293-1325) and
// detect_true_high_frequency_regions (:1166-1271), including the autograd backward of all of them
// (torch.clamp passes gradient only inside [min, max]; torch.min(a, b) splits it on ties; angle/abs
// have zero gradient at 0).
//
// (Pipeline: see "fused frequency-regulariser pipeline" below — six launches forward + backward.)
//
// FFT: Stockham autosort in shared memory, radix 4/2/3/5 butterflies (1080 x 1920 and its pyramid
// factor as 2^a 3^b 5^c) plus a generic O(R^2) butterfly for other prime factors, twiddles from a per-length
// table (host double precision, L1-resident).  Rows: one CTA per PAIR of image rows (two real rows ride one
// complex transform) -> 2 x (W/2+1) complex outputs; the rendered image and the ground truth of a level share
// a launch.  Columns: one CTA per group of adjacent columns, whole column resident in
// shared memory (<= 96 KB).  The spectrum of a level (<= 8.3 MB) stays in L2 between the two passes.
// The inverse used by the backward is the conjugate of the forward (conj -> FFT -> conj) followed by
// a Hermitian completion per row, i.e. an unnormalised C2R.
//
// No dense matmul / tensor-core DFT: at fp32 accuracy (the loss takes log|F| of coefficients down to
// the noise floor) a TF32 DFT-as-GEMM does not meet the 1e-3 loss tolerance, and the FFT work here is
// a few MFLOP per level.
#include "common.cuh"
#include "reduce.cuh"
#include "../../include/hidegs_losses.h"

#include <cmath>
#include <map>
#include <mutex>
#include <utility>
#include <vector>

namespace hg {

namespace {

constexpr int kMaxRadices = 14;
constexpr int kFftThreads = 256;
constexpr float kPi = 3.14159265358979323846f;

// A length-n transform: radix schedule + the device table tw[k] = exp(-2 pi i k / n), k in [0, n), computed once
// per (device, n) on the host in double precision and kept for the life of the process (<= 32 KB per length).
struct Plan {
  int n;
  int nr;
  int radix[kMaxRadices];
  const float2* tw;
};

const float2* twiddle_table(int n) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, float2*> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  auto it = cache.find({dev, n});
  if (it != cache.end()) return it->second;
  std::vector<float2> h((size_t)n);
  for (int k = 0; k < n; ++k) {
    const double a = -2.0 * 3.14159265358979323846 * (double)k / (double)n;
    h[k] = make_float2((float)std::cos(a), (float)std::sin(a));
  }
  float2* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(float2) * (size_t)n) != cudaSuccess) return nullptr;
  // synchronous copy: the table is complete before any stream can launch a kernel that reads it
  if (cudaMemcpy(d, h.data(), sizeof(float2) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(d);
    return nullptr;
  }
  cache[{dev, n}] = d;
  return d;
}

// fewest passes over the radices {10, 9, 8, 6, 5, 4, 3, 2} for m = 2^a 3^b 5^c (0: m has another prime factor)
static int fewest_passes(int m, int* out) {
  static const int kRadices[] = {10, 9, 8, 6, 5, 4, 3, 2};
  if (m == 1) return 0;
  int best = 0, tmp[kMaxRadices], keep[kMaxRadices];
  for (int r : kRadices) {
    if (m % r) continue;
    const int sub = fewest_passes(m / r, tmp + 1);
    if (m / r != 1 && sub == 0) continue;
    if (best == 0 || sub + 1 < best) {
      best = sub + 1;
      keep[0] = r;
      for (int i = 0; i < sub; ++i) keep[1 + i] = tmp[1 + i];
    }
  }
  for (int i = 0; i < best && i < kMaxRadices; ++i) out[i] = keep[i];
  return best;
}

static bool build_plan(int n, Plan* p);

// plans are built once per (device, length) and kept (the pass search and the twiddle upload run once)
bool make_plan(int n, Plan* p) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, Plan> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find({dev, n});
    if (it != cache.end()) { *p = it->second; return true; }
  }
  if (!build_plan(n, p)) return false;
  std::lock_guard<std::mutex> lk(mu);
  cache[{dev, n}] = *p;
  return true;
}

static bool build_plan(int n, Plan* p) {
  p->n = n;
  p->nr = 0;
  int m = n, smooth = 1;
  for (int f : {2, 3, 5})
    while (m % f == 0) { m /= f; smooth *= f; }
  if (smooth > 1) {
    int r[kMaxRadices];
    const int k = fewest_passes(smooth, r);
    if (k <= 0 || k > kMaxRadices) return false;
    for (int i = 0; i < k; ++i) p->radix[p->nr++] = r[i];
  }
  // any other prime factor runs through the generic O(R^2) butterfly (a prime length is one pass of
  // radix n, i.e. the plain DFT): every size works, sizes of the form 2^a 3^b 5^c are the fast path
  for (int f = 7; m > 1 && p->nr < kMaxRadices; f += 2) {
    while (m % f == 0 && p->nr < kMaxRadices) { p->radix[p->nr++] = f; m /= f; }
    if (f * f > m && m > 1) { p->radix[p->nr++] = m; m = 1; }
  }
  if (!(m == 1 && p->nr <= kMaxRadices)) return false;
  p->tw = twiddle_table(n);
  return p->tw != nullptr;
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
// multiply by -i (forward transform convention e^{-i...})
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

template <int R>
__device__ __forceinline__ void dft(float2* v);
template <>
__device__ __forceinline__ void dft<2>(float2* v) {
  const float2 a = v[0], b = v[1];
  v[0] = cadd(a, b);
  v[1] = csub(a, b);
}
template <>
__device__ __forceinline__ void dft<4>(float2* v) {
  const float2 a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
  const float2 c = cadd(v[1], v[3]), d = mul_mi(csub(v[1], v[3]));
  v[0] = cadd(a, c);
  v[1] = cadd(b, d);
  v[2] = csub(a, c);
  v[3] = csub(b, d);
}
template <>
__device__ __forceinline__ void dft<3>(float2* v) {
  const float s = 0.86602540378443864676f;  // sin(2 pi / 3)
  const float2 t1 = cadd(v[1], v[2]);
  const float2 t2 = make_float2(v[0].x - 0.5f * t1.x, v[0].y - 0.5f * t1.y);
  const float2 d = csub(v[1], v[2]);
  const float2 t3 = make_float2(s * d.y, -s * d.x);  // -i * s * d
  v[0] = cadd(v[0], t1);
  v[1] = cadd(t2, t3);
  v[2] = csub(t2, t3);
}
template <>
__device__ __forceinline__ void dft<5>(float2* v) {
  const float c1 = 0.30901699437494742410f, c2 = -0.80901699437494742410f;  // cos(2pi/5), cos(4pi/5)
  const float s1 = 0.95105651629515357212f, s2 = 0.58778525229247312917f;   // sin(2pi/5), sin(4pi/5)
  const float2 a1 = cadd(v[1], v[4]), b1 = csub(v[1], v[4]);
  const float2 a2 = cadd(v[2], v[3]), b2 = csub(v[2], v[3]);
  const float2 x0 = v[0];
  const float2 m1 = make_float2(x0.x + c1 * a1.x + c2 * a2.x, x0.y + c1 * a1.y + c2 * a2.y);
  const float2 m2 = make_float2(x0.x + c2 * a1.x + c1 * a2.x, x0.y + c2 * a1.y + c1 * a2.y);
  // -i * (s1 b1 + s2 b2) and -i * (s2 b1 - s1 b2)
  const float2 n1 = make_float2(s1 * b1.y + s2 * b2.y, -(s1 * b1.x + s2 * b2.x));
  const float2 n2 = make_float2(s2 * b1.y - s1 * b2.y, -(s2 * b1.x - s1 * b2.x));
  v[0] = make_float2(x0.x + a1.x + a2.x, x0.y + a1.y + a2.y);
  v[1] = cadd(m1, n1);
  v[4] = csub(m1, n1);
  v[2] = cadd(m2, n2);
  v[3] = csub(m2, n2);
}

// ---- composite butterflies (Cooley-Tukey inside the registers of one thread): radix R1 * R2 from dft<R1>, dft<R2>
// and the constant twiddles W_N^m = exp(-2 pi i m / N).  Radices 6, 8, 9, 10 bring 1080 / 1920-point transforms (and
// their pyramid halves) down to 3-4 shared-memory passes instead of 5-6.
template <int N>
__device__ __forceinline__ float2 twc(int m);
template <>
__device__ __forceinline__ float2 twc<6>(int m) {
  switch (m) {
    case 1: return make_float2(0.5f, -0.866025404f);
    case 2: return make_float2(-0.5f, -0.866025404f);
    case 3: return make_float2(-1.f, 0.f);
    case 4: return make_float2(-0.5f, 0.866025404f);
    case 5: return make_float2(0.5f, 0.866025404f);
    default: return make_float2(1.f, 0.f);
  }
}
template <>
__device__ __forceinline__ float2 twc<8>(int m) {
  switch (m) {
    case 1: return make_float2(0.707106781f, -0.707106781f);
    case 2: return make_float2(0.f, -1.f);
    case 3: return make_float2(-0.707106781f, -0.707106781f);
    case 4: return make_float2(-1.f, 0.f);
    case 5: return make_float2(-0.707106781f, 0.707106781f);
    case 6: return make_float2(0.f, 1.f);
    case 7: return make_float2(0.707106781f, 0.707106781f);
    default: return make_float2(1.f, 0.f);
  }
}
template <>
__device__ __forceinline__ float2 twc<9>(int m) {
  switch (m) {
    case 1: return make_float2(0.766044443f, -0.64278761f);
    case 2: return make_float2(0.173648178f, -0.984807753f);
    case 3: return make_float2(-0.5f, -0.866025404f);
    case 4: return make_float2(-0.939692621f, -0.342020143f);
    case 5: return make_float2(-0.939692621f, 0.342020143f);
    case 6: return make_float2(-0.5f, 0.866025404f);
    case 7: return make_float2(0.173648178f, 0.984807753f);
    case 8: return make_float2(0.766044443f, 0.64278761f);
    default: return make_float2(1.f, 0.f);
  }
}
template <>
__device__ __forceinline__ float2 twc<10>(int m) {
  switch (m) {
    case 1: return make_float2(0.809016994f, -0.587785252f);
    case 2: return make_float2(0.309016994f, -0.951056516f);
    case 3: return make_float2(-0.309016994f, -0.951056516f);
    case 4: return make_float2(-0.809016994f, -0.587785252f);
    case 5: return make_float2(-1.f, 0.f);
    case 6: return make_float2(-0.809016994f, 0.587785252f);
    case 7: return make_float2(-0.309016994f, 0.951056516f);
    case 8: return make_float2(0.309016994f, 0.951056516f);
    case 9: return make_float2(0.809016994f, 0.587785252f);
    default: return make_float2(1.f, 0.f);
  }
}

template <int R1, int R2>
__device__ __forceinline__ void dft_ct(float2* v) {
  constexpr int N = R1 * R2;
  float2 y[N];  // y[k1 * R2 + i2]
#pragma unroll
  for (int i2 = 0; i2 < R2; ++i2) {
    float2 t[R1];
#pragma unroll
    for (int i1 = 0; i1 < R1; ++i1) t[i1] = v[i1 * R2 + i2];
    dft<R1>(t);
#pragma unroll
    for (int k1 = 0; k1 < R1; ++k1) {
      const int m = (i2 * k1) % N;
      y[k1 * R2 + i2] = (m == 0) ? t[k1] : cmul(t[k1], twc<N>(m));
    }
  }
#pragma unroll
  for (int k1 = 0; k1 < R1; ++k1) {
    float2 t[R2];
#pragma unroll
    for (int i2 = 0; i2 < R2; ++i2) t[i2] = y[k1 * R2 + i2];
    dft<R2>(t);
#pragma unroll
    for (int k2 = 0; k2 < R2; ++k2) v[k1 + R1 * k2] = t[k2];
  }
}
template <>
__device__ __forceinline__ void dft<6>(float2* v) { dft_ct<2, 3>(v); }
template <>
__device__ __forceinline__ void dft<8>(float2* v) { dft_ct<2, 4>(v); }
template <>
__device__ __forceinline__ void dft<9>(float2* v) { dft_ct<3, 3>(v); }
template <>
__device__ __forceinline__ void dft<10>(float2* v) { dft_ct<2, 5>(v); }

// w / d and w % d for 0 <= w < 2^24, 1 <= d <= 2^16 without an integer division (float reciprocal + one correction)
__device__ __forceinline__ void fast_divmod(int w, int d, float inv_d, int& q, int& r) {
  q = (int)((float)w * inv_d);
  r = w - q * d;
  if (r < 0) { r += d; --q; }
  else if (r >= d) { r -= d; ++q; }
}

// One Stockham pass of radix R over `batch` independent length-n sequences stored back to back.
// Twiddle of butterfly input r at position k of a sub-transform of length ns*R:  tw[(k r) * n / (ns R)].
template <int R>
__device__ __forceinline__ void stockham_pass(const float2* __restrict__ src, float2* __restrict__ dst, int n,
                                              int ns, int batch, const float2* __restrict__ tw) {
  const int per = n / R;
  const int tstep = n / (ns * R);
  const float inv_per = 1.0f / (float)per, inv_ns = 1.0f / (float)ns;
  for (int w = threadIdx.x; w < per * batch; w += blockDim.x) {
    int b = 0, j = w;
    if (batch > 1) fast_divmod(w, per, inv_per, b, j);
    const float2* s = src + b * n;
    float2* d = dst + b * n;
    int q = 0, k = 0;
    if (ns > 1) fast_divmod(j, ns, inv_ns, q, k);
    float2 v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = s[j + r * per];
    if (ns > 1) {
      const int t1 = k * tstep;
#pragma unroll
      for (int r = 1; r < R; ++r) v[r] = cmul(v[r], __ldg(tw + t1 * r));
    }
    dft<R>(v);
    const int j0 = (j - k) * R + k;
#pragma unroll
    for (int r = 0; r < R; ++r) d[j0 + r * ns] = v[r];
  }
}

// Generic radix (any R, used for prime factors > 5): one output element per work item.
__device__ __forceinline__ void stockham_pass_generic(const float2* __restrict__ src, float2* __restrict__ dst,
                                                      int n, int ns, int R, int batch,
                                                      const float2* __restrict__ tw) {
  const int per = n / R;
  const int period = ns * R;
  const int tstep = n / period;
  for (int w = threadIdx.x; w < n * batch; w += blockDim.x) {
    const int b = w / n, e = w - b * n;
    const int j = e / R, r = e - j * R;
    const int k = j % ns;
    const float2* s = src + b * n;
    const int step = k + r * ns;  // phase advance per input index, in units of 2 pi / period
    float2 acc = make_float2(0.f, 0.f);
    int ph = 0;
    for (int t = 0; t < R; ++t) {
      acc = cadd(acc, cmul(s[j + t * per], __ldg(tw + ph * tstep)));
      ph += step;
      if (ph >= period) ph -= period;
    }
    dst[b * n + (j - k) * R + k + r * ns] = acc;
  }
}

// Forward FFT of `batch` sequences in shared memory; returns the buffer that holds the result.
__device__ float2* fft_smem(float2* a, float2* b, const Plan& p, int batch) {
  int ns = 1;
  float2 *src = a, *dst = b;
  for (int i = 0; i < p.nr; ++i) {
    const int R = p.radix[i];
    if (R == 8) stockham_pass<8>(src, dst, p.n, ns, batch, p.tw);
    else if (R == 10) stockham_pass<10>(src, dst, p.n, ns, batch, p.tw);
    else if (R == 9) stockham_pass<9>(src, dst, p.n, ns, batch, p.tw);
    else if (R == 6) stockham_pass<6>(src, dst, p.n, ns, batch, p.tw);
    else if (R == 4) stockham_pass<4>(src, dst, p.n, ns, batch, p.tw);
    else if (R == 2) stockham_pass<2>(src, dst, p.n, ns, batch, p.tw);
    else if (R == 3) stockham_pass<3>(src, dst, p.n, ns, batch, p.tw);
    else if (R == 5) stockham_pass<5>(src, dst, p.n, ns, batch, p.tw);
    else stockham_pass_generic(src, dst, p.n, ns, R, batch, p.tw);
    ns *= R;
    __syncthreads();
    float2* t = src; src = dst; dst = t;
  }
  return src;
}

// Up to two images per launch (blockIdx.y): the rendered image and the ground truth of a pyramid level.
struct RealPair { const float* src[2]; float2* dst[2]; };
struct SpecPair { float2* spec[2]; };
struct InvPair { const float2* src[2]; float* dst[2]; };

// ---- rows: real [H][W] (optionally clamped to [0,1]) -> half spectrum [H][W/2+1].
// TWO image rows ride one complex transform (row 2j in the real part, row 2j+1 in the imaginary part):
//   A[k] = (Z[k] + conj Z[N-k]) / 2,   B[k] = (Z[k] - conj Z[N-k]) / (2i).
__global__ void __launch_bounds__(kFftThreads)
fft_rows_r2c_kernel(RealPair io, int H, int W, int clamp01, Plan plan) {
  extern __shared__ float2 sm[];
  float2 *a = sm, *b = sm + W;
  const int r0 = 2 * blockIdx.x, r1 = r0 + 1;
  const float* img = io.src[blockIdx.y];
  float2* spec = io.dst[blockIdx.y];
  const float* s0 = img + (size_t)r0 * W;
  const float* s1 = img + (size_t)r1 * W;
  const bool two = r1 < H;
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    float v0 = __ldg(s0 + i), v1 = two ? __ldg(s1 + i) : 0.f;
    if (clamp01) { v0 = fminf(fmaxf(v0, 0.f), 1.f); v1 = fminf(fmaxf(v1, 0.f), 1.f); }
    a[i] = make_float2(v0, v1);
  }
  __syncthreads();
  const float2* z = fft_smem(a, b, plan, 1);
  const int Wh = W / 2 + 1;
  float2* d0 = spec + (size_t)r0 * Wh;
  float2* d1 = spec + (size_t)r1 * Wh;
  for (int k = threadIdx.x; k < Wh; k += blockDim.x) {
    const float2 p = z[k], q = z[k == 0 ? 0 : W - k];
    d0[k] = make_float2(0.5f * (p.x + q.x), 0.5f * (p.y - q.y));
    if (two) d1[k] = make_float2(0.5f * (p.y + q.y), -0.5f * (p.x - q.x));
  }
}

// ---- columns, in place on [H][Wh]; INVERSE: conj -> FFT -> conj (unnormalised inverse)
template <bool INVERSE>
__global__ void __launch_bounds__(kFftThreads)
fft_cols_kernel(SpecPair io, int H, int Wh, int tc, Plan plan) {
  extern __shared__ float2 sm[];
  float2 *a = sm, *b = sm + (size_t)tc * H;
  float2* spec = io.spec[blockIdx.y];
  const int c0 = blockIdx.x * tc;
  const int nc = min(tc, Wh - c0);
  for (int i = threadIdx.x; i < H * tc; i += blockDim.x) {
    const int y = i / tc, c = i - y * tc;
    float2 v = make_float2(0.f, 0.f);
    if (c < nc) v = spec[(size_t)y * Wh + c0 + c];
    if (INVERSE) v.y = -v.y;
    a[c * H + y] = v;
  }
  __syncthreads();
  const float2* r = fft_smem(a, b, plan, tc);
  for (int i = threadIdx.x; i < H * tc; i += blockDim.x) {
    const int y = i / tc, c = i - y * tc;
    if (c < nc) {
      float2 v = r[c * H + y];
      if (INVERSE) v.y = -v.y;
      spec[(size_t)y * Wh + c0 + c] = v;
    }
  }
}

// ---- rows of the inverse: Hermitian completion of the half rows, inverse transform, real part * scale.
// Two rows ride one transform again: Z = A + i B with A, B the completed (exactly Hermitian: the imaginary parts of
// the DC / Nyquist bins cannot reach the real output and are dropped first) rows; the transform of conj(Z) is
// conj(a + i b) for the real rows a, b.
__global__ void __launch_bounds__(kFftThreads)
fft_rows_c2r_kernel(InvPair io, int H, int W, float scale, Plan plan) {
  extern __shared__ float2 sm[];
  float2 *a = sm, *b = sm + W;
  const int r0 = 2 * blockIdx.x, r1 = r0 + 1;
  const int Wh = W / 2 + 1;
  const float2* spec = io.src[blockIdx.y];
  float* img = io.dst[blockIdx.y];
  const float2* s0 = spec + (size_t)r0 * Wh;
  const float2* s1 = spec + (size_t)r1 * Wh;
  const bool two = r1 < H;
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    const bool upper = i >= Wh;
    const int k = upper ? W - i : i;
    float2 A = s0[k], B = two ? s1[k] : make_float2(0.f, 0.f);
    if (k == 0 || 2 * k == W) { A.y = 0.f; B.y = 0.f; }
    if (upper) { A.y = -A.y; B.y = -B.y; }  // conj(S[N-k])
    // conj(A + i B) = (A.x - B.y) - i (A.y + B.x)
    a[i] = make_float2(A.x - B.y, -(A.y + B.x));
  }
  __syncthreads();
  const float2* r = fft_smem(a, b, plan, 1);
  float* d0 = img + (size_t)r0 * W;
  float* d1 = img + (size_t)r1 * W;
  for (int i = threadIdx.x; i < W; i += blockDim.x) {
    const float2 v = r[i];
    d0[i] = v.x * scale;
    if (two) d1[i] = -v.y * scale;
  }
}

struct FftCfg {
  Plan row, col;
  int tc;
  size_t smem_row, smem_col;
};

int make_cfg(int H, int W, FftCfg* c) {
  if (H <= 0 || W <= 1 || H > 4096 || W > 4096 || !make_plan(W, &c->row) || !make_plan(H, &c->col)) {
    set_error("FFT size %dx%d unsupported (each side must be in [1, 4096])", H, W);
    return HG_ERR_INVALID_ARG;
  }
  // 4 adjacent columns = one 32-byte sector per spectrum row; two resident CTAs per SM up to H = 2048
  c->tc = (int)((96 * 1024) / (16 * (size_t)H));
  if (c->tc > 4) c->tc = 4;
  if (c->tc < 1) c->tc = 1;
  c->smem_row = 2 * (size_t)W * sizeof(float2);
  c->smem_col = 2 * (size_t)c->tc * H * sizeof(float2);
  return HG_OK;
}

int set_smem_attrs() {
  static thread_local bool done[16] = {};
  int dev = 0;
  HG_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 16 && done[dev]) return HG_OK;
  const int maxb = 100 * 1024;
  HG_CUDA_TRY(cudaFuncSetAttribute(fft_rows_r2c_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
  HG_CUDA_TRY(cudaFuncSetAttribute(fft_rows_c2r_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
  HG_CUDA_TRY(cudaFuncSetAttribute(fft_cols_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
  HG_CUDA_TRY(cudaFuncSetAttribute(fft_cols_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
  if (dev >= 0 && dev < 16) done[dev] = true;
  return HG_OK;
}

// Forward 2-D transform of one or two real images of the same size (img1 / spec1 may be NULL).
int fft2_r2c(const float* img0, float2* spec0, const float* img1, float2* spec1, int H, int W, int clamp01,
             cudaStream_t st) {
  FftCfg c;
  int rc = make_cfg(H, W, &c);
  if (rc) return rc;
  rc = set_smem_attrs();
  if (rc) return rc;
  const int Wh = W / 2 + 1;
  const int nimg = img1 ? 2 : 1;
  RealPair rp{{img0, img1}, {spec0, spec1}};
  fft_rows_r2c_kernel<<<dim3((H + 1) / 2, nimg), kFftThreads, c.smem_row, st>>>(rp, H, W, clamp01, c.row);
  HG_POST_LAUNCH(false, st, "fft_rows_r2c");
  SpecPair sp{{spec0, spec1}};
  fft_cols_kernel<false><<<dim3((Wh + c.tc - 1) / c.tc, nimg), kFftThreads, c.smem_col, st>>>(sp, H, Wh, c.tc, c.col);
  HG_POST_LAUNCH(false, st, "fft_cols");
  return HG_OK;
}

// spec is destroyed (the column pass runs in place)
int fft2_c2r(float2* spec, int H, int W, float scale, float* img, cudaStream_t st) {
  FftCfg c;
  int rc = make_cfg(H, W, &c);
  if (rc) return rc;
  rc = set_smem_attrs();
  if (rc) return rc;
  const int Wh = W / 2 + 1;
  SpecPair sp{{spec, nullptr}};
  fft_cols_kernel<true><<<dim3((Wh + c.tc - 1) / c.tc, 1), kFftThreads, c.smem_col, st>>>(sp, H, Wh, c.tc, c.col);
  HG_POST_LAUNCH(false, st, "ifft_cols");
  InvPair ip{{spec, nullptr}, {img, nullptr}};
  fft_rows_c2r_kernel<<<dim3((H + 1) / 2, 1), kFftThreads, c.smem_row, st>>>(ip, H, W, scale, c.row);
  HG_POST_LAUNCH(false, st, "ifft_rows_c2r");
  return HG_OK;
}

// =============================================================== fused frequency-regulariser pipeline
//
// The whole regulariser is six launches (nine with the ground-truth-only high-frequency mask), every intermediate is
// produced by the kernel that has its inputs on chip:
//
//   forward   pyramid_kernel        RGB -> gray pyramid of both images (3 levels from one 40x40 tile in shared memory),
//                                   Sobel / Laplacian sums of all levels, the ground truth's edge score (hf mask)
//             fft_rows_jobs_kernel  row transforms of every level and image in ONE launch (clamp on load, equal work per
//                                   CTA: 1 / 2 / 4 row pairs at levels 0 / 1 / 2)
//             fft_cols_jobs_kernel  column transforms of the rendered AND the ground-truth spectrum of the same columns
//                                   in one CTA -> the spectral sums come out of shared memory (no spectrum is re-read);
//                                   the hf job does forward columns -> high-pass -> inverse columns without leaving the
//                                   SM; the last CTA to finish reduces all partials (fixed order, double) and runs the
//                                   scalar epilogue (no reduce / finalize launches)
//   backward  spectral_grad_cols_kernel   spectral gradient formed on load + inverse columns
//             fft_rows_c2r_jobs_kernel    inverse rows of all levels, clamp gate on store
//             spatial_grad_rgb_kernel     Sobel / Laplacian gradient of all levels + un-pooling down the pyramid +
//                                         gray -> RGB + the upstream scalar, one pass over the image
//   hf mask   fft_rows_c2r_jobs_kernel (|.| + max), hf_combine_kernel (score + min / max), hf_threshold_kernel (mask + count)
constexpr int kMaxLevels = 3;
constexpr int kT0 = 32;   // level-0 tile edge of the image-space kernels
constexpr int kImgThreads = 256;

// Per-level control block written by the epilogue and read by the gradient kernels.
struct LevelCtl {
  float c_sobel, c_lap;      // coefficients of d(loss)/d(response) = c * response
  float c_mag, c_phase;      // spectral coefficients
  float c_band[4];
  float band_count[4];
};

// responses of Sobel-x / Sobel-y / Laplacian (cross-correlation, zero pad 1) around d[0]; rows `stride` floats apart
__device__ __forceinline__ void responses_at(const float* d, int stride, float& sx, float& sy, float& lp) {
  const float a00 = d[-stride - 1], a01 = d[-stride], a02 = d[-stride + 1];
  const float a10 = d[-1], a11 = d[0], a12 = d[1];
  const float a20 = d[stride - 1], a21 = d[stride], a22 = d[stride + 1];
  sx = -a00 + a02 - 2.f * a10 + 2.f * a12 - a20 + a22;
  sy = -a00 - 2.f * a01 - a02 + a20 + 2.f * a21 + a22;
  lp = -a01 - a10 + 4.f * a11 - a12 - a21;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------- image space, forward
struct PyrArgs {
  const float* rgb[2];            // [0] rendered (or the only image), [1] ground truth; NULL: gray[i][0] is read instead
  float* gray[2][kMaxLevels];     // gray pyramid per image (written for images that have an rgb source)
  int n_images, levels;
  int H[kMaxLevels], W[kMaxLevels];
  double* partial;                // [tiles][9] Sobel-x / Sobel-y / Laplacian sums of squares per level, or NULL
  float* hf_score;                // [H][W] 0.6 |grad| + 0.4 |lap| of image `hf_src`, or NULL
  int hf_src;
  unsigned* counters;             // 8 words zeroed by CTA 0 (tickets of the later last-block reductions)
};

template <int NI>
__global__ void __launch_bounds__(kImgThreads) pyramid_kernel(const __grid_constant__ PyrArgs a) {
  __shared__ float g0[2][kT0 + 8][kT0 + 9];
  __shared__ float g1[2][kT0 / 2 + 4][kT0 / 2 + 5];
  __shared__ float g2[2][kT0 / 4 + 2][kT0 / 4 + 3];
  __shared__ float red[kImgThreads / 32][9];
  const int tid = threadIdx.x;
  if (blockIdx.x == 0 && blockIdx.y == 0 && tid < 8 && a.counters) a.counters[tid] = 0u;
  const int x0 = blockIdx.x * kT0, y0 = blockIdx.y * kT0;
  const int H0 = a.H[0], W0 = a.W[0];
  const size_t hw0 = (size_t)H0 * W0;
  constexpr int E0 = kT0 + 8;
  for (int i = tid; i < E0 * E0; i += kImgThreads) {
    const int r = i / E0, c = i - r * E0;
    const int y = y0 + r - 4, x = x0 + c - 4;
    const bool in = y >= 0 && y < H0 && x >= 0 && x < W0;
    _Pragma("unroll") for (int im = 0; im < NI; ++im) {
      float v = 0.f;
      if (in) {
        const size_t p = (size_t)y * W0 + x;
        if (a.rgb[im]) v = (__ldg(a.rgb[im] + p) + __ldg(a.rgb[im] + hw0 + p) + __ldg(a.rgb[im] + 2 * hw0 + p)) / 3.0f;
        else v = __ldg(a.gray[im][0] + p);
      }
      g0[im][r][c] = v;
    }
  }
  __syncthreads();
  for (int i = tid; i < kT0 * kT0; i += kImgThreads) {
    const int r = i / kT0, c = i - r * kT0;
    const int y = y0 + r, x = x0 + c;
    if (y < H0 && x < W0)
      _Pragma("unroll") for (int im = 0; im < NI; ++im)
        if (a.rgb[im]) a.gray[im][0][(size_t)y * W0 + x] = g0[im][r + 4][c + 4];
  }
  const int H1 = a.levels > 1 ? a.H[1] : 0, W1 = a.levels > 1 ? a.W[1] : 0;
  const int H2 = a.levels > 2 ? a.H[2] : 0, W2 = a.levels > 2 ? a.W[2] : 0;
  if (a.levels > 1) {  // 2x2 average pooling (F.avg_pool2d(kernel 2, stride 2)), halo 2
    constexpr int E1 = kT0 / 2 + 4;
    for (int i = tid; i < E1 * E1; i += kImgThreads) {
      const int r = i / E1, c = i - r * E1;
      const int Y = (y0 >> 1) + r - 2, X = (x0 >> 1) + c - 2;
      const bool ok = Y >= 0 && Y < H1 && X >= 0 && X < W1;
      _Pragma("unroll") for (int im = 0; im < NI; ++im)
        g1[im][r][c] = ok ? (g0[im][2 * r][2 * c] + g0[im][2 * r][2 * c + 1] + g0[im][2 * r + 1][2 * c] +
                             g0[im][2 * r + 1][2 * c + 1]) * 0.25f : 0.f;
    }
    __syncthreads();
    {
      const int r = tid >> 4, c = tid & 15;
      const int Y = (y0 >> 1) + r, X = (x0 >> 1) + c;
      if (Y < H1 && X < W1)
        _Pragma("unroll") for (int im = 0; im < NI; ++im)
          if (a.rgb[im]) a.gray[im][1][(size_t)Y * W1 + X] = g1[im][r + 2][c + 2];
    }
  }
  if (a.levels > 2) {
    constexpr int E2 = kT0 / 4 + 2;
    for (int i = tid; i < E2 * E2; i += kImgThreads) {
      const int r = i / E2, c = i - r * E2;
      const int Y = (y0 >> 2) + r - 1, X = (x0 >> 2) + c - 1;
      const bool ok = Y >= 0 && Y < H2 && X >= 0 && X < W2;
      _Pragma("unroll") for (int im = 0; im < NI; ++im)
        g2[im][r][c] = ok ? (g1[im][2 * r][2 * c] + g1[im][2 * r][2 * c + 1] + g1[im][2 * r + 1][2 * c] +
                             g1[im][2 * r + 1][2 * c + 1]) * 0.25f : 0.f;
    }
    __syncthreads();
    if (tid < 64) {
      const int r = tid >> 3, c = tid & 7;
      const int Y = (y0 >> 2) + r, X = (x0 >> 2) + c;
      if (Y < H2 && X < W2)
        _Pragma("unroll") for (int im = 0; im < NI; ++im)
          if (a.rgb[im]) a.gray[im][2][(size_t)Y * W2 + X] = g2[im][r + 1][c + 1];
    }
  }
  // ---- ground-truth edge score of detect_true_high_frequency_regions (:1180-1207)
  if (a.hf_score) {
    const int s = a.hf_src;
    for (int i = tid; i < kT0 * kT0; i += kImgThreads) {
      const int r = i / kT0, c = i - r * kT0;
      const int y = y0 + r, x = x0 + c;
      if (y < H0 && x < W0) {
        float sx, sy, lp;
        responses_at(&g0[s][r + 4][c + 4], kT0 + 9, sx, sy, lp);
        a.hf_score[(size_t)y * W0 + x] = 0.6f * sqrtf(sx * sx + sy * sy + 1e-8f) + 0.4f * fabsf(lp);
      }
    }
  }
  if (!a.partial || NI < 2) return;
  // ---- Sobel / Laplacian of d = gray_rendered - gray_gt at every level (d = 0 outside the image)
  __syncthreads();
  for (int i = tid; i < E0 * E0; i += kImgThreads) {
    const int r = i / E0, c = i - r * E0;
    g0[0][r][c] -= g0[NI - 1][r][c];
  }
  if (a.levels > 1)
    for (int i = tid; i < (kT0 / 2 + 4) * (kT0 / 2 + 4); i += kImgThreads) {
      const int r = i / (kT0 / 2 + 4), c = i - r * (kT0 / 2 + 4);
      g1[0][r][c] -= g1[NI - 1][r][c];
    }
  if (a.levels > 2)
    for (int i = tid; i < (kT0 / 4 + 2) * (kT0 / 4 + 2); i += kImgThreads) {
      const int r = i / (kT0 / 4 + 2), c = i - r * (kT0 / 4 + 2);
      g2[0][r][c] -= g2[NI - 1][r][c];
    }
  __syncthreads();
  float acc[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) acc[q] = 0.f;
  for (int i = tid; i < kT0 * kT0; i += kImgThreads) {
    const int r = i / kT0, c = i - r * kT0;
    if (y0 + r < H0 && x0 + c < W0) {
      float sx, sy, lp;
      responses_at(&g0[0][r + 4][c + 4], kT0 + 9, sx, sy, lp);
      acc[0] += sx * sx; acc[1] += sy * sy; acc[2] += lp * lp;
    }
  }
  if (a.levels > 1) {
    const int r = tid >> 4, c = tid & 15;
    if ((y0 >> 1) + r < H1 && (x0 >> 1) + c < W1) {
      float sx, sy, lp;
      responses_at(&g1[0][r + 2][c + 2], kT0 / 2 + 5, sx, sy, lp);
      acc[3] = sx * sx; acc[4] = sy * sy; acc[5] = lp * lp;
    }
  }
  if (a.levels > 2 && tid < 64) {
    const int r = tid >> 3, c = tid & 7;
    if ((y0 >> 2) + r < H2 && (x0 >> 2) + c < W2) {
      float sx, sy, lp;
      responses_at(&g2[0][r + 1][c + 1], kT0 / 4 + 3, sx, sy, lp);
      acc[6] = sx * sx; acc[7] = sy * sy; acc[8] = lp * lp;
    }
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    const float v = warp_sum(acc[q]);
    if ((tid & 31) == 0) red[tid >> 5][q] = v;
  }
  __syncthreads();
  if (tid < 9) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < kImgThreads / 32; ++w) v += red[w][tid];
    a.partial[((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 9 + tid] = (double)v;
  }
}

// ---------------------------------------------------------------- rows, forward (all levels / images in one launch)
constexpr int kMaxRowJobs = 7;
struct RowJob {
  const float* src;   // [H][W] real
  float2* dst;        // [H][W/2+1]
  int H, W, clamp01, pairs, cta_begin;
  Plan plan;
};
struct RowArgs { int n_jobs; RowJob job[kMaxRowJobs]; };

__global__ void __launch_bounds__(kFftThreads, 5) fft_rows_jobs_kernel(const __grid_constant__ RowArgs a) {
  extern __shared__ float2 sm[];
  int j = 0;
  for (int k = 1; k < a.n_jobs; ++k)
    if ((int)blockIdx.x >= a.job[k].cta_begin) j = k;
  const RowJob& J = a.job[j];
  const int W = J.W, H = J.H, pairs = J.pairs;
  const int p0 = ((int)blockIdx.x - J.cta_begin) * pairs;
  float2 *bufa = sm, *bufb = sm + pairs * W;
  const float inv_w = 1.0f / (float)W;
#pragma unroll 4
  for (int i = threadIdx.x; i < pairs * W; i += blockDim.x) {
    int pr = 0, x = i;
    if (pairs > 1) fast_divmod(i, W, inv_w, pr, x);
    const int r0 = 2 * (p0 + pr), r1 = r0 + 1;
    float v0 = r0 < H ? __ldg(J.src + (size_t)r0 * W + x) : 0.f;
    float v1 = r1 < H ? __ldg(J.src + (size_t)r1 * W + x) : 0.f;
    if (J.clamp01) { v0 = fminf(fmaxf(v0, 0.f), 1.f); v1 = fminf(fmaxf(v1, 0.f), 1.f); }
    bufa[i] = make_float2(v0, v1);
  }
  __syncthreads();
  const float2* z = fft_smem(bufa, bufb, J.plan, pairs);
  const int Wh = W / 2 + 1;
  const float inv_wh = 1.0f / (float)Wh;
  for (int i = threadIdx.x; i < pairs * Wh; i += blockDim.x) {
    int pr = 0, k = i;
    if (pairs > 1) fast_divmod(i, Wh, inv_wh, pr, k);
    const int r0 = 2 * (p0 + pr), r1 = r0 + 1;
    if (r0 >= H) continue;
    const float2* zz = z + pr * W;
    const float2 p = zz[k], q = zz[k == 0 ? 0 : W - k];
    J.dst[(size_t)r0 * Wh + k] = make_float2(0.5f * (p.x + q.x), 0.5f * (p.y - q.y));
    if (r1 < H) J.dst[(size_t)r1 * Wh + k] = make_float2(0.5f * (p.y + q.y), -0.5f * (p.x - q.x));
  }
}

// ---------------------------------------------------------------- spectral terms
__device__ __forceinline__ int signed_freq(int k, int n) { return (k < (n + 1) / 2) ? k : k - n; }

__device__ __forceinline__ int band_of(int ky, int kx, int H, int W) {
  const int fy = signed_freq(ky, H), fx = signed_freq(kx, W);
  const float dist = sqrtf((float)(fy * fy + fx * fx));
  const float md = (float)min(H / 2, W / 2);
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (dist >= (float)i * md / 4.0f && dist < (float)(i + 1) * md / 4.0f) return i;
  return -1;
}

__device__ __forceinline__ bool within(float v, float lo, float hi) { return v >= lo && v <= hi; }
__device__ __forceinline__ float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

// per-CTA partial sums of the column kernel: mag, phase, band_r[4], band_r - band_g [4], count[4] (the loss), then the
// band sums [4] and counts [4] of the level-0 ground-truth spectrum (debug_info['freq_band_energies'])
constexpr int kSpecVals = 14;
constexpr int kColVals = kSpecVals + 8;

struct FinalizeArgs {
  int mode;                        // 0 none, 1 loss epilogue, 2 ground-truth state (band sums only)
  int levels;
  int H[kMaxLevels], W[kMaxLevels];
  const double* spatial_partial;   // [spatial_tiles][9]
  int spatial_tiles;
  int col_begin[kMaxLevels], col_end[kMaxLevels];  // CTAs of the column launch that hold level l's sums
  const double* band0_in;          // cached ground truth: the 8 final band sums; NULL: reduced from the level-0 partials
  double* band0_out;               // mode 2: the 8 final band sums
  LevelCtl* ctl;
  float* stats;
  unsigned* counter;
};

// Sums of ALL columns of a row-major [rows][NC] matrix of per-CTA partials at once, by the whole CTA (256 threads),
// in a fixed order: warp g, lane l adds the rows  g * RPW + l / NC + 8 RPW k  of column l % NC (RPW = 32 / NC rows per
// warp load, eight loads in flight per thread), then the 8 RPW row-slot sums of a column are added in index order.
// The last-CTA epilogue is a serial tail of the launch: its dependent L2 round trips are what has to be short.
template <int NC>
__device__ __forceinline__ void cta_column_sums(const double* __restrict__ p, int rows, double* __restrict__ out,
                                                double* __restrict__ scratch /* [8 * (32 / NC)][NC] */) {
  constexpr int RPW = 32 / NC;
  constexpr int SLOTS = (kFftThreads / 32) * RPW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / NC, col = lane - sub * NC;
  double acc = 0.0;
  if (sub < RPW) {
    double a[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) a[u] = 0.0;
    int r = warp * RPW + sub;
    for (; r + 7 * SLOTS < rows; r += 8 * SLOTS) {
#pragma unroll
      for (int u = 0; u < 8; ++u) a[u] += p[(size_t)(r + u * SLOTS) * NC + col];
    }
    for (int u = 0; r < rows; r += SLOTS, ++u) a[u & 7] += p[(size_t)r * NC + col];
    acc = ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
    scratch[(warp * RPW + sub) * NC + col] = acc;
  }
  __syncthreads();
  if (threadIdx.x < NC) {
    double v = 0.0;
#pragma unroll
    for (int sl = 0; sl < SLOTS; ++sl) v += scratch[sl * NC + threadIdx.x];
    out[threadIdx.x] = v;
  }
  __syncthreads();
}

// Scalar epilogue of compute_true_frequency_loss (:1293-1325, :1362-1401) and the coefficients its backward needs.
// Runs in the last CTA of the column launch; every sum is formed in a fixed order in double.
__device__ void freq_epilogue(const FinalizeArgs& a, const double* __restrict__ col_partial) {
  __shared__ double sums[kMaxLevels * (3 + kSpecVals) + 8];
  __shared__ double scratch[8 * 32];
  __shared__ double colv[kColVals], spv[9];
  constexpr int kPer = 3 + kSpecVals;
  if (a.mode == 1) {
    cta_column_sums<9>(a.spatial_partial, a.spatial_tiles, spv, scratch);
    if (threadIdx.x < 9) sums[(threadIdx.x / 3) * kPer + threadIdx.x % 3] = spv[threadIdx.x];
  }
  for (int l = 0; l < a.levels; ++l) {
    if (a.mode != 1 && l > 0) break;
    cta_column_sums<kColVals>(col_partial + (size_t)a.col_begin[l] * kColVals, a.col_end[l] - a.col_begin[l], colv, scratch);
    if (a.mode == 1 && threadIdx.x < kSpecVals) sums[l * kPer + 3 + threadIdx.x] = colv[threadIdx.x];
    if (l == 0 && threadIdx.x < 8)
      sums[kMaxLevels * kPer + threadIdx.x] = a.band0_in ? a.band0_in[threadIdx.x] : colv[kSpecVals + threadIdx.x];
    __syncthreads();
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  const double* e = sums + kMaxLevels * kPer;
  if (a.band0_out)
    for (int q = 0; q < 8; ++q) a.band0_out[q] = e[q];
  if (a.mode != 1) return;
  const float w_lvl[3] = {0.1f, 0.05f, 0.025f};
  float total = 0.f;
  float lvl_raw[3] = {0, 0, 0};
  float sp_raw[3], fft_raw[3], mag_raw[3], ph_raw[3], band_raw[3];
  float ediff[3][4], cnt[3][4];
  for (int l = 0; l < a.levels; ++l) {
    const double n = (double)a.H[l] * a.W[l];
    const double* s = sums + l * kPer;
    const float gx = (float)(s[0] / n), gy = (float)(s[1] / n), lap = (float)(s[2] / n);
    sp_raw[l] = 0.7f * (gx + gy) + 0.3f * lap;
    const double* v = s + 3;
    mag_raw[l] = (float)(v[0] / n);
    ph_raw[l] = (float)(v[1] / n);
    float bl = 0.f;
    for (int q = 0; q < 4; ++q) {
      cnt[l][q] = (float)v[10 + q];
      ediff[l][q] = cnt[l][q] > 0.f ? (float)(v[6 + q] / (v[10 + q] + 1e-8)) : 0.f;  // E_rendered - E_gt
      bl += ediff[l][q] * ediff[l][q];
    }
    band_raw[l] = bl / 4.f;
    const float mag = clampf(mag_raw[l], 0.f, 10.f), ph = clampf(ph_raw[l], 0.f, kPi), band = clampf(band_raw[l], 0.f, 100.f);
    fft_raw[l] = 0.6f * mag + 0.2f * ph + 0.2f * band;
    const float sp = clampf(sp_raw[l], 0.f, 1.f), ff = clampf(fft_raw[l], 0.f, 10.f);
    lvl_raw[l] = 0.7f * sp + 0.3f * ff;
    const float lvl = clampf(lvl_raw[l], 0.f, 0.1f);
    total += w_lvl[l] * lvl;
    float* st = a.stats + 1 + 6 * l;
    st[0] = sp; st[1] = ff; st[2] = lvl; st[3] = mag; st[4] = ph; st[5] = band;
  }
  a.stats[0] = clampf(total, 0.f, 0.1f);
  const float g_total = within(total, 0.f, 0.1f) ? 1.f : 0.f;
  for (int l = 0; l < a.levels; ++l) {
    const float n = (float)a.H[l] * (float)a.W[l];
    const float g_level = g_total * w_lvl[l] * (within(lvl_raw[l], 0.f, 0.1f) ? 1.f : 0.f);
    const float g_sp = g_level * 0.7f * (within(sp_raw[l], 0.f, 1.f) ? 1.f : 0.f);
    const float g_fft = g_level * 0.3f * (within(fft_raw[l], 0.f, 10.f) ? 1.f : 0.f);
    LevelCtl& c = a.ctl[l];
    c.c_sobel = g_sp * 0.7f * 2.f / n;
    c.c_lap = g_sp * 0.3f * 2.f / n;
    c.c_mag = g_fft * 0.6f * (within(mag_raw[l], 0.f, 10.f) ? 1.f : 0.f) * 2.f / n;
    c.c_phase = g_fft * 0.2f * (within(ph_raw[l], 0.f, kPi) ? 1.f : 0.f) / n;
    const float gb = g_fft * 0.2f * (within(band_raw[l], 0.f, 100.f) ? 1.f : 0.f);
    for (int q = 0; q < 4; ++q) {
      c.band_count[q] = cnt[l][q];
      c.c_band[q] = cnt[l][q] > 0.f ? gb * (2.f / 4.f) * ediff[l][q] / (cnt[l][q] + 1e-8f) : 0.f;
    }
  }
  // band energies of the level-0 ground truth (debug_info)
  for (int q = 0; q < 4; ++q) a.stats[19 + q] = e[4 + q] > 0.0 ? (float)(e[q] / (e[4 + q] + 1e-8)) : 0.f;
}

// ---------------------------------------------------------------- columns, forward
enum ColKind { kColPair = 0, kColPairCached = 1, kColHighpassInverse = 2, kColSingle = 3 };
constexpr int kMaxColJobs = 5;
struct ColJob {
  float2* A;          // spectrum transformed in place
  float2* B;          // second spectrum of the same level: transformed too (pair) or final (pair-cached); else NULL
  int H, W, kind, tc, cta_begin, level;
  Plan plan;
};
struct ColArgs {
  int n_jobs;
  ColJob job[kMaxColJobs];
  double* partial;    // [CTAs][kColVals]
  FinalizeArgs fin;
};

__global__ void __launch_bounds__(kFftThreads, 4) fft_cols_jobs_kernel(const __grid_constant__ ColArgs a) {
  extern __shared__ float2 sm[];
  __shared__ float red[kFftThreads / 32][kColVals];
  __shared__ int s_last;
  const int tid = threadIdx.x;
  int j = 0;
  for (int k = 1; k < a.n_jobs; ++k)
    if ((int)blockIdx.x >= a.job[k].cta_begin) j = k;
  const ColJob& J = a.job[j];
  const int H = J.H, W = J.W, Wh = W / 2 + 1, tc = J.tc, kind = J.kind;
  const int c0 = ((int)blockIdx.x - J.cta_begin) * tc;
  const int nc = min(tc, Wh - c0);
  const int nseq = kind == kColPair ? 2 * tc : tc;
  const int tsh = tc == 4 ? 2 : (tc == 2 ? 1 : 0);  // tc is 1, 2 or 4
  float2 *bufa = sm, *bufb = sm + (size_t)nseq * H;
#pragma unroll 4
  for (int i = tid; i < H * tc; i += kFftThreads) {
    const int y = i >> tsh, c = i - (y << tsh);
    float2 v = make_float2(0.f, 0.f), w = v;
    if (c < nc) {
      v = J.A[(size_t)y * Wh + c0 + c];
      if (kind == kColPair) w = J.B[(size_t)y * Wh + c0 + c];
    }
    bufa[c * H + y] = v;
    if (kind == kColPair) bufa[(tc + c) * H + y] = w;
  }
  __syncthreads();
  float2* r = fft_smem(bufa, bufb, J.plan, nseq);
  if (kind == kColHighpassInverse) {
    // high-pass of detect_true_high_frequency_regions (:1229-1243: keep dist > 0.3 * min(h//2, w//2)), then the inverse
    // column transform (conj -> FFT -> conj) of the same columns
    const float radius = (float)((double)min(H / 2, W / 2) * 0.3);
    for (int i = tid; i < H * tc; i += kFftThreads) {
      const int y = i >> tsh, c = i - (y << tsh);
      const int fy = signed_freq(y, H), fx = signed_freq(c0 + c, W);
      const float dist = sqrtf((float)(fy * fy + fx * fx));
      float2 v = r[c * H + y];
      if (!(dist > radius)) v = make_float2(0.f, 0.f);
      v.y = -v.y;
      r[c * H + y] = v;
    }
    __syncthreads();
    float2* r2 = fft_smem(r, r == bufa ? bufb : bufa, J.plan, tc);
    for (int i = tid; i < H * tc; i += kFftThreads) {
      const int y = i >> tsh, c = i - (y << tsh);
      if (c < nc) {
        float2 v = r2[c * H + y];
        v.y = -v.y;
        J.A[(size_t)y * Wh + c0 + c] = v;
      }
    }
  } else {
    float acc[kColVals];
#pragma unroll
    for (int q = 0; q < kColVals; ++q) acc[q] = 0.f;
    const bool sums = kind != kColSingle;
    const bool band0 = J.level == 0;
    for (int i = tid; i < H * tc; i += kFftThreads) {
      const int y = i >> tsh, c = i - (y << tsh);
      if (c >= nc) continue;
      const int kx = c0 + c;
      const size_t p = (size_t)y * Wh + kx;
      const float2 fa = r[c * H + y];
      J.A[p] = fa;
      float2 fb = make_float2(0.f, 0.f);
      if (kind == kColPair) {
        fb = r[(tc + c) * H + y];
        J.B[p] = fb;
      } else if (kind == kColPairCached) {
        fb = J.B[p];
      }
      const float w = (kx == 0 || (2 * kx == W)) ? 1.f : 2.f;  // Hermitian twin outside the half spectrum
      const int band = (sums || band0) ? band_of(y, kx, H, W) : -1;
      if (sums) {
        const float ma = hypotf(fa.x, fa.y), mb = hypotf(fb.x, fb.y);
        const float dl = logf(ma + 1e-6f) - logf(mb + 1e-6f);
        acc[0] += w * dl * dl;
        const float pa = atan2f(fa.y, fa.x), pb = atan2f(fb.y, fb.x);
        const float ad = fabsf(pa - pb);
        acc[1] += w * fminf(ad, 2.f * kPi - ad);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (band == q) {
            acc[2 + q] += w * ma;
            acc[6 + q] += w * (ma - mb);  // difference summed directly: E_r - E_g cancels
            acc[10 + q] += w;
            if (band0) { acc[kSpecVals + q] += w * mb; acc[kSpecVals + 4 + q] += w; }
          }
      } else if (band0 && band >= 0) {  // ground-truth state: band sums of this (ground-truth) spectrum
        const float m = hypotf(fa.x, fa.y);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (band == q) { acc[kSpecVals + q] += w * m; acc[kSpecVals + 4 + q] += w; }
      }
    }
#pragma unroll
    for (int q = 0; q < kColVals; ++q) {
      const float v = warp_sum(acc[q]);
      if ((tid & 31) == 0) red[tid >> 5][q] = v;
    }
    __syncthreads();
    if (tid < kColVals) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < kFftThreads / 32; ++w) v += red[w][tid];
      a.partial[(size_t)blockIdx.x * kColVals + tid] = (double)v;
    }
  }
  if (a.fin.mode == 0) return;
  // ---- the last CTA to arrive reduces every partial and runs the scalar epilogue
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(a.fin.counter, 1u) == gridDim.x - 1 ? 1 : 0;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  freq_epilogue(a.fin, a.partial);
}

// ---------------------------------------------------------------- backward: spectral gradient + inverse columns
struct GradJob {
  float2* A;          // rendered spectrum in, inverse-column-transformed gradient out
  const float2* B;    // ground-truth spectrum
  int H, W, tc, cta_begin, level;
  Plan plan;
};
struct GradArgs { int n_jobs; GradJob job[kMaxLevels]; const LevelCtl* ctl; };

__global__ void __launch_bounds__(kFftThreads, 4) spectral_grad_cols_kernel(const __grid_constant__ GradArgs a) {
  extern __shared__ float2 sm[];
  const int tid = threadIdx.x;
  int j = 0;
  for (int k = 1; k < a.n_jobs; ++k)
    if ((int)blockIdx.x >= a.job[k].cta_begin) j = k;
  const GradJob& J = a.job[j];
  const int H = J.H, W = J.W, Wh = W / 2 + 1, tc = J.tc;
  const int c0 = ((int)blockIdx.x - J.cta_begin) * tc;
  const int nc = min(tc, Wh - c0);
  const LevelCtl* ctl = a.ctl + J.level;
  const float c_mag = ctl->c_mag, c_phase = ctl->c_phase;
  const int tsh = tc == 4 ? 2 : (tc == 2 ? 1 : 0);  // tc is 1, 2 or 4
  float2 *bufa = sm, *bufb = sm + (size_t)tc * H;
#pragma unroll 2
  for (int i = tid; i < H * tc; i += kFftThreads) {
    const int ky = i >> tsh, c = i - (ky << tsh);
    float gre = 0.f, gim = 0.f;
    if (c < nc) {
      const int kx = c0 + c;
      const float2 fa = J.A[(size_t)ky * Wh + kx], fb = J.B[(size_t)ky * Wh + kx];
      const float ma = hypotf(fa.x, fa.y), mb = hypotf(fb.x, fb.y);
      if (ma > 0.f) {
        float dmag = c_mag * (logf(ma + 1e-6f) - logf(mb + 1e-6f)) / (ma + 1e-6f);
        const int band = band_of(ky, kx, H, W);
        if (band >= 0) dmag += ctl->c_band[band];
        gre = dmag * fa.x / ma;
        gim = dmag * fa.y / ma;
        const float pa = atan2f(fa.y, fa.x), pb = atan2f(fb.y, fb.x);
        const float dlt = pa - pb, ad = fabsf(dlt);
        const float sgn = dlt > 0.f ? 1.f : (dlt < 0.f ? -1.f : 0.f);
        const float other = 2.f * kPi - ad;
        // d min(|D|, 2pi - |D|) / dD, with torch.min's even split on ties
        const float dw = ad < other ? sgn : (ad > other ? -sgn : 0.f);
        const float gp = c_phase * dw / (ma * ma);
        gre += gp * (-fa.y);
        gim += gp * fa.x;
      }
    }
    bufa[c * H + ky] = make_float2(gre, -gim);  // conj: the inverse is conj -> FFT -> conj
  }
  __syncthreads();
  const float2* r = fft_smem(bufa, bufb, J.plan, tc);
  for (int i = tid; i < H * tc; i += kFftThreads) {
    const int y = i >> tsh, c = i - (y << tsh);
    if (c < nc) {
      float2 v = r[c * H + y];
      v.y = -v.y;
      J.A[(size_t)y * Wh + c0 + c] = v;
    }
  }
}

// ---------------------------------------------------------------- rows, inverse (all levels in one launch)
struct InvJob {
  const float2* spec;  // [H][W/2+1], columns already inverse-transformed
  float* dst;          // [H][W]
  const float* gate;   // mode 0: gray image whose clamp(0, 1) gates the gradient (NULL: no gate)
  int H, W, pairs, cta_begin, mode;  // mode 0: dst = x * scale (gated); mode 1: dst = |x * scale| and the maximum of it
  float scale;
  Plan plan;
};
struct InvArgs {
  int n_jobs;
  InvJob job[kMaxLevels];
  float* pmax;         // mode 1: per-CTA maxima, then mm[1] = the global maximum (last CTA)
  float* mm;
  unsigned* counter;
};

__global__ void __launch_bounds__(kFftThreads, 5) fft_rows_c2r_jobs_kernel(const __grid_constant__ InvArgs a) {
  extern __shared__ float2 sm[];
  __shared__ float smax[kFftThreads / 32];
  __shared__ int s_last;
  int j = 0;
  for (int k = 1; k < a.n_jobs; ++k)
    if ((int)blockIdx.x >= a.job[k].cta_begin) j = k;
  const InvJob& J = a.job[j];
  const int W = J.W, H = J.H, Wh = W / 2 + 1, pairs = J.pairs;
  const int p0 = ((int)blockIdx.x - J.cta_begin) * pairs;
  float2 *bufa = sm, *bufb = sm + pairs * W;
  const float inv_w = 1.0f / (float)W;
#pragma unroll 4
  for (int i = threadIdx.x; i < pairs * W; i += blockDim.x) {
    int pr = 0, x = i;
    if (pairs > 1) fast_divmod(i, W, inv_w, pr, x);
    const int r0 = 2 * (p0 + pr), r1 = r0 + 1;
    const bool upper = x >= Wh;
    const int k = upper ? W - x : x;
    float2 A = r0 < H ? J.spec[(size_t)r0 * Wh + k] : make_float2(0.f, 0.f);
    float2 B = r1 < H ? J.spec[(size_t)r1 * Wh + k] : make_float2(0.f, 0.f);
    if (k == 0 || 2 * k == W) { A.y = 0.f; B.y = 0.f; }
    if (upper) { A.y = -A.y; B.y = -B.y; }  // conj(S[N-k])
    bufa[i] = make_float2(A.x - B.y, -(A.y + B.x));  // conj(A + i B)
  }
  __syncthreads();
  const float2* r = fft_smem(bufa, bufb, J.plan, pairs);
  float vmax = 0.f;
  for (int i = threadIdx.x; i < pairs * W; i += blockDim.x) {
    int pr = 0, x = i;
    if (pairs > 1) fast_divmod(i, W, inv_w, pr, x);
    const int r0 = 2 * (p0 + pr), r1 = r0 + 1;
    if (r0 >= H) continue;
    const float2 v = r[i];
    float o0 = v.x * J.scale, o1 = -v.y * J.scale;
    const size_t q0 = (size_t)r0 * W + x, q1 = (size_t)r1 * W + x;
    if (J.mode == 1) {
      o0 = fabsf(o0);
      o1 = fabsf(o1);
      vmax = fmaxf(vmax, o0);
      if (r1 < H) vmax = fmaxf(vmax, o1);
    } else if (J.gate) {  // d clamp(gray, 0, 1) / d gray
      const float g0 = __ldg(J.gate + q0);
      if (!(g0 >= 0.f && g0 <= 1.f)) o0 = 0.f;
      if (r1 < H) {
        const float g1 = __ldg(J.gate + q1);
        if (!(g1 >= 0.f && g1 <= 1.f)) o1 = 0.f;
      }
    }
    J.dst[q0] = o0;
    if (r1 < H) J.dst[q1] = o1;
  }
  if (J.mode != 1) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = vmax;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kFftThreads / 32; ++w) vmax = fmaxf(vmax, smax[w]);
    a.pmax[blockIdx.x] = vmax;
    __threadfence();
    s_last = atomicAdd(a.counter, 1u) == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  float m = 0.f;  // |.| >= 0; the maximum is order independent
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) m = fmaxf(m, a.pmax[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < kFftThreads / 32; ++w) m = fmaxf(m, smax[w]);
    a.mm[1] = m;
  }
}

// ---------------------------------------------------------------- image space, backward
struct BwdArgs {
  int levels;
  int H[kMaxLevels], W[kMaxLevels];
  const float* gr[kMaxLevels];   // gray pyramids
  const float* gg[kMaxLevels];
  const float* dg[kMaxLevels];   // gradient w.r.t. the gray level from the spectral terms (clamp-gated)
  const LevelCtl* ctl;
  const float* gscale;           // device scalar multiplied into the result (NULL: 1)
  float* out;                    // [3][H0][W0]
};

// scaled responses c * R on tile + halo 1 from d on tile + halo 2 (zero where the centre is outside the image)
template <int T>
__device__ __forceinline__ void scaled_responses(const float (*d)[T + 5], float (*rx)[T + 3], float (*ry)[T + 3],
                                                 float (*rl)[T + 3], int y0, int x0, int H, int W, float cs, float cl,
                                                 int tid) {
  for (int i = tid; i < (T + 2) * (T + 2); i += kImgThreads) {
    const int r = i / (T + 2), c = i - r * (T + 2);
    const int y = y0 + r - 1, x = x0 + c - 1;
    float sx = 0.f, sy = 0.f, lp = 0.f;
    if (y >= 0 && y < H && x >= 0 && x < W) responses_at(&d[r + 1][c + 1], T + 5, sx, sy, lp);
    rx[r][c] = cs * sx; ry[r][c] = cs * sy; rl[r][c] = cl * lp;
  }
}

// dL/dd(p) = sum_{u,v} K[u][v] * R(p - (u-1, v-1)); (r, c) index the response arrays (tile + halo 1)
template <int T>
__device__ __forceinline__ float stencil_adjoint(const float (*rx)[T + 3], const float (*ry)[T + 3],
                                                 const float (*rl)[T + 3], int r, int c) {
  const float gx = -rx[r + 1][c + 1] + rx[r + 1][c - 1] - 2.f * rx[r][c + 1] + 2.f * rx[r][c - 1] - rx[r - 1][c + 1] + rx[r - 1][c - 1];
  const float gy = -ry[r + 1][c + 1] - 2.f * ry[r + 1][c] - ry[r + 1][c - 1] + ry[r - 1][c + 1] + 2.f * ry[r - 1][c] + ry[r - 1][c - 1];
  const float gl = -rl[r + 1][c] - rl[r][c + 1] + 4.f * rl[r][c] - rl[r][c - 1] - rl[r - 1][c];
  return gx + gy + gl;
}

template <int T>
__device__ __forceinline__ void load_diff(float (*d)[T + 5], const float* __restrict__ gr, const float* __restrict__ gg,
                                          int y0, int x0, int H, int W, int tid) {
  for (int i = tid; i < (T + 4) * (T + 4); i += kImgThreads) {
    const int r = i / (T + 4), c = i - r * (T + 4);
    const int y = y0 + r - 2, x = x0 + c - 2;
    const bool in = y >= 0 && y < H && x >= 0 && x < W;
    d[r][c] = in ? (__ldg(gr + (size_t)y * W + x) - __ldg(gg + (size_t)y * W + x)) : 0.f;
  }
}

__global__ void __launch_bounds__(kImgThreads) spatial_grad_rgb_kernel(const __grid_constant__ BwdArgs a) {
  constexpr int T0 = kT0, T1 = kT0 / 2, T2 = kT0 / 4;
  __shared__ float d0[T0 + 4][T0 + 5], d1[T1 + 4][T1 + 5], d2[T2 + 4][T2 + 5];
  __shared__ float rx0[T0 + 2][T0 + 3], ry0[T0 + 2][T0 + 3], rl0[T0 + 2][T0 + 3];
  __shared__ float rx1[T1 + 2][T1 + 3], ry1[T1 + 2][T1 + 3], rl1[T1 + 2][T1 + 3];
  __shared__ float rx2[T2 + 2][T2 + 3], ry2[T2 + 2][T2 + 3], rl2[T2 + 2][T2 + 3];
  __shared__ float t1[T1][T1 + 1], t2[T2][T2 + 1];
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * T0, y0 = blockIdx.y * T0;
  const int H0 = a.H[0], W0 = a.W[0];
  const int H1 = a.levels > 1 ? a.H[1] : 0, W1 = a.levels > 1 ? a.W[1] : 0;
  const int H2 = a.levels > 2 ? a.H[2] : 0, W2 = a.levels > 2 ? a.W[2] : 0;
  load_diff<T0>(d0, a.gr[0], a.gg[0], y0, x0, H0, W0, tid);
  if (a.levels > 1) load_diff<T1>(d1, a.gr[1], a.gg[1], y0 >> 1, x0 >> 1, H1, W1, tid);
  if (a.levels > 2) load_diff<T2>(d2, a.gr[2], a.gg[2], y0 >> 2, x0 >> 2, H2, W2, tid);
  __syncthreads();
  scaled_responses<T0>(d0, rx0, ry0, rl0, y0, x0, H0, W0, a.ctl[0].c_sobel, a.ctl[0].c_lap, tid);
  if (a.levels > 1) scaled_responses<T1>(d1, rx1, ry1, rl1, y0 >> 1, x0 >> 1, H1, W1, a.ctl[1].c_sobel, a.ctl[1].c_lap, tid);
  if (a.levels > 2) scaled_responses<T2>(d2, rx2, ry2, rl2, y0 >> 2, x0 >> 2, H2, W2, a.ctl[2].c_sobel, a.ctl[2].c_lap, tid);
  __syncthreads();
  if (a.levels > 2 && tid < T2 * T2) {
    const int r = tid / T2, c = tid - r * T2;
    const int Y = (y0 >> 2) + r, X = (x0 >> 2) + c;
    float t = 0.f;
    if (Y < H2 && X < W2) t = __ldg(a.dg[2] + (size_t)Y * W2 + X) + stencil_adjoint<T2>(rx2, ry2, rl2, r + 1, c + 1);
    t2[r][c] = t;
  }
  __syncthreads();
  if (a.levels > 1) {
    const int r = tid / T1, c = tid - r * T1;
    const int Y = (y0 >> 1) + r, X = (x0 >> 1) + c;
    float t = 0.f;
    if (Y < H1 && X < W1) {
      t = __ldg(a.dg[1] + (size_t)Y * W1 + X) + stencil_adjoint<T1>(rx1, ry1, rl1, r + 1, c + 1);
      if (a.levels > 2 && (Y >> 1) < H2 && (X >> 1) < W2) t += 0.25f * t2[r >> 1][c >> 1];  // avg_pool2d backward
    }
    t1[r][c] = t;
  }
  __syncthreads();
  const float gs = a.gscale ? __ldg(a.gscale) : 1.0f;
  const size_t hw0 = (size_t)H0 * W0;
  for (int i = tid; i < T0 * T0; i += kImgThreads) {
    const int r = i / T0, c = i - r * T0;
    const int y = y0 + r, x = x0 + c;
    if (y >= H0 || x >= W0) continue;
    const size_t p = (size_t)y * W0 + x;
    float t = __ldg(a.dg[0] + p) + stencil_adjoint<T0>(rx0, ry0, rl0, r + 1, c + 1);
    if (a.levels > 1 && (y >> 1) < H1 && (x >> 1) < W1) t += 0.25f * t1[r >> 1][c >> 1];
    const float g = (t / 3.0f) * gs;
    a.out[p] = g; a.out[hw0 + p] = g; a.out[2 * hw0 + p] = g;
  }
}

// ---------------------------------------------------------------- high-frequency mask tail
constexpr int kSumBlocks = 148 * 8;

// score = clamp(0.7 spatial + 0.3 hs / max(hs), 0, 5) in place over hs; min / max of it -> mm2[0..1] (last CTA)
__global__ void __launch_bounds__(256)
hf_combine_kernel(float* __restrict__ hs, const float* __restrict__ spatial, const float* __restrict__ mm, int64_t n,
                  float* __restrict__ pmin, float* __restrict__ pmax, float* __restrict__ mm2, unsigned* counter) {
  __shared__ float smin[8], smax[8];
  __shared__ int s_last;
  float lo = __int_as_float(0x7f800000), hi = -__int_as_float(0x7f800000);
  const float mx = mm[1];
#pragma unroll 4
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float h = hs[i];
    if (mx > 1e-8f) h = h / mx;
    const float v = fminf(fmaxf(0.7f * spatial[i] + 0.3f * h, 0.f), 5.0f);
    hs[i] = v;
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { lo = fminf(lo, smin[i]); hi = fmaxf(hi, smax[i]); }
    pmin[blockIdx.x] = lo;
    pmax[blockIdx.x] = hi;
    __threadfence();
    s_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  lo = __int_as_float(0x7f800000);
  hi = -__int_as_float(0x7f800000);
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) { lo = fminf(lo, pmin[i]); hi = fmaxf(hi, pmax[i]); }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) { smin[threadIdx.x >> 5] = lo; smax[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) { lo = fminf(lo, smin[i]); hi = fmaxf(hi, smax[i]); }
    mm2[0] = lo;
    mm2[1] = hi;
  }
}

// mask = (score - min) / (max - min) > thresh; count = sum(mask) (0 / 1 values: the float sum is exact below 2^24 per
// CTA, the cross-CTA sum runs in double)
__global__ void __launch_bounds__(256)
hf_threshold_kernel(const float* __restrict__ score, const float* __restrict__ mm, int64_t n, float thresh,
                    float* __restrict__ mask, double* __restrict__ partial, float* __restrict__ count, unsigned* counter) {
  __shared__ float red[8];
  __shared__ double sm[32];
  __shared__ int s_last;
  float cnt = 0.f;
  const float lo = mm[0], range = mm[1] - mm[0];
#pragma unroll 4
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = range > 1e-6f ? (score[i] - lo) / range : 0.f;
    const float m = s > thresh ? 1.f : 0.f;
    mask[i] = m;
    cnt += m;
  }
  cnt = warp_sum(cnt);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
  __syncthreads();
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += red[i];
    partial[blockIdx.x] = (double)s;
    __threadfence();
    s_last = atomicAdd(counter, 1u) == gridDim.x - 1 ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const double s = cta_sum_strided(partial, (int)gridDim.x, 1, 0, sm);
  if (threadIdx.x == 0) count[0] = (float)s;
}

// ---------------------------------------------------------------- host side: workspace, job tables, launches
struct Carver {
  char* p;
  explicit Carver(void* base) : p((char*)(((uintptr_t)base + 255) / 256 * 256)) {}
  template <typename T>
  T* take(size_t count) {
    T* r = (T*)p;
    p += (count * sizeof(T) + 255) / 256 * 256;
    return r;
  }
};

static inline size_t pad256(size_t b) { return (b + 255) / 256 * 256; }

// columns per CTA of the column kernels: the pair job keeps 2 * tc columns twice (ping-pong) in shared memory
static int cols_per_cta(int H) {  // 1, 2 or 4 (a power of two: the kernels index with shifts)
  const int tc = (int)((72 * 1024) / (32 * (size_t)H));
  return tc >= 4 ? 4 : (tc >= 2 ? 2 : 1);
}
// row pairs per CTA of the row kernels: about 2048 points per CTA (1 / 2 / 4 pairs at 1920 / 960 / 480 columns)
static int pairs_per_cta(int W) {
  int p = 2048 / W;
  return p > 8 ? 8 : (p < 1 ? 1 : p);
}

struct Dims {
  int levels;
  int H[kMaxLevels], W[kMaxLevels];
  Plan row[kMaxLevels], col[kMaxLevels];
};

static int make_dims(int32_t H, int32_t W, int32_t levels, Dims* d) {
  if (levels < 1 || levels > kMaxLevels || H < 4 || W < 4) {
    set_error("frequency loss: bad size / level count");
    return HG_ERR_INVALID_ARG;
  }
  d->levels = levels;
  d->H[0] = H; d->W[0] = W;
  for (int l = 1; l < levels; ++l) { d->H[l] = d->H[l - 1] / 2; d->W[l] = d->W[l - 1] / 2; }
  for (int l = 0; l < levels; ++l) {
    if (d->H[l] <= 0 || d->W[l] <= 1 || d->H[l] > 4096 || d->W[l] > 4096 || !make_plan(d->W[l], &d->row[l]) ||
        !make_plan(d->H[l], &d->col[l])) {
      set_error("FFT size %dx%d unsupported (each side must be in [1, 4096])", d->H[l], d->W[l]);
      return HG_ERR_INVALID_ARG;
    }
  }
  return HG_OK;
}

static int set_fused_attrs() {
  static thread_local bool done[64] = {};
  int dev = 0;
  HG_CUDA_TRY(cudaGetDevice(&dev));
  if (dev >= 0 && dev < 64 && done[dev]) return HG_OK;
  const int maxb = 200 * 1024;
  HG_CUDA_TRY(cudaFuncSetAttribute(fft_rows_jobs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
  HG_CUDA_TRY(cudaFuncSetAttribute(fft_cols_jobs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
  HG_CUDA_TRY(cudaFuncSetAttribute(spectral_grad_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
  HG_CUDA_TRY(cudaFuncSetAttribute(fft_rows_c2r_jobs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxb));
  if (dev >= 0 && dev < 64) done[dev] = true;
  return HG_OK;
}

// Ground-truth side of the regulariser for one image: gray pyramid, its spectra, the 8 level-0 band sums.
struct GtState {
  float* gg[kMaxLevels];
  float2* fg[kMaxLevels];
  double* band0;   // 8 final sums
};

static void carve_gt_state(void* base, const Dims& d, GtState* g) {
  Carver cv(base);
  for (int l = 0; l < d.levels; ++l) {
    g->gg[l] = cv.take<float>((size_t)d.H[l] * d.W[l]);
    g->fg[l] = cv.take<float2>((size_t)d.H[l] * (d.W[l] / 2 + 1));
  }
  g->band0 = cv.take<double>(8);
}

static size_t gt_state_bytes(const Dims& d) {
  size_t total = 512 + pad256(8 * sizeof(double));
  for (int l = 0; l < d.levels; ++l)
    total += pad256((size_t)d.H[l] * d.W[l] * 4) + pad256((size_t)d.H[l] * (d.W[l] / 2 + 1) * 8);
  return total;
}

// The workspace of one call (forward state kept for the backward).
struct Work {
  LevelCtl* ctl;
  unsigned* counters;  // [0] column epilogue, [1] hf maximum, [2] hf min / max, [3] hf count
  float* mm;           // [0..1] hf maximum, [2..3] score min / max
  float *gr[kMaxLevels], *dg[kMaxLevels];
  float2* fr[kMaxLevels];
  double* spatial_partial;
  double* col_partial;
  int tiles, col_ctas_max;
  // high-frequency mask (uncached ground truth)
  float *hf_score, *hf_hs, *hf_pmin, *hf_pmax;
  float2* hf_spec;
  double* hf_part;
  GtState gt;          // uncached: carved behind everything else
  char* end;
};

static int col_ctas(const Dims& d, int l) { return (d.W[l] / 2 + 1 + cols_per_cta(d.H[l]) - 1) / cols_per_cta(d.H[l]); }

static void carve_work(void* ws, const Dims& d, bool with_gt, bool with_hf, Work* w) {
  Carver cv(ws);
  w->ctl = cv.take<LevelCtl>(kMaxLevels);
  w->counters = cv.take<unsigned>(8);
  w->mm = cv.take<float>(8);
  w->tiles = ((d.W[0] + kT0 - 1) / kT0) * ((d.H[0] + kT0 - 1) / kT0);
  int ctas = 0;
  for (int l = 0; l < d.levels; ++l) {
    const size_t hw = (size_t)d.H[l] * d.W[l];
    w->gr[l] = cv.take<float>(hw);
    w->dg[l] = cv.take<float>(hw);
    w->fr[l] = cv.take<float2>((size_t)d.H[l] * (d.W[l] / 2 + 1));
    ctas += col_ctas(d, l);
  }
  const int hf_ctas = (d.W[0] / 2 + 1 + 3) / 4 + 8;
  w->col_ctas_max = ctas + hf_ctas;
  w->spatial_partial = cv.take<double>((size_t)w->tiles * 9);
  w->col_partial = cv.take<double>((size_t)w->col_ctas_max * kColVals);
  w->hf_score = w->hf_hs = w->hf_pmin = w->hf_pmax = nullptr;
  w->hf_spec = nullptr;
  w->hf_part = nullptr;
  if (with_hf) {
    const size_t hw = (size_t)d.H[0] * d.W[0];
    w->hf_score = cv.take<float>(hw);
    w->hf_hs = cv.take<float>(hw);
    w->hf_spec = cv.take<float2>((size_t)d.H[0] * (d.W[0] / 2 + 1));
    w->hf_pmin = cv.take<float>(2048 > kSumBlocks ? 2048 : kSumBlocks);
    w->hf_pmax = cv.take<float>(2048 > kSumBlocks ? 2048 : kSumBlocks);
    w->hf_part = cv.take<double>(kSumBlocks);
  }
  if (with_gt) {
    carve_gt_state(cv.p, d, &w->gt);
    cv.p += gt_state_bytes(d);
  }
  w->end = cv.p;
}

static size_t work_bytes(const Dims& d, bool with_gt, bool with_hf) {
  Work w;
  carve_work((void*)(uintptr_t)4096, d, with_gt, with_hf, &w);
  return (size_t)(w.end - (char*)(uintptr_t)4096) + 512;
}

static RowJob row_job(const float* src, float2* dst, const Dims& d, int l, int clamp01, int* cta) {
  RowJob j;
  j.src = src; j.dst = dst; j.H = d.H[l]; j.W = d.W[l]; j.clamp01 = clamp01;
  j.pairs = pairs_per_cta(d.W[l]);
  j.cta_begin = *cta;
  j.plan = d.row[l];
  *cta += ((d.H[l] + 1) / 2 + j.pairs - 1) / j.pairs;
  return j;
}

static size_t row_smem(const Dims& d, int l) { return 2 * (size_t)pairs_per_cta(d.W[l]) * d.W[l] * sizeof(float2); }

static int launch_rows(const RowArgs& ra, int ctas, size_t smem, cudaStream_t st) {
  fft_rows_jobs_kernel<<<ctas, kFftThreads, smem, st>>>(ra);
  HG_POST_LAUNCH(false, st, "fft_rows_jobs");
  return HG_OK;
}

// the tail of the high-frequency mask: inverse rows (|.|, max) -> score, min / max -> mask, count
static int launch_hf_tail(const Dims& d, const Work& w, float thresh, float* mask, float* count, cudaStream_t st) {
  InvArgs ia{};
  ia.n_jobs = 1;
  InvJob& j = ia.job[0];
  j.spec = w.hf_spec; j.dst = w.hf_hs; j.gate = nullptr; j.H = d.H[0]; j.W = d.W[0];
  j.pairs = pairs_per_cta(d.W[0]); j.cta_begin = 0; j.mode = 1;
  j.scale = 1.0f / ((float)d.H[0] * (float)d.W[0]);
  j.plan = d.row[0];
  ia.pmax = w.hf_pmax; ia.mm = w.mm; ia.counter = w.counters + 1;
  const int ctas = ((d.H[0] + 1) / 2 + j.pairs - 1) / j.pairs;
  if (ctas > 2048) { set_error("hf mask: image too tall"); return HG_ERR_INVALID_ARG; }
  fft_rows_c2r_jobs_kernel<<<ctas, kFftThreads, row_smem(d, 0), st>>>(ia);
  HG_POST_LAUNCH(false, st, "ifft_rows_hf");
  const int64_t hw = (int64_t)d.H[0] * d.W[0];
  hf_combine_kernel<<<kSumBlocks, 256, 0, st>>>(w.hf_hs, w.hf_score, w.mm, hw, w.hf_pmin, w.hf_pmax, w.mm + 2, w.counters + 2);
  HG_POST_LAUNCH(false, st, "hf_combine");
  hf_threshold_kernel<<<kSumBlocks, 256, 0, st>>>(w.hf_hs, w.mm + 2, hw, thresh, mask, w.hf_part, count, w.counters + 3);
  HG_POST_LAUNCH(false, st, "hf_threshold");
  return HG_OK;
}

static void fill_pyr_dims(PyrArgs* pa, const Dims& d) {
  pa->levels = d.levels;
  for (int l = 0; l < kMaxLevels; ++l) { pa->H[l] = l < d.levels ? d.H[l] : 0; pa->W[l] = l < d.levels ? d.W[l] : 0; }
}

// Forward of the regulariser (+ optionally the high-frequency mask of the same ground truth).
static int freq_forward(const float* rendered, const float* gt, void* gt_state, const Dims& d, float hf_thresh,
                        float* hf_mask, float* hf_count, float* stats, void* ws, cudaStream_t st) {
  const bool cached = gt_state != nullptr;
  const bool hf = hf_mask != nullptr;
  int rc = set_fused_attrs();
  if (rc) return rc;
  Work w;
  carve_work(ws, d, !cached, hf, &w);
  GtState G;
  if (cached) carve_gt_state(gt_state, d, &G); else G = w.gt;
  // ---- image space
  PyrArgs pa{};
  pa.rgb[0] = rendered; pa.rgb[1] = cached ? nullptr : gt;
  for (int l = 0; l < d.levels; ++l) { pa.gray[0][l] = w.gr[l]; pa.gray[1][l] = G.gg[l]; }
  pa.n_images = 2;
  fill_pyr_dims(&pa, d);
  pa.partial = w.spatial_partial;
  pa.hf_score = hf ? w.hf_score : nullptr;
  pa.hf_src = 1;
  pa.counters = w.counters;
  const dim3 tiles((d.W[0] + kT0 - 1) / kT0, (d.H[0] + kT0 - 1) / kT0);
  pyramid_kernel<2><<<tiles, kImgThreads, 0, st>>>(pa);
  HG_POST_LAUNCH(false, st, "pyramid");
  // ---- rows
  RowArgs ra{};
  int cta = 0;
  size_t smem = 0;
  for (int l = 0; l < d.levels; ++l) {
    ra.job[ra.n_jobs++] = row_job(w.gr[l], w.fr[l], d, l, 1, &cta);
    if (!cached) ra.job[ra.n_jobs++] = row_job(G.gg[l], G.fg[l], d, l, 1, &cta);
    if (row_smem(d, l) > smem) smem = row_smem(d, l);
  }
  if (hf) ra.job[ra.n_jobs++] = row_job(G.gg[0], w.hf_spec, d, 0, 0, &cta);  // unclamped here (:1221)
  rc = launch_rows(ra, cta, smem, st);
  if (rc) return rc;
  // ---- columns + sums + epilogue (heaviest CTAs first: the hf job runs two transforms per column group; the
  // scheduler hands out CTAs in index order, so the light level-2 CTAs fill the tail of the launch)
  ColArgs ca{};
  cta = 0;
  smem = 0;
  if (hf) {
    ColJob& j = ca.job[ca.n_jobs++];
    j.A = w.hf_spec; j.B = nullptr; j.H = d.H[0]; j.W = d.W[0]; j.kind = kColHighpassInverse;
    j.tc = 2 * cols_per_cta(d.H[0]) > 4 ? 4 : 2 * cols_per_cta(d.H[0]);
    j.cta_begin = cta; j.level = -1; j.plan = d.col[0];
    cta += (d.W[0] / 2 + 1 + j.tc - 1) / j.tc;
    const size_t b = 2 * (size_t)j.tc * d.H[0] * sizeof(float2);
    if (b > smem) smem = b;
  }
  for (int l = 0; l < d.levels; ++l) {
    ColJob& j = ca.job[ca.n_jobs++];
    j.A = w.fr[l]; j.B = G.fg[l]; j.H = d.H[l]; j.W = d.W[l];
    j.kind = cached ? kColPairCached : kColPair;
    j.tc = cols_per_cta(d.H[l]); j.cta_begin = cta; j.level = l; j.plan = d.col[l];
    ca.fin.col_begin[l] = cta;
    cta += col_ctas(d, l);
    ca.fin.col_end[l] = cta;
    const size_t b = 2 * (size_t)(cached ? 1 : 2) * j.tc * d.H[l] * sizeof(float2);
    if (b > smem) smem = b;
  }
  ca.partial = w.col_partial;
  FinalizeArgs& f = ca.fin;
  f.mode = 1; f.levels = d.levels;
  for (int l = 0; l < d.levels; ++l) { f.H[l] = d.H[l]; f.W[l] = d.W[l]; }
  f.spatial_partial = w.spatial_partial; f.spatial_tiles = w.tiles;
  f.band0_in = cached ? G.band0 : nullptr; f.band0_out = nullptr;
  f.ctl = w.ctl; f.stats = stats; f.counter = w.counters;
  fft_cols_jobs_kernel<<<cta, kFftThreads, smem, st>>>(ca);
  HG_POST_LAUNCH(false, st, "fft_cols_jobs");
  if (hf) return launch_hf_tail(d, w, hf_thresh, hf_mask, hf_count, st);
  return HG_OK;
}

// Backward of the regulariser on the state the forward left in the workspace: grad = gscale * d freq_loss / d rendered.
static int freq_backward(void* gt_state, const Dims& d, bool forward_had_hf, const float* gscale, float* grad, void* ws,
                         cudaStream_t st) {
  const bool cached = gt_state != nullptr;
  Work w;
  carve_work(ws, d, !cached, forward_had_hf, &w);
  GtState G;
  if (cached) carve_gt_state(gt_state, d, &G); else G = w.gt;
  GradArgs ga{};
  int cta = 0;
  size_t smem = 0;
  for (int l = 0; l < d.levels; ++l) {
    GradJob& j = ga.job[ga.n_jobs++];
    j.A = w.fr[l]; j.B = G.fg[l]; j.H = d.H[l]; j.W = d.W[l];
    j.tc = 2 * cols_per_cta(d.H[l]) > 4 ? 4 : 2 * cols_per_cta(d.H[l]);
    j.cta_begin = cta; j.level = l; j.plan = d.col[l];
    cta += (d.W[l] / 2 + 1 + j.tc - 1) / j.tc;
    const size_t b = 2 * (size_t)j.tc * d.H[l] * sizeof(float2);
    if (b > smem) smem = b;
  }
  ga.ctl = w.ctl;
  spectral_grad_cols_kernel<<<cta, kFftThreads, smem, st>>>(ga);
  HG_POST_LAUNCH(false, st, "spectral_grad_cols");
  InvArgs ia{};
  cta = 0;
  smem = 0;
  for (int l = 0; l < d.levels; ++l) {
    InvJob& j = ia.job[ia.n_jobs++];
    j.spec = w.fr[l]; j.dst = w.dg[l]; j.gate = w.gr[l]; j.H = d.H[l]; j.W = d.W[l];
    j.pairs = pairs_per_cta(d.W[l]); j.cta_begin = cta; j.mode = 0; j.scale = 1.0f; j.plan = d.row[l];
    cta += ((d.H[l] + 1) / 2 + j.pairs - 1) / j.pairs;
    if (row_smem(d, l) > smem) smem = row_smem(d, l);
  }
  fft_rows_c2r_jobs_kernel<<<cta, kFftThreads, smem, st>>>(ia);
  HG_POST_LAUNCH(false, st, "ifft_rows_jobs");
  BwdArgs ba{};
  ba.levels = d.levels;
  for (int l = 0; l < d.levels; ++l) {
    ba.H[l] = d.H[l]; ba.W[l] = d.W[l];
    ba.gr[l] = w.gr[l]; ba.gg[l] = G.gg[l]; ba.dg[l] = w.dg[l];
  }
  ba.ctl = w.ctl; ba.gscale = gscale; ba.out = grad;
  const dim3 tiles((d.W[0] + kT0 - 1) / kT0, (d.H[0] + kT0 - 1) / kT0);
  spatial_grad_rgb_kernel<<<tiles, kImgThreads, 0, st>>>(ba);
  HG_POST_LAUNCH(false, st, "spatial_grad_rgb");
  return HG_OK;
}

}  // namespace
}  // namespace hg

using namespace hg;

extern "C" {

size_t hg_fft2_workspace_bytes(int32_t H, int32_t W) { return (size_t)H * (W / 2 + 1) * 8 + 512; }

int hg_fft2_r2c(const float* img, int32_t H, int32_t W, float* spec, void* ws, void* st) {
  (void)ws;
  if (!img || !spec) { set_error("hg_fft2_r2c: NULL pointer"); return HG_ERR_INVALID_ARG; }
  return fft2_r2c(img, (float2*)spec, nullptr, nullptr, H, W, 0, (cudaStream_t)st);
}

int hg_fft2_c2r(const float* spec, int32_t H, int32_t W, float* img, int scale_inv, void* ws, void* st_) {
  if (!img || !spec || !ws) { set_error("hg_fft2_c2r: NULL pointer"); return HG_ERR_INVALID_ARG; }
  cudaStream_t st = (cudaStream_t)st_;
  Carver cv(ws);
  const size_t n = (size_t)H * (W / 2 + 1);
  float2* tmp = cv.take<float2>(n);
  HG_CUDA_TRY(cudaMemcpyAsync(tmp, spec, n * sizeof(float2), cudaMemcpyDeviceToDevice, st));
  return fft2_c2r(tmp, H, W, scale_inv ? 1.0f / ((float)H * (float)W) : 1.0f, img, st);
}

size_t hg_freq_loss_workspace_bytes(int32_t H, int32_t W, int32_t levels) {
  Dims d;
  d.levels = levels < 1 ? 1 : (levels > kMaxLevels ? kMaxLevels : levels);
  d.H[0] = H < 4 ? 4 : H; d.W[0] = W < 4 ? 4 : W;
  for (int l = 1; l < d.levels; ++l) { d.H[l] = d.H[l - 1] / 2; d.W[l] = d.W[l - 1] / 2; }
  return work_bytes(d, true, true);
}

size_t hg_freq_gt_state_bytes(int32_t H, int32_t W, int32_t levels) {
  Dims d;
  d.levels = levels < 1 ? 1 : (levels > kMaxLevels ? kMaxLevels : levels);
  d.H[0] = H < 4 ? 4 : H; d.W[0] = W < 4 ? 4 : W;
  for (int l = 1; l < d.levels; ++l) { d.H[l] = d.H[l - 1] / 2; d.W[l] = d.W[l - 1] / 2; }
  // (+ the scratch of the preparation itself: per-CTA partial sums and a ticket)
  int ctas = 8;
  for (int l = 0; l < d.levels; ++l) ctas += col_ctas(d, l);
  return gt_state_bytes(d) + pad256((size_t)ctas * kColVals * 8) + 1024;
}

int hg_freq_gt_prepare(const float* gt, int32_t H, int32_t W, int32_t levels, void* gt_state, void* st_) {
  if (!gt || !gt_state) {
    set_error("hg_freq_gt_prepare: NULL pointer");
    return HG_ERR_INVALID_ARG;
  }
  Dims d;
  int rc = make_dims(H, W, levels, &d);
  if (rc) return rc;
  rc = set_fused_attrs();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)st_;
  GtState G;
  carve_gt_state(gt_state, d, &G);
  Carver scratch((char*)gt_state + gt_state_bytes(d));
  unsigned* counters = scratch.take<unsigned>(8);
  int n_ctas = 8;
  for (int l = 0; l < d.levels; ++l) n_ctas += col_ctas(d, l);
  double* partial = scratch.take<double>((size_t)n_ctas * kColVals);
  PyrArgs pa{};
  pa.rgb[0] = gt; pa.rgb[1] = nullptr;
  for (int l = 0; l < d.levels; ++l) pa.gray[0][l] = G.gg[l];
  pa.n_images = 1;
  fill_pyr_dims(&pa, d);
  pa.counters = counters;
  const dim3 tiles((d.W[0] + kT0 - 1) / kT0, (d.H[0] + kT0 - 1) / kT0);
  pyramid_kernel<1><<<tiles, kImgThreads, 0, st>>>(pa);
  HG_POST_LAUNCH(false, st, "pyramid");
  RowArgs ra{};
  int cta = 0;
  size_t smem = 0;
  for (int l = 0; l < d.levels; ++l) {
    ra.job[ra.n_jobs++] = row_job(G.gg[l], G.fg[l], d, l, 1, &cta);
    if (row_smem(d, l) > smem) smem = row_smem(d, l);
  }
  rc = launch_rows(ra, cta, smem, st);
  if (rc) return rc;
  ColArgs ca{};
  cta = 0;
  smem = 0;
  for (int l = 0; l < d.levels; ++l) {
    ColJob& j = ca.job[ca.n_jobs++];
    j.A = G.fg[l]; j.B = nullptr; j.H = d.H[l]; j.W = d.W[l]; j.kind = kColSingle;
    j.tc = cols_per_cta(d.H[l]); j.cta_begin = cta; j.level = l; j.plan = d.col[l];
    if (l == 0) { ca.fin.col_begin[0] = cta; ca.fin.col_end[0] = cta + col_ctas(d, 0); }
    cta += col_ctas(d, l);
    const size_t b = 2 * (size_t)j.tc * d.H[l] * sizeof(float2);
    if (b > smem) smem = b;
  }
  ca.partial = partial;
  ca.fin.mode = 2; ca.fin.levels = 1;  // the last CTA reduces the level-0 band sums into the state
  ca.fin.band0_in = nullptr; ca.fin.band0_out = G.band0; ca.fin.counter = counters;
  fft_cols_jobs_kernel<<<cta, kFftThreads, smem, st>>>(ca);
  HG_POST_LAUNCH(false, st, "fft_cols_jobs");
  return HG_OK;
}

int hg_freq_forward(const float* rendered, const float* gt, void* gt_state, int32_t H, int32_t W, int32_t levels,
                    float hf_thresh, float* hf_mask, float* hf_count, float* stats, void* ws, void* st_) {
  if (!rendered || (!gt && !gt_state) || !stats || !ws || (hf_mask && !hf_count)) {
    set_error("hg_freq_forward: NULL pointer");
    return HG_ERR_INVALID_ARG;
  }
  Dims d;
  int rc = make_dims(H, W, levels, &d);
  if (rc) return rc;
  return freq_forward(rendered, gt_state ? nullptr : gt, gt_state, d, hf_thresh, hf_mask, hf_count, stats, ws,
                      (cudaStream_t)st_);
}

int hg_freq_backward(void* gt_state, int32_t H, int32_t W, int32_t levels, int32_t forward_had_hf_mask,
                     const float* gscale, float* grad_rendered, void* ws, void* st_) {
  if (!grad_rendered || !ws) {
    set_error("hg_freq_backward: NULL pointer");
    return HG_ERR_INVALID_ARG;
  }
  Dims d;
  int rc = make_dims(H, W, levels, &d);
  if (rc) return rc;
  return freq_backward(gt_state, d, forward_had_hf_mask != 0, gscale, grad_rendered, ws, (cudaStream_t)st_);
}

int hg_freq_loss(const float* rendered, const float* gt, int32_t H, int32_t W, int32_t levels, float* stats,
                 float* grad_rendered, void* ws, void* st_) {
  if (!gt) {
    set_error("hg_freq_loss: NULL pointer");
    return HG_ERR_INVALID_ARG;
  }
  int rc = hg_freq_forward(rendered, gt, nullptr, H, W, levels, 0.f, nullptr, nullptr, stats, ws, st_);
  if (rc || !grad_rendered) return rc;
  return hg_freq_backward(nullptr, H, W, levels, 0, nullptr, grad_rendered, ws, st_);
}

int hg_freq_loss_cached(const float* rendered, void* gt_state, int32_t H, int32_t W, int32_t levels, float* stats,
                        float* grad_rendered, void* ws, void* st_) {
  if (!gt_state) {
    set_error("hg_freq_loss_cached: NULL ground-truth state");
    return HG_ERR_INVALID_ARG;
  }
  int rc = hg_freq_forward(rendered, nullptr, gt_state, H, W, levels, 0.f, nullptr, nullptr, stats, ws, st_);
  if (rc || !grad_rendered) return rc;
  return hg_freq_backward(gt_state, H, W, levels, 0, nullptr, grad_rendered, ws, st_);
}

size_t hg_hf_mask_workspace_bytes(int32_t H, int32_t W) {
  Dims d;
  d.levels = 1;
  d.H[0] = H < 4 ? 4 : H; d.W[0] = W < 4 ? 4 : W;
  return work_bytes(d, true, true);
}

int hg_hf_mask(const float* gt, int32_t H, int32_t W, float thresh, float* mask, float* count, void* ws, void* st_) {
  if (!gt || !mask || !count || !ws || H < 4 || W < 4) {
    set_error("hg_hf_mask: bad argument");
    return HG_ERR_INVALID_ARG;
  }
  Dims d;
  int rc = make_dims(H, W, 1, &d);
  if (rc) return rc;
  rc = set_fused_attrs();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)st_;
  Work w;
  carve_work(ws, d, true, true, &w);
  PyrArgs pa{};
  pa.rgb[0] = gt; pa.rgb[1] = nullptr;
  pa.gray[0][0] = w.gt.gg[0];
  pa.n_images = 1;
  fill_pyr_dims(&pa, d);
  pa.hf_score = w.hf_score;
  pa.hf_src = 0;
  pa.counters = w.counters;
  const dim3 tiles((W + kT0 - 1) / kT0, (H + kT0 - 1) / kT0);
  pyramid_kernel<1><<<tiles, kImgThreads, 0, st>>>(pa);
  HG_POST_LAUNCH(false, st, "pyramid");
  RowArgs ra{};
  int cta = 0;
  ra.job[ra.n_jobs++] = row_job(w.gt.gg[0], w.hf_spec, d, 0, 0, &cta);  // unclamped (:1221)
  rc = launch_rows(ra, cta, row_smem(d, 0), st);
  if (rc) return rc;
  ColArgs ca{};
  ColJob& j = ca.job[ca.n_jobs++];
  j.A = w.hf_spec; j.B = nullptr; j.H = H; j.W = W; j.kind = kColHighpassInverse;
  j.tc = 2 * cols_per_cta(H) > 4 ? 4 : 2 * cols_per_cta(H);
  j.cta_begin = 0; j.level = -1; j.plan = d.col[0];
  ca.partial = w.col_partial;
  ca.fin.mode = 0;
  fft_cols_jobs_kernel<<<(W / 2 + 1 + j.tc - 1) / j.tc, kFftThreads, 2 * (size_t)j.tc * H * sizeof(float2), st>>>(ca);
  HG_POST_LAUNCH(false, st, "fft_cols_jobs");
  return launch_hf_tail(d, w, thresh, mask, count, st);
}

}  // extern "C"
