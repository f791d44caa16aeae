// hier_half.cpp — host-side float <-> half array conversion for the .hier file format (hier_io.cu).
// IEEE binary16, round-to-nearest-even (what half.hpp 2.2 does with HALF_ROUND_STYLE 1, and numpy).  The bulk runs on
// the F16C unit eight values at a time when the CPU has one (checked at run time); the scalar code is the fallback and
// the tail.  Compiled by the host compiler only (no CUDA in this file).
#include <cstddef>
#include <cstdint>
#include <cstring>

#if defined(__x86_64__) || defined(__i386__)
#include <immintrin.h>
#define HG_HAVE_X86 1
#else
#define HG_HAVE_X86 0
#endif

namespace hg {

static inline float half_to_float_scalar(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else {  // subnormal: normalise
      int sh = 0;
      while (!(man & 0x400u)) {
        man <<= 1;
        ++sh;
      }
      man &= 0x3ffu;
      bits = sign | ((uint32_t)(127 - 15 - sh + 1) << 23) | (man << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | (man << 13);
  } else {
    bits = sign | ((exp + 112u) << 23) | (man << 13);
  }
  float f;
  memcpy(&f, &bits, 4);
  return f;
}

static inline uint16_t float_to_half_scalar(float f) {
  uint32_t x;
  memcpy(&x, &f, 4);
  const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
  x &= 0x7fffffffu;
  if (x >= 0x7f800000u) return sign | (x > 0x7f800000u ? 0x7e00u : 0x7c00u);  // NaN (quiet) / inf
  if (x >= 0x477ff000u) return sign | 0x7c00u;                                // rounds to >= 65520 -> inf
  if (x < 0x33000001u) return sign;                                           // <= 2^-25 rounds to zero (tie to even)
  const uint32_t exp = x >> 23, man = (x & 0x7fffffu) | 0x800000u;
  int shift;
  uint32_t half_exp;
  if (exp < 113) {  // subnormal half
    shift = 13 + (113 - (int)exp);
    half_exp = 0;
  } else {
    shift = 13;
    half_exp = exp - 112;
  }
  const uint32_t keep = man >> shift, rem = man & ((1u << shift) - 1), halfway = 1u << (shift - 1);
  uint32_t r = (half_exp ? ((half_exp << 10) | (keep & 0x3ffu)) : keep);
  if (rem > halfway || (rem == halfway && (keep & 1u))) ++r;  // carries propagate into the exponent correctly
  return sign | (uint16_t)r;
}

#if HG_HAVE_X86
__attribute__((target("avx,f16c"))) static size_t widen_f16c(const uint16_t* src, float* dst, size_t n) {
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    _mm256_storeu_ps(dst + i, _mm256_cvtph_ps(_mm_loadu_si128(reinterpret_cast<const __m128i*>(src + i))));
  return i;
}
__attribute__((target("avx,f16c"))) static size_t narrow_f16c(const float* src, uint16_t* dst, size_t n) {
  size_t i = 0;
  for (; i + 8 <= n; i += 8)
    _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + i),
                     _mm256_cvtps_ph(_mm256_loadu_ps(src + i), _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
  return i;
}
static bool have_f16c() {
  static const bool ok = __builtin_cpu_supports("avx") && __builtin_cpu_supports("f16c");
  return ok;
}
#endif

void half_to_float_array(const uint16_t* src, float* dst, size_t n) {
  size_t i = 0;
#if HG_HAVE_X86
  if (have_f16c()) i = widen_f16c(src, dst, n);
#endif
  for (; i < n; ++i) dst[i] = half_to_float_scalar(src[i]);
}

void float_to_half_array(const float* src, uint16_t* dst, size_t n) {
  size_t i = 0;
#if HG_HAVE_X86
  if (have_f16c()) i = narrow_f16c(src, dst, n);
#endif
  for (; i < n; ++i) dst[i] = float_to_half_scalar(src[i]);
}

// (exposed for the tests: the scalar path must agree with the F16C path bit for bit)
void float_to_half_array_scalar(const float* src, uint16_t* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = float_to_half_scalar(src[i]);
}
void half_to_float_array_scalar(const uint16_t* src, float* dst, size_t n) {
  for (size_t i = 0; i < n; ++i) dst[i] = half_to_float_scalar(src[i]);
}

}  // namespace hg

extern "C" {
// test hook: 0 = dispatching implementation, 1 = scalar fallback
__attribute__((visibility("default"))) void hg_test_float_to_half(const float* src, uint16_t* dst, int64_t n, int scalar) {
  if (scalar) hg::float_to_half_array_scalar(src, dst, (size_t)n);
  else hg::float_to_half_array(src, dst, (size_t)n);
}
__attribute__((visibility("default"))) void hg_test_half_to_float(const uint16_t* src, float* dst, int64_t n, int scalar) {
  if (scalar) hg::half_to_float_array_scalar(src, dst, (size_t)n);
  else hg::half_to_float_array(src, dst, (size_t)n);
}
}
