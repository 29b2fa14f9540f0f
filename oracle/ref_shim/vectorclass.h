// Minimal stand-in for Agner Fog's VCL Vec8f (absent from this image, un-vendored by the reference): the lane-wise AVX operations
// that /root/reference/Color.hpp:47-73 uses, defined as VCL's vectorf256.h publishes them (_mm256_{add,sub,mul,div,min,max}_ps).
// TEST INFRASTRUCTURE ONLY (oracle/Makefile `ref`).
#pragma once
#include <immintrin.h>
struct Vec8f {
	__m256 v;
	Vec8f() : v(_mm256_setzero_ps()) {}
	Vec8f(float s) : v(_mm256_set1_ps(s)) {}
	Vec8f(__m256 x) : v(x) {}
	operator __m256() const { return v; }
	Vec8f& load(const float* p) { v = _mm256_loadu_ps(p); return *this; }
	void store(float* p) const { _mm256_storeu_ps(p, v); }
};
static inline Vec8f operator+(Vec8f a, Vec8f b) { return _mm256_add_ps(a, b); }
static inline Vec8f operator-(Vec8f a, Vec8f b) { return _mm256_sub_ps(a, b); }
static inline Vec8f operator*(Vec8f a, Vec8f b) { return _mm256_mul_ps(a, b); }
static inline Vec8f operator/(Vec8f a, Vec8f b) { return _mm256_div_ps(a, b); }
static inline Vec8f operator+(Vec8f a, float b) { return a + Vec8f(b); }
static inline Vec8f operator-(Vec8f a, float b) { return a - Vec8f(b); }
static inline Vec8f operator*(Vec8f a, float b) { return a * Vec8f(b); }
static inline Vec8f operator*(float a, Vec8f b) { return Vec8f(a) * b; }
static inline Vec8f min(Vec8f a, Vec8f b) { return _mm256_min_ps(a, b); }
static inline Vec8f max(Vec8f a, Vec8f b) { return _mm256_max_ps(a, b); }
