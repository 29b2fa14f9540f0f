// Stand-in for the reference's VectorMath.hpp, whose SIMD wrapper classes use MSVC-only syntax (deducing-this, `operator##OP##=`).
// The parts the renderer uses — VectorMath.hpp:7-19 (integer type maps) and :581-662 (scalar fast math) — are included VERBATIM
// from line ranges cut out at build time; SimdVec / mul_add are only named by Color.hpp's unused luminance() overload. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <functional>
#include "Core.hpp"
#include "vm_ints.inc"     // VectorMath.hpp:7-19: UnsignedIntType / SignedIntType (DataStreams.hpp's bitset)
#include "vm_scalar.inc"
template <class V, size_t N> struct SimdVec { V v[N]; V& operator[](size_t i) { return v[i]; } const V& operator[](size_t i) const { return v[i]; } };
static inline Vec8f mul_add(Vec8f a, Vec8f b, Vec8f c) { return _mm256_fmadd_ps(a, b, c); }
static inline Vec8f mul_add(Vec8f a, float b, Vec8f c) { return _mm256_fmadd_ps(a, _mm256_set1_ps(b), c); }
