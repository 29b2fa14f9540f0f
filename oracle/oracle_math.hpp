// oracle_math.hpp — TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// CPU restatement of the scalar math the reference's hot path calls. Every function cites the
// reference file:line it follows (paths are relative to /root/reference). Nothing in the product
// (cpu-raytracing-experiments_b200/, include/) may include or link this file; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use the oracle.
//
// Arithmetic convention (DESIGN.md "Numerics"): IEEE binary32, no contraction (-ffp-contract=off),
// explicit fmaf only where the reference uses _mm256_f(n)madd_ps (BVH.hpp:252-260).
// glm functions that live in the un-vendored third-party glm (no pinned version; any 0.9.9+/1.0)
// are restated from their published scalar definitions and marked [glm].
#pragma once
#include <cstdint>
#include <cstring>
#include <cmath>
#include <cfloat>

namespace orc {

struct V3 { float x, y, z; };
struct Quat { float w, x, y, z; };  // glm::quat ctor order is (w,x,y,z): Sampling.hpp:157, Renderer.hpp:275

static inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
static inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }

// std::max/std::min and glm::max/glm::min share these exact select semantics (NaN and -0 behaviour).
static inline float smax(float a, float b) { return (a < b) ? b : a; }
static inline float smin(float a, float b) { return (b < a) ? b : a; }

static inline uint32_t f2u(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }

// [glm] dot(vec3): tmp = a*b; tmp.x + tmp.y + tmp.z
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
// [glm] cross
static inline V3 cross(V3 x, V3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
// [glm] normalize(v) = v * inversesqrt(dot(v,v)); inversesqrt(x) = 1/sqrt(x)
static inline V3 normalize(V3 v) { float inv = 1.0f / std::sqrt(dot(v, v)); return v * inv; }
// [glm] operator*(quat, vec3) (== gtx rotate(quat, vec3)): v + ((uv*w) + uuv) * 2
static inline V3 qrotate(Quat q, V3 v) {
	V3 qv{q.x, q.y, q.z};
	V3 uv = cross(qv, v);
	V3 uuv = cross(qv, uv);
	return v + ((uv * q.w) + uuv) * 2.0f;
}
// [glm] quatLookAt == quatLookAtRH(direction, up) -> quat_cast(mat3{right, up', -direction})
static inline Quat quat_look_at(V3 direction, V3 up) {
	V3 c2 = -direction;
	V3 right = cross(up, c2);
	V3 c0 = right * (1.0f / std::sqrt(smax(0.00001f, dot(right, right))));
	V3 c1 = cross(c2, c0);
	// quat_cast(mat3 m) with m[col][row]
	const float m00 = c0.x, m01 = c0.y, m02 = c0.z;
	const float m10 = c1.x, m11 = c1.y, m12 = c1.z;
	const float m20 = c2.x, m21 = c2.y, m22 = c2.z;
	float fx = m00 - m11 - m22, fy = m11 - m00 - m22, fz = m22 - m00 - m11, fw = m00 + m11 + m22;
	int biggest = 0; float fb = fw;
	if (fx > fb) { fb = fx; biggest = 1; }
	if (fy > fb) { fb = fy; biggest = 2; }
	if (fz > fb) { fb = fz; biggest = 3; }
	float bv = std::sqrt(fb + 1.0f) * 0.5f;
	float mult = 0.25f / bv;
	switch (biggest) {
	case 0: return {bv, (m12 - m21) * mult, (m20 - m02) * mult, (m01 - m10) * mult};
	case 1: return {(m12 - m21) * mult, bv, (m01 + m10) * mult, (m20 + m02) * mult};
	case 2: return {(m20 - m02) * mult, (m01 + m10) * mult, bv, (m12 + m21) * mult};
	default: return {(m01 - m10) * mult, (m20 + m02) * mult, (m12 + m21) * mult, bv};
	}
}

// ---------------------------------------------------------------- Random.hpp (whole file)
static inline float make_unit_float(uint32_t x) { return static_cast<float>(x) * 0x1p-32f; }  // Random.hpp:5 (can be 1.0f)
static inline uint32_t pcg_state_transition(uint32_t v) { return v * 747796405u + 2891336453u; }  // :10-13
static inline uint32_t pcg_output(uint32_t v) {  // :14-18
	v = ((v >> ((v >> 28u) + 4u)) ^ v) * 277803737u;
	return (v >> 22u) ^ v;
}
static inline uint32_t pcg_generate(uint32_t* s) { uint32_t p = *s; *s = pcg_state_transition(p); return pcg_output(p); }  // :20-24
static inline float rand_unit_float(uint32_t* s) { return make_unit_float(pcg_generate(s)); }  // :26-29
static inline uint32_t rand_bounded_int(uint32_t* s, uint32_t range) {  // :31-34
	uint32_t v = static_cast<uint32_t>(rand_unit_float(s) * static_cast<float>(range));
	return (v < range - 1) ? v : range - 1;
}
static inline uint32_t hash_u32(uint32_t i) {  // :36-43
	i ^= i >> 16; i *= 0x21f0aaadu; i ^= i >> 15; i *= 0xd35a2d97u; i ^= i >> 15;
	return i ^ 0xe6fe3bebu;
}
static inline uint32_t hash_2d(uint32_t x, uint32_t y) {  // :45-50
	const uint32_t qx = 0x41c64e6du * ((x >> 1u) ^ y);
	const uint32_t qy = 0x41c64e6du * ((y >> 1u) ^ x);
	return 0x41c64e6du * (qx ^ (qy >> 3u));
}
// Bitmanip.hpp:200-233 (value is computed by Renderer.hpp:80 but never used, Q5)
static inline uint32_t bitreverse32(uint32_t v) {
	v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
	v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
	v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
	return __builtin_bswap32(v);
}

// ---------------------------------------------------------------- VectorMath.hpp:581-662
static inline float f_xor(float a, float b) { return u2f(f2u(a) ^ f2u(b)); }  // :611-613
static inline float f_and(float a, float b) { return u2f(f2u(a) & f2u(b)); }  // :617-619
static inline float fast_abs(float v) { return u2f(f2u(v) & 0x7fffffffu); }  // :581-584
static inline float fast_copysign(float v, float s) { return u2f((f2u(v) & 0x7fffffffu) | (f2u(s) & 0x80000000u)); }  // :589-592

static constexpr float kPi = 3.14159265358979323846264338327950288f;
static constexpr float kTwoPi = 6.28318530717958647692528676655900576f;
static constexpr float kHalfPi = 1.57079632679489661923132169163975144f;
static constexpr float kOneOverPi = 0.318309886183790671537767526745028724f;
static constexpr float kOneOverTwoPi = 0.159154943091895335768883763372514362f;

// VectorMath.hpp:644-662. _mm_round_ss(NEAREST) == nearbyintf under the default rounding mode.
static inline void fast_sincos(float x, float* sine, float* cosine) {
	const float qf = std::nearbyintf(x * kOneOverPi);
	const uint32_t sign_mask = static_cast<uint32_t>(static_cast<int32_t>(qf)) << 31;
	x += qf * (-0.78515625f * 4);
	x += qf * (-0.00024187564849853515625f * 4);
	x += qf * (-3.7747668102383613586e-08f * 4);
	x += qf * (-1.2816720341285448015e-12f * 4);
	x = kHalfPi - (kHalfPi - x);
	float x2 = x * x;
	x = u2f(f2u(x) ^ sign_mask);
	float su = 2.6083159809786593541503e-06f; float cu = -2.71811842367242206819355e-07f;
	su = su * x2 - 0.0001981069071916863322258f; cu = (cu * x2 + 2.47990446951007470488548e-05f);
	su = su * x2 + 0.00833307858556509017944336f; cu = (cu * x2 - 0.00138888787478208541870117f);
	su = su * x2 - 0.166666597127914428710938f; cu = (cu * x2 + 0.0416666641831398010253906f);
	su = x2 * (su * x) + x; cu = (cu * x2 - 0.5f); cu = (cu * x2 + 1.0f);
	cu = u2f(f2u(cu) ^ sign_mask);
	if (fast_abs(su) > 1.0f) su = 0.0f;
	if (fast_abs(cu) > 1.0f) cu = 0.0f;
	*sine = su; *cosine = cu;
}
// VectorMath.hpp:625-630
static inline float fast_asin(float x) {
	float f = fast_abs(x);
	f = (f < 1.0f) ? 1.0f - (1.0f - f) : 1.0f;
	f = kHalfPi - std::sqrt(1.0f - f) * (1.5707963267f + f * (-0.213300989f + f * (0.077980478f + f * -0.02164095f)));
	return fast_copysign(f, x);
}
// VectorMath.hpp:632-642
static inline float fast_atan2(float y, float x) {
	const float a = fast_abs(x); const float b = fast_abs(y);
	float lo = smin(a, b), hi = smax(a, b);
	float k = hi == 0.0f ? 0.0f : lo / hi;
	k = 1.0f - (1.0f - k);
	const float k2 = k * k;
	float r = k * (0.43157974f * k2 + 1.0f) / ((0.05831938f * k2 + 0.76443945f) * k2 + 1.0f);
	if (b > a) r = kHalfPi - r;
	if (x < 0.0f) r = kPi - r;
	return fast_copysign(r, y);
}

// ---------------------------------------------------------------- Sampling.hpp
static inline float median3(float a, float b, float c) { return smax(smin(a, b), smin(smax(a, b), c)); }  // :8-12
static inline float median5(float a, float b, float c, float d, float e) {  // :13-21
	return median3(smax(smin(a, b), smin(c, d)), smin(smax(a, b), smax(c, d)), e);
}
static inline V3 spherical_to_cartesian(float phi_over_2pi, float sin_theta, float cos_theta) {  // :77-84
	float cos_phi, sin_phi; fast_sincos(phi_over_2pi * kTwoPi, &sin_phi, &cos_phi);
	return {sin_theta * cos_phi, sin_theta * sin_phi, cos_theta};
}
static inline V3 hemisphere(float t, float s) {  // :92-94
	return spherical_to_cartesian(s, std::sqrt(t), std::sqrt(smax(0.0f, 1.0f - t)));
}
static inline void orthonormal_basis(V3 n, V3* v2, V3* v3) {  // :116-131 (branch-free variant, #if true)
	float sign = f_and(-0.0f, n.z);
	float s = f_xor(1.0f, sign);
	float z = -1.0f / (s + n.z);
	float s_nx = f_xor(sign, n.x);
	float ny_z = n.y * z;
	float t = n.x * ny_z;
	*v2 = V3{1.0f + (s_nx * n.x) * z, f_xor(sign, t), -s_nx};
	*v3 = V3{t, s + ny_z * n.y, -n.y};
}
static inline Quat tangent_space(V3 N) {  // :150-159
	if (N.z < -1.0f + FLT_EPSILON) return {0.0f, 0.0f, 1.0f, 0.0f};
	float s = std::sqrt(2.0f * (N.z + 1.0f));
	float invs = 1.0f / s;
	return {s * 0.5f, -N.y * invs, N.x * invs, 0.0f};
}
static inline V3 to_local(Quat T, V3 v) {  // :161-169
	float temp = 2.0f * (v.z * T.w + v.x * T.y - T.x * v.y);
	return {v.x - T.y * temp, v.y + T.x * temp, temp * T.w - v.z};
}
static inline V3 to_world(Quat T, V3 v) {  // :171-179
	float temp = 2.0f * (v.z * T.w - v.x * T.y + T.x * v.y);
	return {v.x + T.y * temp, v.y - T.x * temp, temp * T.w - v.z};
}
static inline float conePdf(float cosThetaMax) { return kOneOverTwoPi / smax(1e-6f, 1.0f - cosThetaMax); }  // :192-194
static inline float spherePdf(float radius_sq, float dist_sq) {  // :196-200
	float sinThetaMax2 = radius_sq / dist_sq;
	float cosThetaMax = std::sqrt(smax(0.0f, 1.0f - sinThetaMax2));
	return conePdf(cosThetaMax);
}
static inline V3 sample_direction_to_sphere(V3 Wc, float sinThetaMax2, float center_dist, float radius2,
                                            float t, float s, float* out_distance, float* out_pdf) {  // :220-239
	float cosThetaMax = std::sqrt(smax(0.0f, 1.0f - sinThetaMax2));
	*out_pdf = conePdf(cosThetaMax);
	float cosTheta = 1.0f - t * (1.0f - cosThetaMax);
	float sinTheta = std::sqrt(sinThetaMax2 * t);
	float src_blend = (sinThetaMax2 < 0.00068523f ? sinTheta : cosTheta);
	float invert = std::sqrt(smax(0.0f, 1.0f - src_blend * src_blend));
	cosTheta = (sinThetaMax2 < 0.00068523f ? invert : cosTheta);
	sinTheta = (sinThetaMax2 < 0.00068523f ? sinTheta : invert);
	float temp = center_dist * sinTheta;
	*out_distance = center_dist * cosTheta - std::sqrt(smax(0.0f, radius2 - temp * temp)) - 1e-5f;
	V3 Ll = spherical_to_cartesian(s, sinTheta, cosTheta);
	V3 wcX, wcY; orthonormal_basis(Wc, &wcX, &wcY);
	return {wcX.x * Ll.x + wcY.x * Ll.y + Wc.x * Ll.z,
	        wcX.y * Ll.x + wcY.y * Ll.y + Wc.y * Ll.z,
	        wcX.z * Ll.x + wcY.z * Ll.y + Wc.z * Ll.z};
}
// ---- GGX (the reference's `#define BRDF 1` build: Closure<ClosureType::GGX>, DataStreams.hpp:184-219) — Sampling.hpp:102-104,249-309
static inline float gmix(float a, float b, float t) { return a * (1.0f - t) + b * t; }  // [glm] mix(x, y, a) = x * (1 - a) + y * a
static inline void disk(float t, float s, float* x, float* y) {  // :102-104 via polar_to_cartesian :85-91
	float cos_phi, sin_phi; fast_sincos(s * kTwoPi, &sin_phi, &cos_phi);
	const float rho = std::sqrt(t);
	*x = rho * cos_phi; *y = rho * sin_phi;
}
static inline V3 distribution_visible_normals(V3 Vlocal, float alpha, float u, float v) {  // :253-270
	V3 V = normalize(V3{alpha * Vlocal.x, alpha * Vlocal.y, Vlocal.z});
	float sx, sy; disk(u, v, &sx, &sy);
	const float t = 1.0f - sx * sx;
	sy = gmix(std::sqrt(t), sy, V.z * 0.5f + 0.5f);
	V3 X, Y; orthonormal_basis(V, &X, &Y);
	V3 H = X * sx + Y * sy + V * std::sqrt(smax(0.0f, t - sy * sy));
	return normalize(V3{alpha * H.x, alpha * H.y, smax(0.0f, H.z)});
}
static inline float pow5(float x) { float t = x * x; t *= t; return x * t; }  // :272
static inline V3 Fresnel(V3 F0, float HdotV) {  // :273-275: glm::mix(F0, Spectrum{1}, pow5(clamp(1 - HdotV, 0, 1)))
	float c = 1.0f - HdotV; c = c < 0.0f ? 0.0f : (1.0f < c ? 1.0f : c);  // std::clamp(v, lo, hi) = v < lo ? lo : hi < v ? hi : v
	const float a = pow5(c);
	return V3{F0.x * (1.0f - a) + 1.0f * a, F0.y * (1.0f - a) + 1.0f * a, F0.z * (1.0f - a) + 1.0f * a};
}
static inline float GGX_D(float alpha2, float NdotH2) {  // :278-281
	float temp = (1.0f + (alpha2 - 1.0f) * NdotH2);
	return alpha2 / (kPi * temp * temp);
}
static inline float Smith_G2_Height_Correlated_GGX_Lagarde(float alpha2, float NdotL, float NdotV) {  // :287-291
	float a = NdotV * std::sqrt(alpha2 + NdotL * (NdotL - alpha2 * NdotL));
	float b = NdotL * std::sqrt(alpha2 + NdotV * (NdotV - alpha2 * NdotV));
	return 0.5f / (a + b);
}
static inline V3 microfacet_brdf(V3 F0, float alpha, float NdotV, float NdotL, float NdotH, float HdotV) {  // :293-296
	const float alpha2 = alpha * alpha;
	return Fresnel(F0, HdotV) * (NdotL * GGX_D(smax(0.00001f, alpha2), NdotH * NdotH) * Smith_G2_Height_Correlated_GGX_Lagarde(alpha2, NdotL, NdotV));
}
static inline float G1_GGX(float alpha2, float NdotS2) { return 2.0f / (1.0f + std::sqrt(((alpha2 * (1.0f - NdotS2)) + NdotS2) / NdotS2)); }  // :297-299
static inline float Smith_G2_Over_G1_Height_Correlated(float alpha2, float NdotL, float NdotV) {  // :301-305
	float G1V = G1_GGX(alpha2, NdotV * NdotV);
	float G1L = G1_GGX(alpha2, NdotL * NdotL);
	return G1L / (G1V + G1L - G1V * G1L);
}
static inline V3 vndf_estimator(V3 F0, float alpha, float NdotV, float NdotL, float HdotV) {  // :307-309
	return Fresnel(F0, HdotV) * Smith_G2_Over_G1_Height_Correlated(alpha * alpha, NdotL, NdotV);
}
// Closure<GGX>::eval, DataStreams.hpp:189-195
static inline V3 ggx_eval(V3 F0, float alpha, V3 Llocal, V3 Vlocal) {
	float NdotL = smax(0.0f, Llocal.z);
	float NdotV = smax(0.0f, Vlocal.z);
	V3 Hn = normalize(Llocal + Vlocal);
	float NdotH = smax(0.0f, Hn.z);
	float HdotV = smax(0.0f, dot(Hn, Vlocal));
	return microfacet_brdf(F0, alpha, NdotV, NdotL, NdotH, HdotV);
}
// Closure<GGX>::sample, DataStreams.hpp:199-218: direction (local frame) and estimator
static inline void ggx_sample(V3 F0, float alpha, V3 Vlocal, float u0, float u1, V3* dir, V3* estimator) {
	float NdotV = smax(0.0f, Vlocal.z);
	float HdotV;
	if (alpha == 0.0f) { *dir = V3{-Vlocal.x, -Vlocal.y, Vlocal.z}; HdotV = NdotV; }
	else {
		V3 Hlocal = distribution_visible_normals(Vlocal, alpha, u0, u1);
		HdotV = dot(Hlocal, Vlocal);
		*dir = Hlocal * (2.0f * HdotV) - Vlocal;   // (2.0f * HdotV) * Hlocal - Vlocal
		HdotV = smax(0.0f, HdotV);
	}
	float NdotL = smax(0.0f, dir->z);
	*estimator = vndf_estimator(F0, alpha, NdotV, NdotL, HdotV);
}
static inline float powerHeuristic(float f, float g) { float f2 = f * f; return f2 / smax(1e-6f, f2 + g * g); }  // :241-244
static inline float powerHeuristic_over_f(float f, float g) { return f / smax(1e-6f, f * f + g * g); }  // :245-247

// ---------------------------------------------------------------- Color.hpp:47-49, 66-73 (Vec8f lane-wise, no FMA)
static inline float aces_fit(float x) { return (x * (x + 0.0245786f) - 0.000090537f) / (x * (0.983729f * x + 0.4329510f) + 0.238081f); }
static inline void tonemapping(float& r, float& g, float& b) {
	float x = aces_fit(r * 0.59719f + g * 0.35458f + b * 0.04823f);
	float y = aces_fit(r * 0.07600f + g * 0.90834f + b * 0.01566f);
	float z = aces_fit(r * 0.02840f + g * 0.13383f + b * 0.83777f);
	// VCL min(a,b)=_mm256_min_ps(a,b) -> (a<b)?a:b ; max(a,b)=_mm256_max_ps(a,b) -> (a>b)?a:b ; second operand on NaN
	auto vmin = [](float a, float b) { return (a < b) ? a : b; };
	auto vmax = [](float a, float b) { return (a > b) ? a : b; };
	r = vmin(1.0f, vmax(0.0f, x * 1.604750f + y * -0.53108f + z * -0.07367f));
	g = vmin(1.0f, vmax(0.0f, x * -0.10208f + y * 1.10813f + z * -0.00605f));
	b = vmin(1.0f, vmax(0.0f, x * -0.00327f + y * -0.07276f + z * 1.07602f));
}

}  // namespace orc
