// Minimal stand-in for the parts of glm that /root/reference/Sampling.hpp, Color.hpp:30-74, VectorMath.hpp:581-662 and
// Camera.hpp:5-59,81-87 use, so that
// those reference files can be compiled VERBATIM by oracle/Makefile `ref` (glm itself is absent from this image and un-vendored
// by the reference). Only component-wise scalar definitions, as glm publishes them for its non-SIMD vec3/quat types:
// no arithmetic OF THE REFERENCE is replaced — every formula, constant and operation order under test comes from the
// reference's own source text. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
namespace glm {
struct vec2 { union { float x, r; }; union { float y, g; }; vec2() : x(0), y(0) {} vec2(float a, float b) : x(a), y(b) {} };
struct vec3 {
	union { float x, r; }; union { float y, g; }; union { float z, b; };
	vec3() : x(0), y(0), z(0) {}
	explicit vec3(float s) : x(s), y(s), z(s) {}
	vec3(float a, float b_, float c) : x(a), y(b_), z(c) {}
	vec3& operator*=(float s) { x *= s; y *= s; z *= s; return *this; }
	vec3& operator*=(const vec3& o) { x *= o.x; y *= o.y; z *= o.z; return *this; }
	vec3& operator+=(const vec3& o) { x += o.x; y += o.y; z += o.z; return *this; }
	float& operator[](int i) { return (&x)[i]; }
	const float& operator[](int i) const { return (&x)[i]; }
};
template <int L, class T> struct vec_of;               // glm::vec<3, float> as BVH.hpp:21 spells it
template <> struct vec_of<3, float> { using type = vec3; };
template <int L, class T> using vec = typename vec_of<L, T>::type;
struct vec4 { float x, y, z, w; };
struct mat4 { float m[16]; };
struct quat {  // constructor order (w, x, y, z) as in glm::qua
	float x, y, z, w;
	quat() : x(0), y(0), z(0), w(1) {}
	quat(float w_, float x_, float y_, float z_) : x(x_), y(y_), z(z_), w(w_) {}
	explicit quat(const vec3& euler);
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator+(vec3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline vec3 operator-(vec3 a, float s) { return {a.x - s, a.y - s, a.z - s}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 min(vec3 a, vec3 b) { return {(b.x < a.x) ? b.x : a.x, (b.y < a.y) ? b.y : a.y, (b.z < a.z) ? b.z : a.z}; }  // glm::min(x, y) = (y < x) ? y : x
inline vec3 max(vec3 a, vec3 b) { return {(a.x < b.x) ? b.x : a.x, (a.y < b.y) ? b.y : a.y, (a.z < b.z) ? b.z : a.z}; }  // glm::max(x, y) = (x < y) ? y : x
inline float min(float a, float b) { return (b < a) ? b : a; }
inline float max(float a, float b) { return (a < b) ? b : a; }
inline vec3 abs(vec3 a) { return {std::fabs(a.x), std::fabs(a.y), std::fabs(a.z)}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }            // glm compute_dot<vec3>: (x*x + y*y) + z*z, left to right
inline float length(vec3 v) { return std::sqrt(dot(v, v)); }
inline float distance(vec3 a, vec3 b) { return length(b - a); }
inline vec3 normalize(vec3 v) { return v * (1.0f / std::sqrt(dot(v, v))); }                // v * inversesqrt(dot(v, v))
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }             // glm: x * (1 - a) + y * a
inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
// ---- what Camera.hpp:5-59,81-87 needs (glm/geometric.inl, gtc/quaternion.inl, gtx/quaternion.inl as published)
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 cross(vec3 x, vec3 y) { return {x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y}; }
inline float atan(float x) { return std::atan(x); }
inline vec3 operator*(quat q, vec3 v) {  // v + ((uv * q.w) + uuv) * 2
	const vec3 qv(q.x, q.y, q.z); const vec3 uv(cross(qv, v)); const vec3 uuv(cross(qv, uv));
	return v + ((uv * q.w) + uuv) * 2.0f;
}
inline vec3 rotate(quat q, vec3 v) { return q * v; }
inline quat conjugate(quat q) { return quat(q.w, -q.x, -q.y, -q.z); }
inline quat operator*(quat p, quat q) {
	return quat(p.w * q.w - p.x * q.x - p.y * q.y - p.z * q.z, p.w * q.x + p.x * q.w + p.y * q.z - p.z * q.y,
	            p.w * q.y + p.y * q.w + p.z * q.x - p.x * q.z, p.w * q.z + p.z * q.w + p.x * q.y - p.y * q.x);
}
inline quat normalize(quat q) {
	// length(qua) = sqrt(dot(q, q)) with glm's quaternion dot (detail/type_quat.inl, compute_dot<qua>): tmp(w*w, x*x, y*y, z*z); (tmp.x + tmp.y) + (tmp.z + tmp.w)
	const float len = std::sqrt((q.w * q.w + q.x * q.x) + (q.y * q.y + q.z * q.z));
	if (len <= 0.0f) return quat(1, 0, 0, 0);
	const float inv = 1.0f / len; return quat(q.w * inv, q.x * inv, q.y * inv, q.z * inv);
}
inline quat quat_from_euler(vec3 e);
inline quat::quat(const vec3& euler) { *this = quat_from_euler(euler); }
inline quat quat_from_euler(vec3 e) {  // qua(vec3 eulerAngle)
	const vec3 c(std::cos(e.x * 0.5f), std::cos(e.y * 0.5f), std::cos(e.z * 0.5f)), s(std::sin(e.x * 0.5f), std::sin(e.y * 0.5f), std::sin(e.z * 0.5f));
	return quat(c.x * c.y * c.z + s.x * s.y * s.z, s.x * c.y * c.z - c.x * s.y * s.z, c.x * s.y * c.z + s.x * c.y * s.z, c.x * c.y * s.z - s.x * s.y * c.z);
}
inline quat quatLookAt(vec3 direction, vec3 up) {  // quatLookAtRH + quat_cast(mat3), columns {right, up', -direction}
	const vec3 c2 = -direction;
	const vec3 right = cross(up, c2);
	const float d = dot(right, right);
	const vec3 c0 = right * (1.0f / std::sqrt(d > 0.00001f ? d : 0.00001f));
	const vec3 c1 = cross(c2, c0);
	const float fx = c0.x - c1.y - c2.z, fy = c1.y - c0.x - c2.z, fz = c2.z - c0.x - c1.y, fw = c0.x + c1.y + c2.z;
	int big = 0; float fb = fw;
	if (fx > fb) { fb = fx; big = 1; }
	if (fy > fb) { fb = fy; big = 2; }
	if (fz > fb) { fb = fz; big = 3; }
	const float v = std::sqrt(fb + 1.0f) * 0.5f, m = 0.25f / v;
	switch (big) {
	case 0: return quat(v, (c1.z - c2.y) * m, (c2.x - c0.z) * m, (c0.y - c1.x) * m);
	case 1: return quat((c1.z - c2.y) * m, v, (c0.y + c1.x) * m, (c2.x + c0.z) * m);
	case 2: return quat((c2.x - c0.z) * m, (c0.y + c1.x) * m, v, (c1.z + c2.y) * m);
	default: return quat((c0.y - c1.x) * m, (c2.x + c0.z) * m, (c1.z + c2.y) * m, v);
	}
}
template <class T> constexpr T pi() { return T(3.14159265358979323846264338327950288); }
template <class T> constexpr T two_pi() { return T(6.28318530717958647692528676655900576); }
template <class T> constexpr T half_pi() { return T(1.57079632679489661923132169163975144); }
template <class T> constexpr T one_over_pi() { return T(0.318309886183790671537767526745028724); }
template <class T> constexpr T one_over_two_pi() { return T(0.159154943091895335768883763372514362); }
}  // namespace glm
