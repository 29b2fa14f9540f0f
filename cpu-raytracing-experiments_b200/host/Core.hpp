// Core.hpp — vector types of the C++ mirror. The reference uses glm (Core.hpp:13-23, un-vendored); define B2R_USE_GLM
// before including to keep glm::vec3/quat in an application that already has it, otherwise these minimal PODs with the same
// member names are used. Only storage and the few operators the mirrored interface needs — no arithmetic of the hot path
// lives on the host.
#pragma once
#include <cstdint>
#include <cmath>
#ifdef B2R_USE_GLM
#include <glm/vec3.hpp>
#include <glm/vec4.hpp>
#include <glm/gtc/quaternion.hpp>
namespace b2r_host { using vec3 = glm::vec3; using vec4 = glm::vec4; using quat = glm::quat; }
#else
namespace b2r_host {
struct vec3 {
	float x = 0, y = 0, z = 0;
	vec3() = default;
	explicit vec3(float s) : x(s), y(s), z(s) {}
	vec3(double X, double Y, double Z) : x(static_cast<float>(X)), y(static_cast<float>(Y)), z(static_cast<float>(Z)) {}
	float& operator[](int i) { return (&x)[i]; }
	const float& operator[](int i) const { return (&x)[i]; }
};
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return a * s; }
struct vec4 { float x = 0, y = 0, z = 0, w = 0; };
struct quat {  // (w, x, y, z) constructor order like glm::quat
	float w = 1, x = 0, y = 0, z = 0;
	quat() = default;
	quat(float W, float X, float Y, float Z) : w(W), x(X), y(Y), z(Z) {}
};
}  // namespace b2r_host
#endif
using Spectrum = b2r_host::vec3;  // Core.hpp:34
