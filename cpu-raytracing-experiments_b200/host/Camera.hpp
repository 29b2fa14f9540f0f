// Camera.hpp — pinhole camera of the reference (Camera.hpp:5-88): quaternion view, focal-length projection. The orientation
// maths (glm::quatLookAt) is done by libb2r (b2r_camera_lookat) so host and device agree bit-for-bit.
#pragma once
#include "Core.hpp"
#include "../../include/b2r.h"

struct Projection {  // Camera.hpp:5-45
	Projection(uint32_t W, uint32_t H, float focal_length, float focus_distance, float f_number)
		: focal_length(focal_length), focus_distance(focus_distance), f_number(f_number) { UpdateLensOnly(); Resize(W, H); }
	static float calc_field_of_view(float focal_length, float sensor_size = 24.0f) { return 2.0f * std::atan((sensor_size / 2.0f) / focal_length); }
	static float calc_aperture(float focal_length, float f_number) { return focal_length / (2.0f * f_number); }
	void UpdateLens() { UpdateLensOnly(); z = half_height * inv_half_tan; }
	void Resize(uint32_t W, uint32_t H) { half_height = static_cast<float>(H) * 0.5f; half_width = static_cast<float>(W) * 0.5f; z = half_height * inv_half_tan; }
	float half_height = 0.5f, half_width = 0.5f, z = 0, aperture_radius = 0, inv_half_tan = 0;
	float focal_length, focus_distance, f_number;
	float near = 0.0f, far = 1000.0f;
private:
	void UpdateLensOnly() { inv_half_tan = (-2.0f / 24.0f) * focal_length; aperture_radius = calc_aperture(focal_length, f_number); }
};

struct View {  // Camera.hpp:47-59
	View(b2r_host::vec3 eye, b2r_host::vec3 forward) : pos(eye) {
		const float e[3] = {eye.x, eye.y, eye.z}, d[3] = {forward.x, forward.y, forward.z};
		float out[11]; b2r_camera_lookat(e, d, 16, 16, 50.0f, 1.0f, out);
		orient = b2r_host::quat(out[3], out[4], out[5], out[6]);
	}
	// orient = conjugate(normalize(quat(angles) * conjugate(orient)))  (Camera.hpp:51-53)
	void Rotate(b2r_host::vec3 angles) {
		const float cx = std::cos(angles.x * 0.5f), cy = std::cos(angles.y * 0.5f), cz = std::cos(angles.z * 0.5f);
		const float sx = std::sin(angles.x * 0.5f), sy = std::sin(angles.y * 0.5f), sz = std::sin(angles.z * 0.5f);
		const b2r_host::quat a(cx * cy * cz + sx * sy * sz, sx * cy * cz - cx * sy * sz, cx * sy * cz + sx * cy * sz, cx * cy * sz - sx * sy * cz);
		const b2r_host::quat b(orient.w, -orient.x, -orient.y, -orient.z);
		b2r_host::quat p(a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
		                 a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x);
		// glm::normalize(qua): length = sqrt(dot), glm's quaternion dot adds pairwise, (w*w + x*x) + (y*y + z*z); identity for a zero length
		const float len = std::sqrt((p.w * p.w + p.x * p.x) + (p.y * p.y + p.z * p.z));
		if (len <= 0.0f) { orient = b2r_host::quat(1.0f, 0.0f, 0.0f, 0.0f); return; }
		const float inv = 1.0f / len;
		orient = b2r_host::quat(p.w * inv, -(p.x * inv), -(p.y * inv), -(p.z * inv));
	}
	void Translate(b2r_host::vec3 t) {  // pos += orient * t
		const b2r_host::vec3 q{orient.x, orient.y, orient.z};
		auto cross = [](b2r_host::vec3 a, b2r_host::vec3 b) { return b2r_host::vec3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; };
		const b2r_host::vec3 uv = cross(q, t), uuv = cross(q, uv);
		pos = pos + (t + (uv * orient.w + uuv) * 2.0f);
	}
	b2r_host::quat orient;
	b2r_host::vec3 pos;
};

struct Camera {  // Camera.hpp:61-88
	Camera(b2r_host::vec3 eye = {0, 0, 0}, b2r_host::vec3 direction = {0, 0, -1}, uint32_t width = 1, uint32_t height = 1,
	       float focal_length = 50.0f, float focus_distance = 1.0f, float f_number = 16.0f, float exposure = 1.0f)
		: view(eye, direction), projection(width, height, focal_length, focus_distance, f_number), exp(exposure) {}
	void Resize(uint32_t width, uint32_t height) { projection.Resize(width, height); }
	void RotateLocal(b2r_host::vec3 angles) { view.Rotate(angles); }
	void TranslateLocal(b2r_host::vec3 t) { view.Translate(t); }
	View view;
	Projection projection;
	float exp;

	struct Ray { b2r_host::vec3 origin, dir; };
	// Camera.hpp:80-88: the ray through pixel (x, y) with sub-pixel position samples[0..1]; computed by libb2r's camera routine (the one the
	// kernels run per pixel), so `const auto [orig, dir] = scene.camera.generate_ray(x, y, no_pixel_jitter)` (Application.cpp:288) compiles
	// unchanged and gives the renderer's own ray bit for bit.
	Ray generate_ray(int32_t x, int32_t y, const float* samples) const noexcept {
		const float p[3] = {view.pos.x, view.pos.y, view.pos.z}, q[4] = {view.orient.w, view.orient.x, view.orient.y, view.orient.z};
		float o[3], d[3];
		b2r_camera_ray(p, q, projection.half_width, projection.half_height, projection.z, x, y, samples, o, d);
		return {view.pos, {d[0], d[1], d[2]}};
	}
};
