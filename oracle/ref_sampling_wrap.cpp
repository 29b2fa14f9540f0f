// C wrappers around the reference's OWN sampling / scalar-math / tonemapping source, compiled verbatim from /root/reference
// (never copied into this repo): Sampling.hpp whole, plus two line ranges that oracle/Makefile `ref` extracts at build time
// into a temporary directory because their files as a whole use MSVC-only constructs g++ rejects:
//   vm_scalar.inc     = VectorMath.hpp:581-662  (fast_abs/sign/copysign/round, f_xor/f_or/f_and, fast_asin, fast_atan2, fast_sincos)
//   color_tonemap.inc = Color.hpp:30-74         (ACES_input, ACES_rtt_odt_fit, ACES_output, tonemapping scalar and Vec8f)
//   camera_projection_view.inc = Camera.hpp:5-59, camera_generate_ray_body.inc = Camera.hpp:81-87
// glm and VCL are replaced by the minimal stand-ins under ref_shim/ (component-wise definitions only). Used by
// tests/test_oracle_ref_sampling.py (live, where /root/reference exists) and tests/gen_golden.py (-> tests/golden/sampling_kat.json).
// One platform note: the reference calls sqrt()/abs() unqualified on floats. MSVC's <cmath> declares the float overloads in the
// global namespace, so there `sqrt(float)` is a float operation; libstdc++'s <cmath> leaves only the C double versions global
// (which would turn `center_dist * cosTheta - sqrt(..) - 1e-5f`, Sampling.hpp:231, into double arithmetic and `abs(v.x)` into
// abs(int)). <math.h> / <stdlib.h> are libstdc++'s way to get MSVC's overload set, so they are included first.
#include <math.h>
#include <stdlib.h>
#include <cfloat>
#include <cstdint>
#include <algorithm>
#include <immintrin.h>
#define __vectorcall
#include "Core.hpp"
#include "Random.hpp"
#include "vm_scalar.inc"
#include "Sampling.hpp"
#include "color_tonemap.inc"
// Camera.hpp as a whole defines a struct inside a return type (line 80), which g++ rejects; Projection (:5-45), View (:47-59) and the
// body of generate_ray (:81-87) are cut out verbatim instead and the rejected declarator is the only line written here.
#include "camera_projection_view.inc"
struct RefRay { glm::vec3 origin, dir; };
struct RefCamera {
	View view; Projection projection;
	RefCamera(glm::vec3 eye, glm::vec3 direction, uint32_t width, uint32_t height, float focal_length) : view(eye, direction), projection(width, height, focal_length, 1.0f, 16.0f) {}
	RefRay generate_ray(int32_t x, int32_t y, const float* const __restrict samples) const noexcept {
#include "camera_generate_ray_body.inc"
	}
};
extern "C" {
// out: origin xyz, dir xyz; also returns orient (w,x,y,z), half_width, half_height, z through cam_out[7]
void ref_generate_ray(const float eye[3], const float dir[3], uint32_t w, uint32_t h, float focal_mm, int32_t x, int32_t y, const float samples[2], float out[6], float cam_out[7]) {
	RefCamera c(glm::vec3{eye[0], eye[1], eye[2]}, glm::vec3{dir[0], dir[1], dir[2]}, w, h, focal_mm);
	RefRay r = c.generate_ray(x, y, samples);
	out[0] = r.origin.x; out[1] = r.origin.y; out[2] = r.origin.z; out[3] = r.dir.x; out[4] = r.dir.y; out[5] = r.dir.z;
	cam_out[0] = c.view.orient.w; cam_out[1] = c.view.orient.x; cam_out[2] = c.view.orient.y; cam_out[3] = c.view.orient.z;
	cam_out[4] = c.projection.half_width; cam_out[5] = c.projection.half_height; cam_out[6] = c.projection.z;
}
// Camera::RotateLocal / TranslateLocal (Camera.hpp:51-56,74-79; the app's mouse / WASD handlers, Application.cpp:236-247,299): starting from
// Camera{eye, dir}, `n_moves` x {Rotate(angles), Translate(offset)}; out7 = pos xyz, orient wxyz
void ref_camera_move(const float eye[3], const float dir[3], const float* angles3, const float* offsets3, uint32_t n_moves, float out7[7]) {
	View v(glm::vec3{eye[0], eye[1], eye[2]}, glm::vec3{dir[0], dir[1], dir[2]});
	for (uint32_t i = 0; i < n_moves; i++) {
		v.Rotate(glm::vec3{angles3[3 * i], angles3[3 * i + 1], angles3[3 * i + 2]});
		v.Translate(glm::vec3{offsets3[3 * i], offsets3[3 * i + 1], offsets3[3 * i + 2]});
	}
	out7[0] = v.pos.x; out7[1] = v.pos.y; out7[2] = v.pos.z; out7[3] = v.orient.w; out7[4] = v.orient.x; out7[5] = v.orient.y; out7[6] = v.orient.z;
}
void ref_fast_sincos(float x, float* s, float* c) { fast_sincos(x, s, c); }
float ref_fast_asin(float x) { return fast_asin(x); }
float ref_fast_atan2(float y, float x) { return fast_atan2(y, x); }
float ref_median3(float a, float b, float c) { return median<float>(a, b, c); }
float ref_median5(const float v[5]) { return median<float>(v[0], v[1], v[2], v[3], v[4]); }
void ref_hemisphere(float t, float s, float out[3]) { glm::vec3 v = hemisphere(t, s); out[0] = v.x; out[1] = v.y; out[2] = v.z; }
void ref_orthonormal_basis(const float n[3], float out[6]) { glm::vec3 a, b; orthonormal_basis(glm::vec3{n[0], n[1], n[2]}, &a, &b); out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = b.x; out[4] = b.y; out[5] = b.z; }
void ref_tangent_space(const float n[3], float out_wxyz[4]) { glm::quat q = tangent_space(glm::vec3{n[0], n[1], n[2]}); out_wxyz[0] = q.w; out_wxyz[1] = q.x; out_wxyz[2] = q.y; out_wxyz[3] = q.z; }
void ref_to_local(const float q[4], const float v[3], float out[3]) { glm::vec3 r = to_local(glm::quat{q[0], q[1], q[2], q[3]}, glm::vec3{v[0], v[1], v[2]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void ref_to_world(const float q[4], const float v[3], float out[3]) { glm::vec3 r = to_world(glm::quat{q[0], q[1], q[2], q[3]}, glm::vec3{v[0], v[1], v[2]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
float ref_cone_pdf(float c) { return conePdf(c); }
float ref_sphere_pdf(float r2, float d2) { return spherePdf(r2, d2); }
void ref_sample_direction_to_sphere(const float wc[3], float s2, float cd, float r2, float t, float s, float out[5]) {
	float dist, pdf; glm::vec3 L = sample_direction_to_sphere(glm::vec3{wc[0], wc[1], wc[2]}, s2, cd, r2, t, s, &dist, &pdf);
	out[0] = L.x; out[1] = L.y; out[2] = L.z; out[3] = dist; out[4] = pdf;
}
float ref_power_heuristic(float f, float g) { return powerHeuristic(f, g); }
float ref_power_heuristic_over_f(float f, float g) { return powerHeuristic_over_f(f, g); }
void ref_tonemap_scalar(float rgb[3]) { tonemapping(rgb[0], rgb[1], rgb[2], rgb, rgb + 1, rgb + 2); }
void ref_tonemap_vec8(float r[8], float g[8], float b[8]) { Vec8f R, G, B; R.load(r); G.load(g); B.load(b); tonemapping(R, G, B); R.store(r); G.store(g); B.store(b); }
}
