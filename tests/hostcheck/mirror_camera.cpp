// mirror_camera.cpp — TEST HARNESS (part of tests/hostcheck/libhostcheck.so): the C++ mirror's Camera (cpu-raytracing-experiments_b200/host/
// Camera.hpp — what a user of the mirror moves around with the reference's RotateLocal / TranslateLocal calls) behind a C tap, so that
// tests/test_oracle_ref_sampling.py can compare it with the reference's own Camera.hpp compiled into oracle/_ref/librefsampling.so.
#include "Camera.hpp"

extern "C" void hc_mirror_camera_move(const float eye[3], const float dir[3], const float* angles3, const float* offsets3, uint32_t n_moves, float out7[7]) {
	Camera c({eye[0], eye[1], eye[2]}, {dir[0], dir[1], dir[2]}, 16, 16, 50.0f);
	for (uint32_t i = 0; i < n_moves; i++) {
		c.RotateLocal({angles3[3 * i], angles3[3 * i + 1], angles3[3 * i + 2]});      // Camera.hpp:74-76 (Application.cpp:299)
		c.TranslateLocal({offsets3[3 * i], offsets3[3 * i + 1], offsets3[3 * i + 2]}); // Camera.hpp:77-79 (Application.cpp:236-247)
	}
	out7[0] = c.view.pos.x; out7[1] = c.view.pos.y; out7[2] = c.view.pos.z;
	out7[3] = c.view.orient.w; out7[4] = c.view.orient.x; out7[5] = c.view.orient.y; out7[6] = c.view.orient.z;
}

// Camera::generate_ray of the mirror (Camera.hpp:80-88; Application.cpp:288 calls it for focus picking): out6 = origin, dir; cam7 = orient wxyz,
// half_width, half_height, z — the same row layout as ref_generate_ray of oracle/ref_sampling_wrap.cpp
extern "C" void hc_mirror_generate_ray(const float eye[3], const float dir[3], uint32_t w, uint32_t h, float focal_mm, int32_t x, int32_t y, const float samples[2], float out6[6], float cam7[7]) {
	Camera c({eye[0], eye[1], eye[2]}, {dir[0], dir[1], dir[2]}, w, h, focal_mm);
	const auto [orig, d] = c.generate_ray(x, y, samples);
	out6[0] = orig.x; out6[1] = orig.y; out6[2] = orig.z; out6[3] = d.x; out6[4] = d.y; out6[5] = d.z;
	cam7[0] = c.view.orient.w; cam7[1] = c.view.orient.x; cam7[2] = c.view.orient.y; cam7[3] = c.view.orient.z;
	cam7[4] = c.projection.half_width; cam7[5] = c.projection.half_height; cam7[6] = c.projection.z;
}
