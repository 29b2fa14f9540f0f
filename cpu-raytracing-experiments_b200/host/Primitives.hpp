// Primitives.hpp — Sphere / Material / Sky, layout-identical to the reference (Primitives.hpp:7-47) and to the C ABI PODs.
#pragma once
#include "Core.hpp"
#include "../../include/b2r.h"

struct alignas(16) Sphere {  // Primitives.hpp:7-17
	b2r_host::vec3 position;
	float radius_sq = 0;
	int32_t material_ID = 0;
	Sphere() = default;
	Sphere(b2r_host::vec3 p, float r2, int32_t m) : position(p), radius_sq(r2), material_ID(m) {}
};
static_assert(sizeof(Sphere) == sizeof(b2r_sphere), "Sphere must match b2r_sphere (32 B)");

struct alignas(32) Material {  // Primitives.hpp:18-27
	b2r_host::vec3 albedo, F0, F80, emission, transmission;
	float roughness = 0, IOR_minus_one = 0;
};
static_assert(sizeof(Material) == sizeof(b2r_material), "Material must match b2r_material (96 B)");

struct Sky {  // Primitives.hpp:29-47: equirect HDRI (RGBA32F, nearest texel) x ambient colour; evaluated on the GPU
	b2r_host::vec3 ambient_color;
	int32_t hdri_width = 0, hdri_height = 0, hdri_channels = 0;
	float* hdri_data = nullptr;
	float hdri_fwidth = 0, hdri_fheight = 0;
	// Application.cpp:225-231: hdri_data = stbi_loadf(path, &hdri_width, &hdri_height, &hdri_channels, 4); terminate on failure;
	// hdri_fwidth/fheight = size - 1. Here through b2r_read_hdr (the Radiance decoder of stb_image restated); returns false where the
	// reference prints stbi_failure_reason() and terminates. The array is owned by `hdri_storage`.
	std::vector<float> hdri_storage;
	bool Load(const char* path) {
		int32_t w = 0, h = 0;
		if (b2r_read_hdr(path, nullptr, &w, &h) != B2R_OK) return false;
		hdri_storage.assign(static_cast<size_t>(w) * h * 4, 0.0f);
		if (b2r_read_hdr(path, hdri_storage.data(), &w, &h) != B2R_OK) return false;
		hdri_data = hdri_storage.data(); hdri_width = w; hdri_height = h; hdri_channels = 3;
		hdri_fwidth = static_cast<float>(w - 1); hdri_fheight = static_cast<float>(h - 1);
		return true;
	}
};
