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
};
