// Scene.hpp — Scene aggregate and LightingAcceleration of the reference (Scene.hpp:9-26).
#pragma once
#include <vector>
#include "Camera.hpp"
#include "Primitives.hpp"
#include "BVH.hpp"

struct LightingAcceleration {  // Scene.hpp:9-17: indices (into scene.geometry) of the emissive spheres
	std::vector<int32_t> prims;
	LightingAcceleration() {}
	LightingAcceleration(const std::vector<Sphere>& src_prims, const std::vector<Material>& material) {
		uint32_t n = 0;
		prims.resize(src_prims.size());
		b2r_find_lights(reinterpret_cast<const b2r_sphere*>(src_prims.data()), static_cast<uint32_t>(src_prims.size()),
		                reinterpret_cast<const b2r_material*>(material.data()), static_cast<uint32_t>(material.size()), prims.data(), &n);
		prims.resize(n);
	}
};

struct Scene {  // Scene.hpp:19-26
	std::vector<Sphere> geometry;
	std::vector<Material> material;
	LightingAcceleration lighting_acceleration;
	Camera camera;
	Sky sky;
	BoundingVolumeHierarchy<Sphere> acceleration_structure;
};
