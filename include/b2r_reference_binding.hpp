// b2r_reference_binding.hpp — the binding a maintainer of Borx25/CPU-Raytracing-experiments would add: a drop-in for
// `Renderer<Policy>` (Renderer.hpp:28-479) that works on the REFERENCE'S OWN `Scene` object (Scene.hpp:19-26) and forwards the hot
// path to libb2r.so through the C ABI (include/b2r.h). Include it after the reference's Scene.hpp:
//
//     #include "Scene.hpp"                       // the reference's: Sphere, Material, Camera, Sky, BoundingVolumeHierarchy, Scene
//     #include "b2r_reference_binding.hpp"
//     b2r::ReferenceRenderer<Scene, glm::vec4> renderer{scene};   // instead of  Renderer<> renderer{scene};
//
// Same public surface as the reference's class as the app uses it (Application.cpp:247,256,274,297-299,332,367,375-381,402,420,508-514):
// RequiredTiling(), Resize, ResetAccumulator, Accumulate, Render, framebuffer, width, height, accumulations, h_tiles, v_tiles.
// The reference's PODs are passed as they are: Sphere (32 B), Material (96 B) and BVH::Node (32 B) are layout-identical to
// b2r_sphere / b2r_material / b2r_bvh_node (static_asserts below). Differences from the reference class:
// the scene is snapshotted at upload, so after the app edits geometry or materials — where it already rebuilds the BVH and resets,
// Application.cpp:508-510 — it calls SceneChanged(), or SceneMoved() when spheres only moved (GPU refit of the traversal tree instead of a rebuild); camera state is re-read at every ResetAccumulator()/Resize().
// tests/refbinding/ compiles this header against the reference's real headers and runs it next to the reference's own Renderer<>.
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <type_traits>
#include <string>
#include <vector>
#include "b2r.h"

namespace b2r {

struct ReferencePolicy {  // RendererPolicy's run-time twin (Renderer.hpp:19-26); K = AccumulationBuckets (:41)
	uint32_t max_bounces = 16, buckets = 5, flags = 0; int32_t device = 0;
};

template <class SceneT, class Vec4T>
struct ReferenceRenderer {
	static constexpr size_t TileRoot = 16, TileSize = 256, StreamSize = 256;
	static constexpr size_t RequiredTiling() { return TileRoot; }
	const SceneT& scene;
	ReferencePolicy policy;
	std::vector<Vec4T> framebuffer;                    // RGBA32F, tonemapped, row 0 = y 0 (what Renderer.hpp:447-473 writes)
	uint32_t width = 0, height = 0, accumulations = 0, h_tiles = 0, v_tiles = 0;

	explicit ReferenceRenderer(const SceneT& s, ReferencePolicy p = {}) : scene(s), policy(p) { static_assert(sizeof(Vec4T) == 16, "framebuffer texel = 4 floats"); }
	~ReferenceRenderer() { if (!framebuffer.empty()) b2r_host_unregister(framebuffer.data()); if (ctx_) b2r_destroy(ctx_); }
	ReferenceRenderer(const ReferenceRenderer&) = delete;
	ReferenceRenderer& operator=(const ReferenceRenderer&) = delete;

	void Resize(uint32_t new_width, uint32_t new_height) {                   // Renderer.hpp:53-63
		width = new_width; height = new_height; h_tiles = width / TileRoot; v_tiles = height / TileRoot;
		if (!framebuffer.empty()) b2r_host_unregister(framebuffer.data());
		framebuffer.resize(static_cast<size_t>(width) * height);
		b2r_host_register(framebuffer.data(), framebuffer.size() * sizeof(framebuffer[0]));  // Render() then copies into it at DMA speed (best effort: failure is ignored)
		if (!ctx_) {
			b2r_config cfg{}; cfg.width = width; cfg.height = height; cfg.max_bounces = policy.max_bounces; cfg.buckets = policy.buckets;
			cfg.flags = policy.flags; cfg.device = policy.device;
			check(b2r_create(&ctx_, &cfg));
			SceneChanged();
		} else check(b2r_resize(ctx_, width, height));
		ResetAccumulator();
	}
	void ResetAccumulator() {                                                  // Renderer.hpp:64-67 (the app calls it after every camera move)
		accumulations = 0;
		if (!ctx_) return;
		check(b2r_reset(ctx_)); upload_camera();
	}
	// scene.geometry / material / sky / acceleration_structure / lighting_acceleration were changed and rebuilt by the app
	void SceneChanged() {
		using Sphere = typename std::remove_cvref_t<decltype(scene.geometry)>::value_type;
		using Material = typename std::remove_cvref_t<decltype(scene.material)>::value_type;
		using Node = typename std::remove_cvref_t<decltype(scene.acceleration_structure.nodes)>::value_type;
		static_assert(sizeof(Sphere) == sizeof(b2r_sphere) && sizeof(Material) == sizeof(b2r_material) && sizeof(Node) == sizeof(b2r_bvh_node), "reference PODs are layout-identical to the ABI's");
		const auto& bvh = scene.acceleration_structure; const auto& lights = scene.lighting_acceleration.prims;
		const float ambient[3] = {scene.sky.ambient_color.x, scene.sky.ambient_color.y, scene.sky.ambient_color.z};
		check(b2r_upload_scene(ctx_, reinterpret_cast<const b2r_sphere*>(bvh.prims.data()), reinterpret_cast<const b2r_bvh_node*>(bvh.nodes.data()),
		                       static_cast<uint32_t>(bvh.prims.size()), static_cast<uint32_t>(bvh.nodes.size()),
		                       reinterpret_cast<const b2r_material*>(scene.material.data()), static_cast<uint32_t>(scene.material.size()),
		                       lights.data(), static_cast<uint32_t>(lights.size()),
		                       reinterpret_cast<const b2r_sphere*>(scene.geometry.data()), static_cast<uint32_t>(scene.geometry.size()),
		                       ambient, scene.sky.hdri_data, scene.sky.hdri_width, scene.sky.hdri_height));
		upload_camera();
	}
	// Geometry was dragged in the editor and the app has rebuilt acceleration_structure / lighting_acceleration (Application.cpp:508-509):
	// instead of a full SceneChanged() the GPU keeps its traversal tree's topology, re-links its leaves to the rebuilt leaf order and
	// refits its boxes (b2r_refit_scene). Falls back to SceneChanged() when the sphere count changed or the kept topology has degraded
	// past `rebuild_above` (sum of inner-box areas relative to when the tree was built). Returns that ratio (1 after a rebuild).
	float SceneMoved(float rebuild_above = 1.5f) {
		const auto& bvh = scene.acceleration_structure; const auto& lights = scene.lighting_acceleration.prims;
		float quality = 1.0f;
		const int rc = b2r_refit_scene(ctx_, reinterpret_cast<const b2r_sphere*>(bvh.prims.data()), static_cast<uint32_t>(bvh.prims.size()),
		                               reinterpret_cast<const b2r_material*>(scene.material.data()), static_cast<uint32_t>(scene.material.size()),
		                               lights.data(), static_cast<uint32_t>(lights.size()),
		                               reinterpret_cast<const b2r_sphere*>(scene.geometry.data()), static_cast<uint32_t>(scene.geometry.size()), &quality);
		if (rc == B2R_ERR_ARG || rc == B2R_ERR_STATE || (rc == B2R_OK && quality > rebuild_above)) { SceneChanged(); return 1.0f; }  // spheres added / removed, or moved too far
		check(rc);
		return quality;
	}
	void Accumulate() { check(b2r_accumulate(ctx_, 1)); ++accumulations; }   // Renderer.hpp:73-434
	void Render() {                                                            // Renderer.hpp:436-478: silently does nothing unless accumulations % K == 0
		const int rc = b2r_resolve(ctx_, reinterpret_cast<float*>(framebuffer.data()), 1);
		if (rc != B2R_OK && rc != B2R_ERR_NOT_READY) check(rc);
	}
	auto& GetFrame() { return framebuffer; }                                   // the app keeps its own Image and uploads framebuffer.data() (Renderer.hpp:477)
	b2r_ctx* context() { return ctx_; }

private:
	b2r_ctx* ctx_ = nullptr;
	void upload_camera() {  // Camera.hpp:61-88: view.pos, view.orient, projection.half_width / half_height / z, exp
		const auto& c = scene.camera;
		const float pos[3] = {c.view.pos.x, c.view.pos.y, c.view.pos.z}, q[4] = {c.view.orient.w, c.view.orient.x, c.view.orient.y, c.view.orient.z};
		check(b2r_set_camera(ctx_, pos, q, c.projection.half_width, c.projection.half_height, c.projection.z, c.exp));
	}
	static void check(int rc) { if (rc != B2R_OK) throw std::runtime_error(std::string("libb2r: ") + b2r_last_error()); }
};

}  // namespace b2r
