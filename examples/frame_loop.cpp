// frame_loop.cpp — the reference's per-frame driver (RaytracingApp::UIRender, Application.cpp:361-406) against the drop-in
// Renderer: build Scenes::Default (Application.cpp:33-101), pad the viewport to the tiling, then Accumulate -> Render each frame
// and print the same read-out ("[W X H] : ms : fps : Msamples/s", Application.cpp:400-403).
//   g++ -std=c++20 -O2 examples/frame_loop.cpp -I cpu-raytracing-experiments_b200/host -L cpu-raytracing-experiments_b200 -lb2r -Wl,-rpath,$PWD/cpu-raytracing-experiments_b200 -o frame_loop
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <memory>
#include "Renderer.hpp"

using b2r_host::vec3;

static void default_scene(Scene& scene) {  // Application.cpp:33-101
	scene.camera = Camera{{-0.2, 0.3, 1}, {0.1, -0.4, -1}, 1, 1, 40.0, 0.0f, 16.0f, 1.0f};
	auto add = [&](vec3 p, float r2, auto&& edit) { scene.material.push_back(Material{}); edit(scene.material.back()); scene.geometry.push_back(Sphere{p, r2, (int32_t)scene.material.size() - 1}); };
	add({0.3, -1.47, 0.0}, 1.5f * 1.5f, [](Material& m) { m.albedo = vec3{1.0f}; m.F0 = vec3{0.8f}; m.F80 = vec3{0.9f}; m.roughness = 0.2f; });
	add({0.29999, 0.0801, 0.0}, 0.05f * 0.05f, [](Material& m) { m.emission = 0.1f * vec3{25.0, 25.0, 200.0}; m.albedo = vec3{1.0f}; m.roughness = 1.0f; });
	add({0.3302, 0.36165, 0.7119}, 0.05f * 0.05f, [](Material& m) { m.emission = 0.1f * vec3{150.0, 150.0, 150.0}; m.albedo = vec3{1.0f}; m.roughness = 1.0f; });
	add({-0.4857, -0.0242, -0.41383}, 0.05f * 0.05f, [](Material& m) { m.emission = vec3{200.0, 17.0, 25.0}; m.albedo = vec3{1.0f}; m.roughness = 1.0f; });
	add({0.3, 1.7, 0.0}, 1.5f * 1.5f, [](Material& m) { m.albedo = vec3{0.793, 0.793, 0.664}; m.F0 = vec3{0.04f}; m.F80 = vec3{0.5f}; m.roughness = 0.85f; });
	add({0.018, 0.022f, 0.07}, 0.02f * 0.02f, [](Material& m) { m.albedo = vec3{0.05f}; m.F0 = vec3{0.03f}; m.F80 = vec3{0.5f}; m.transmission = vec3{0.95, 0.95, 0.95}; m.IOR_minus_one = 0.44f; m.roughness = 0.05; });
	add({-0.037, 0.022f, 0.00}, 0.03f * 0.03f, [](Material& m) { m.albedo = vec3{1.0f}; m.F0 = vec3{0.944, 0.776, 0.373}; m.F80 = vec3{0.8, 0.8, 0.6}; m.roughness = 0.15; });
	add({-0.0846, -0.0334, 0.283}, 0.012f * 0.012f, [](Material& m) { m.albedo = vec3{1.0f}; m.F0 = vec3{0.076288, 0.077375, 0.078887}; m.F80 = vec3{0.47990, 0.48028, 0.48080}; m.transmission = vec3{0.670, 0.764, 0.855}; m.IOR_minus_one = 0.762f; m.roughness = 0.1; });
	add({0.03863, -0.00788, 0.2835}, 0.012f * 0.012f, [](Material& m) { m.albedo = vec3{1.0f}; m.F0 = vec3{0.04f}; m.F80 = vec3{0.5f}; m.roughness = 0.8f; });
	scene.sky.ambient_color = vec3{0.0f, 0.0f, 0.0f};
	scene.acceleration_structure = decltype(scene.acceleration_structure){scene.geometry};          // Application.cpp:233
	scene.lighting_acceleration = decltype(scene.lighting_acceleration){scene.geometry, scene.material};  // :234
}

int main(int argc, char** argv) {
	uint32_t viewport_width = argc > 1 ? atoi(argv[1]) : 1920, viewport_height = argc > 2 ? atoi(argv[2]) : 1080, frames = argc > 3 ? atoi(argv[3]) : 65;
	Scene scene; default_scene(scene);
	Renderer renderer(scene);
	const uint32_t tiling = static_cast<uint32_t>(Renderer::RequiredTiling());
	viewport_width = (viewport_width + tiling - 1) / tiling * tiling; viewport_height = (viewport_height + tiling - 1) / tiling * tiling;  // Application.cpp:367-372
	scene.camera.Resize(viewport_width, viewport_height);
	renderer.Resize(viewport_width, viewport_height);
	auto t0 = std::chrono::steady_clock::now();
	for (uint32_t f = 0; f < frames; f++) { renderer.Accumulate(); renderer.Render(); }  // Application.cpp:379-380
	const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / frames;
	// focus picking exactly as Application.cpp:282-298 writes it: a hand-built RayStream<8>, one ray, BVH::Traverse<8>
	{
		auto depth_ray = std::make_unique<RayStream<8>>();
		depth_ray->hit.matID[0] = -1; depth_ray->hit.primID[0] = -1; depth_ray->hit.tfar[0] = FLT_MAX;
		auto* raygen_buffer = depth_ray->path.input;
		static constexpr float no_pixel_jitter[2]{0.5f, 0.5f};
		const auto [orig, dir] = scene.camera.generate_ray(static_cast<int32_t>(viewport_width / 2), static_cast<int32_t>(viewport_height / 2), no_pixel_jitter);  // Application.cpp:287-288
		raygen_buffer->dir.x[0] = dir.x; raygen_buffer->dir.y[0] = dir.y; raygen_buffer->dir.z[0] = dir.z;
		raygen_buffer->p.x[0] = orig.x; raygen_buffer->p.y[0] = orig.y; raygen_buffer->p.z[0] = orig.z;
		scene.acceleration_structure.Traverse<8>(*depth_ray->path.input, depth_ray->hit, 1);
		float t2; int32_t p2; const float r6[6] = {orig.x, orig.y, orig.z, dir.x, dir.y, dir.z};
		renderer.Traverse(r6, 1, &t2, &p2);
		std::printf("focus pick: prim %d mat %d depth %.6f (renderer path: prim %d depth %.6f)\n", depth_ray->hit.primID[0], depth_ray->hit.matID[0], depth_ray->hit.tfar[0], p2, t2);
	}
	double sum = 0; for (auto& p : renderer.framebuffer) sum += p.x + p.y + p.z;
	std::printf("[%u X %u] : %.3fms : %.1ffps : %.1fMsamples/s  (mean tonemapped value %.5f after %u accumulations)\n", viewport_width, viewport_height, ms, 1000.0 / ms,
	            viewport_width * viewport_height * 1e-3 / ms, sum / (3.0 * renderer.framebuffer.size()), renderer.accumulations);
	// a geometry drag as the editor does it (Application.cpp:508-510): rebuild the BVH and the light list, reset — and tell the renderer the
	// spheres only moved, so that the GPU refits its traversal tree instead of rebuilding it
	scene.geometry[1].position.x += 0.25f; scene.geometry[1].position.y += 0.125f;
	scene.acceleration_structure = decltype(scene.acceleration_structure){scene.geometry};
	scene.lighting_acceleration = decltype(scene.lighting_acceleration){scene.geometry, scene.material};
	const float quality = renderer.SceneMoved(); renderer.ResetAccumulator();
	for (uint32_t f = 0; f < 5; f++) { renderer.Accumulate(); renderer.Render(); }
	sum = 0; for (auto& p : renderer.framebuffer) sum += p.x + p.y + p.z;
	std::printf("after a geometry drag: tree quality ratio %.4f, mean tonemapped value %.5f after %u accumulations\n", quality, sum / (3.0 * renderer.framebuffer.size()), renderer.accumulations);
	// an edit that adds a sphere: same two rebuild lines, same call — SceneMoved() notices that the count changed and uploads the scene anew
	scene.material.push_back(Material{}); scene.material.back().albedo = vec3{0.2f, 0.6f, 0.3f};
	scene.geometry.push_back(Sphere{vec3{0.125f, 0.0625f, 0.375f}, 0.0625f * 0.0625f, (int32_t)scene.material.size() - 1});
	scene.acceleration_structure = decltype(scene.acceleration_structure){scene.geometry};
	scene.lighting_acceleration = decltype(scene.lighting_acceleration){scene.geometry, scene.material};
	const float quality2 = renderer.SceneMoved(); renderer.ResetAccumulator();
	for (uint32_t f = 0; f < 5; f++) { renderer.Accumulate(); renderer.Render(); }
	sum = 0; for (auto& p : renderer.framebuffer) sum += p.x + p.y + p.z;
	std::printf("after adding a sphere: %zu spheres, tree quality ratio %.4f, mean tonemapped value %.5f after %u accumulations\n", scene.geometry.size(), quality2, sum / (3.0 * renderer.framebuffer.size()), renderer.accumulations);
	return 0;
}
