// tests/refbinding — include/b2r_reference_binding.hpp compiled against the REFERENCE'S OWN headers (through the temporary include
// tree of oracle/ref_renderer_build.sh) and run next to the reference's own Renderer<>: ONE reference `Scene` object, built the way
// Application.cpp:35-101,233-234 builds it, is rendered by `Renderer<>` on the host cores and by `b2r::ReferenceRenderer` on the GPU;
// the program prints the divergent-pixel fraction of the bucket-free comparison available through the public surface: the
// tonemapped framebuffer after 5, and after `spp`, samples. TEST INFRASTRUCTURE ONLY (built only where /root/reference exists).
#include <math.h>
#include <stdlib.h>
#include <cfloat>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cassert>
#include <algorithm>
#include <array>
#include <bit>
#include <format>
#include <limits>
#include <memory>
#include <memory_resource>
#include <new>
#include <numeric>
#include <ranges>
#include <span>
#include <vector>
#include <immintrin.h>
void* operator new(std::size_t n) { void* p = aligned_alloc(32, (n + 31) & ~static_cast<std::size_t>(31)); if (!p) throw std::bad_alloc(); return p; }
void* operator new[](std::size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { free(p); }
void operator delete[](void* p) noexcept { free(p); }
void operator delete(void* p, std::size_t) noexcept { free(p); }
void operator delete[](void* p, std::size_t) noexcept { free(p); }
namespace std { inline constexpr unsigned long long min(unsigned long a, unsigned long long b) { return a < b ? a : b; } }
#include "Renderer.hpp"                     // the reference's (and through it Scene.hpp, Camera.hpp, BVH.hpp, ...)
#include "b2r_reference_binding.hpp"

static void default_scene(Scene& scene, uint32_t w, uint32_t h) {  // a few spheres in the spirit of Scenes::Default; values are this test's own
	using glm::vec3;
	scene.camera = Camera{{-0.2f, 0.3f, 1.0f}, {0.1f, -0.4f, -1.0f}, w, h, 40.0f, 1.0f, 16.0f, 1.0f};
	auto mat = [&](vec3 albedo, vec3 emission) { Material m{}; m.albedo = albedo; m.emission = emission; scene.material.push_back(m); return static_cast<int32_t>(scene.material.size() - 1); };
	scene.geometry.push_back(Sphere{vec3{0.0f, -100.5f, 0.0f}, 100.0f * 100.0f, mat(vec3{0.7f, 0.7f, 0.7f}, vec3{0.0f, 0.0f, 0.0f})});
	scene.geometry.push_back(Sphere{vec3{0.0f, 0.0f, 0.0f}, 0.5f * 0.5f, mat(vec3{0.8f, 0.3f, 0.3f}, vec3{0.0f, 0.0f, 0.0f})});
	scene.geometry.push_back(Sphere{vec3{1.0f, 0.1f, -0.4f}, 0.6f * 0.6f, mat(vec3{0.3f, 0.8f, 0.4f}, vec3{0.0f, 0.0f, 0.0f})});
	scene.geometry.push_back(Sphere{vec3{-1.1f, 0.0f, -0.2f}, 0.5f * 0.5f, mat(vec3{0.3f, 0.4f, 0.9f}, vec3{0.0f, 0.0f, 0.0f})});
	scene.geometry.push_back(Sphere{vec3{0.3f, 2.5f, 0.5f}, 0.4f * 0.4f, mat(vec3{1.0f, 1.0f, 1.0f}, vec3{30.0f, 28.0f, 25.0f})});
	scene.geometry.push_back(Sphere{vec3{-1.5f, 1.2f, 1.0f}, 0.2f * 0.2f, mat(vec3{1.0f, 1.0f, 1.0f}, vec3{10.0f, 20.0f, 40.0f})});
	scene.sky.ambient_color = vec3{0.0f, 0.0f, 0.0f};
	scene.acceleration_structure = decltype(scene.acceleration_structure){scene.geometry};                   // Application.cpp:233
	scene.lighting_acceleration = decltype(scene.lighting_acceleration){scene.geometry, scene.material};    // Application.cpp:234
}

int main(int argc, char** argv) {
	const uint32_t w = 320, h = 192; const uint32_t spp = argc > 1 ? static_cast<uint32_t>(atoi(argv[1])) : 200;
	const bool exact = argc > 2 && !strcmp(argv[2], "exact");  // B2R_FLAG_REFERENCE_EXACT: every framebuffer must then be bit-identical
	Scene scene; default_scene(scene, w, h);
	Renderer<> cpu{scene};                                   // the reference
	b2r::ReferencePolicy policy; if (exact) policy.flags = B2R_FLAG_REFERENCE_EXACT;
	b2r::ReferenceRenderer<Scene, glm::vec4> gpu{scene, policy};     // the drop-in, same Scene object
	static_assert(decltype(cpu)::RequiredTiling() == decltype(gpu)::RequiredTiling());
	cpu.Resize(w, h); gpu.Resize(w, h);
	bool all_identical = true;
	auto compare = [&](const char* what) {
		double se = 0; size_t divergent = 0, identical = 0; const size_t n = static_cast<size_t>(w) * h;
		for (size_t i = 0; i < n; i++) {
			const float* a = reinterpret_cast<const float*>(&cpu.framebuffer[i]); const float* b = reinterpret_cast<const float*>(&gpu.framebuffer[i]);
			bool div = false, same = true;
			for (int c = 0; c < 3; c++) { const double d = double(a[c]) - b[c]; se += d * d; div |= fabs(d) > 1e-4 * fmax(fabs(a[c]), 1e-6); same &= std::bit_cast<uint32_t>(a[c]) == std::bit_cast<uint32_t>(b[c]); }
			divergent += div; identical += same;
		}
		printf("%s: accumulations %u/%u, tonemapped pixels more than 1e-4 (relative) apart %.3e, bit-identical pixels %.4f, RMSE %.3e\n", what, cpu.accumulations, gpu.accumulations,
		       double(divergent) / n, double(identical) / n, sqrt(se / (3.0 * n)));
		all_identical &= identical == n;
		return sqrt(se / (3.0 * n));
	};
	for (uint32_t i = 0; i < 5; i++) { cpu.Accumulate(); gpu.Accumulate(); }
	cpu.Render(); gpu.Render();
	compare("after 5 samples");
	for (uint32_t i = 5; i < spp; i++) { cpu.Accumulate(); gpu.Accumulate(); cpu.Render(); gpu.Render(); }  // Render() acts every 5th sample, as in the app's frame loop
	const double rmse = compare("converged");
	// a camera move, as Application.cpp:247,299 does it
	scene.camera.TranslateLocal({0.1f, 0.0f, -0.2f}); cpu.ResetAccumulator(); gpu.ResetAccumulator();
	for (uint32_t i = 0; i < 5; i++) { cpu.Accumulate(); gpu.Accumulate(); }
	cpu.Render(); gpu.Render();
	const double rmse_moved = compare("after a camera move + 5 samples");
	// a geometry drag, as Application.cpp:508-510 does it: the app rebuilds the BVH (new leaf order) and the light list, then resets. The
	// reference's Renderer<> reads the Scene live; the binding refits the GPU's traversal tree instead of rebuilding it (SceneMoved).
	scene.geometry[1].position += glm::vec3{0.35f, 0.2f, 0.3f}; scene.geometry[3].position += glm::vec3{0.2f, 0.6f, -0.5f}; scene.geometry[2].radius_sq = 0.45f * 0.45f;
	scene.geometry[4].position += glm::vec3{-0.4f, 0.0f, 0.2f};   // a light moves too
	scene.acceleration_structure = decltype(scene.acceleration_structure){scene.geometry};
	scene.lighting_acceleration = decltype(scene.lighting_acceleration){scene.geometry, scene.material};
	cpu.ResetAccumulator();
	const float quality = gpu.SceneMoved(1e9f); gpu.ResetAccumulator();   // (threshold out of reach: this run must take the refit path)
	for (uint32_t i = 0; i < 10; i++) { cpu.Accumulate(); gpu.Accumulate(); }
	cpu.Render(); gpu.Render();
	printf("geometry drag: traversal tree refitted on the GPU, quality ratio %.3f\n", quality);
	const double rmse_edit = compare("after a geometry drag (GPU refit) + 10 samples");
	const bool ok = cpu.accumulations == gpu.accumulations && rmse < 1e-3 && rmse_moved < 2e-2 && rmse_edit < 2e-2 && (!exact || all_identical);
	if (exact) printf("reference-exact mode: every compared framebuffer bit-identical: %s\n", all_identical ? "yes" : "NO");
	printf("%s\n", ok ? "REFBINDING OK" : "REFBINDING FAILED");
	return ok ? 0 : 1;
}
