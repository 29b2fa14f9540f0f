// C wrappers around the reference's OWN Renderer<>: Renderer.hpp (Accumulate, Render) with DataStreams.hpp, BVH.hpp, Scene.hpp,
// Camera.hpp, Sampling.hpp, Primitives.hpp, ... compiled by g++ from /root/reference — see ref_renderer_build.sh for the temporary
// include tree, the stand-ins (ppl.h, Image.h, glm, VCL, VectorMath.hpp's / Color.hpp's SIMD parts) and the token-level syntax edits.
// TEST INFRASTRUCTURE ONLY: tests/test_oracle_ref_renderer.py runs it next to the oracle (bucket sums and tonemapped frames must be
// identical bit for bit) and tests/gen_golden.py stores its outputs in tests/golden/renderer_kat.json.
//
// Platform notes (none changes an expression of the reference):
//  * <math.h>/<stdlib.h> first: MSVC's <cmath> has the float overloads of sqrt/abs in the global namespace (ref_sampling_wrap.cpp);
//  * std::min(size_t, 32ull) (BVH.hpp:152) needs LLP64; the overload below gives LP64 the same call;
//  * the reference issues 32-byte ALIGNED AVX loads/stores on std::vector storage (accumulator, framebuffer: Renderer.hpp:447-474).
//    MSVC's allocator aligns every block >= 4 KiB to 32 bytes; glibc's aligns to 16. operator new below (bound inside this library
//    only, -Bsymbolic) returns 32-byte aligned blocks;
//  * FP contraction is off (-ffp-contract=off), as in the oracle: FMA only where the reference writes FMA intrinsics;
//  * max_bounces is a template argument of Renderer<> (RendererPolicy): one instantiation per value used by the tests.
#include <math.h>
#include <stdlib.h>
#include <cfloat>
#include <climits>
#include <cstdint>
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
#include "Renderer.hpp"

namespace {
struct RefBase {
	Scene scene;
	virtual ~RefBase() {}
	virtual void resize(uint32_t w, uint32_t h) = 0;
	virtual void accumulate() = 0;
	virtual void render() = 0;
	virtual uint32_t accumulations() const = 0;
	virtual void set_accumulations(uint32_t a) = 0;
	virtual void read_buckets(float* out) const = 0;   // [5][3][npix], pixel index = tile * 256 + ID (the oracle's layout)
	virtual const float* framebuffer() const = 0;
};
template <size_t MB> struct RefImpl : RefBase {
	using R = Renderer<RendererPolicy{4, 1, 64, MB, 1e2f}>;
	R renderer;
	RefImpl() : renderer(scene) {}
	void resize(uint32_t w, uint32_t h) override { renderer.Resize(w, h); }
	void accumulate() override { renderer.Accumulate(); }
	void render() override { renderer.Render(); }
	uint32_t accumulations() const override { return renderer.accumulations; }
	void set_accumulations(uint32_t a) override { renderer.accumulations = a; }
	void read_buckets(float* out) const override {
		const size_t tiles = renderer.accumulator.size(), npix = tiles * R::TileSize;
		for (size_t t = 0; t < tiles; t++) for (size_t k = 0; k < R::AccumulationBuckets; k++) {
			const auto& c = renderer.accumulator[t].color[k];
			std::memcpy(out + (k * 3 + 0) * npix + t * R::TileSize, c.r, sizeof c.r);
			std::memcpy(out + (k * 3 + 1) * npix + t * R::TileSize, c.g, sizeof c.g);
			std::memcpy(out + (k * 3 + 2) * npix + t * R::TileSize, c.b, sizeof c.b);
		}
	}
	const float* framebuffer() const override { return reinterpret_cast<const float*>(renderer.framebuffer.data()); }
};
}  // namespace

extern "C" {
// geometry: n x 32-byte Sphere records, materials: n_mat x 96-byte Material records (the reference's own layouts). hdri: RGBA32F or null.
// Scene set-up as Application.cpp:225-234: camera, sky, then the BVH and the light list from the geometry.
void* ref_renderer_create(const void* geometry, uint32_t n, const void* materials, uint32_t n_mat, const float eye[3], const float dir[3],
                          float focal_length, float exposure, const float ambient[3], const float* hdri, int32_t hdri_w, int32_t hdri_h,
                          uint32_t width, uint32_t height, uint32_t max_bounces) {
	static_assert(sizeof(Sphere) == 32 && sizeof(Material) == 96, "record sizes");
	RefBase* r = nullptr;
	switch (max_bounces) {
	case 1: r = new RefImpl<1>(); break;
	case 2: r = new RefImpl<2>(); break;
	case 4: r = new RefImpl<4>(); break;
	case 8: r = new RefImpl<8>(); break;
	case 16: r = new RefImpl<16>(); break;
	default: return nullptr;
	}
	Scene& sc = r->scene;
	sc.geometry.assign(static_cast<const Sphere*>(geometry), static_cast<const Sphere*>(geometry) + n);
	sc.material.assign(static_cast<const Material*>(materials), static_cast<const Material*>(materials) + n_mat);
	sc.camera = Camera{glm::vec3{eye[0], eye[1], eye[2]}, glm::vec3{dir[0], dir[1], dir[2]}, width, height, focal_length, 1.0f, 16.0f, exposure};
	sc.sky.ambient_color = glm::vec3{ambient[0], ambient[1], ambient[2]};
	if (hdri) {
		sc.sky.hdri_data = const_cast<float*>(hdri); sc.sky.hdri_width = hdri_w; sc.sky.hdri_height = hdri_h; sc.sky.hdri_channels = 4;
		sc.sky.hdri_fwidth = static_cast<float>(hdri_w - 1); sc.sky.hdri_fheight = static_cast<float>(hdri_h - 1);
	}
	sc.acceleration_structure = decltype(sc.acceleration_structure){sc.geometry};
	sc.lighting_acceleration = decltype(sc.lighting_acceleration){sc.geometry, sc.material};
	r->resize(width, height);
	return r;
}
void ref_renderer_destroy(void* h) { delete static_cast<RefBase*>(h); }
void ref_renderer_accumulate(void* h, uint32_t n) { for (uint32_t i = 0; i < n; i++) static_cast<RefBase*>(h)->accumulate(); }
void ref_renderer_set_accumulations(void* h, uint32_t a) { static_cast<RefBase*>(h)->set_accumulations(a); }
uint32_t ref_renderer_accumulations(void* h) { return static_cast<RefBase*>(h)->accumulations(); }
void ref_renderer_read_buckets(void* h, float* out) { static_cast<RefBase*>(h)->read_buckets(out); }
// Renderer::Render (acts only when accumulations % 5 == 0, Q21); copies the RGBA32F framebuffer (raster order, row 0 = y 0)
int ref_renderer_render(void* h, float* rgba_out, uint32_t n_floats) {
	RefBase* r = static_cast<RefBase*>(h);
	if (r->accumulations() % 5) return 0;
	r->render();
	std::memcpy(rgba_out, r->framebuffer(), static_cast<size_t>(n_floats) * sizeof(float));
	return 1;
}
uint32_t ref_renderer_light_count(void* h) { return static_cast<uint32_t>(static_cast<RefBase*>(h)->scene.lighting_acceleration.prims.size()); }
}
