// C wrappers around the reference's OWN BVH builder and sphere-intersection loops, compiled verbatim from /root/reference (never
// copied into this repo). BVH.hpp as a whole needs MSVC (deducing-this at :88, `typename const` at :237/:310), so oracle/Makefile
// `ref` cuts these line ranges out at build time into a temporary directory:
//   bvh_class_head.inc      = BVH.hpp:17-87   (struct BoundingVolumeHierarchy {, Node with half_area / largest_axis / centroid, SplitHeuristic, members)
//   bvh_ctor_body.inc       = BVH.hpp:91-206  (the constructor's body: per-axis centroid sort, SAH sweeps, partition, recursion, leaf reorder)
//                             Two declarator lines are NOT taken from the reference: `template<typename Primitive>` (:16 — as a template,
//                             `Node::Vector` without `typename` at :81,:87 is MSVC-only syntax; Primitive is an alias of Sphere instead) and
//                             the constructor's signature (:90 — its default argument `= SplitHeuristic{}` is ill-formed inside a
//                             non-template class); they are written out below. No arithmetic is on those lines.
//   bvh_intersect_body.inc  = BVH.hpp:239-287  (body of intersect_prims: the AVX2+FMA block of 8 and the scalar tail)
//   bvh_shadow_body.inc     = BVH.hpp:292-304  (body of intersect_prims_shadow, 292-304 inside its braces)
// Primitives.hpp and DataStructures.hpp are included whole; glm comes from ref_shim/. Two platform notes, neither touching the
// reference's arithmetic: (1) <math.h> first, for MSVC's global float overloads of sqrt/abs (see ref_sampling_wrap.cpp);
// (2) `std::min(i - begin, 32ull)` (BVH.hpp:152) only compiles where size_t is unsigned long long (LLP64); the overload below gives
// LP64 the same call. Used by tests/test_oracle_ref_bvh.py and tests/gen_golden.py (-> tests/golden/bvh_kat.json).
#include <math.h>
#include <stdlib.h>
#include <cfloat>
#include <cstdint>
#include <cstring>
#include <cassert>
#include <algorithm>
#include <array>
#include <bit>
#include <limits>
#include <memory>
#include <memory_resource>
#include <numeric>
#include <ranges>
#include <span>
#include <vector>
#include <immintrin.h>
namespace std { inline constexpr unsigned long long min(unsigned long a, unsigned long long b) { return a < b ? a : b; } }
#define __vectorcall
#include "Core.hpp"
#include "vm_scalar.inc"       // VectorMath.hpp:581-662: Primitives.hpp's Sky calls fast_atan2 / fast_asin
#include "Primitives.hpp"
#include "DataStructures.hpp"
using Primitive = Sphere;
#include "bvh_class_head.inc"
	BoundingVolumeHierarchy() {}
	BoundingVolumeHierarchy(std::span<const Primitive> primitives, SplitHeuristic heuristic);
};
BoundingVolumeHierarchy::BoundingVolumeHierarchy(std::span<const Primitive> primitives, SplitHeuristic heuristic) {
#include "bvh_ctor_body.inc"


template <size_t N> struct alignas(32) RefSoA3 { float x[N], y[N], z[N]; };
template <size_t N> struct alignas(32) RefRaysIn { RefSoA3<N> p, dir; };
template <size_t N> struct alignas(32) RefHitOut { float tfar[N]; int32_t primID[N]; };
template <size_t N> struct RefOccluded { uint8_t bit[N]; void set(size_t i) { bit[i] = 1; } };
template <size_t N> struct alignas(32) RefShadowIn { RefSoA3<N> p, dir; float tfar[N]; RefOccluded<N> occluded; };
template <size_t N>
static void ref_intersect_prims(const std::vector<Sphere>& prims, const RefRaysIn<N>& in, RefHitOut<N>& out, size_t begin_ray, size_t end_ray, size_t begin_prim, size_t end_prim) {
#include "bvh_intersect_body.inc"
}
template <size_t N>
static void ref_intersect_prims_shadow(const std::vector<Sphere>& prims, RefShadowIn<N>& in, size_t begin_ray, size_t end_ray, size_t begin_prim, size_t end_prim) {
#include "bvh_shadow_body.inc"
}

extern "C" {
// geometry: n records of {float pos[3]; float radius_sq; int32 material_ID; pad} (32 bytes, = sizeof(Sphere)); nodes_out: 2n-1 x 32 bytes
uint32_t ref_bvh_build(const void* geometry, uint32_t n, void* nodes_out, void* prims_out) {
	static_assert(sizeof(Sphere) == 32 && sizeof(BoundingVolumeHierarchy::Node) == 32, "record sizes");
	BoundingVolumeHierarchy b{std::span<const Sphere>(static_cast<const Sphere*>(geometry), n), BoundingVolumeHierarchy::SplitHeuristic{}};
	std::memcpy(nodes_out, b.nodes.data(), b.nodes.size() * 32);
	std::memcpy(prims_out, b.prims.data(), b.prims.size() * 32);
	return static_cast<uint32_t>(b.nodes.size());
}
// the same constructor with a caller-chosen SplitHeuristic (BVH.hpp:70-83)
uint32_t ref_bvh_build_h(const void* geometry, uint32_t n, uint32_t log_cluster_size, float cost_ratio, void* nodes_out, void* prims_out) {
	BoundingVolumeHierarchy::SplitHeuristic h; h.log_cluster_size = log_cluster_size; h.cost_ratio = cost_ratio;
	BoundingVolumeHierarchy b{std::span<const Sphere>(static_cast<const Sphere*>(geometry), n), h};
	std::memcpy(nodes_out, b.nodes.data(), b.nodes.size() * 32);
	std::memcpy(prims_out, b.prims.data(), b.prims.size() * 32);
	return static_cast<uint32_t>(b.nodes.size());
}
float ref_node_half_area(const float lo[3], const float hi[3]) { BoundingVolumeHierarchy::Node nd(glm::vec3{lo[0], lo[1], lo[2]}, glm::vec3{hi[0], hi[1], hi[2]}); return nd.half_area(); }
// rays: n x 6 floats (origin, dir), n <= 256; all spheres against all rays as the reference does (begin/end = whole ranges).
// n rays: the first n & ~7 go through the AVX2 block, the rest through the scalar tail, exactly as in the reference.
void ref_intersect_closest(const void* spheres, uint32_t n_spheres, const float* rays, uint32_t n, float* tfar_out, int32_t* prim_out) {
	std::vector<Sphere> prims(static_cast<const Sphere*>(spheres), static_cast<const Sphere*>(spheres) + n_spheres);
	static RefRaysIn<256> in; static RefHitOut<256> out;
	for (uint32_t i = 0; i < n; i++) { in.p.x[i] = rays[6 * i]; in.p.y[i] = rays[6 * i + 1]; in.p.z[i] = rays[6 * i + 2]; in.dir.x[i] = rays[6 * i + 3]; in.dir.y[i] = rays[6 * i + 4]; in.dir.z[i] = rays[6 * i + 5]; out.tfar[i] = FLT_MAX; out.primID[i] = -1; }
	ref_intersect_prims<256>(prims, in, out, 0, n, 0, prims.size());
	for (uint32_t i = 0; i < n; i++) { tfar_out[i] = out.tfar[i]; prim_out[i] = out.primID[i]; }
}
void ref_intersect_shadow(const void* spheres, uint32_t n_spheres, const float* rays, const float* tfar, uint32_t n, uint8_t* occluded_out) {
	std::vector<Sphere> prims(static_cast<const Sphere*>(spheres), static_cast<const Sphere*>(spheres) + n_spheres);
	static RefShadowIn<256> in;
	for (uint32_t i = 0; i < n; i++) { in.p.x[i] = rays[6 * i]; in.p.y[i] = rays[6 * i + 1]; in.p.z[i] = rays[6 * i + 2]; in.dir.x[i] = rays[6 * i + 3]; in.dir.y[i] = rays[6 * i + 4]; in.dir.z[i] = rays[6 * i + 5]; in.tfar[i] = tfar[i]; in.occluded.bit[i] = 0; }
	ref_intersect_prims_shadow<256>(prims, in, 0, n, 0, prims.size());
	for (uint32_t i = 0; i < n; i++) occluded_out[i] = in.occluded.bit[i];
}
}
