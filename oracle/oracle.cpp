// oracle.cpp — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the reference's hot path: Renderer::Accumulate (Renderer.hpp:73-434),
// Renderer::Render (Renderer.hpp:436-478), the BVH builder (BVH.hpp:90-206), the brute-force and
// stream-BVH intersection routines (BVH.hpp:219-404) and the stream machinery (DataStreams.hpp).
// PARITY IS PINNED against the reference itself. The reference is MSVC-only C++ with un-vendored dependencies (glm, Agner Fog VCL,
// PPL, Vulkan ...), but its renderer is header-only, and with stand-ins for those dependencies (oracle/ref_shim/:
// component-wise glm/VCL definitions, a thread-pool parallel_for, an Image stub) and six token-level syntax edits made on a
// temporary copy at build time, g++ compiles it from /root/reference (oracle/Makefile `ref`, oracle/ref_renderer_build.sh; nothing
// is copied into this repo, outputs go to the git-ignored oracle/_ref/):
//   * _ref/librefrenderer.so — Renderer<>::Accumulate / Render with DataStreams.hpp, BVH.hpp, Scene.hpp, Camera.hpp, Sampling.hpp,
//     Primitives.hpp, Random.hpp ...: THE REFERENCE, runnable. This file in its slot-exact mode (ORC_SLOT_EXACT) reproduces its bucket
//     sums and tonemapped frames bit for bit (tests/test_oracle_ref_renderer.py, tests/golden/renderer_kat.json);
//   * _ref/librefbvh.so — the BVH constructor and the three sphere loops (bit-identical node arrays and leaf order up to 100k spheres,
//     tests/test_oracle_ref_bvh.py); _ref/librefsampling.so — Sampling.hpp, the scalar fast math, ACES tonemapping, Camera
//     (tests/test_oracle_ref_sampling.py); _ref/librefrng.so — Random.hpp / Bitmanip.hpp (tests/test_oracle_rng.py).
// What that does not pin: the author's MSVC binary (its /fp mode and std::sort tie order are unknown; see the canonical choices
// below) and the stream-BVH traversal (`USEBVH`, disabled in the reference: BVH.hpp:320-358 is restated here, checked against
// brute force).
//
// Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may load this
// library. The product never links it.
//
// Canonical choices where the reference is compiler- or slot-dependent (DESIGN.md "Numerics"):
//   * no FP contraction; FMA exactly where BVH.hpp:252-260 has fmadd/fnmadd intrinsics;
//   * closest-hit sphere test: the SIMD-FMA formula for every ray (flag ORC_SLOT_EXACT switches to the
//     reference's slot-dependent mix of SIMD blocks of 8 + scalar tail, Q7/Q16);
//   * per-axis centroid sort is stable (ties keep the lower original index) — what MSVC's std::sort
//     (insertion sort for N<=32) does for the default scene; implementation-defined beyond that (Q19).
#include "oracle_math.hpp"

#include <algorithm>
#include <atomic>
#include <memory>
#include <numeric>
#include <thread>
#include <vector>

using namespace orc;

namespace {

// ---------------------------------------------------------------- PODs, layout-identical to the reference
struct Sphere {  // Primitives.hpp:7-17 (alignas(16), 32 B)
	float px, py, pz; float radius_sq; int32_t material_ID; int32_t pad[3];
};
static_assert(sizeof(Sphere) == 32);
struct Material {  // Primitives.hpp:18-27 (alignas(32), 96 B)
	float albedo[3], F0[3], F80[3], emission[3], transmission[3]; float roughness, IOR_minus_one; float pad[7];
};
static_assert(sizeof(Material) == 96);
struct Node {  // BVH.hpp:18-27 (32 B)
	float min_bound[3]; uint32_t first_id; float max_bound[3]; uint32_t prim_count;
};
static_assert(sizeof(Node) == 32);

constexpr int kTileRoot = 16;   // Renderer.hpp:20,32 (log_tile = 4)
constexpr int kTileSize = 256;  // Renderer.hpp:33-34 (StreamSize == TileSize, 1 spp per call)

enum Flags : uint32_t {
	ORC_BVH = 1u,         // USEBVH true: stream-BVH traversal of BVH.hpp:320-358 / 368-403 (reference ships false)
	ORC_SLOT_EXACT = 2u,  // closest-hit: SIMD blocks of 8 + scalar tail exactly as BVH.hpp:250-286 (default: FMA formula everywhere)
	ORC_NO_MIS = 4u,      // this repo's "MIS off" definition: no NEE, radiance += throughput*emission
	ORC_GGX = 8u,         // the reference's `#define BRDF 1` build (Renderer.hpp:70,207-213): Closure<GGX> (DataStreams.hpp:184-219) with F0 and
	                      // alpha = roughness^2 + (1 - roughness^2) * gloss_decay_table[bounce]. The reference never declares gloss_decay_table (BRDF 1
	                      // does not compile as shipped); oracle/ref_renderer_build.sh supplies an all-zero table ("no decay") and so does this port.
	                      // Closure<GGX>::pdf is `return 0.0f; //TODO` in the reference and is restated as that.
};

// ---------------------------------------------------------------- BVH node helpers (BVH.hpp:28-67)
static inline Node node_empty() {
	Node n; for (int i = 0; i < 3; i++) { n.min_bound[i] = +FLT_MAX; n.max_bound[i] = -FLT_MAX; }
	n.first_id = 0; n.prim_count = 0; return n;
}
static inline void node_merge(Node& a, const Node& b) {  // operator|= :35-39 with glm::min/max
	for (int i = 0; i < 3; i++) { a.min_bound[i] = smin(a.min_bound[i], b.min_bound[i]); a.max_bound[i] = smax(a.max_bound[i], b.max_bound[i]); }
}
static inline size_t node_largest_axis(const Node& n) {  // :48-54
	float d[3] = {n.max_bound[0] - n.min_bound[0], n.max_bound[1] - n.min_bound[1], n.max_bound[2] - n.min_bound[2]};
	size_t ret = 0; for (size_t i = 1; i < 3; ++i) if (d[ret] < d[i]) ret = i; return ret;
}
static inline float node_half_area(const Node& n) {  // :58-67 — loop stops before d.x: returns d.y*d.z only (Q17)
	float d[3] = {n.max_bound[0] - n.min_bound[0], n.max_bound[1] - n.min_bound[1], n.max_bound[2] - n.min_bound[2]};
	float area = 0.0f; int32_t i = 2;
	for (float accum = d[i--]; i > 0; i--) { area += d[i] * accum; accum += d[i]; }
	return area;
}

struct BVH {
	std::vector<Node> nodes;
	std::vector<Sphere> prims;
	std::vector<uint32_t> prim_ids;  // prims[i] = geometry[prim_ids[i]]
};

// BVH.hpp:90-206, SplitHeuristic{} defaults (log_cluster_size 0, cost_ratio 1): :70-83
static void build_bvh(const Sphere* primitives, size_t primnum, BVH& out) {
	struct StackFrame { size_t ID, begin, count; };
	struct Split { size_t pos, axis; float cost; };
	std::vector<uint32_t> primIDs(primnum * 3);
	std::vector<Node> bboxes(primnum);
	std::vector<V3> centroids(primnum);
	std::vector<float> accum_cost(primnum);
	std::vector<uint8_t> marks(primnum);
	out.prims.assign(primnum, Sphere{});
	out.nodes.clear();
	out.nodes.reserve(2 * (primnum + 1));
	if (primnum == 0) { out.prim_ids.clear(); out.nodes.push_back(node_empty()); return; }

	for (size_t i = 0; i < primnum; i++) {  // :115-117 (Sphere::bounds, Primitives.hpp:13-16; centroid :55-57)
		float r = std::sqrt(primitives[i].radius_sq);
		Node b; b.first_id = 0; b.prim_count = 0;
		b.min_bound[0] = primitives[i].px - r; b.min_bound[1] = primitives[i].py - r; b.min_bound[2] = primitives[i].pz - r;
		b.max_bound[0] = primitives[i].px + r; b.max_bound[1] = primitives[i].py + r; b.max_bound[2] = primitives[i].pz + r;
		bboxes[i] = b;
		centroids[i] = V3{(b.max_bound[0] + b.min_bound[0]) * 0.5f, (b.max_bound[1] + b.min_bound[1]) * 0.5f, (b.max_bound[2] + b.min_bound[2]) * 0.5f};
	}
	for (size_t axis = 0; axis < 3; ++axis) {  // :118-122 (canonical: stable, see header)
		uint32_t* ids = primIDs.data() + axis * primnum;
		std::iota(ids, ids + primnum, 0u);
		std::stable_sort(ids, ids + primnum, [&](uint32_t a, uint32_t b) { return (&centroids[a].x)[axis] < (&centroids[b].x)[axis]; });
	}
	{  // :125 root = fold of all boxes
		Node root = node_empty();
		for (size_t i = 0; i < primnum; i++) node_merge(root, bboxes[i]);
		out.nodes.push_back(root);
	}
	auto reduce_bboxes = [&](size_t from, size_t to) {  // :109-113 — reads the axis-0 list
		Node res = node_empty();
		for (size_t i = from; i < to; ++i) node_merge(res, bboxes[primIDs[i]]);
		return res;
	};
	std::vector<StackFrame> stack;  // Stack<StackFrame,64> :127
	stack.push_back({0, 0, primnum});
	while (!stack.empty()) {
		StackFrame item = stack.back(); stack.pop_back();
		if (item.count <= 1) {  // :133-137
			out.nodes[item.ID].first_id = static_cast<uint32_t>(item.begin);
			out.nodes[item.ID].prim_count = static_cast<uint32_t>(item.count);
			continue;
		}
		const size_t first_child = out.nodes.size();
		out.nodes[item.ID].first_id = static_cast<uint32_t>(first_child);
		const Node node = out.nodes[item.ID];
		const size_t begin = item.begin, end = item.begin + item.count;
		// :144 non_split_cost = half_area * (prim_count(size) - cost_ratio)
		Split best{begin + (item.count + 1) / 2, node_largest_axis(node), node_half_area(node) * (static_cast<float>(item.count) - 1.0f)};
		for (size_t axis = 0; axis < 3; ++axis) {  // :146-171
			const uint32_t* ids = primIDs.data() + axis * primnum;
			size_t first_right = 0;
			Node right_bbox = node_empty();
			for (size_t i = end - 1; i > begin;) {  // Q18: the "chunked" inner loop runs all the way down to begin
				float right_cost = 0.0f;
				for (; i > i - std::min<size_t>(i - begin, 32); --i) {
					node_merge(right_bbox, bboxes[ids[i]]);
					accum_cost[i] = right_cost = node_half_area(right_bbox) * static_cast<float>(end - i);
				}
				if (right_cost > best.cost) { first_right = i; break; }
			}
			Node left_bbox = node_empty();
			for (size_t i = begin; i < end - 1; i++) {
				node_merge(left_bbox, bboxes[ids[i]]);
				if (i < first_right) break;
				float left_cost = node_half_area(left_bbox) * static_cast<float>(i + 1 - begin);
				if (left_cost > best.cost) break;
				float cost = left_cost + accum_cost[i + 1];
				if (cost < best.cost) best = Split{i + 1, axis, cost};
			}
		}
		{  // :173-184
			const uint32_t* ids = primIDs.data() + best.axis * primnum;
			for (size_t i = begin; i < best.pos; ++i) marks[ids[i]] = 1;
			for (size_t i = best.pos; i < end; ++i) marks[ids[i]] = 0;
			for (size_t axis = 0; axis < 3; ++axis) {
				if (axis == best.axis) continue;
				uint32_t* a = primIDs.data() + axis * primnum;
				std::stable_partition(a + begin, a + end, [&](uint32_t id) { return marks[id] != 0; });
			}
		}
		{  // :185-198
			const size_t rb[2] = {begin, best.pos}, re[2] = {best.pos, end};
			const Node children[2] = {reduce_bboxes(rb[0], re[0]), reduce_bboxes(rb[1], re[1])};
			size_t sort_area = static_cast<size_t>(node_half_area(children[0]) < node_half_area(children[1]));
			size_t sort_size = static_cast<size_t>(re[0] - rb[0] < re[1] - rb[1]);
			size_t combined = sort_area ^ sort_size;
			out.nodes.push_back(children[sort_area]);
			out.nodes.push_back(children[1 - sort_area]);
			stack.push_back({first_child + combined, rb[sort_size], re[sort_size] - rb[sort_size]});
			stack.push_back({first_child + (1 - combined), rb[1 - sort_size], re[1 - sort_size] - rb[1 - sort_size]});
		}
	}
	out.prim_ids.assign(primIDs.begin(), primIDs.begin() + primnum);  // :201-205 final order = axis-0 list
	for (size_t i = 0; i < primnum; i++) out.prims[i] = primitives[primIDs[i]];
}

// ---------------------------------------------------------------- streams (DataStreams.hpp:74-157)
struct Buffer {
	float px[kTileSize], py[kTileSize], pz[kTileSize], dx[kTileSize], dy[kTileSize], dz[kTileSize];
	float rr[kTileSize], rg[kTileSize], rb[kTileSize], tr[kTileSize], tg[kTileSize], tb[kTileSize];
	float pdf[kTileSize]; uint32_t pixelID[kTileSize];
};
struct Hit { float tfar[kTileSize]; int32_t primID[kTileSize]; int32_t matID[kTileSize]; };
struct ShadowStream {
	float px[kTileSize], py[kTileSize], pz[kTileSize], dx[kTileSize], dy[kTileSize], dz[kTileSize];
	float tfar[kTileSize]; float r[kTileSize], g[kTileSize], b[kTileSize];
	uint8_t occluded[kTileSize];
};
struct ShaderData {
	float Px[kTileSize], Py[kTileSize], Pz[kTileSize], Vx[kTileSize], Vy[kTileSize], Vz[kTileSize];
	float Tx[kTileSize], Ty[kTileSize], Tz[kTileSize], Tw[kTileSize];
	float albedo[kTileSize][3];  // Closure<LambertianDiffuse>::albedo (DataStreams.hpp:165-167); ORC_GGX: Closure<GGX>::F0 (:186)
	float alpha[kTileSize];      // ORC_GGX: Closure<GGX>::alpha (:187)
	uint8_t is_emissive[kTileSize];
};

struct Counters {
	std::atomic<uint64_t> extension_rays{0}, shadow_rays{0}, shaded_hits{0}, terminated{0}, sphere_tests{0}, box_tests{0}, dropped{0};
};

struct Ctx {
	uint32_t width = 0, height = 0, h_tiles = 0, v_tiles = 0;
	uint32_t max_bounces = 16, K = 5, flags = 0;
	uint32_t accumulations = 0;
	std::vector<Sphere> geometry; std::vector<Material> material;
	std::vector<int32_t> lights;  // LightingAcceleration::prims (Scene.hpp:9-17): indices into geometry
	BVH bvh;
	// camera (Camera.hpp)
	V3 cam_pos{0, 0, 0}; Quat cam_orient{1, 0, 0, 0}; float half_width = 0.5f, half_height = 0.5f, cam_z = -1.0f, exposure = 1.0f;
	// sky (Primitives.hpp:29-47)
	float ambient[3] = {0, 0, 0}; int32_t hdri_w = 0, hdri_h = 0; std::vector<float> hdri; float hdri_fw = 0, hdri_fh = 0;
	std::vector<float> accumulator;  // [tile][K][3][256] (Renderer.hpp:43-46)
	Counters counters;
};

// ---------------------------------------------------------------- sphere tests (BVH.hpp:236-305)
// SIMD lane formula, BVH.hpp:251-267. Returns true and updates (tfar, primID) when the lane's mask bit is set.
static inline void sphere_closest_fma(const Sphere& s, int32_t prim_ID, float ox, float oy, float oz, float dx, float dy, float dz, float& tfar, int32_t& primID) {
	float temp_x = s.px - ox;
	float b = dx * temp_x;
	float disc = std::fmaf(-temp_x, temp_x, s.radius_sq);
	float temp_y = s.py - oy;
	b = std::fmaf(dy, temp_y, b);
	disc = std::fmaf(-temp_y, temp_y, disc);
	float temp_z = s.pz - oz;
	b = std::fmaf(dz, temp_z, b);
	disc = std::fmaf(-temp_z, temp_z, disc);
	disc = std::fmaf(b, b, disc);
	// :261 early-out can never fire (8-bit movemask vs 0xFFFFFFFF, Q6)
	float sq = std::sqrt(disc);           // disc < 0 -> NaN (fails the ordered compare below); disc == -0 -> -0 (sign bit set)
	bool sq_sign = std::signbit(disc);    // sign bit of sqrt(disc) for every non-NaN input
	float dist = b - sq;
	if (std::signbit(dist)) dist = b + sq;           // blendv on the sign bit of (b - sqrt)
	// mask = (dist < tfar) & ~(sign(sqrt) | sign(dist)); NaN dist fails the ordered compare
	if ((dist < tfar) && !sq_sign && !std::signbit(dist)) { tfar = dist; primID = prim_ID; }
}
// scalar tail, BVH.hpp:270-286
static inline void sphere_closest_scalar(const Sphere& s, int32_t prim_ID, float ox, float oy, float oz, float dx, float dy, float dz, float& tfar, int32_t& primID) {
	float b = 0.0f; float disc = s.radius_sq;
	const float c[3] = {s.px, s.py, s.pz}, o[3] = {ox, oy, oz}, d[3] = {dx, dy, dz};
	for (int dim = 0; dim < 3; dim++) { float temp = c[dim] - o[dim]; b += d[dim] * temp; disc -= temp * temp; }
	disc += b * b;
	if (disc < 0.0f) return;
	disc = std::sqrt(disc);
	float dist = (b >= disc ? b - disc : b + disc);
	if (dist < 0.0f || dist >= tfar) return;
	tfar = dist; primID = prim_ID;
}
// shadow any-hit, BVH.hpp:294-300
static inline bool sphere_shadow(const Sphere& s, float ox, float oy, float oz, float dx, float dy, float dz, float tfar) {
	V3 P{s.px - ox, s.py - oy, s.pz - oz};
	float b = dot(V3{dx, dy, dz}, P);
	float disc = b * b - dot(P, P) + s.radius_sq;
	if (disc < 0.0f) return false;
	disc = std::sqrt(disc);
	float dist = (b >= disc ? b - disc : b + disc);
	if (dist < 0.0f || dist >= tfar) return false;
	return true;
}

// BVH.hpp:236-288 over rays [begin_ray,end_ray) and prims [begin_prim,end_prim)
static void intersect_prims(const Ctx& c, const Buffer& in, Hit& out, size_t begin_ray, size_t end_ray, size_t begin_prim, size_t end_prim, uint64_t& tests) {
	const bool slot_exact = (c.flags & ORC_SLOT_EXACT) != 0;
	for (size_t prim_ID = begin_prim; prim_ID < end_prim; prim_ID++) {
		const Sphere& s = c.bvh.prims[prim_ID];
		size_t ID = begin_ray;
		const size_t simd_end = slot_exact ? begin_ray + ((end_ray - begin_ray) / 8) * 8 : end_ray;
		for (; ID < simd_end; ID++) sphere_closest_fma(s, static_cast<int32_t>(prim_ID), in.px[ID], in.py[ID], in.pz[ID], in.dx[ID], in.dy[ID], in.dz[ID], out.tfar[ID], out.primID[ID]);
		for (; ID < end_ray; ID++) sphere_closest_scalar(s, static_cast<int32_t>(prim_ID), in.px[ID], in.py[ID], in.pz[ID], in.dx[ID], in.dy[ID], in.dz[ID], out.tfar[ID], out.primID[ID]);
	}
	tests += static_cast<uint64_t>(end_prim - begin_prim) * (end_ray - begin_ray);
}
// BVH.hpp:290-305
static void intersect_prims_shadow(const Ctx& c, ShadowStream& in, size_t begin_ray, size_t end_ray, size_t begin_prim, size_t end_prim, uint64_t& tests) {
	for (size_t ID = begin_ray; ID < end_ray; ID++) {
		for (size_t prim_ID = begin_prim; prim_ID < end_prim; prim_ID++) {
			tests++;
			if (sphere_shadow(c.bvh.prims[prim_ID], in.px[ID], in.py[ID], in.pz[ID], in.dx[ID], in.dy[ID], in.dz[ID], in.tfar[ID])) { in.occluded[ID] = 1; break; }
		}
	}
}

// BVH.hpp:208-234
struct AABBAccel { float mx[kTileSize], my[kTileSize], mz[kTileSize], nx[kTileSize], ny[kTileSize], nz[kTileSize], t[kTileSize]; };
static inline bool test_AABB(const AABBAccel& a, const Node& node, size_t i) {
	float lo = node.min_bound[0] * a.mx[i] - a.nx[i];
	float hi = node.max_bound[0] * a.mx[i] - a.nx[i];
	float tmin = smax(1e-4f, smin(lo, hi));
	float tmax = smin(a.t[i], smax(lo, hi));
	lo = node.min_bound[1] * a.my[i] - a.ny[i];
	hi = node.max_bound[1] * a.my[i] - a.ny[i];
	tmin = smax(tmin, smin(lo, hi));
	tmax = smin(tmax, smax(lo, hi));
	lo = node.min_bound[2] * a.mz[i] - a.nz[i];
	hi = node.max_bound[2] * a.mz[i] - a.nz[i];
	tmin = smax(tmin, smin(lo, hi));
	tmax = smin(tmax, smax(lo, hi));
	return tmax >= tmin;
}

// BVH.hpp:309-360
static void traverse(const Ctx& c, const Buffer& in, Hit& out, size_t size, uint64_t& sphere_tests, uint64_t& box_tests) {
	const auto& prims = c.bvh.prims; const auto& nodes = c.bvh.nodes;
	auto finish = [&] { for (size_t i = 0; i < size; i++) if (out.primID[i] >= 0) out.matID[i] = prims[out.primID[i]].material_ID; };
	if (!(c.flags & ORC_BVH)) {  // :311-318 (as shipped)
		intersect_prims(c, in, out, 0, size, 0, prims.size(), sphere_tests);
		finish(); return;
	}
	struct StackFrame { size_t ID, head; };
	std::vector<StackFrame> stack; stack.reserve(64);
	StackFrame frame{0, 0};
	AABBAccel accel;
	for (size_t i = 0; i < size; i++) {  // :327-333
		accel.nx[i] = in.px[i] * (accel.mx[i] = 1.0f / in.dx[i]);
		accel.ny[i] = in.py[i] * (accel.my[i] = 1.0f / in.dy[i]);
		accel.nz[i] = in.pz[i] * (accel.mz[i] = 1.0f / in.dz[i]);
		accel.t[i] = out.tfar[i];
	}
	for (;;) {  // :335-358
	restart:
		const Node& node = nodes[frame.ID];
		for (; frame.head < size; frame.head++) {
			box_tests++;
			if (test_AABB(accel, node, frame.head)) {
				if (node.prim_count == 0) {
					stack.push_back({static_cast<size_t>(node.first_id) + 1, frame.head});
					frame.ID = static_cast<size_t>(node.first_id);
					goto restart;
				}
				intersect_prims(c, in, out, frame.head, size, node.first_id, node.first_id + node.prim_count, sphere_tests);
				break;
			}
		}
		if (stack.empty()) { finish(); return; }
		frame = stack.back(); stack.pop_back();
	}
}
// BVH.hpp:362-404
static void traverse_shadow(const Ctx& c, ShadowStream& in, size_t size, uint64_t& sphere_tests, uint64_t& box_tests) {
	const auto& nodes = c.bvh.nodes;
	if (!(c.flags & ORC_BVH)) { intersect_prims_shadow(c, in, 0, size, 0, c.bvh.prims.size(), sphere_tests); return; }
	struct StackFrame { size_t ID, head; };
	std::vector<StackFrame> stack; stack.reserve(64);
	StackFrame frame{0, 0};
	AABBAccel accel;
	for (size_t i = 0; i < size; i++) {
		accel.nx[i] = in.px[i] * (accel.mx[i] = 1.0f / in.dx[i]);
		accel.ny[i] = in.py[i] * (accel.my[i] = 1.0f / in.dy[i]);
		accel.nz[i] = in.pz[i] * (accel.mz[i] = 1.0f / in.dz[i]);
		accel.t[i] = in.tfar[i];
	}
	for (;;) {
	restart:
		const Node& node = nodes[frame.ID];
		for (; frame.head < size; frame.head++) {
			box_tests++;
			if (test_AABB(accel, node, frame.head)) {
				if (node.prim_count == 0) {
					stack.push_back({static_cast<size_t>(node.first_id) + 1, frame.head});
					frame.ID = static_cast<size_t>(node.first_id);
					goto restart;
				}
				intersect_prims_shadow(c, in, frame.head, size, node.first_id, node.first_id + node.prim_count, sphere_tests);
				break;
			}
		}
		if (stack.empty()) return;
		frame = stack.back(); stack.pop_back();
	}
}

// ---------------------------------------------------------------- Camera.hpp:80-88
static inline void generate_ray(const Ctx& c, int32_t x, int32_t y, const float* samples, V3* origin, V3* dir) {
	*origin = c.cam_pos;
	*dir = normalize(qrotate(c.cam_orient, V3{static_cast<float>(x) + samples[0] - c.half_width,
	                                         static_cast<float>(y) + samples[1] - c.half_height, c.cam_z}));
}
// Primitives.hpp:35-46
static inline V3 sky_eval(const Ctx& c, float x, float y, float z) {
	float u = c.hdri_fw * (0.5f + kOneOverTwoPi * fast_atan2(z, x));
	float v = c.hdri_fh * (0.5f - kOneOverPi * fast_asin(y));
	const float* texel = c.hdri.data() + 4 * (static_cast<int32_t>(v) * c.hdri_w + static_cast<int32_t>(u));
	return {texel[0] * c.ambient[0], texel[1] * c.ambient[1], texel[2] * c.ambient[2]};
}

// ---------------------------------------------------------------- Renderer::Accumulate body for one tile (Renderer.hpp:83-433)
struct TileScratch { Buffer buffers[2]; Hit hit; ShadowStream shadow; ShaderData sd; uint32_t seed[kTileSize]; uint32_t RayID[kTileSize]; uint8_t termination[kTileSize], has_shadowray[kTileSize]; std::vector<uint32_t> sort_buffer; };

static void accumulate_tile(Ctx& c, uint32_t LaunchIndex, TileScratch& S) {
	const uint32_t accumulations = c.accumulations;
	const uint32_t light_count = static_cast<uint32_t>(c.lights.size());                       // :77
	const float light_selection_pdf = 1.0f / static_cast<float>(c.lights.size());              // :78
	const bool has_ambient = smax(c.ambient[0], smax(c.ambient[1], c.ambient[2])) > 0.0f;      // :79
	const uint32_t bucket_index = accumulations % c.K;                                         // :82 (Q1)
	const bool mis = !(c.flags & ORC_NO_MIS);
	const bool ggx = (c.flags & ORC_GGX) != 0;
	float* out_r = c.accumulator.data() + (static_cast<size_t>(LaunchIndex) * c.K + bucket_index) * 3 * kTileSize;  // :84
	float* out_g = out_r + kTileSize; float* out_b = out_g + kTileSize;
	const int32_t tile_x = kTileRoot * static_cast<int32_t>(LaunchIndex % c.h_tiles);          // :85-88
	const int32_t tile_y = kTileRoot * static_cast<int32_t>(LaunchIndex / c.h_tiles);
	uint64_t n_ext = 0, n_shadow = 0, n_shaded = 0, n_term = 0, n_sphere = 0, n_box = 0, n_dropped = 0;

	Buffer* in = &S.buffers[0]; Buffer* out = &S.buffers[1];
	for (int i = 0; i < kTileSize; i++) {  // :97-109
		in->rr[i] = in->rg[i] = in->rb[i] = 0.0f;
		in->tr[i] = in->tg[i] = in->tb[i] = 1.0f;
		in->pixelID[i] = static_cast<uint32_t>(i);
		S.seed[i] = static_cast<uint32_t>(static_cast<int32_t>((static_cast<size_t>(LaunchIndex) * kTileSize + i) * (c.max_bounces * 2 + 1)));  // Q2
	}
	for (int ID = 0; ID < kTileSize; ID++) {  // :113-127
		int32_t x = tile_x + ID % kTileRoot, y = tile_y + ID / kTileRoot;
		uint32_t rng_state = hash_2d(accumulations, S.seed[ID]);
		const float camera_samples[2] = {rand_unit_float(&rng_state), rand_unit_float(&rng_state)};
		V3 o, d; generate_ray(c, x, y, camera_samples, &o, &d);
		in->dx[ID] = d.x; in->dy[ID] = d.y; in->dz[ID] = d.z; in->px[ID] = o.x; in->py[ID] = o.y; in->pz[ID] = o.z;
	}
	S.sort_buffer.assign(std::max<size_t>(64, c.material.size() + 2), 0);

	size_t active_rays = kTileSize;
	for (size_t bounce = 0; bounce < c.max_bounces && active_rays > 0; bounce++, std::swap(in, out)) {  // :131
		std::memset(S.termination, 0, sizeof S.termination); std::memset(S.has_shadowray, 0, sizeof S.has_shadowray);  // :135-140
		std::memset(S.shadow.occluded, 0, sizeof S.shadow.occluded); std::memset(S.sd.is_emissive, 0, sizeof S.sd.is_emissive);
		std::fill(S.sort_buffer.begin(), S.sort_buffer.end(), 0u);  // :141-149
		for (size_t i = 0; i < ((active_rays + 7) / 8) * 8; i++) { S.hit.tfar[i] = FLT_MAX; S.hit.matID[i] = -1; S.hit.primID[i] = -1; }  // :150-158

		n_ext += active_rays;
		traverse(c, *in, S.hit, active_rays, n_sphere, n_box);  // :165

		for (size_t ID = 0; ID < active_rays; ID++) {  // closest-hit shader :169-214
			const int32_t mat_ID = S.hit.matID[ID];
			if (mat_ID == -1) continue;
			const int32_t prim_ID = S.hit.primID[ID];
			const float depth = S.hit.tfar[ID];
			const V3 D{in->dx[ID], in->dy[ID], in->dz[ID]};
			V3 hit_point{in->px[ID] + D.x * depth, in->py[ID] + D.y * depth, in->pz[ID] + D.z * depth};
			const Sphere& sp = c.bvh.prims[prim_ID];
			V3 N{hit_point.x - sp.px, hit_point.y - sp.py, hit_point.z - sp.pz};
			N = normalize(N);
			if (dot(N, D) >= 0.0f) N = -N;
			Quat T = tangent_space(N);
			V3 Vlocal = to_local(T, -D);
			S.sd.Px[ID] = hit_point.x + N.x * 1e-4f; S.sd.Py[ID] = hit_point.y + N.y * 1e-4f; S.sd.Pz[ID] = hit_point.z + N.z * 1e-4f;
			S.sd.Vx[ID] = Vlocal.x; S.sd.Vy[ID] = Vlocal.y; S.sd.Vz[ID] = Vlocal.z;
			S.sd.Tx[ID] = T.x; S.sd.Ty[ID] = T.y; S.sd.Tz[ID] = T.z; S.sd.Tw[ID] = T.w;
			const Material& m = c.material[mat_ID];
			if (smax(m.emission[0], smax(m.emission[1], m.emission[2])) > FLT_EPSILON) S.sd.is_emissive[ID] = 1;  // :201-203
			if (!ggx) { S.sd.albedo[ID][0] = m.albedo[0]; S.sd.albedo[ID][1] = m.albedo[1]; S.sd.albedo[ID][2] = m.albedo[2]; }  // :208
			else {  // :210-212 with gloss_decay_table == 0
				S.sd.albedo[ID][0] = m.F0[0]; S.sd.albedo[ID][1] = m.F0[1]; S.sd.albedo[ID][2] = m.F0[2];
				float alpha = m.roughness; alpha *= alpha;
				S.sd.alpha[ID] = alpha + (1.0f - alpha) * 0.0f;
			}
		}

		size_t miss_count;  // sort_rayID, DataStreams.hpp:221-253 (call Renderer.hpp:235-241)
		{
			uint32_t* sb = S.sort_buffer.data();
			const uint32_t k = static_cast<uint32_t>(c.material.size());
			for (size_t i = 0; i < active_rays; i++) ++sb[1 + S.hit.matID[i]];   // histogram into sort_buffer+1 (key -1 -> slot 0)
			miss_count = sb[0];
			for (uint32_t i = 1; i < k + 1; i++) sb[i] += sb[i - 1];                         // prefix_sum(sort_buffer, k+1)
			for (int32_t i = static_cast<int32_t>(active_rays) - 1; i >= 0; i--) S.RayID[--sb[1 + S.hit.matID[i]]] = static_cast<uint32_t>(i);  // counting_sort
		}
		const size_t hit_count = active_rays - miss_count;  // :243
		n_shaded += hit_count;

		if (mis && light_count > 0) {  // zero lights is undefined in the reference (1/0, prims[0] of an empty list, Q15): defined here as "no light sampling"
			size_t shadow_index = 0;
			for (size_t i = 0; i < hit_count; i++) {  // NEE :249-298
				const int32_t ID = static_cast<int32_t>(S.RayID[miss_count + i]);
				uint32_t rng_state = hash_2d(accumulations, S.seed[in->pixelID[ID]] + static_cast<uint32_t>(bounce) * 2);  // :255 (Q3)
				const float LightSamples[2] = {rand_unit_float(&rng_state), rand_unit_float(&rng_state)};
				int32_t selected_light = static_cast<int32_t>(rand_bounded_int(&rng_state, light_count));
				int32_t light_primID = c.lights[selected_light];
				const Sphere& light_prim = c.geometry[light_primID];
				if (light_primID == S.hit.primID[ID]) continue;  // :263 (Q9: different index spaces)
				V3 Wc = V3{light_prim.px, light_prim.py, light_prim.pz} - V3{S.sd.Px[ID], S.sd.Py[ID], S.sd.Pz[ID]};
				float center_dist2 = dot(Wc, Wc);
				if (center_dist2 <= light_prim.radius_sq) continue;
				float center_dist = std::sqrt(center_dist2);
				Wc = Wc * (1.0f / center_dist);
				float sinThetaMax2 = light_prim.radius_sq / center_dist2;
				{
					float NdotW = (2.0f * S.sd.Tw[ID]) * (Wc.z * S.sd.Tw[ID] + Wc.x * S.sd.Ty[ID] - S.sd.Tx[ID] * Wc.y) - Wc.z;  // :271
					if (NdotW < 0.0f && sinThetaMax2 < NdotW * NdotW) continue;
				}
				float light_distance, light_pdf;
				V3 L = sample_direction_to_sphere(Wc, sinThetaMax2, center_dist, light_prim.radius_sq, LightSamples[0], LightSamples[1], &light_distance, &light_pdf);
				Quat T{S.sd.Tw[ID], S.sd.Tx[ID], S.sd.Ty[ID], S.sd.Tz[ID]};
				V3 Llocal = to_local(T, L);
				if (Llocal.z < 0.0f) continue;
				const Material& lm = c.material[light_prim.material_ID];
				V3 radiance = V3{lm.emission[0], lm.emission[1], lm.emission[2]} * V3{in->tr[ID], in->tg[ID], in->tb[ID]};  // :277
				if (!ggx) {  // Closure<Lambertian>::eval, DataStreams.hpp:169-172
					float NdotL = smax(0.0f, Llocal.z);
					float f = kOneOverPi * NdotL;
					radiance = radiance * V3{S.sd.albedo[ID][0] * f, S.sd.albedo[ID][1] * f, S.sd.albedo[ID][2] * f};
				} else radiance = radiance * ggx_eval(V3{S.sd.albedo[ID][0], S.sd.albedo[ID][1], S.sd.albedo[ID][2]}, S.sd.alpha[ID], Llocal, V3{S.sd.Vx[ID], S.sd.Vy[ID], S.sd.Vz[ID]});  // :189-195
				light_pdf *= light_selection_pdf;
				float brdf_pdf = ggx ? 0.0f : kOneOverPi * smax(0.0f, Llocal.z);  // DataStreams.hpp:173-176; GGX :196-198 (`return 0.0f; //TODO`)
				radiance = radiance * powerHeuristic_over_f(light_pdf, brdf_pdf);
				if (smax(smax(radiance.x, radiance.y), radiance.z) <= 0.0f) continue;
				S.shadow.dx[shadow_index] = L.x; S.shadow.dy[shadow_index] = L.y; S.shadow.dz[shadow_index] = L.z;
				S.shadow.px[shadow_index] = S.sd.Px[ID]; S.shadow.py[shadow_index] = S.sd.Py[ID]; S.shadow.pz[shadow_index] = S.sd.Pz[ID];
				S.shadow.tfar[shadow_index] = light_distance;
				S.shadow.r[shadow_index] = radiance.x; S.shadow.g[shadow_index] = radiance.y; S.shadow.b[shadow_index] = radiance.z;
				S.has_shadowray[ID] = 1;
				++shadow_index;
			}
			n_shadow += shadow_index;
			traverse_shadow(c, S.shadow, shadow_index, n_sphere, n_box);  // :302
			for (size_t i = miss_count, shadow_ID = 0; i < active_rays; i++) {  // :304-314
				const int32_t ID = static_cast<int32_t>(S.RayID[i]);
				if (S.has_shadowray[ID]) {
					if (!S.shadow.occluded[shadow_ID]) { in->rr[ID] += S.shadow.r[shadow_ID]; in->rg[ID] += S.shadow.g[shadow_ID]; in->rb[ID] += S.shadow.b[shadow_ID]; }
					++shadow_ID;
				}
			}
		}
		// emissive primitive hit :319-353
		if (mis && bounce > 0) {
			for (size_t ID = 0; ID < active_rays; ID++) {
				if (!S.sd.is_emissive[ID]) continue;
				V3 throughput{in->tr[ID], in->tg[ID], in->tb[ID]};
				const Sphere& light_prim = c.bvh.prims[S.hit.primID[ID]];
				const float radius2 = light_prim.radius_sq;
				const float depth = S.hit.tfar[ID];
				const float NdotV = S.sd.Vz[ID];
				float center_dist2 = depth * (depth + NdotV * (2.0f * std::sqrt(radius2))) + radius2;  // :331
				float weight = powerHeuristic(in->pdf[ID], light_selection_pdf * spherePdf(radius2, center_dist2));
				throughput = throughput * weight;
				const float* em = c.material[S.hit.matID[ID]].emission;
				in->rr[ID] += throughput.x * em[0]; in->rg[ID] += throughput.y * em[1]; in->rb[ID] += throughput.z * em[2];
			}
		} else if (mis) {  // bounce 0: raw emission (:344-352)
			for (size_t ID = 0; ID < active_rays; ID++) {
				if (!S.sd.is_emissive[ID]) continue;
				const float* em = c.material[S.hit.matID[ID]].emission;
				in->rr[ID] += em[0]; in->rg[ID] += em[1]; in->rb[ID] += em[2];
			}
		} else {  // ORC_NO_MIS: this repo's definition (Q23), NOT the reference's `MIS false` build
			for (size_t ID = 0; ID < active_rays; ID++) {
				if (!S.sd.is_emissive[ID]) continue;
				const float* em = c.material[S.hit.matID[ID]].emission;
				in->rr[ID] += in->tr[ID] * em[0]; in->rg[ID] += in->tg[ID] * em[1]; in->rb[ID] += in->tb[ID] * em[2];
			}
		}
		// BRDF sampling :357-404
		size_t output_index = 0;
		if (bounce < c.max_bounces - 1) {
			for (size_t i = 0; i < hit_count; i++) {
				const int32_t ID = static_cast<int32_t>(S.RayID[miss_count + i]);
				uint32_t rng_state = hash_2d(accumulations, S.seed[in->pixelID[ID]] + static_cast<uint32_t>(bounce) * 2 + 1);  // :362
				const float brdf_samples[2] = {rand_unit_float(&rng_state), rand_unit_float(&rng_state)};
				V3 sdir, estimator;
				if (!ggx) { sdir = hemisphere(brdf_samples[0], brdf_samples[1]); estimator = V3{S.sd.albedo[ID][0], S.sd.albedo[ID][1], S.sd.albedo[ID][2]}; }  // DataStreams.hpp:177-181
				else ggx_sample(V3{S.sd.albedo[ID][0], S.sd.albedo[ID][1], S.sd.albedo[ID][2]}, S.sd.alpha[ID], V3{S.sd.Vx[ID], S.sd.Vy[ID], S.sd.Vz[ID]}, brdf_samples[0], brdf_samples[1], &sdir, &estimator);  // :199-218
				V3 throughput{in->tr[ID], in->tg[ID], in->tb[ID]};
				throughput = throughput * estimator;
				{  // Russian roulette :377-384 (Q13)
					float q = 1.0f - smax(throughput.x, smax(throughput.y, throughput.z));
					if (rand_unit_float(&rng_state) < q) { S.termination[ID] = 1; continue; }
					throughput = throughput * (1.0f / smax(FLT_EPSILON, 1.0f - q));
				}
				Quat T{S.sd.Tw[ID], S.sd.Tx[ID], S.sd.Ty[ID], S.sd.Tz[ID]};
				sdir = to_world(T, sdir);  // :386
				out->px[output_index] = S.sd.Px[ID]; out->py[output_index] = S.sd.Py[ID]; out->pz[output_index] = S.sd.Pz[ID];
				out->dx[output_index] = sdir.x; out->dy[output_index] = sdir.y; out->dz[output_index] = sdir.z;
				out->tr[output_index] = throughput.x; out->tg[output_index] = throughput.y; out->tb[output_index] = throughput.z;
				out->rr[output_index] = in->rr[ID]; out->rg[output_index] = in->rg[ID]; out->rb[output_index] = in->rb[ID];
				out->pixelID[output_index] = in->pixelID[ID];
				out->pdf[output_index] = ggx ? 0.0f : kOneOverPi * smax(0.0f, sdir.z);  // :401 — pdf of the WORLD-space dir (Q10); GGX: 0 (:196-198)
				output_index++;
			}
		} else {
			n_dropped += hit_count;  // Q11: last-bounce survivors are neither continued nor accumulated
		}
		// miss shader :408-420
		for (size_t i = 0; i < miss_count; i++) S.termination[S.RayID[i]] = 1;
		if (has_ambient) {
			for (size_t i = 0; i < miss_count; i++) {
				const int32_t ID = static_cast<int32_t>(S.RayID[i]);
				V3 sky = sky_eval(c, in->dx[ID], in->dy[ID], in->dz[ID]);
				in->rr[ID] += in->tr[ID] * sky.x; in->rg[ID] += in->tr[ID] * sky.y; in->rb[ID] += in->tr[ID] * sky.z;  // Q14: throughput.r for all three
			}
		}
		// accumulation :424-431
		for (size_t ID = 0; ID < active_rays; ID++) {
			if (!S.termination[ID]) continue;
			const uint32_t px = in->pixelID[ID];
			out_r[px] += in->rr[ID]; out_g[px] += in->rg[ID]; out_b[px] += in->rb[ID];
			n_term++;
		}
		active_rays = output_index;
	}
	c.counters.extension_rays += n_ext; c.counters.shadow_rays += n_shadow; c.counters.shaded_hits += n_shaded;
	c.counters.terminated += n_term; c.counters.sphere_tests += n_sphere; c.counters.box_tests += n_box; c.counters.dropped += n_dropped;
}

static void run_tiles(Ctx& c, const uint32_t* tiles, size_t n_tiles, int threads) {
	if (threads < 1) threads = 1;
	std::atomic<size_t> next{0};
	auto worker = [&] {
		auto S = std::make_unique<TileScratch>();
		for (;;) {
			size_t i = next.fetch_add(1);
			if (i >= n_tiles) break;
			accumulate_tile(c, tiles ? tiles[i] : static_cast<uint32_t>(i), *S);
		}
	};
	if (threads == 1) { worker(); return; }
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; t++) pool.emplace_back(worker);
	for (auto& t : pool) t.join();
}

// median of K bucket sums. K == 5: the reference's network (Sampling.hpp:13-21 with VCL min/max). Other K are this
// repo's definition: odd K -> middle order statistic, even K -> mean of the two middle ones.
static inline float vmin(float a, float b) { return (a < b) ? a : b; }
static inline float vmax(float a, float b) { return (a > b) ? a : b; }
static inline float median3_vcl(float a, float b, float c) { return vmax(vmin(a, b), vmin(vmax(a, b), c)); }
static float median_k(const float* v, uint32_t K) {
	if (K == 5) return median3_vcl(vmax(vmin(v[0], v[1]), vmin(v[2], v[3])), vmin(vmax(v[0], v[1]), vmax(v[2], v[3])), v[4]);
	if (K == 3) return median3_vcl(v[0], v[1], v[2]);
	if (K == 1) return v[0];
	float s[64]; for (uint32_t i = 0; i < K; i++) s[i] = v[i];
	std::sort(s, s + K);
	return (K & 1) ? s[K / 2] : (s[K / 2 - 1] + s[K / 2]) * 0.5f;
}

}  // namespace

// ================================================================ C API (ctypes)
extern "C" {

Ctx* orc_create(uint32_t width, uint32_t height, uint32_t max_bounces, uint32_t K, uint32_t flags) {
	if (width % kTileRoot || height % kTileRoot || K < 1 || K > 64 || max_bounces < 1) return nullptr;
	Ctx* c = new Ctx();
	c->width = width; c->height = height; c->h_tiles = width / kTileRoot; c->v_tiles = height / kTileRoot;  // Renderer.hpp:53-63
	c->max_bounces = max_bounces; c->K = K; c->flags = flags;
	c->accumulator.assign(static_cast<size_t>(c->h_tiles) * c->v_tiles * K * 3 * kTileSize, 0.0f);
	return c;
}
void orc_destroy(Ctx* c) { delete c; }

// Scene.hpp:19-26 + Application.cpp:233-234: builds the BVH and the light list
int orc_set_scene(Ctx* c, const void* geometry, uint32_t n_geom, const void* materials, uint32_t n_mat,
                  const float ambient[3], const float* hdri_rgba, int32_t hdri_w, int32_t hdri_h) {
	c->geometry.assign(static_cast<const Sphere*>(geometry), static_cast<const Sphere*>(geometry) + n_geom);
	c->material.assign(static_cast<const Material*>(materials), static_cast<const Material*>(materials) + n_mat);
	build_bvh(c->geometry.data(), n_geom, c->bvh);
	c->lights.clear();  // Scene.hpp:12-16
	for (int32_t i = 0; i < static_cast<int32_t>(n_geom); i++) {
		const float* em = c->material[c->geometry[i].material_ID].emission;
		if (em[0] * em[0] + em[1] * em[1] + em[2] * em[2] > 0.0f) c->lights.push_back(i);
	}
	for (int i = 0; i < 3; i++) c->ambient[i] = ambient ? ambient[i] : 0.0f;
	c->hdri_w = hdri_w; c->hdri_h = hdri_h;
	if (hdri_rgba && hdri_w > 0 && hdri_h > 0) c->hdri.assign(hdri_rgba, hdri_rgba + static_cast<size_t>(hdri_w) * hdri_h * 4);
	else c->hdri.clear();
	c->hdri_fw = static_cast<float>(hdri_w - 1); c->hdri_fh = static_cast<float>(hdri_h - 1);  // Application.cpp:230-231
	return static_cast<int>(c->lights.size());
}
// Camera{eye, direction, w, h, focal_length, ...} (Camera.hpp:5-32,47-50,61-68). Resized to the ctx image.
void orc_set_camera_lookat(Ctx* c, const float eye[3], const float dir[3], float focal_length, float exposure) {
	c->cam_pos = V3{eye[0], eye[1], eye[2]};
	c->cam_orient = quat_look_at(normalize(V3{dir[0], dir[1], dir[2]}), V3{0.0f, 1.0f, 0.0f});
	float inv_half_tan = (-2.0f / 24.0f) * focal_length;  // Camera.hpp:21-26
	c->half_height = static_cast<float>(c->height) * 0.5f; c->half_width = static_cast<float>(c->width) * 0.5f;
	c->cam_z = c->half_height * inv_half_tan;
	c->exposure = exposure;
}
void orc_set_camera_raw(Ctx* c, const float pos[3], const float orient_wxyz[4], float half_w, float half_h, float z, float exposure) {
	c->cam_pos = V3{pos[0], pos[1], pos[2]}; c->cam_orient = Quat{orient_wxyz[0], orient_wxyz[1], orient_wxyz[2], orient_wxyz[3]};
	c->half_width = half_w; c->half_height = half_h; c->cam_z = z; c->exposure = exposure;
}
void orc_get_camera_raw(const Ctx* c, float out[11]) {
	out[0] = c->cam_pos.x; out[1] = c->cam_pos.y; out[2] = c->cam_pos.z;
	out[3] = c->cam_orient.w; out[4] = c->cam_orient.x; out[5] = c->cam_orient.y; out[6] = c->cam_orient.z;
	out[7] = c->half_width; out[8] = c->half_height; out[9] = c->cam_z; out[10] = c->exposure;
}
// one camera ray for caller-supplied sub-pixel samples (Camera::generate_ray, Camera.hpp:80-88): known-answer tap
void orc_generate_ray_at(const Ctx* c, int32_t x, int32_t y, const float samples[2], float out[6]) {
	V3 o, d; generate_ray(*c, x, y, samples, &o, &d);
	out[0] = o.x; out[1] = o.y; out[2] = o.z; out[3] = d.x; out[4] = d.y; out[5] = d.z;
}
void orc_reset(Ctx* c) { c->accumulations = 0; std::fill(c->accumulator.begin(), c->accumulator.end(), 0.0f); }  // Renderer.hpp:64-67
void orc_set_accumulations(Ctx* c, uint32_t acc) { c->accumulations = acc; }
uint32_t orc_get_accumulations(const Ctx* c) { return c->accumulations; }

// Renderer::Accumulate (Renderer.hpp:73-434) n times, all tiles, `threads` workers (PPL parallel_for stand-in)
void orc_accumulate(Ctx* c, uint32_t n_samples, int threads) {
	for (uint32_t s = 0; s < n_samples; s++) { ++c->accumulations; run_tiles(*c, nullptr, static_cast<size_t>(c->h_tiles) * c->v_tiles, threads); }
}
// Same, restricted to a list of tiles (full-size spot checks): every tile is independent (Renderer.hpp:84)
void orc_accumulate_tiles(Ctx* c, const uint32_t* tiles, uint32_t n_tiles, uint32_t n_samples, int threads) {
	for (uint32_t s = 0; s < n_samples; s++) { ++c->accumulations; run_tiles(*c, tiles, n_tiles, threads); }
}
// bucket sums as [K][3][npix] with npix in tile order (t = tile*256 + ID)
void orc_read_buckets(const Ctx* c, float* out) {
	const size_t tiles = static_cast<size_t>(c->h_tiles) * c->v_tiles, npix = tiles * kTileSize;
	for (size_t t = 0; t < tiles; t++) for (uint32_t k = 0; k < c->K; k++) for (int ch = 0; ch < 3; ch++)
		std::memcpy(out + (static_cast<size_t>(k) * 3 + ch) * npix + t * kTileSize, c->accumulator.data() + ((t * c->K + k) * 3 + ch) * kTileSize, kTileSize * sizeof(float));
}
// Renderer::Render (Renderer.hpp:436-478), generalised to K buckets. rgba_out: width*height*4, row-major, row 0 first
// (the reference displays it V-flipped, Application.cpp:381). tonemap=0 returns the linear median-of-means instead.
int orc_render(const Ctx* c, float* rgba_out, int tonemap) {
	if (c->accumulations % c->K) return 1;  // :437
	const float scale = c->exposure / static_cast<float>(c->accumulations / c->K);  // :439
	const size_t tiles = static_cast<size_t>(c->h_tiles) * c->v_tiles;
	for (size_t t = 0; t < tiles; t++) {
		const float* src = c->accumulator.data() + t * c->K * 3 * kTileSize;
		for (int ID = 0; ID < kTileSize; ID++) {
			float v[3][64];
			for (uint32_t k = 0; k < c->K; k++) for (int ch = 0; ch < 3; ch++) v[ch][k] = src[(k * 3 + ch) * kTileSize + ID];
			float r = scale * median_k(v[0], c->K), g = scale * median_k(v[1], c->K), b = scale * median_k(v[2], c->K);  // :453-455
			if (tonemap) tonemapping(r, g, b);  // :461
			size_t x = kTileRoot * (t % c->h_tiles) + ID % kTileRoot, y = kTileRoot * (t / c->h_tiles) + ID / kTileRoot;
			float* dst = rgba_out + (y * c->width + x) * 4;
			dst[0] = r; dst[1] = g; dst[2] = b; dst[3] = 1.0f;  // :465-473
		}
	}
	return 0;
}
void orc_read_counters(const Ctx* c, uint64_t out[7]) {
	out[0] = c->counters.extension_rays; out[1] = c->counters.shadow_rays; out[2] = c->counters.shaded_hits; out[3] = c->counters.terminated;
	out[4] = c->counters.sphere_tests; out[5] = c->counters.box_tests; out[6] = c->counters.dropped;
}
void orc_reset_counters(Ctx* c) {
	c->counters.extension_rays = 0; c->counters.shadow_rays = 0; c->counters.shaded_hits = 0; c->counters.terminated = 0;
	c->counters.sphere_tests = 0; c->counters.box_tests = 0; c->counters.dropped = 0;
}

// BVH taps
uint32_t orc_bvh_node_count(const Ctx* c) { return static_cast<uint32_t>(c->bvh.nodes.size()); }
void orc_bvh_read(const Ctx* c, void* nodes_out, void* prims_out, uint32_t* prim_ids_out) {
	if (nodes_out) std::memcpy(nodes_out, c->bvh.nodes.data(), c->bvh.nodes.size() * sizeof(Node));
	if (prims_out) std::memcpy(prims_out, c->bvh.prims.data(), c->bvh.prims.size() * sizeof(Sphere));
	if (prim_ids_out) std::memcpy(prim_ids_out, c->bvh.prim_ids.data(), c->bvh.prim_ids.size() * sizeof(uint32_t));
}
uint32_t orc_light_count(const Ctx* c) { return static_cast<uint32_t>(c->lights.size()); }
void orc_read_lights(const Ctx* c, int32_t* out) { std::memcpy(out, c->lights.data(), c->lights.size() * sizeof(int32_t)); }

// primary rays of sample `acc` for every pixel, tile order: out[t*6 + {ox,oy,oz,dx,dy,dz}] (Renderer.hpp:113-127)
void orc_generate_rays(const Ctx* c, uint32_t acc, float* out) {
	const size_t tiles = static_cast<size_t>(c->h_tiles) * c->v_tiles;
	for (size_t tile = 0; tile < tiles; tile++) for (int ID = 0; ID < kTileSize; ID++) {
		size_t t = tile * kTileSize + ID;
		uint32_t seed = static_cast<uint32_t>(static_cast<int32_t>(t * (c->max_bounces * 2 + 1)));
		int32_t x = kTileRoot * static_cast<int32_t>(tile % c->h_tiles) + ID % kTileRoot, y = kTileRoot * static_cast<int32_t>(tile / c->h_tiles) + ID / kTileRoot;
		uint32_t st = hash_2d(acc, seed);
		const float cs[2] = {rand_unit_float(&st), rand_unit_float(&st)};
		V3 o, d; generate_ray(*c, x, y, cs, &o, &d);
		float* p = out + t * 6; p[0] = o.x; p[1] = o.y; p[2] = o.z; p[3] = d.x; p[4] = d.y; p[5] = d.z;
	}
}
// brute-force closest hit (FMA formula) / any hit for arbitrary rays: rays[n*6], tfar_io[n], prim_out[n]
void orc_trace_closest(const Ctx* c, const float* rays, uint32_t n, float* tfar_out, int32_t* prim_out) {
	for (uint32_t i = 0; i < n; i++) {
		const float* r = rays + static_cast<size_t>(i) * 6; float tf = FLT_MAX; int32_t pid = -1;
		for (size_t p = 0; p < c->bvh.prims.size(); p++) sphere_closest_fma(c->bvh.prims[p], static_cast<int32_t>(p), r[0], r[1], r[2], r[3], r[4], r[5], tf, pid);
		tfar_out[i] = tf; prim_out[i] = pid;
	}
}
// the scalar-tail formula of the closest-hit loop (BVH.hpp:270-286) for every ray: known-answer tap
void orc_trace_closest_scalar(const Ctx* c, const float* rays, uint32_t n, float* tfar_out, int32_t* prim_out) {
	for (uint32_t i = 0; i < n; i++) {
		const float* r = rays + static_cast<size_t>(i) * 6; float tf = FLT_MAX; int32_t pid = -1;
		for (size_t p = 0; p < c->bvh.prims.size(); p++) sphere_closest_scalar(c->bvh.prims[p], static_cast<int32_t>(p), r[0], r[1], r[2], r[3], r[4], r[5], tf, pid);
		tfar_out[i] = tf; prim_out[i] = pid;
	}
}
void orc_trace_shadow(const Ctx* c, const float* rays, const float* tfar, uint32_t n, uint8_t* occluded_out) {
	for (uint32_t i = 0; i < n; i++) {
		const float* r = rays + static_cast<size_t>(i) * 6; uint8_t occ = 0;
		for (size_t p = 0; p < c->bvh.prims.size() && !occ; p++) occ = sphere_shadow(c->bvh.prims[p], r[0], r[1], r[2], r[3], r[4], r[5], tfar[i]);
		occluded_out[i] = occ;
	}
}

// ---- scalar taps for known-answer tests (tests/test_oracle_kat.py, tests/golden/)
uint32_t orc_hash_u32(uint32_t i) { return hash_u32(i); }
uint32_t orc_hash_2d(uint32_t x, uint32_t y) { return hash_2d(x, y); }
uint32_t orc_pcg_generate(uint32_t* s) { return pcg_generate(s); }
float orc_rand_unit_float(uint32_t* s) { return rand_unit_float(s); }
uint32_t orc_rand_bounded_int(uint32_t* s, uint32_t range) { return rand_bounded_int(s, range); }
float orc_make_unit_float(uint32_t x) { return make_unit_float(x); }
uint32_t orc_bitreverse(uint32_t x) { return bitreverse32(x); }
void orc_fast_sincos(float x, float* s, float* c) { fast_sincos(x, s, c); }
float orc_fast_asin(float x) { return fast_asin(x); }
float orc_fast_atan2(float y, float x) { return fast_atan2(y, x); }
float orc_median5(const float v[5]) { return median5(v[0], v[1], v[2], v[3], v[4]); }
float orc_median_k(const float* v, uint32_t K) { return median_k(v, K); }
void orc_hemisphere(float t, float s, float out[3]) { V3 v = hemisphere(t, s); out[0] = v.x; out[1] = v.y; out[2] = v.z; }
void orc_tangent_space(const float n[3], float out_wxyz[4]) { Quat q = tangent_space(V3{n[0], n[1], n[2]}); out_wxyz[0] = q.w; out_wxyz[1] = q.x; out_wxyz[2] = q.y; out_wxyz[3] = q.z; }
void orc_to_local(const float q[4], const float v[3], float out[3]) { V3 r = to_local(Quat{q[0], q[1], q[2], q[3]}, V3{v[0], v[1], v[2]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_to_world(const float q[4], const float v[3], float out[3]) { V3 r = to_world(Quat{q[0], q[1], q[2], q[3]}, V3{v[0], v[1], v[2]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_orthonormal_basis(const float n[3], float out[6]) { V3 a, b; orthonormal_basis(V3{n[0], n[1], n[2]}, &a, &b); out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = b.x; out[4] = b.y; out[5] = b.z; }
void orc_sample_direction_to_sphere(const float Wc[3], float s2, float cd, float r2, float t, float s, float out[5]) {
	float dist, pdf; V3 L = sample_direction_to_sphere(V3{Wc[0], Wc[1], Wc[2]}, s2, cd, r2, t, s, &dist, &pdf);
	out[0] = L.x; out[1] = L.y; out[2] = L.z; out[3] = dist; out[4] = pdf;
}
float orc_sphere_pdf(float r2, float d2) { return spherePdf(r2, d2); }
float orc_power_heuristic(float f, float g) { return powerHeuristic(f, g); }
void orc_distribution_visible_normals(const float v[3], float alpha, float u0, float u1, float out[3]) { V3 h = distribution_visible_normals(V3{v[0], v[1], v[2]}, alpha, u0, u1); out[0] = h.x; out[1] = h.y; out[2] = h.z; }
void orc_microfacet_brdf(const float f0[3], float alpha, float ndv, float ndl, float ndh, float hdv, float out[3]) { V3 r = microfacet_brdf(V3{f0[0], f0[1], f0[2]}, alpha, ndv, ndl, ndh, hdv); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_vndf_estimator(const float f0[3], float alpha, float ndv, float ndl, float hdv, float out[3]) { V3 r = vndf_estimator(V3{f0[0], f0[1], f0[2]}, alpha, ndv, ndl, hdv); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_ggx_eval(const float f0[3], float alpha, const float l[3], const float v[3], float out[3]) { V3 r = ggx_eval(V3{f0[0], f0[1], f0[2]}, alpha, V3{l[0], l[1], l[2]}, V3{v[0], v[1], v[2]}); out[0] = r.x; out[1] = r.y; out[2] = r.z; }
void orc_ggx_sample(const float f0[3], float alpha, const float v[3], float u0, float u1, float out[6]) { V3 d, e; ggx_sample(V3{f0[0], f0[1], f0[2]}, alpha, V3{v[0], v[1], v[2]}, u0, u1, &d, &e); out[0] = d.x; out[1] = d.y; out[2] = d.z; out[3] = e.x; out[4] = e.y; out[5] = e.z; }
float orc_power_heuristic_over_f(float f, float g) { return powerHeuristic_over_f(f, g); }
void orc_tonemap(float rgb[3]) { tonemapping(rgb[0], rgb[1], rgb[2]); }
void orc_quat_look_at(const float dir[3], float out_wxyz[4]) { Quat q = quat_look_at(normalize(V3{dir[0], dir[1], dir[2]}), V3{0, 1, 0}); out_wxyz[0] = q.w; out_wxyz[1] = q.x; out_wxyz[2] = q.y; out_wxyz[3] = q.z; }
// standalone BVH build for structural tests: nodes_out must hold 2n-1 (n>0) entries
uint32_t orc_build_bvh(const void* geometry, uint32_t n, void* nodes_out, void* prims_out, uint32_t* prim_ids_out) {
	BVH b; build_bvh(static_cast<const Sphere*>(geometry), n, b);
	std::memcpy(nodes_out, b.nodes.data(), b.nodes.size() * sizeof(Node));
	if (prims_out) std::memcpy(prims_out, b.prims.data(), b.prims.size() * sizeof(Sphere));
	if (prim_ids_out) std::memcpy(prim_ids_out, b.prim_ids.data(), b.prim_ids.size() * sizeof(uint32_t));
	return static_cast<uint32_t>(b.nodes.size());
}

}  // extern "C"
