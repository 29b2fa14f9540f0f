// BVH.hpp — BoundingVolumeHierarchy<Sphere> of the reference (BVH.hpp:16-206): same public members (nodes, prims), built by
// libb2r's host builder with bit-identical node and leaf order. Traversal happens on the GPU inside Renderer; the public
// Traverse / Traverse_shadow entry points (used by the app's focus picking, Application.cpp:282-298) go through a Renderer.
#pragma once
#include <cfloat>
#include <memory>
#include <span>
#include <stdexcept>
#include <string>
#include <vector>
#include "DataStreams.hpp"
#include "Primitives.hpp"

template <typename Primitive> struct BoundingVolumeHierarchy;
template <> struct BoundingVolumeHierarchy<Sphere> {
	using Node = b2r_bvh_node;              // {vec3 min_bound; u32 first_id; vec3 max_bound; u32 prim_count}, BVH.hpp:18-27
	std::vector<Node> nodes;
	std::vector<Sphere> prims;              // leaf order (BVH.hpp:201-205)
	std::vector<uint32_t> prim_ids;         // prims[i] == geometry[prim_ids[i]]
	struct SplitHeuristic {                 // BVH.hpp:70-83 (the arithmetic lives in b2r_bvh_build_ex)
		size_t log_cluster_size = 0;        // log2 of the size of primitive clusters
		float cost_ratio = 1.0f;            // cost of a ray-box test over the cost of a primitive test
	};
	BoundingVolumeHierarchy() {}
	// BVH.hpp:90 has one constructor with `SplitHeuristic heuristic = SplitHeuristic{}`; in a non-template class that default argument is
	// ill-formed (the nested struct's member initialisers are not complete yet), hence the delegating pair
	BoundingVolumeHierarchy(std::span<const Sphere> primitives) : BoundingVolumeHierarchy(primitives, SplitHeuristic()) {}
	BoundingVolumeHierarchy(std::span<const Sphere> primitives, SplitHeuristic heuristic) {
		const uint32_t n = static_cast<uint32_t>(primitives.size());
		nodes.resize(n ? 2 * n - 1 : 1); prims.resize(n); prim_ids.resize(n);
		uint32_t n_nodes = 0;
		b2r_bvh_build_ex(reinterpret_cast<const b2r_sphere*>(primitives.data()), n, static_cast<uint32_t>(heuristic.log_cluster_size), heuristic.cost_ratio,
		                 nodes.data(), reinterpret_cast<b2r_sphere*>(prims.data()), prim_ids.data(), &n_nodes);
		nodes.resize(n_nodes);
	}
	const Node& root() const { return nodes.front(); }

	// BoundingVolumeHierarchy::Traverse<N> / Traverse_shadow<N> (BVH.hpp:309-404) with the reference's signatures, for callers
	// outside the renderer (focus picking, Application.cpp:282-298). The rays are traced on the GPU by a small private context
	// holding just this tree (created on first use); out.matID is filled like BVH.hpp:313-317.
	template <size_t N> void Traverse(const typename RayStream<N>::Buffer& in, typename RayStream<N>::Hit& out, size_t size) const {
		if (size == 0 || prims.empty()) return;
		std::vector<float> rays(6 * size); std::vector<float> t(size); std::vector<int32_t> id(size);
		for (size_t i = 0; i < size; i++) { float* r = &rays[6 * i]; r[0] = in.p.x[i]; r[1] = in.p.y[i]; r[2] = in.p.z[i]; r[3] = in.dir.x[i]; r[4] = in.dir.y[i]; r[5] = in.dir.z[i]; }
		check(b2r_trace_closest(tracer(), rays.data(), static_cast<uint32_t>(size), t.data(), id.data()));
		for (size_t i = 0; i < size; i++) if (id[i] >= 0 && t[i] < out.tfar[i]) { out.tfar[i] = t[i]; out.primID[i] = id[i]; out.matID[i] = prims[id[i]].material_ID; }
	}
	template <size_t N> void Traverse_shadow(typename RayStream<N>::ShadowStream& in, size_t size) const {
		if (size == 0 || prims.empty()) return;
		std::vector<float> rays(6 * size); std::vector<uint8_t> occ(size);
		for (size_t i = 0; i < size; i++) { float* r = &rays[6 * i]; r[0] = in.p.x[i]; r[1] = in.p.y[i]; r[2] = in.p.z[i]; r[3] = in.dir.x[i]; r[4] = in.dir.y[i]; r[5] = in.dir.z[i]; }
		check(b2r_trace_shadow(tracer(), rays.data(), in.tfar, static_cast<uint32_t>(size), occ.data()));
		for (size_t i = 0; i < size; i++) if (occ[i]) in.occluded_flag[i] = true;
	}

private:
	struct CtxDeleter { void operator()(b2r_ctx* c) const { b2r_destroy(c); } };
	mutable std::shared_ptr<b2r_ctx> trace_ctx;
	static void check(int rc) { if (rc < 0) throw std::runtime_error(std::string("libb2r: ") + b2r_last_error()); }
	b2r_ctx* tracer() const {
		if (!trace_ctx) {
			b2r_config cfg{}; cfg.width = 16; cfg.height = 16; cfg.max_bounces = 1; cfg.buckets = 1; cfg.flags = B2R_FLAG_NO_MIS; cfg.samples_in_flight = 1;
			b2r_ctx* c = nullptr; check(b2r_create(&c, &cfg));
			trace_ctx = std::shared_ptr<b2r_ctx>(c, CtxDeleter{});
			int32_t max_mat = 0; for (const Sphere& s : prims) max_mat = s.material_ID > max_mat ? s.material_ID : max_mat;
			std::vector<Material> mats(static_cast<size_t>(max_mat) + 1);  // traversal never reads them
			const float amb[3] = {0, 0, 0};
			check(b2r_upload_scene(c, reinterpret_cast<const b2r_sphere*>(prims.data()), nodes.data(), static_cast<uint32_t>(prims.size()), static_cast<uint32_t>(nodes.size()),
			                       reinterpret_cast<const b2r_material*>(mats.data()), static_cast<uint32_t>(mats.size()), nullptr, 0,
			                       reinterpret_cast<const b2r_sphere*>(prims.data()), static_cast<uint32_t>(prims.size()), amb, nullptr, 0, 0));
		}
		return trace_ctx.get();
	}
};
