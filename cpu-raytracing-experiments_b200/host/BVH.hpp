// BVH.hpp — BoundingVolumeHierarchy<Sphere> of the reference (BVH.hpp:16-206): same public members (nodes, prims), built by
// libb2r's host builder with bit-identical node and leaf order. Traversal happens on the GPU inside Renderer; the public
// Traverse / Traverse_shadow entry points (used by the app's focus picking, Application.cpp:282-298) go through a Renderer.
#pragma once
#include <span>
#include <vector>
#include "Primitives.hpp"

template <typename Primitive> struct BoundingVolumeHierarchy;
template <> struct BoundingVolumeHierarchy<Sphere> {
	using Node = b2r_bvh_node;              // {vec3 min_bound; u32 first_id; vec3 max_bound; u32 prim_count}, BVH.hpp:18-27
	std::vector<Node> nodes;
	std::vector<Sphere> prims;              // leaf order (BVH.hpp:201-205)
	std::vector<uint32_t> prim_ids;         // prims[i] == geometry[prim_ids[i]]
	BoundingVolumeHierarchy() {}
	BoundingVolumeHierarchy(std::span<const Sphere> primitives) {
		const uint32_t n = static_cast<uint32_t>(primitives.size());
		nodes.resize(n ? 2 * n - 1 : 1); prims.resize(n); prim_ids.resize(n);
		uint32_t n_nodes = 0;
		b2r_bvh_build(reinterpret_cast<const b2r_sphere*>(primitives.data()), n, nodes.data(), reinterpret_cast<b2r_sphere*>(prims.data()), prim_ids.data(), &n_nodes);
		nodes.resize(n_nodes);
	}
	const Node& root() const { return nodes.front(); }
};
