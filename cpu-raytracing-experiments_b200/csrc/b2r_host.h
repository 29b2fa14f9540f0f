// b2r_host.h — host-side scene preparation shared by the C ABI (b2r.cu): BVH construction with the reference's
// node/leaf order, and its flattening into the 128-byte 4-wide node layout the traversal kernels read.
#pragma once
#include <stdint.h>
#include <vector>
#include <vector_types.h>
#include "../../include/b2r.h"

namespace b2r {

// 128-byte traversal node: four 32-byte child slots, each two float4.
//   inner child : {lo.x, lo.y, lo.z, hi.x} {hi.y, hi.z, link = wide-node index (>= 0), 0}
//   leaf child  : {c.x,  c.y,  c.z,  r^2 } {0,    0,    link = ~prim_index      (< 0),  0}   (sphere inlined: no leaf fetch)
//   empty slot  : link = kEmptyLink
// Child boxes are the reference's binary-node boxes (BVH.hpp:18-27) padded outward by kBoxPad so that the float slab
// test stays conservative; the spheres are the reference's prims (BVH leaf order), bit-exact.
struct alignas(128) WideNode { float slot[4][8]; };
static_assert(sizeof(WideNode) == 128, "one node = one 128-byte line");
constexpr int32_t kEmptyLink = INT32_MIN;
constexpr int kTraversalStack = 64;  // entries per thread; upload fails with B2R_ERR_BVH if a tree needs more

struct WideBvh {
	std::vector<WideNode> nodes;   // nodes[0] = root, breadth-first (top of the tree is contiguous)
	uint32_t max_stack = 0;        // worst-case traversal stack occupancy for this tree
	uint32_t tn_bits = 10;         // low bits of a 32-bit stack entry that carry the entry distance (rest: node index)
	uint32_t depth = 0;
	std::vector<uint32_t> level_first;  // nodes of BFS level l are [level_first[l], level_first[l+1]); children always sit on a deeper level
	double cost = 0.0;             // sum of the inner-slot half areas (what a refit is compared against)
	std::vector<uint32_t> geom_of_prim;  // for b2r_refit_scene: geometry index of the sphere each BVH-order leaf index stands for (empty: unknown)
};

// BoundingVolumeHierarchy<Sphere> constructor (BVH.hpp:90-206), bit-identical node and leaf order.
void build_reference_bvh(const b2r_sphere* geometry, uint32_t n, std::vector<b2r_bvh_node>& nodes,
                         std::vector<b2r_sphere>& prims, std::vector<uint32_t>& prim_ids, uint32_t log_cluster_size = 0u, float cost_ratio = 1.0f);
// Checks the invariants the flattening relies on (children adjacent, leaf size 1, indices in range). Returns false if malformed.
bool validate_reference_bvh(const b2r_bvh_node* nodes, uint32_t n_nodes, uint32_t n_prims);
// A second binary tree over the same spheres (reference leaf order kept in the leaf links), built for traversal speed.
void build_traversal_tree(const b2r_sphere* prims_bvh_order, uint32_t n, std::vector<b2r_bvh_node>& nodes);
// Collapse the binary tree 2 -> 4 wide and inline the leaf spheres.
void flatten_bvh(const b2r_bvh_node* nodes, uint32_t n_nodes, const b2r_sphere* prims, uint32_t n_prims, WideBvh& out);
// Sum of the inner-slot half areas of a flattened tree (WideBvh::cost; k_tree_cost computes the same sum on the device after a refit).
double wide_cost(const WideBvh& tree);
// prims (BVH leaf order) as a permutation of geometry (original order), matched by value — the reference's BVH keeps reordered COPIES
// of the spheres and no index map (BVH.hpp:201-205). geom_of_prim[i] = index into geometry of prims[i]; equal spheres are paired in
// index order. Returns false when prims is not a permutation of geometry.
bool match_prims_to_geometry(const b2r_sphere* prims, const b2r_sphere* geometry, uint32_t n, std::vector<uint32_t>& geom_of_prim);


// Scene arrays in the packed form the kernels read (SceneDev): spheres {c.xyz, r^2} in BVH leaf order, per-material
// {albedo, emissive flag} and {emission}, per-light {sphere of scene.geometry[light]} and {emission, light_primID}.
struct PackedScene {
	std::vector<float4> prims, mat_albedo, mat_emission, light_sphere, light_emit;
	std::vector<int32_t> prim_mat;
};
void pack_scene(const b2r_sphere* prims, uint32_t n_prims, const b2r_material* materials, uint32_t n_mat,
                const int32_t* light_geom_idx, uint32_t n_lights, const b2r_sphere* geometry, PackedScene& out);

}  // namespace b2r
