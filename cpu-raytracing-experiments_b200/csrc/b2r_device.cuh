// b2r_device.cuh — sm_100a kernels of the wavefront path tracer.
//
// Data layout in HBM (DESIGN.md "Layout"):
//   path queue (x2, ping-pong), SoA planes over the queue slot i:
//       A[i] = {o.x, o.y, o.z, d.x}   B[i] = {d.y, d.z, pdf, as_float(pid)}   T[c*cap + i] = throughput channel c
//     = 44 B per path (RayStream<>::Buffer, DataStreams.hpp:75-88, minus radiance). pid = slot << 26 | t with t the
//     pixel's tile-order index (tile*256 + ID) and slot the sample-in-flight index inside the batch.
//   radiance RAD[(slot*3 + c)*npix + t]: a path's running radiance lives at its pixel (one path per pixel per sample)
//     and is touched only when a contribution arrives (unoccluded light sample, emissive hit, sky).
//   hit queue H[i] = {tfar, as_float(prim)} and shadow queue SA/SB/SL: BVH pipeline only.
//   buckets ACC[(k*3 + c)*npix + t]: running sums per median-of-means bucket (AccumulationTile, Renderer.hpp:43-46).
// All kernels are persistent (grid = SM count x resident CTAs, looping over a device-side count), so a whole batch —
// max_bounces rounds — is enqueued, or replayed as one CUDA graph, without host synchronisation.
#pragma once
#include <cuda_runtime.h>
#include "b2r_shade.h"

namespace b2r {

constexpr int kBruteTile = 1024;       // spheres staged per shared-memory tile in the brute-force kernels
constexpr int kBlock = 256;            // threads per CTA, shading kernels
constexpr int kTravBlock = 128;        // threads per CTA, traversal kernels

// ---------------------------------------------------------------------------------------------- device-only helpers
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }

__device__ __forceinline__ void stat_add(unsigned long long* stats, int which, uint32_t v) {
	// one atomic per warp: REDUX.SUM over the lanes, lane 0 publishes
	const uint32_t s = __reduce_add_sync(0xffffffffu, v);
	if (lane_id() == 0 && s) atomicAdd(stats + which, static_cast<unsigned long long>(s));
}

__device__ __forceinline__ PathState load_path(const QueueDev& q, int side, uint32_t i) {
	const float4 a = q.A[side][i], b = q.B[side][i];
	PathState s; s.ox = a.x; s.oy = a.y; s.oz = a.z; s.dx = a.w; s.dy = b.x; s.dz = b.y; s.pdf = b.z; s.pid = __float_as_uint(b.w);
	const float* t = q.T[side];
	s.tr = t[i]; s.tg = t[q.cap + i]; s.tb = t[2u * q.cap + i];
	return s;
}

__device__ __forceinline__ void store_path(const QueueDev& q, int side, uint32_t i, const PathState& s) {
	q.A[side][i] = make_float4(s.ox, s.oy, s.oz, s.dx);
	q.B[side][i] = make_float4(s.dy, s.dz, s.pdf, __uint_as_float(s.pid));
	float* t = q.T[side];
	t[i] = s.tr; t[q.cap + i] = s.tg; t[2u * q.cap + i] = s.tb;
}

// Append `keep` threads of the CTA to a queue: one atomic per CTA (warp ballots -> shared prefix -> single atomicAdd).
// Returns the destination index (valid when keep). Must be called by all threads of the CTA.
__device__ __forceinline__ uint32_t block_append(bool keep, uint32_t* counter, uint32_t* s_warp /*[blockDim/32 + 1]*/) {
	const uint32_t ballot = __ballot_sync(0xffffffffu, keep);
	const uint32_t warp = threadIdx.x >> 5, lane = lane_id(), n_warps = blockDim.x >> 5;
	if (lane == 0) s_warp[warp] = __popc(ballot);
	__syncthreads();
	if (threadIdx.x == 0) {
		uint32_t run = 0;
		for (uint32_t w = 0; w < n_warps; w++) { const uint32_t c = s_warp[w]; s_warp[w] = run; run += c; }
		s_warp[n_warps] = run ? atomicAdd(counter, run) : 0u;
	}
	__syncthreads();
	const uint32_t dst = s_warp[n_warps] + s_warp[warp] + __popc(ballot & ((1u << lane) - 1u));
	__syncthreads();  // s_warp is reused by the next call
	return dst;
}

// ---------------------------------------------------------------------------------------------- brute-force pipeline
// One fused kernel per bounce: intersect all spheres (BVH.hpp:311-318 as shipped, USEBVH false) -> closest-hit shader
// -> light sample + inline any-hit -> emissive -> BRDF sample / roulette -> compacted append to the next queue.
// Spheres are staged in shared memory (tiles of kBruteTile) and read as warp-uniform broadcasts.
template <bool FIRST, bool COUNT>
__global__ void __launch_bounds__(kBlock, 2) k_bounce_brute(const Params p, const uint32_t bounce) {
	__shared__ float4 s_prim[kBruteTile];
	__shared__ uint32_t s_warp[kBlock / 32 + 1];
	const SceneDev& sc = p.scene;
	const uint32_t n_in = FIRST ? p.batch->n_slots * p.frame.npix : p.cnt.paths[bounce];
	const int side = bounce & 1;
	const uint32_t n_tiles = (sc.n_prims + kBruteTile - 1) / kBruteTile;
	const bool mis = !(p.frame.flags & B2R_FLAG_NO_MIS);
	const bool last = bounce + 1 >= p.frame.max_bounces;
	uint32_t c_shadow = 0, c_hits = 0, c_term = 0, c_drop = 0, c_events = 0, c_sphere = 0;

	if (n_tiles == 1) {  // whole scene fits: stage once per CTA
		for (uint32_t j = threadIdx.x; j < sc.n_prims; j += blockDim.x) s_prim[j] = ldg4(sc.prims + j);
		__syncthreads();
	}
	for (uint32_t base = blockIdx.x * kBlock; base < n_in; base += gridDim.x * kBlock) {
		const uint32_t i = base + threadIdx.x;
		const bool live = i < n_in;
		PathState s;
		if (live) s = FIRST ? primary_path(p.frame, p.batch->acc[i / p.frame.npix], i / p.frame.npix, i % p.frame.npix) : load_path(p.q, side, i);
		// ---- closest hit over every sphere, ties to the lowest BVH-order index (strict <, Q6)
		float best = FLT_MAX; int32_t prim = -1;
		for (uint32_t tile = 0; tile < n_tiles; tile++) {
			const uint32_t first = tile * kBruteTile, cnt = min(static_cast<uint32_t>(kBruteTile), sc.n_prims - first);
			if (n_tiles > 1) {
				__syncthreads();
				for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) s_prim[j] = ldg4(sc.prims + first + j);
				__syncthreads();
			}
			if (live) {
				for (uint32_t j = 0; j < cnt; j++) {
					const float4 sp = s_prim[j]; float d;
					if (sphere_hit_closest(sp.x, sp.y, sp.z, sp.w, s.ox, s.oy, s.oz, s.dx, s.dy, s.dz, &d) && d < best) { best = d; prim = static_cast<int32_t>(first + j); }
				}
				if (COUNT) c_sphere += cnt;
			}
		}
		// ---- shade
		bool keep = false, want_shadow = false, emissive = false, hit = live && prim >= 0;
		Surface sf; ShadowRay sr; f3 e_add{0.0f, 0.0f, 0.0f};
		uint32_t acc = 0, seed = 0;
		if (hit) {
			acc = p.batch->acc[s.pid >> 26]; seed = pixel_seed(s.pid & kPixMask, p.frame.max_bounces);
			sf = shade_surface(sc, s, best, prim);
			c_hits++;
			if (!last) {  // at the last bounce the whole radiance of a surviving hit path is dropped (Q11): nothing to add
				if (mis) want_shadow = shade_light_sample(sc, sf, s, prim, acc, seed, bounce, &sr);
				emissive = sf.emissive;
				if (emissive) e_add = shade_emission(sc, sf, s, best, bounce, mis);
			}
		}
		// ---- shadow ray: any hit along [0, tfar) (BVH.hpp:290-305)
		if (n_tiles == 1) {
			if (want_shadow) {
				c_shadow++;
				if (COUNT) c_sphere += sc.n_prims;
				for (uint32_t j = 0; j < sc.n_prims; j++) {
					const float4 sp = s_prim[j];
					if (sphere_hit_any(sp.x, sp.y, sp.z, sp.w, sr.o.x, sr.o.y, sr.o.z, sr.d.x, sr.d.y, sr.d.z, sr.tfar)) { want_shadow = false; break; }
				}
			}
		} else {
			if (want_shadow) c_shadow++;
			for (uint32_t tile = 0; tile < n_tiles; tile++) {
				const uint32_t first = tile * kBruteTile, cnt = min(static_cast<uint32_t>(kBruteTile), sc.n_prims - first);
				__syncthreads();
				for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) s_prim[j] = ldg4(sc.prims + first + j);
				__syncthreads();
				if (want_shadow) {
					if (COUNT) c_sphere += cnt;
					for (uint32_t j = 0; j < cnt; j++) {
						const float4 sp = s_prim[j];
						if (sphere_hit_any(sp.x, sp.y, sp.z, sp.w, sr.o.x, sr.o.y, sr.o.z, sr.d.x, sr.d.y, sr.d.z, sr.tfar)) { want_shadow = false; break; }
					}
				}
			}
		}
		// ---- contributions: unoccluded light sample first, then emission (order of Renderer.hpp:304-353)
		if (hit) {
			if (last) { rad_zero(p.rad, p.frame.npix, s.pid); c_drop++; }
			else {
				if (want_shadow || emissive) {
					const f3 l = want_shadow ? sr.L : f3{0.0f, 0.0f, 0.0f};
					rad_add(p.rad, p.frame.npix, s.pid, l, e_add); c_events++;
				}
				keep = shade_continue(sf, &s, acc, seed, bounce);
				if (!keep) c_term++;  // roulette: path ends here, radiance stays at the pixel (Renderer.hpp:379-381,424-430)
			}
		} else if (live) {  // miss (Renderer.hpp:408-420)
			c_term++;
			if (sc.has_ambient) { rad_add(p.rad, p.frame.npix, s.pid, shade_sky(sc, s), f3{0.0f, 0.0f, 0.0f}); c_events++; }
		}
		const uint32_t dst = block_append(keep, p.cnt.paths + bounce + 1, s_warp);
		if (keep) store_path(p.q, side ^ 1, dst, s);
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.cnt.stats + ST_EXT, static_cast<unsigned long long>(n_in));
	stat_add(p.cnt.stats, ST_SHADOW, c_shadow); stat_add(p.cnt.stats, ST_HITS, c_hits); stat_add(p.cnt.stats, ST_TERM, c_term);
	stat_add(p.cnt.stats, ST_DROPPED, c_drop); stat_add(p.cnt.stats, ST_EVENTS, c_events);
	if (COUNT) stat_add(p.cnt.stats, ST_SPHERE, c_sphere);
}

// camera rays of a batch -> queue side 0 (Renderer.hpp:97-127)
__global__ void __launch_bounds__(kBlock) k_generate(const Params p) {
	const uint32_t n = p.batch->n_slots * p.frame.npix;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
		store_path(p.q, 0, i, primary_path(p.frame, p.batch->acc[i / p.frame.npix], i / p.frame.npix, i % p.frame.npix));
	if (blockIdx.x == 0 && threadIdx.x == 0) p.cnt.paths[0] = n;
}
// closest-hit traversal of queue side (bounce & 1): persistent warps fetch 32 rays at a time
template <bool COUNT>
__global__ void __launch_bounds__(kTravBlock) k_intersect_closest(const Params p, const uint32_t bounce) {
	const uint32_t n_in = p.cnt.paths[bounce];
	const int side = bounce & 1;
	uint32_t c_sphere = 0, c_box = 0;
	for (;;) {
		uint32_t base = 0;
		if (lane_id() == 0) base = atomicAdd(p.cnt.work_a + bounce, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= n_in) break;
		const uint32_t i = base + lane_id();
		if (i < n_in) {
			const float4 a = p.q.A[side][i], b = p.q.B[side][i];
			const Ray r{a.x, a.y, a.z, a.w, b.x, b.y};
			float best; int32_t prim;
			traverse_closest<COUNT>(p.scene.wide, r, &best, &prim, &c_sphere, &c_box);
			p.q.H[i] = make_float2(best, __int_as_float(prim));
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.cnt.stats + ST_EXT, static_cast<unsigned long long>(n_in));
	if (COUNT) { stat_add(p.cnt.stats, ST_SPHERE, c_sphere); stat_add(p.cnt.stats, ST_BOX, c_box); }
}
// shade the hit records: light sample -> shadow queue, emission, BRDF sample / roulette -> next path queue
__global__ void __launch_bounds__(kBlock, 2) k_shade(const Params p, const uint32_t bounce) {
	__shared__ uint32_t s_warp[kBlock / 32 + 1];
	const SceneDev& sc = p.scene;
	const uint32_t n_in = p.cnt.paths[bounce];
	const int side = bounce & 1;
	const bool mis = !(p.frame.flags & B2R_FLAG_NO_MIS);
	const bool last = bounce + 1 >= p.frame.max_bounces;
	uint32_t c_hits = 0, c_term = 0, c_drop = 0, c_events = 0;
	for (uint32_t base = blockIdx.x * kBlock; base < n_in; base += gridDim.x * kBlock) {
		const uint32_t i = base + threadIdx.x;
		const bool live = i < n_in;
		PathState s; float depth = FLT_MAX; int32_t prim = -1;
		if (live) { s = load_path(p.q, side, i); const float2 h = p.q.H[i]; depth = h.x; prim = __float_as_int(h.y); }
		bool keep = false, want_shadow = false; ShadowRay sr;
		const bool hit = live && prim >= 0;
		if (hit) {
			const uint32_t acc = p.batch->acc[s.pid >> 26], seed = pixel_seed(s.pid & kPixMask, p.frame.max_bounces);
			const Surface sf = shade_surface(sc, s, depth, prim);
			c_hits++;
			if (last) { rad_zero(p.rad, p.frame.npix, s.pid); c_drop++; }  // Q11
			else {
				if (mis) want_shadow = shade_light_sample(sc, sf, s, prim, acc, seed, bounce, &sr);
				if (sf.emissive) { rad_add(p.rad, p.frame.npix, s.pid, shade_emission(sc, sf, s, depth, bounce, mis), f3{0.0f, 0.0f, 0.0f}); c_events++; }
				keep = shade_continue(sf, &s, acc, seed, bounce);
				if (!keep) c_term++;
			}
		} else if (live) {
			c_term++;
			if (sc.has_ambient) { rad_add(p.rad, p.frame.npix, s.pid, shade_sky(sc, s), f3{0.0f, 0.0f, 0.0f}); c_events++; }
		}
		const uint32_t pid = s.pid;  // shade_continue keeps pid
		const uint32_t sdst = block_append(want_shadow, p.cnt.shadow + bounce, s_warp);
		if (want_shadow) {
			p.q.SA[sdst] = make_float4(sr.o.x, sr.o.y, sr.o.z, sr.d.x);
			p.q.SB[sdst] = make_float4(sr.d.y, sr.d.z, sr.tfar, __uint_as_float(pid));
			p.q.SL[sdst] = sr.L.x; p.q.SL[p.q.cap + sdst] = sr.L.y; p.q.SL[2u * p.q.cap + sdst] = sr.L.z;
		}
		const uint32_t dst = block_append(keep, p.cnt.paths + bounce + 1, s_warp);
		if (keep) store_path(p.q, side ^ 1, dst, s);
	}
	stat_add(p.cnt.stats, ST_HITS, c_hits); stat_add(p.cnt.stats, ST_TERM, c_term);
	stat_add(p.cnt.stats, ST_DROPPED, c_drop); stat_add(p.cnt.stats, ST_EVENTS, c_events);
}
// shadow rays of this bounce: any-hit traversal, unoccluded light samples are added to the pixel's radiance
template <bool COUNT>
__global__ void __launch_bounds__(kTravBlock) k_intersect_shadow(const Params p, const uint32_t bounce) {
	const uint32_t n_in = p.cnt.shadow[bounce];
	uint32_t c_sphere = 0, c_box = 0, c_events = 0;
	for (;;) {
		uint32_t base = 0;
		if (lane_id() == 0) base = atomicAdd(p.cnt.work_b + bounce, 32u);
		base = __shfl_sync(0xffffffffu, base, 0);
		if (base >= n_in) break;
		const uint32_t i = base + lane_id();
		if (i < n_in) {
			const float4 a = p.q.SA[i], b = p.q.SB[i];
			const Ray r{a.x, a.y, a.z, a.w, b.x, b.y};
			if (!traverse_any<COUNT>(p.scene.wide, r, b.z, &c_sphere, &c_box)) {
				const f3 L{p.q.SL[i], p.q.SL[p.q.cap + i], p.q.SL[2u * p.q.cap + i]};
				rad_add(p.rad, p.frame.npix, __float_as_uint(b.w), L, f3{0.0f, 0.0f, 0.0f}); c_events++;
			}
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.cnt.stats + ST_SHADOW, static_cast<unsigned long long>(n_in));
	stat_add(p.cnt.stats, ST_EVENTS, c_events);
	if (COUNT) { stat_add(p.cnt.stats, ST_SPHERE, c_sphere); stat_add(p.cnt.stats, ST_BOX, c_box); }
}

// ---------------------------------------------------------------------------------------------- accumulate + resolve
// Fold the batch's per-sample radiance into its median-of-means bucket, in sample order (Renderer.hpp:424-430 adds one
// sample at a time; bucket = acc % K, :82), and clear RAD for the next batch.
__global__ void __launch_bounds__(kBlock) k_accumulate(const Params p) {
	const uint32_t n = 3u * p.frame.npix, npix = p.frame.npix, slots = p.batch->n_slots;
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const uint32_t c = i / npix, t = i - c * npix;
		for (uint32_t s = 0; s < slots; s++) {
			const uint32_t k = p.batch->acc[s] % p.frame.buckets;
			float* src = p.rad + (static_cast<size_t>(s) * 3u + c) * npix + t;
			float* dst = p.acc + (static_cast<size_t>(k) * 3u + c) * npix + t;
			*dst += *src; *src = 0.0f;
		}
	}
}
// median of K bucket sums. K == 5 is the reference's network (Sampling.hpp:13-21); other K are this repo's definition:
// odd K -> middle order statistic, even K -> mean of the two middle ones.
__device__ __forceinline__ float median_buckets(const float* acc, uint32_t K, uint32_t stride, uint32_t t) {
	if (K == 5) return median_of_5(acc[t], acc[stride + t], acc[2u * stride + t], acc[3u * stride + t], acc[4u * stride + t]);
	if (K == 3) return median_of_3(acc[t], acc[stride + t], acc[2u * stride + t]);
	if (K == 1) return acc[t];
	float v[64];
	for (uint32_t k = 0; k < K; k++) {  // insertion sort
		float x = acc[static_cast<size_t>(k) * stride + t]; int j = static_cast<int>(k) - 1;
		while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; j--; }
		v[j + 1] = x;
	}
	return (K & 1u) ? v[K / 2u] : (v[K / 2u - 1u] + v[K / 2u]) * 0.5f;
}
// Renderer::Render, Renderer.hpp:436-478: one thread per pixel (tile order in, raster RGBA out)
__global__ void __launch_bounds__(kBlock) k_resolve(const Params p, float4* __restrict__ fb, const float scale, const int tonemap) {
	const uint32_t npix = p.frame.npix, K = p.frame.buckets;
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < npix; t += gridDim.x * blockDim.x) {
		float r = scale * median_buckets(p.acc, K, 3u * npix, t);
		float g = scale * median_buckets(p.acc + npix, K, 3u * npix, t);
		float b = scale * median_buckets(p.acc + 2u * npix, K, 3u * npix, t);
		if (tonemap) aces_tonemap(&r, &g, &b);
		int32_t x, y; pixel_xy(t, p.frame.h_tiles, &x, &y);
		fb[static_cast<size_t>(y) * p.frame.width + x] = make_float4(r, g, b, 1.0f);
	}
}

// ---------------------------------------------------------------------------------------------- taps
__global__ void k_tap_generate(const Params p, const uint32_t acc, float* __restrict__ out) {
	for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < p.frame.npix; t += gridDim.x * blockDim.x) {
		Pcg rng{hash_2d(acc, pixel_seed(t, p.frame.max_bounces))};
		const float s0 = rng.next_unit(), s1 = rng.next_unit();
		int32_t x, y; pixel_xy(t, p.frame.h_tiles, &x, &y);
		const f3 d = camera_dir(p.frame.cam, x, y, s0, s1);
		float* o = out + static_cast<size_t>(t) * 6u;
		o[0] = p.frame.cam.px; o[1] = p.frame.cam.py; o[2] = p.frame.cam.pz; o[3] = d.x; o[4] = d.y; o[5] = d.z;
	}
}
// Traverse / Traverse_shadow on caller-supplied rays (Application.cpp:282-298)
__global__ void k_tap_trace(const SceneDev sc, const float* __restrict__ rays, const float* __restrict__ tfar_in, const uint32_t n,
                            const int use_bvh, const int shadow, float* __restrict__ tfar_out, int32_t* __restrict__ prim_out, uint8_t* __restrict__ occ_out) {
	for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		const float* rr = rays + static_cast<size_t>(i) * 6u;
		const Ray r{rr[0], rr[1], rr[2], rr[3], rr[4], rr[5]};
		uint32_t cs = 0, cb = 0;
		if (!shadow) {
			float best = FLT_MAX; int32_t prim = -1;
			if (use_bvh) traverse_closest<false>(sc.wide, r, &best, &prim, &cs, &cb);
			else for (uint32_t j = 0; j < sc.n_prims; j++) {
				const float4 sp = ldg4(sc.prims + j); float d;
				if (sphere_hit_closest(sp.x, sp.y, sp.z, sp.w, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, &d) && d < best) { best = d; prim = static_cast<int32_t>(j); }
			}
			tfar_out[i] = best; prim_out[i] = prim;
		} else {
			bool occ = false;
			if (use_bvh) occ = traverse_any<false>(sc.wide, r, tfar_in[i], &cs, &cb);
			else for (uint32_t j = 0; j < sc.n_prims && !occ; j++) {
				const float4 sp = ldg4(sc.prims + j);
				occ = sphere_hit_any(sp.x, sp.y, sp.z, sp.w, r.ox, r.oy, r.oz, r.dx, r.dy, r.dz, tfar_in[i]);
			}
			occ_out[i] = occ ? 1 : 0;
		}
	}
}

}  // namespace b2r
