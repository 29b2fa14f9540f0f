// DataStreams.hpp — the parts of the reference's stream types (DataStreams.hpp:74-128) that cross the drop-in boundary: the app
// builds a RayStream<8> by hand for focus picking (Application.cpp:282-298) and passes its Buffer / Hit to
// BoundingVolumeHierarchy::Traverse. Inside the renderer these streams became HBM queues; here they are plain host SoA structs
// with the reference's member names.
#pragma once
#include <cstddef>
#include <cstdint>

template <size_t Size> struct alignas(64) RayStream {
	struct Buffer {
		struct { float x[Size], y[Size], z[Size]; } p, dir;
		struct { float r[Size], g[Size], b[Size]; } radiance, throughput;
		float pdf[Size];
		uint32_t pixelID[Size];
	};
	struct Path {
		Path() : input(&buffers[0]), output(&buffers[1]) {}
		Buffer* input; Buffer* output;
		void swap() noexcept { Buffer* t = input; input = output; output = t; }
		Buffer buffers[2];
	} path;
	struct Hit { float tfar[Size]; int32_t primID[Size]; int32_t matID[Size]; } hit;
	struct ShadowStream {
		struct { float x[Size], y[Size], z[Size]; } p, dir;
		float tfar[Size];
		struct { float r[Size], g[Size], b[Size]; } radiance;
		bool occluded_flag[Size];
	} shadow_rays;
};
