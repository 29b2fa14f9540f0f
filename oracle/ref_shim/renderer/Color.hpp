// Stand-in for the reference's Color.hpp, most of which (sRGB packing, Vec8i paths) needs VCL and VectorMath.hpp's SIMD classes and is
// not on the renderer's path. The part Renderer::Render calls — Color.hpp:30-74: ACES_input, ACES_rtt_odt_fit, ACES_output,
// tonemapping (scalar and Vec8f) — is included VERBATIM from the line range cut out at build time. TEST INFRASTRUCTURE ONLY.
#pragma once
#include "VectorMath.hpp"
#include <algorithm>
#include "color_tonemap.inc"
