#!/bin/bash
# Builds oracle/_ref/librefrenderer.so: the reference's OWN Renderer<>::Accumulate / Render (Renderer.hpp), DataStreams.hpp, BVH.hpp,
# Scene.hpp, Camera.hpp, Sampling.hpp, Color.hpp, Primitives.hpp, DataStructures.hpp, Random.hpp, Bitmanip.hpp, Core.hpp compiled by
# g++ from /root/reference. Nothing is copied into the repo: a temporary directory of symlinks to the reference's files is made so
# that quote-includes find the stand-ins for what cannot exist here (ppl.h, Image.h, VectorMath.hpp's SIMD classes, glm, VCL), and
# the files that use MSVC-only SYNTAX get the token-level edits listed below (sed on a temporary copy; no expression is changed).
set -e
REF=${1:-/root/reference}; HERE=$(cd "$(dirname "$0")" && pwd); OUT=$HERE/_ref/librefrenderer.so
T=$(mktemp -d); trap 'rm -rf "$T"' EXIT
for f in Renderer.hpp Scene.hpp Core.hpp Sampling.hpp Random.hpp Bitmanip.hpp Primitives.hpp DataStructures.hpp iacaMarks.h; do ln -s "$REF/$f" "$T/$f"; done
cp "$HERE"/ref_shim/renderer/*.h "$HERE"/ref_shim/renderer/*.hpp "$T/"
sed -n '581,662p' "$REF/VectorMath.hpp" > "$T/vm_scalar.inc"; sed -n '7,19p' "$REF/VectorMath.hpp" > "$T/vm_ints.inc"; sed -n '30,74p' "$REF/Color.hpp" > "$T/color_tonemap.inc"
# BVH.hpp: :88 deducing-this accessor removed (unused); `typename const X` -> `const typename X` (:237,:310,...); `Node::Vector` inside the
# class template needs `typename` (:81,:87)
sed -e '88d' -e 's/typename const /const typename /g' -e 's/sizeof(Node::Vector)/sizeof(typename Node::Vector)/' -e 's/std::pmr::vector<Node::Vector>/std::pmr::vector<typename Node::Vector>/' "$REF/BVH.hpp" > "$T/BVH.hpp"
# Camera.hpp:80 defines a struct in a return type: it gets a name
sed -e 's/^\tstruct { glm::vec3 origin, dir; } generate_ray(/\tstruct RayOD { glm::vec3 origin, dir; }; RayOD generate_ray(/' "$REF/Camera.hpp" > "$T/Camera.hpp"
# DataStreams.hpp: token-level edits are added here as the compiler asks for them
sed -e 's/typename const /const typename /g' "$REF/DataStreams.hpp" > "$T/DataStreams.hpp"
g++ -std=c++23 -fPIC -shared -O2 -mavx2 -mfma -mbmi -mbmi2 -mlzcnt -ffp-contract=off -Wno-attributes -fpermissive -w -pthread -Wl,-Bsymbolic \
    '-D__assume(x)=' -D__vectorcall= -I "$T" -I "$HERE/ref_shim" "$HERE/ref_renderer_wrap.cpp" -o "$OUT"
echo "built $OUT from $REF/Renderer.hpp and the files it includes"
# The reference's GGX build (`#define BRDF 1`, Renderer.hpp:70): the same sources with that one token changed. As shipped it does not
# compile — Renderer.hpp:212 reads `gloss_decay_table[bounce]`, which no file of the reference declares — so the table is supplied here:
# all zeros ("gloss is not reduced on later bounces"), the one input of this build that is not the reference's.
OUT_GGX=$HERE/_ref/librefrenderer_ggx.so
rm "$T/Renderer.hpp"; sed -e 's/^#define BRDF 0/#define BRDF 1/' "$REF/Renderer.hpp" > "$T/Renderer.hpp"
grep -q '^#define BRDF 1' "$T/Renderer.hpp"
echo 'static constexpr float gloss_decay_table[1025] = {};' > "$T/gloss_decay_table.h"
g++ -std=c++23 -fPIC -shared -O2 -mavx2 -mfma -mbmi -mbmi2 -mlzcnt -ffp-contract=off -Wno-attributes -fpermissive -w -pthread -Wl,-Bsymbolic \
    '-D__assume(x)=' -D__vectorcall= -include "$T/gloss_decay_table.h" -I "$T" -I "$HERE/ref_shim" "$HERE/ref_renderer_wrap.cpp" -o "$OUT_GGX"
echo "built $OUT_GGX: the same with #define BRDF 1 and an all-zero gloss_decay_table"
