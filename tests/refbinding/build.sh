#!/bin/bash
# Builds tests/refbinding/refbinding: include/b2r_reference_binding.hpp + the reference's own headers (same temporary include tree
# as oracle/ref_renderer_build.sh) linked against libb2r.so. Only where /root/reference exists; the binary travels to the GPU box.
set -e
REF=${1:-/root/reference}; HERE=$(cd "$(dirname "$0")" && pwd); ROOT=$(cd "$HERE/../.." && pwd); OR=$ROOT/oracle
T=$(mktemp -d); trap 'rm -rf "$T"' EXIT
for f in Renderer.hpp Scene.hpp Core.hpp Sampling.hpp Random.hpp Bitmanip.hpp Primitives.hpp DataStructures.hpp iacaMarks.h; do ln -s "$REF/$f" "$T/$f"; done
cp "$OR"/ref_shim/renderer/*.h "$OR"/ref_shim/renderer/*.hpp "$T/"
sed -n '581,662p' "$REF/VectorMath.hpp" > "$T/vm_scalar.inc"; sed -n '7,19p' "$REF/VectorMath.hpp" > "$T/vm_ints.inc"; sed -n '30,74p' "$REF/Color.hpp" > "$T/color_tonemap.inc"
sed -e '88d' -e 's/typename const /const typename /g' -e 's/sizeof(Node::Vector)/sizeof(typename Node::Vector)/' -e 's/std::pmr::vector<Node::Vector>/std::pmr::vector<typename Node::Vector>/' "$REF/BVH.hpp" > "$T/BVH.hpp"
sed -e 's/^\tstruct { glm::vec3 origin, dir; } generate_ray(/\tstruct RayOD { glm::vec3 origin, dir; }; RayOD generate_ray(/' "$REF/Camera.hpp" > "$T/Camera.hpp"
sed -e 's/typename const /const typename /g' "$REF/DataStreams.hpp" > "$T/DataStreams.hpp"
g++ -std=c++23 -O2 -mavx2 -mfma -mbmi -mbmi2 -mlzcnt -ffp-contract=off -Wno-attributes -fpermissive -w -pthread '-D__assume(x)=' -D__vectorcall= \
    -I "$T" -I "$OR/ref_shim" -I "$ROOT/include" "$HERE/refbinding.cpp" -o "$HERE/refbinding" \
    -L "$ROOT/cpu-raytracing-experiments_b200" -lb2r -Wl,-rpath,'$ORIGIN/../../cpu-raytracing-experiments_b200'
echo "built $HERE/refbinding"
