#!/bin/sh
# oracle/_ref: the reference's OWN native code for this path, usable as a checker.
# The only native component of ParallelTemperingMonteCarlo.jl is the binary-only Fortran library
# MachineLearningPotential/lib/librunnerjulia.so (`forward`).  It needs libblas.so.3 / libgfortran.so.5 / libquadmath,
# none of which are on the system path; the copies bundled with the image's Python wheels are ABI compatible
# (SURVEY §8c).  Nothing is compiled: the binary is copied out of the read-only reference tree next to the shims.
# The Julia part of the reference cannot be built or run here (no julia binary, no network).
set -e
here=$(cd "$(dirname "$0")" && pwd)
ref=${PTMC_REFERENCE:-/root/reference}
out="$here/_ref"
src="$ref/MachineLearningPotential/lib/librunnerjulia.so"
[ -f "$src" ] || { echo "build_ref: $src not present (fine on the GPU box: oracle/_ref travels prebuilt)"; exit 0; }
mkdir -p "$out"
cp -f "$src" "$out/librunnerjulia.so"
chmod u+w "$out/librunnerjulia.so"
site=$(python3 -c 'import sysconfig; print(sysconfig.get_paths()["purelib"])')
libs="$site/opencv_python_headless.libs"
# symlinks into the image's site-packages (same image, same paths on the GPU box)
ln -sf "$(ls "$libs"/libopenblasp-r0-*.so | head -1)" "$out/libblas.so.3"
ln -sf "$libs/libgfortran-83c28eba.so.5.0.0" "$out/libgfortran.so.5"
ln -sf "$libs/libgfortran-83c28eba.so.5.0.0" "$out/libgfortran-83c28eba.so.5.0.0"
for q in "$libs"/libquadmath-*.so.*; do ln -sf "$q" "$out/$(basename "$q")"; done
echo "build_ref: $out ready"
