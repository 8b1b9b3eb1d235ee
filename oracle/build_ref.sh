#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference sources (read in place from $REF/Cuda, default
# /root/reference) plus oracle/ref_wrap.cu into oracle/_ref/libref_qr.so.
# Outputs go ONLY into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).
# We do not run the reference's CMake; the recipe is three nvcc invocations.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/Cuda" ]; then
    echo "build_ref: $REF/Cuda not present (GPU box?) - keeping prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT" "$OUT/log"
# qr.cu:52 includes a CMake-generated qr_config.h (Cuda/qr_config.h.in:2); the data set is
# an absent LFS blob, so the path is a dummy.
printf '#define QR_JACOBIAN_PATH "%s/jacobians"\n' "$OUT" > "$OUT/qr_config.h"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-std=c++20 -O2 -w -gencode arch=compute_100a,code=sm_100a -rdc=true
       -Xcompiler -fPIC -I"$OUT" -I"$REF/Cuda")
"$NVCC" "${FLAGS[@]}" -c "$REF/Cuda/qr.cu"    -o "$OUT/qr.o" &
"$NVCC" "${FLAGS[@]}" -c "$REF/Cuda/mmult.cu" -o "$OUT/mmult.o" &
"$NVCC" "${FLAGS[@]}" -c "$HERE/ref_wrap.cu"  -o "$OUT/ref_wrap.o" &
wait
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC \
    "$OUT/qr.o" "$OUT/mmult.o" "$OUT/ref_wrap.o" -o "$OUT/libref_qr.so" -lcudart
rm -f "$OUT"/*.o
echo "built $OUT/libref_qr.so"
# Eigen::HouseholderQR timing driver (the library call of C++/main.cpp:54) against the vendored Eigen 3.4.0 headers
if [ -d "$REF/Cuda/QR/Solver/Eigen" ]; then
    g++ -O3 -march=x86-64-v3 -DNDEBUG -std=c++17 -I"$REF/Cuda/QR/Solver" "$HERE/eigen_qr_bench.cpp" -o "$OUT/eigen_qr" \
        || g++ -O3 -DNDEBUG -std=c++17 -I"$REF/Cuda/QR/Solver" "$HERE/eigen_qr_bench.cpp" -o "$OUT/eigen_qr"
    echo "built $OUT/eigen_qr"
fi
