#!/bin/bash
# Builds libt3d_b200.so (sm_100a only) in-tree next to the Python package.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libt3d_b200.so"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr \
     -Xcompiler -fPIC -shared ${T3D_NVCC_EXTRA} "$HERE/t3d_api.cu" -o "$OUT"
echo "built $OUT"
