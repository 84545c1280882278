#!/usr/bin/env bash
# Build libnngp_b200.so (sm_100a only) in-tree.  Usage: nngp-src_b200/build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
out="$here/libnngp_b200.so"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -shared -cudart static \
  "$@" -o "$out" "$here/csrc/capi.cu" "$here/csrc/encoder.cc"
echo "built $out"
