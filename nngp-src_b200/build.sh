#!/usr/bin/env bash
# Build libnngp_b200.so (sm_100a only) in-tree.  Usage: nngp-src_b200/build.sh [extra nvcc flags]
# The hash of the sources (csrc/*, include/*.h, this script) is embedded as nngp_build_id().
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
PY="${PYTHON:-python3}"
out="${OUT:-$here/libnngp_b200.so}"
id="$("$PY" "$here/nngp_b200/_build.py" --hash)"
"$NVCC" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
  -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xcompiler -pthread -shared -cudart static \
  -DNNGP_BUILD_ID="\"$id\"" \
  "$@" -o "$out.tmp" "$here/csrc/capi.cu" "$here/csrc/encoder.cc"
mv -f "$out.tmp" "$out"
echo "built $out (build id $id)"
