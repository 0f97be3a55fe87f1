#!/bin/bash
# Builds libnsb200.so (sm_100a only) next to the package.  Usage: build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libnsb200.so"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC -shared --cudart=shared "$@" \
     -o "$OUT" "$HERE/nsb200.cu"
echo "built $OUT"
