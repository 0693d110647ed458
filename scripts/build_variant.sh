#!/bin/bash
# usage: scripts/build_variant.sh <name> [extra nvcc defines...]  -> ott-packager_b200/variants/libottgpu_<name>.so (experiments only)
set -e
NAME=$1; shift
HERE=$(cd "$(dirname "$0")/../ott-packager_b200/csrc" && pwd)
OUT=$HERE/../variants; mkdir -p $OUT
TMP=$(mktemp -d)
ARCH="-gencode arch=compute_100a,code=sm_100a"
FLAGS="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-ffp-contract=off"
nvcc $ARCH $FLAGS "$@" -c $HERE/kernels.cu -o $TMP/kernels.o
for f in polyphase plan runtime; do nvcc $ARCH $FLAGS "$@" -x cu -c $HERE/$f.cpp -o $TMP/$f.o & done; wait
nvcc $ARCH -shared -o $OUT/libottgpu_$NAME.so $TMP/*.o -cudart static
rm -rf $TMP
echo built $OUT/libottgpu_$NAME.so
