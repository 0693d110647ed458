#!/bin/bash
# usage: scripts/gpu_test_variants.sh "<pytest args>" <variant names...>
ARGS=$1; shift
for v in "$@"; do
  LIB=$PWD/ott-packager_b200/variants/libottgpu_$v.so
  [ "$v" = "main" ] && LIB=$PWD/ott-packager_b200/libottgpu.so
  echo "== $v"; OTTGPU_LIB=$LIB timeout 600 python -m pytest $ARGS -q -m gpu 2>&1 | tail -3
done
