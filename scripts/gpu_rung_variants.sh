#!/bin/bash
# usage: scripts/gpu_rung_variants.sh "<rung_cost args>" <variants...>
ARGS=$1; shift
for v in "$@"; do
  LIB=$PWD/ott-packager_b200/variants/libottgpu_$v.so
  [ "$v" = "main" ] && LIB=$PWD/ott-packager_b200/libottgpu.so
  echo "== $v"; OTTGPU_LIB=$LIB timeout 300 python scripts/rung_cost.py $ARGS 2>&1 | tail -8
done
