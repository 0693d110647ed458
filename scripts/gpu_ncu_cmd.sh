#!/bin/bash
# usage: scripts/gpu_ncu_cmd.sh <tag> <kernel-regex> <skip> <command...>   -- plain run, then one full ncu capture of one launch
TAG=$1; K=$2; SKIP=$3; shift 3
OUT=gpurun_out; mkdir -p $OUT
timeout 300 "$@" > $OUT/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o $OUT/${TAG} "$@" > $OUT/${TAG}_ncu.log 2>&1
echo "ncu exit $?"
