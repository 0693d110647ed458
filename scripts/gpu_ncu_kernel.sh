#!/bin/bash
# usage: scripts/gpu_ncu_kernel.sh <tag> <kernel-regex> <bench args...>   -- one full ncu capture of one kernel
TAG=$1; K=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
CMD="python bench.py --steps 4 --warmup 3 --channels ${CHANNELS:-8} --no-cpu $*"
timeout 300 $CMD > $OUT/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o $OUT/${TAG} $CMD > $OUT/${TAG}_ncu.log 2>&1
echo "ncu exit $?"
