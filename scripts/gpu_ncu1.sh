#!/bin/bash
# one full ncu capture of the main ladder launch: scripts/gpu_ncu1.sh <tag> <workload> [extra bench args]
TAG=$1; W=$2; shift 2
OUT=gpurun_out; mkdir -p $OUT
SMALL="python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu --no-extra --workload $W $*"
timeout 300 $SMALL > $OUT/${TAG}_plain_$W.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_kernel -s 1 -c 1 -f -o $OUT/${TAG}_${W}_ladder $SMALL > $OUT/${TAG}_ncu_$W.log 2>&1
echo "ncu ladder $W exit $?"
