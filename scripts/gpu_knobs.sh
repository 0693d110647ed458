#!/bin/bash
# usage: scripts/gpu_knobs.sh <tag> "<ENV=.. ENV=..>" ...   bench cfg2 + cfg1 kernel times under different planner knobs
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
i=0
for KN in "$@"; do
  i=$((i+1))
  for W in cfg2 cfg1; do
    env $KN timeout 300 python bench.py --no-cpu --no-extra --steps 100 --repeats 3 --workload $W > $OUT/${TAG}_${i}_$W.json 2> $OUT/${TAG}_${i}_$W.err
    python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_${i}_$W.json"))
    k=d["roofline"]["kernels"]
    print("$KN $W: value %.0f  " % d["value"], {n:round(v["ms_per_launch"],4) for n,v in k.items()})
except Exception as e:
    print("$KN $W failed", e); print(open("$OUT/${TAG}_${i}_$W.err").read()[-800:])
PY
  done
done
