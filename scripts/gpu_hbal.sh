#!/bin/bash
# usage: scripts/gpu_hbal.sh <mode:value ...>  -- ladder kernel time per OTT_HBAL_MODE (0 = this tile's own V shape, 1 = plan-wide
# average) and OTT_HBAL (weight of the V share in the H-pair dealing, plan.cpp)
mkdir -p gpurun_out
for W in ${WORKLOADS:-cfg2 cfg4}; do
for mv in "$@"; do
m=${mv%%:*}; v=${mv##*:}
OTT_HBAL_MODE=$m OTT_HBAL=$v python bench.py --no-cpu --no-extra --steps 50 --repeats 3 --workload $W 2>/dev/null > gpurun_out/hbal_${W}_${m}_$v.json
python - <<PY
import json
d=json.load(open("gpurun_out/hbal_${W}_${m}_$v.json")); print("$W mode=$m hbal=$v", round(d["value"]), {k:round(x["ms_per_launch"],4) for k,x in d["roofline"]["kernels"].items()})
PY
done; done
