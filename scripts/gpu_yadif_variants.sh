#!/bin/bash
# usage: scripts/gpu_yadif_variants.sh <tag> <variants...>  -- GPU parity of the channel path, then yadif / ladder kernel times of each build on cfg4
TAG=$1; shift
timeout 300 python -m pytest tests/test_parity_channel.py -x -q -m gpu 2>&1 | tail -1
BENCH_ARGS="--no-extra --workload cfg4 --repeats 3" bash scripts/gpu_variants.sh $TAG "$@" > /dev/null
for v in "$@"; do python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_$v.json")); print("$v", round(d["value"]), {k:round(x["ms_per_launch"],4) for k,x in d["roofline"]["kernels"].items()})
PY
done
