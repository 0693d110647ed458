#!/bin/bash
# round-3 iteration: GPU parity tests of the ladder + channel, then the ladder kernel's time on cfg2 / cfg4 / cfg1 (+ ncu with NCU=1)
TAG=${1:-v3}
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_parity_ladder.py tests/test_parity_channel.py -x -q -m gpu 2>&1 | tail -2
for W in ${WORKLOADS:-cfg2 cfg4 cfg1}; do
python bench.py --no-cpu --no-extra --steps 50 --repeats 3 --workload $W > $OUT/${TAG}_$W.json 2>$OUT/${TAG}_$W.err
python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_$W.json"))
    print("$W", round(d["value"]), round(d["ms_per_step"],4), {k:(round(x["ms_per_launch"],4), round(x["frac"],3)) for k,x in d["roofline"]["kernels"].items()}, d["run"].get("parity"))
except Exception as e:
    print("$W failed", e); print(open("$OUT/${TAG}_$W.err").read()[-1500:])
PY
done
if [ "${NCU:-0}" = "1" ]; then bash scripts/gpu_ncu1.sh $TAG ${NCUW:-cfg2}; fi
