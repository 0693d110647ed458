#!/bin/bash
# 16-bit path check: GPU parity tests of the ladder + cfg3 bench (+ optional ncu capture with NCU=1)
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_parity_ladder.py tests/test_parity_channel.py -x -q -m gpu 2>&1 | tail -2
python bench.py --no-cpu --no-extra --steps 50 --repeats 3 --workload cfg3 > $OUT/c3b.json 2>/dev/null
python - <<PY
import json; d=json.load(open("$OUT/c3b.json"))
print("cfg3", round(d["value"]), round(d["ms_per_step"],4), {k:(round(x["ms_per_launch"],4), round(x["frac"],3)) for k,x in d["roofline"]["kernels"].items()}, d["config"].get("channels_per_gpu"), d["run"].get("parity"))
PY
if [ "${NCU:-0}" = "1" ]; then bash scripts/gpu_ncu1.sh c3b cfg3; fi
