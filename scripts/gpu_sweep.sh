#!/bin/bash
# planner-knob sweep on the B200 box: prints kernel ms per launch for each setting
OUT=gpurun_out; mkdir -p $OUT
run() { env "$@" python bench.py --no-cpu --steps 200 > $OUT/sweep.json 2>$OUT/sweep.err; python - "$*" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/sweep.json")); print("%-70s kernel %.3f ms  value %.0f  parity %s" % (sys.argv[1], d["roofline"]["ms_per_launch"], d["value"], d["config"]["parity"][:9]))
except Exception as e: print(sys.argv[1], "FAILED", e, open("gpurun_out/sweep.err").read()[-300:])
PY
}
run OTTGPU_SRC_CAP_KB=36 OTTGPU_MID_CAP_KB=36 OTTGPU_MAX_TILE_ROWS=128
run OTTGPU_SRC_CAP_KB=36 OTTGPU_MID_CAP_KB=36 OTTGPU_MAX_TILE_ROWS=256
run OTTGPU_SRC_CAP_KB=36 OTTGPU_MID_CAP_KB=36 OTTGPU_MAX_TILE_ROWS=64
run OTTGPU_SRC_CAP_KB=36 OTTGPU_MID_CAP_KB=36 OTTGPU_MAX_TILE_ROWS=96
run OTTGPU_SRC_CAP_KB=30 OTTGPU_MID_CAP_KB=40 OTTGPU_MAX_TILE_ROWS=128
run OTTGPU_SRC_CAP_KB=24 OTTGPU_MID_CAP_KB=24 OTTGPU_MAX_TILE_ROWS=128
run OTTGPU_SRC_CAP_KB=16 OTTGPU_MID_CAP_KB=16 OTTGPU_MAX_TILE_ROWS=128
run OTTGPU_SRC_CAP_KB=64 OTTGPU_MID_CAP_KB=64 OTTGPU_MAX_TILE_ROWS=256
