#!/bin/bash
# planner-knob sweep on the B200 box: prints kernel ms per launch for each setting.
# usage: scripts/gpu_sweep.sh [workload] [channels]     (explicit OTTGPU_SRC_CAP_KB + OTTGPU_MID_CAP_KB override the planner's
# step-down list; keep src + 2*mid + ~9 KB of tables under 112 KB or the launch drops to one CTA per SM)
WL=${1:-cfg2}; CH=${2:-32}
OUT=gpurun_out; mkdir -p $OUT
run() { env "$@" python bench.py --workload $WL --channels $CH --cpu-seconds 0.5 --steps 200 > $OUT/sweep.json 2>$OUT/sweep.err; python - "$*" <<'PY'
import json,sys
try:
    d=json.load(open("gpurun_out/sweep.json")); print("%-70s kernel %.4f ms  value %.0f  parity %s" % (sys.argv[1], d["roofline"]["ms_per_launch"], d["value"], str(d["config"]["parity"])[:9]))
except Exception as e: print(sys.argv[1], "FAILED", e, open("gpurun_out/sweep.err").read()[-300:])
PY
}
run OTTGPU_DEFAULTS=1
run OTTGPU_SRC_CAP_KB=48 OTTGPU_MID_CAP_KB=27
run OTTGPU_SRC_CAP_KB=36 OTTGPU_MID_CAP_KB=32
run OTTGPU_SRC_CAP_KB=56 OTTGPU_MID_CAP_KB=24
run OTTGPU_SRC_CAP_KB=64 OTTGPU_MID_CAP_KB=20
run OTTGPU_SRC_CAP_KB=68 OTTGPU_MID_CAP_KB=18
run OTTGPU_SRC_CAP_KB=48 OTTGPU_MID_CAP_KB=27 OTTGPU_MAX_TILE_ROWS=96
run OTTGPU_SRC_CAP_KB=48 OTTGPU_MID_CAP_KB=27 OTTGPU_MAX_TILE_ROWS=192
