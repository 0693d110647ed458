#!/bin/bash
# Round-2 multi-GPU record on one box with N GPUs: the default bench line (cfg4) and the mixed 64-channel box (cfg5)
N=${1:-8}; OUT=gpurun_out; mkdir -p $OUT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-extra > $OUT/r2f_cfg4_${N}gpu.json 2> $OUT/r2f_cfg4_${N}gpu.err; echo "cfg4 exit $?"
timeout 400 $TR --master-port 29612 scripts/bench_mixed.py --gpus $N --steps 60 > $OUT/r2f_mixed_${N}gpu.json 2> $OUT/r2f_mixed_${N}gpu.err; echo "mixed exit $?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/r2f_cfg4_${N}gpu.json"))
    print("cfg4 N=$N: value %.0f  e2e %.0f (%s, %.2f ms)  dev_out %.0f" % (d["value"], d["e2e"]["value"], d["e2e"]["variant"], d["e2e"]["ms_per_step"], d["e2e_device_out"]["value"]))
except Exception as e: print("cfg4 failed", e); print(open("$OUT/r2f_cfg4_${N}gpu.err").read()[-1200:])
try:
    print("mixed N=$N:", open("$OUT/r2f_mixed_${N}gpu.json").read()[:700])
except Exception as e: print("mixed failed", e)
PY
