#!/bin/bash
# JPEG + link-floor check on one B200
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_parity_jpeg.py tests/test_host_glue.py -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python scripts/bench_jpeg.py > $OUT/r2f_jpeg.jsonl 2>/dev/null; cut -c1-380 $OUT/r2f_jpeg.jsonl
python bench.py --no-cpu --no-extra --steps 20 --warmup 5 > $OUT/lf2.json 2>/dev/null
python - <<PY
import json; d=json.load(open("$OUT/lf2.json"))
for k in ("e2e_device_out","e2e_host_out"):
    e=d[k]; print(k, round(e["value"]), round(e["ms_per_step"],3), round(e["link_floor_ms_per_step"],3), round(e["link_floor_share"],3))
PY
