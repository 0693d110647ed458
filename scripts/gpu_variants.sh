#!/bin/bash
# usage: scripts/gpu_variants.sh <tag> <variant names...>   (bench each ott-packager_b200/variants/libottgpu_<name>.so; "main" = the product build)
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for v in "$@"; do
  LIB=$PWD/ott-packager_b200/variants/libottgpu_$v.so
  [ "$v" = "main" ] && LIB=$PWD/ott-packager_b200/libottgpu.so
  OTTGPU_LIB=$LIB timeout 300 python bench.py --no-cpu --steps 200 ${BENCH_ARGS:-} > $OUT/${TAG}_$v.json 2> $OUT/${TAG}_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_$v.json"))
    r=d["roofline"]
    print("$v: value %.0f fps  kernel %.4f ms  frac %.3f  %s" % (d["value"], r["ms_per_launch"], r["frac"], d["run"].get("parity")))
except Exception as e:
    print("$v: failed", e); print(open("$OUT/${TAG}_$v.err").read()[-1500:])
PY
done
