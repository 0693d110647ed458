#!/bin/bash
# usage: scripts/gpu_clocks.sh <workload> <variant...>  -- phase clock printouts of experiment builds (build_variant.sh <name> -DOTTGPU_PHASE_CLOCKS ...)
W=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  OTTGPU_LIB=$PWD/ott-packager_b200/variants/libottgpu_$v.so python bench.py --no-cpu --no-extra --steps 3 --warmup 3 --repeats 1 --workload $W 2>&1 | grep -E "^cta" > gpurun_out/clk_$v.txt
  python - <<PY
import re
acc={}; n=0; prod={}; np_=0
for l in open("gpurun_out/clk_$v.txt"):
    kv=re.findall(r"(\w+) (\d+)", l)
    if "PRODUCER" in l:
        np_+=1
        for k,x in kv:
            if k!="cta": prod[k]=prod.get(k,0)+int(x)
    else:
        n+=1
        for k,x in kv:
            if k not in ("cta","warp"): acc[k]=acc.get(k,0)+int(x)
print("$v compute:", {k:round(x/max(n,1)) for k,x in acc.items()}, " producer:", {k:round(x/max(np_,1)) for k,x in prod.items()})
PY
done
