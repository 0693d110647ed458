#!/bin/bash
# Round-2 check run on one B200 under gpurun: GPU test subset, bench (both arms optional), full ncu captures of the MAIN ladder launch
set -u
TAG=${1:-r2b}
OUT=gpurun_out; mkdir -p $OUT
if [ "${FULL:-0}" = "1" ]; then
  timeout 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
else
  timeout 600 python -m pytest tests/test_parity_ladder.py tests/test_parity_channel.py -x -q -m gpu -k "not digest" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
fi
tail -3 $OUT/${TAG}_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS:-} > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench.json"))
    k=d["roofline"]["kernels"]
    print("value %.0f fps  ms/step %.3f  e2e %.0f (%s)  dev_out %.0f host_out %.0f  parity %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("variant"), d["e2e_device_out"]["value"], d["e2e_host_out"]["value"], d["run"]["parity"]))
    print({n:(round(v["ms_per_launch"],4), round(v["frac"],3)) for n,v in k.items()})
    for w,v in d.get("workloads",{}).items(): print(w, round(v["value"]), {n:(round(x["ms_per_launch"],4), round(x["frac"],3)) for n,x in v["kernels"].items()}, v.get("parity"))
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/${TAG}_bench.err").read()[-2000:])
PY
if [ "${NCU:-1}" = "1" ]; then
SMALL="python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu --no-extra"
for W in cfg4 cfg2; do
timeout 300 $SMALL --workload $W > $OUT/${TAG}_plain_$W.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_kernel -s 1 -c 1 -f -o $OUT/${TAG}_${W}_ladder $SMALL --workload $W > $OUT/${TAG}_ncu_$W.log 2>&1
echo "ncu ladder $W exit $?"
done
fi
