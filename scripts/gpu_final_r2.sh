#!/bin/bash
# Round-2 final evidence on one B200 under gpurun: GPU test tier, smoke, both bench arms, launch list, full ncu captures of the
# main ladder launch (cfg2, cfg4) and of yadif, JPEG and single-channel figures.
set -u
OUT=gpurun_out; mkdir -p $OUT; T=${TAG:-r2f}
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/${T}_pytest.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/${T}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${T}_smoke.log 2>&1; echo "smoke exit $?"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/${T}_bench.json 2> $OUT/${T}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${T}_bench_reference.json 2> $OUT/${T}_bench_reference.err; echo "reference exit $?"
timeout 300 python scripts/bench_jpeg.py > $OUT/${T}_jpeg.jsonl 2> $OUT/${T}_jpeg.err; echo "jpeg exit $?"; cat $OUT/${T}_jpeg.jsonl | cut -c1-330
timeout 300 python scripts/bench_mixed.py --steps 60 > $OUT/${T}_mixed_1gpu.json 2> $OUT/${T}_mixed_1gpu.err; echo "mixed exit $?"
for W in cfg2 cfg1; do timeout 300 python scripts/bench_single_channel.py --workload $W > $OUT/${T}_single_$W.json 2>/dev/null; done
SMALL="python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu --no-extra"
timeout 300 $SMALL > $OUT/${T}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/${T}_launches.csv $SMALL > $OUT/${T}_ncu1.log 2>&1
echo "ncu launches exit $?"
for W in cfg4 cfg2; do
timeout 300 $SMALL --workload $W > $OUT/${T}_plain_$W.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_kernel -s 1 -c 1 -f -o $OUT/${T}_${W}_ladder $SMALL --workload $W > $OUT/${T}_ncu_$W.log 2>&1
echo "ncu ladder $W exit $?"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:yadif_kernel -s 3 -c 1 -f -o $OUT/${T}_cfg4_yadif $SMALL > $OUT/${T}_ncu_yadif.log 2>&1; echo "ncu yadif exit $?"
python - <<PY
import json
d=json.load(open("$OUT/${T}_bench.json")); r=json.load(open("$OUT/${T}_bench_reference.json"))
print("value %.0f e2e %.0f (%s) dev_out %.0f  ref %.0f  ratio e2e %.1f value %.1f" % (d["value"], d["e2e"]["value"], d["e2e"]["variant"], d["e2e_device_out"]["value"], r["value"], d["e2e"]["value"]/r["value"], d["value"]/r["value"]))
print({n:(round(v["ms_per_launch"],4), round(v["frac"],3)) for n,v in d["roofline"]["kernels"].items()}, d["clocks"])
PY
# 16-bit ladder (cfg3) capture as well
timeout 300 $SMALL --workload cfg3 > $OUT/${T}_plain_cfg3.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_kernel -s 1 -c 1 -f -o $OUT/${T}_cfg3_ladder $SMALL --workload cfg3 > $OUT/${T}_ncu_cfg3.log 2>&1
echo "ncu ladder cfg3 exit $?"
