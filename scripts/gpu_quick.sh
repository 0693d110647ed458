#!/bin/bash
# quick iteration on the B200 box: GPU tests (fast subset unless FULL=1), short bench, ncu full capture of ladder_kernel
set -u
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
if [ "${FULL:-0}" = "1" ]; then
  timeout 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
else
  timeout 600 python -m pytest tests/test_parity_ladder.py tests/test_parity_channel.py -x -q -m gpu -k "not digest" > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?"
fi
tail -3 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --cpu-seconds 0.5 --steps 300 ${BENCH_ARGS:-} > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
python - <<PY
import json
try:
    d=json.load(open("$OUT/${TAG}_bench.json"))
    print("value %.0f fps  ms/step %.3f  kernel %.3f ms  roofline %.1f GB/s (%.3f)  e2e %.0f  e2e_host %.0f  parity %s clocks %s" % (d["value"], d["ms_per_step"], d["roofline"]["ms_per_launch"], d["roofline"]["achieved"], d["roofline"]["frac"], d["e2e"]["value"], d["e2e_host_out"]["value"], d["run"]["parity"], d["clocks"]))
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/${TAG}_bench.err").read()[-2000:])
PY
if [ "${NCU:-1}" = "1" ]; then
BENCH_SMALL="python bench.py --steps 4 --warmup 3 --channels 8 --no-cpu ${BENCH_ARGS:-}"
timeout 300 $BENCH_SMALL > $OUT/${TAG}_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_kernel -s 3 -c 1 -f -o $OUT/${TAG}_ladder $BENCH_SMALL > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"
fi
