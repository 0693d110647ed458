#!/bin/bash
# Round-2 evidence run on one B200 under gpurun: GPU test tier, smoke, both bench arms, ncu launch list + full captures.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/r2_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/r2_pytest.log; tail -3 $OUT/r2_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r2_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/r2_smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/r2_bench.json 2> $OUT/r2_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/r2_bench_reference.json 2> $OUT/r2_bench_reference.err; echo "reference exit $?"
SMALL="python bench.py --steps 2 --warmup 3 --repeats 1 --no-cpu --no-extra"
timeout 300 $SMALL > $OUT/r2_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $OUT/r2_launches.csv $SMALL > $OUT/r2_ncu1.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_kernel -s 2 -c 4 -f -o $OUT/r2_cfg4_ladder $SMALL > $OUT/r2_ncu2.log 2>&1; echo "ncu ladder cfg4 exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:yadif_kernel -s 3 -c 1 -f -o $OUT/r2_cfg4_yadif $SMALL > $OUT/r2_ncu3.log 2>&1; echo "ncu yadif exit $?"
timeout 300 $SMALL --workload cfg2 > $OUT/r2_plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_kernel -s 2 -c 4 -f -o $OUT/r2_cfg2_ladder $SMALL --workload cfg2 > $OUT/r2_ncu4.log 2>&1; echo "ncu ladder cfg2 exit $?"
ls -la $OUT | grep r2_
