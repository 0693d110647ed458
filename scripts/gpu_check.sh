#!/bin/bash
# Runs on the B200 box under gpurun: GPU test tier, smoke, bench, then ncu launch list + one full capture.
# usage: scripts/gpu_check.sh [tag]
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/${TAG}_box.txt 2>&1
nproc >> $OUT/${TAG}_box.txt; lscpu | grep -E "Model name|Socket|Core|Thread|NUMA node\(s\)" >> $OUT/${TAG}_box.txt
nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current,clocks.max.sm --format=csv >> $OUT/${TAG}_box.txt 2>&1
ls /usr/lib/x86_64-linux-gnu 2>/dev/null | grep -E "nvidia-encode|nvcuvid" >> $OUT/${TAG}_box.txt
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/${TAG}_smoke.log
timeout 600 python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench exit $?"
cat $OUT/${TAG}_bench.json; tail -3 $OUT/${TAG}_bench.err
timeout 300 python bench.py --workload cfg1 --cpu-seconds 0.5 --steps 100 > $OUT/${TAG}_bench_cfg1.json 2> $OUT/${TAG}_bench_cfg1.err; cat $OUT/${TAG}_bench_cfg1.json
BENCH_SMALL="python bench.py --steps 4 --warmup 3 --no-cpu"
timeout 300 $BENCH_SMALL > $OUT/${TAG}_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv $BENCH_SMALL > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ladder_kernel -s 3 -c 2 -f -o $OUT/${TAG}_ladder $BENCH_SMALL > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full exit $?"
ls -la $OUT
