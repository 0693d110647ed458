#!/bin/bash
# converter on the B200 box: full GPU test tier, then the converter micro-benchmark
OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/cv_pytest.log 2>&1; echo "pytest exit $?"
tail -4 $OUT/cv_pytest.log
timeout 600 python scripts/bench_convert.py > $OUT/cv_bench.jsonl 2> $OUT/cv_bench.err; echo "bench_convert exit $?"
cat $OUT/cv_bench.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('%-12s %7.1f us/frame  %6.0f GB/s (%.2f)  e2e %6.0f fps' % (d['pix_fmt'], d['us_per_frame'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value']))
"
tail -5 $OUT/cv_bench.err
