#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT
for sz in 1920x1080 3840x2160; do
timeout 600 python scripts/bench_convert.py --size $sz > $OUT/cv_bench_$sz.jsonl 2> $OUT/cv_bench.err; echo "bench_convert $sz exit $?"
cat $OUT/cv_bench_$sz.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('%-12s %7.1f us/frame  %6.0f GB/s (%.2f)  e2e %6.0f fps' % (d['pix_fmt'], d['us_per_frame'], d['roofline']['achieved'], d['roofline']['frac'], d['e2e']['value']))
"
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:convert_kernel -c 400 --csv --log-file $OUT/cv_launches.csv python scripts/bench_convert.py --steps 20 --pictures 24 > $OUT/cv_ncu.log 2>&1; echo "ncu exit $?"
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open("gpurun_out/cv_launches.csv")) if len(r) > 10]
hdr = rows[0]; i_id, i_m, i_v = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value")
per = collections.defaultdict(dict)
for r in rows[1:]:
    per[int(r[i_id])][r[i_m]] = float(r[i_v].replace(",", ""))
ids = sorted(per)
# launches come in groups per pixel format: 10 warm-up + 20 timed + 6 + 10 e2e = 46 per format
n = len(ids) // 4
for k, name in enumerate(["yuv420p", "yuv422p", "yuv420p10le", "yuv422p10le"]):
    grp = [per[i] for i in ids[k * n + 10:k * n + 30]]
    t = sorted(g["gpu__time_duration.sum"] for g in grp)[len(grp) // 2]
    rd = sum(g["dram__bytes_read.sum"] for g in grp) / len(grp); wr = sum(g["dram__bytes_write.sum"] for g in grp) / len(grp)
    print("%-12s median kernel %.2f us  dram read %.2f MB write %.2f MB" % (name, t / 1000.0, rd / 1e6, wr / 1e6))
PY
