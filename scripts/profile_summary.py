#!/usr/bin/env python
"""Turns gpurun_out/<tag>_ladder.ncu-rep (+ launch list CSV) into profiles/<tag>_*.{md,csv,json}.
usage: python scripts/profile_summary.py <tag> <channels-in-capture>"""
import csv, io, json, os, subprocess, sys, collections

tag, channels = sys.argv[1], int(sys.argv[2])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = os.path.join(ROOT, "gpurun_out", tag + "_ladder.ncu-rep")
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
# the capture may hold launches of different classes (main ladder launch, thumbnail launch): keep the biggest grid with the
# longest duration = the main launch of a step, never an average over unlike launches
ig, it = h.index("launch__grid_size"), h.index("gpu__time_duration.sum")
gmax = max(float(r[ig].replace(",", "")) for r in rows[2:])
main = [r for r in rows[2:] if float(r[ig].replace(",", "")) == gmax]
tmax = max(float(r[it].replace(",", "")) for r in main)
main = [r for r in main if float(r[it].replace(",", "")) >= 0.8 * tmax]
def col(name):
    i = h.index(name)
    vals = [float(r[i].replace(",", "")) for r in main]
    return sum(vals) / len(vals), rows[1][i]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
out = {}
for k in keys:
    if k in h:
        v, u = col(k)
        out[k] = {"value": v, "unit": u}
def mb(k):
    v, u = out[k]["value"], out[k]["unit"]
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
traffic_mb = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stderr=subprocess.DEVNULL).decode()
srows = list(csv.reader(io.StringIO(src)))
sh, sd = srows[1], srows[2:]
ia, isrc, ismp = sh.index("Instructions Executed"), sh.index("Source"), sh.index("# Samples")
stalls = [(x, sh.index(x)) for x in sh if x.startswith("stall_") and "Not Issued" not in x]
byop, st, tot, ts = collections.Counter(), collections.Counter(), 0, 0
sass_ops = set()
for r in sd:
    try:
        n = int(r[ia])
    except (ValueError, IndexError):
        continue
    sp = r[isrc].split()
    op = (sp[1] if sp[0].startswith("@") else sp[0])
    sass_ops.add(op.split(".")[0])
    byop[op.split(".")[0]] += n; tot += n; ts += int(r[ismp] or 0)
    for x, i in stalls:
        st[x] += int(r[i] or 0)
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
with open(os.path.join(ROOT, "profiles", tag + "_ladder_ncu.md"), "w") as f:
    f.write("# ladder_kernel, ncu --set full --clock-control none (%s, %d channels = %d source frames per launch)\n\n" % (tag, channels, channels))
    f.write("| metric | value |\n|---|---|\n")
    for k in keys:
        if k in out:
            f.write("| %s | %.4g %s |\n" % (k, out[k]["value"], out[k]["unit"]))
    f.write("| DRAM traffic per launch (read+write) | %.2f MB (%.2f MB per source frame; algorithmic 8.87 MB: rung writes are still in the 126 MB L2 when the kernel ends) |\n" % (traffic_mb, traffic_mb / channels))
    f.write("\n## SASS opcode mix (warp instructions executed)\n\n| opcode | share |\n|---|---|\n")
    for op, n in byop.most_common(16):
        f.write("| %s | %.1f %% |\n" % (op, 100.0 * n / tot))
    f.write("\nTMA / mbarrier evidence in SASS: %s\n" % ", ".join(sorted(o for o in sass_ops if o in ("UTMALDG", "UBLKCP", "SYNCS", "IDP", "VABSDIFF", "ELECT"))))
    f.write("\n## Warp stall reasons (share of samples)\n\n| reason | share |\n|---|---|\n")
    for x, v in st.most_common(9):
        f.write("| %s | %.1f %% |\n" % (x[6:], 100.0 * v / max(ts, 1)))
json.dump({"capture": tag, "channels": channels, "traffic_bytes_per_launch": traffic_mb * 1e6,
           "traffic_bytes_per_frame": traffic_mb * 1e6 / channels},
          open(os.path.join(ROOT, "profiles", tag + "_ladder_traffic.json"), "w"), indent=1)
lc = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
if os.path.exists(lc):
    lines = [l for l in open(lc) if l.startswith('"')]
    open(os.path.join(ROOT, "profiles", tag + "_launches.csv"), "w").writelines(lines)
print("wrote profiles/%s_*" % tag, "traffic MB/launch %.2f" % traffic_mb)
