#!/usr/bin/env python
"""Turns a one-launch `ncu --set full --import-source on` capture into profiles/<name>_ncu.md (+ <name>_traffic.json):
headline metrics, SASS opcode mix, stall reasons, and for ladder_kernel the per-phase instruction budget.
usage: python scripts/profile_r2.py <report.ncu-rep> <out-name> <kernel> <frames-in-capture> <algorithmic-bytes-per-frame>"""
import csv, io, json, os, subprocess, sys, collections

rep, name, kernel, frames, alg = sys.argv[1], sys.argv[2], sys.argv[3], int(sys.argv[4]), float(sys.argv[5])
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
# several launches may have been captured (a step can also run a small thumbnail-class launch of the same kernel): keep
# the longest one = the step's main launch
it_ = h.index("gpu__time_duration.sum")
r = max(rows[2:], key=lambda x: float(x[it_].replace(",", "")))
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
val = {}
for k in keys:
    if k in h:
        i = h.index(k)
        val[k] = (float(r[i].replace(",", "")), units[i])
def mb(k):
    v, u = val[k]
    return v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}[u]
traffic = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stderr=subprocess.DEVNULL).decode()
allrows = list(csv.reader(io.StringIO(src)))
blocks, cur = [], None
for rr in allrows:
    if rr and rr[0] == "Kernel Name":
        cur = []
        blocks.append(cur)
    elif cur is not None:
        cur.append(rr)
def block_instr(b):
    i = b[0].index("Instructions Executed")
    return sum(int(x[i]) for x in b[1:] if len(x) > i and x[i].isdigit())
best = max(blocks, key=block_instr)
srows = [None] + best
sh = srows[1]
ii, isrc, ismp = sh.index("Instructions Executed"), sh.index("Source"), sh.index("# Samples")
stalls = [(x, sh.index(x)) for x in sh if x.startswith("stall_") and "Not Issued" not in x]
ops, st, ti, ts = collections.Counter(), collections.Counter(), 0, 0
seen = set()
for rr in srows[2:]:
    if len(rr) <= ismp or not rr[ii].isdigit():
        continue
    n = int(rr[ii]); ti += n; ts += int(rr[ismp] or 0)
    sp = rr[isrc].split()
    op = (sp[1] if sp[0].startswith("@") else sp[0]).split(".")[0]
    ops[op] += n; seen.add(op)
    for x, i in stalls:
        st[x] += int(rr[i] or 0)
out = os.path.join(ROOT, "profiles", name + "_ncu.md")
with open(out, "w") as f:
    f.write("# %s, `ncu --set full --clock-control none --import-source on`, one launch = %d source frames (%s)\n\n" % (kernel, frames, name))
    f.write("| metric | value |\n|---|---|\n")
    for k in keys:
        if k in val:
            f.write("| %s | %.5g %s |\n" % (k, val[k][0], val[k][1]))
    us = val["gpu__time_duration.sum"][0] * {"usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "us": 1.0, "ms": 1e3, "ns": 1e-3}.get(val["gpu__time_duration.sum"][1], 1.0)
    f.write("| warp instructions per source frame | %.0f k |\n" % (ti / frames / 1e3))
    f.write("| DRAM traffic per launch (read + write) | %.1f MB = %.2f MB per source frame; algorithmic %.2f MB per frame (writes still resident in the 126 MB L2 when the kernel ends are not counted) |\n"
            % (traffic, traffic / frames, alg / 1e6))
    f.write("| duration under ncu (cold cache, serialised) | %.1f us = %.2f us per source frame |\n" % (us, us / frames))
    f.write("\n## SASS opcode mix (warp instructions executed)\n\n| opcode | share |\n|---|---|\n")
    for op, n in ops.most_common(18):
        f.write("| %s | %.1f %% |\n" % (op, 100.0 * n / ti))
    ev = sorted(o for o in seen if o in ("UTMALDG", "UBLKCP", "SYNCS", "IMMA", "IDP", "VABSDIFF4", "ELECT", "I2IP", "LDSM", "UTMASTG"))
    f.write("\nSASS evidence (TMA / mbarrier / tensor / packed ops): %s\n" % ", ".join(ev))
    f.write("\n## Warp stall reasons (share of samples)\n\n| reason | share |\n|---|---|\n")
    for x, v in st.most_common(9):
        f.write("| %s | %.1f %% |\n" % (x[6:], 100.0 * v / max(ts, 1)))
    if "ladder" in kernel:
        f.write("\n## Per-phase instruction budget (nvdisasm line info x executed counts, scripts/phase_budget.py)\n\n")
        pb = subprocess.check_output([sys.executable, os.path.join(ROOT, "scripts", "phase_budget.py"), rep,
                                      os.path.join(ROOT, "ott-packager_b200", "libottgpu.so"), str(frames)]).decode()
        f.write(pb)
json.dump({"capture": name, "kernel": kernel, "channels": frames, "traffic_bytes_per_launch": traffic * 1e6,
           "traffic_bytes_per_frame": traffic * 1e6 / frames}, open(os.path.join(ROOT, "profiles", name + "_traffic.json"), "w"), indent=1)
print("wrote", out)
