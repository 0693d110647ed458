#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` export: opcode mix, stall reasons, hottest regions."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ia, isrc, ismp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
stalls = [(h, hdr.index(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = ts = 0
byop, smp, st = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    try:
        n = int(r[ia])
    except (ValueError, IndexError):
        continue
    sp = r[isrc].split()
    op = (sp[1] if sp[0].startswith('@') else sp[0]).split('.')[0]
    byop[op] += n; tot += n
    k = int(r[ismp] or 0); smp[op] += k; ts += k
    for h, i in stalls:
        st[h] += int(r[i] or 0)
print("total warp-instr", tot, "samples", ts)
for op, n in byop.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 16):
    print("%-10s %12d %5.1f%%  samples %5.1f%%" % (op, n, 100.0 * n / tot, 100.0 * smp[op] / max(ts, 1)))
print("stalls:", ", ".join("%s %.1f%%" % (h[6:], 100.0 * v / max(ts, 1)) for h, v in st.most_common(8)))
blk = []
for i in range(0, len(data), 64):
    seg = [r for r in data[i:i + 64] if len(r) > max(ia, ismp)]
    blk.append((sum(int(r[ia] or 0) for r in seg), sum(int(r[ismp] or 0) for r in seg), i))
print("hottest 64-instruction regions (index, %instr, %samples, top ops):")
for n, s, i in sorted(blk, key=lambda b: -b[1])[:12]:
    ops = collections.Counter()
    for r in data[i:i + 64]:
        if len(r) <= isrc or not r[isrc].split():
            continue
        sp = r[isrc].split()
        ops[(sp[1] if sp[0].startswith('@') else sp[0]).split('.')[0]] += 1
    print("  %6d %5.1f%% %5.1f%%  %s" % (i, 100.0 * n / tot, 100.0 * s / max(ts, 1), ops.most_common(5)))
