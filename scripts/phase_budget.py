#!/usr/bin/env python
"""Per-phase instruction / stall budget of ladder_kernel from an `ncu --set full --import-source on` capture.

Joins ncu's SASS page (instructions executed + stall samples per instruction) with `nvdisasm -g` line info of the same
libottgpu.so, buckets every instruction by the source function of its innermost line, and prints a table.
usage: python scripts/phase_budget.py <report.ncu-rep> [libottgpu.so] [frames-in-capture] [out.md]"""
import collections, csv, io, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep = sys.argv[1]
lib = os.path.abspath(sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "ott-packager_b200", "libottgpu.so"))
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 8
out_md = sys.argv[4] if len(sys.argv) > 4 else None
KERNEL = os.environ.get("OTT_PHASE_KERNEL", "ladder_kernelILi1")   # instance in the cubin (mangled): ILi1 = 8-bit (ncu: ladder_kernel<1>), ILi2 = 16-bit
NCU_KERNEL = "ladder_kernel"

# ---- source line -> enclosing function (top-level definitions of the .cuh / .cu files)
def function_map(path):
    m, cur = {}, "?"
    rx = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:OTT_DEV|__global__|__device__|inline|static)\b.*?\b([A-Za-z_][A-Za-z0-9_]*)\s*\(")
    for i, line in enumerate(open(path), 1):
        if line and not line[0].isspace():
            g = rx.match(line)
            if g and not line.rstrip().endswith(";"):
                cur = g.group(1)
        m[i] = cur
    return m
fmaps = {}
def func_of(path, line):
    if path not in fmaps:
        fmaps[path] = function_map(path) if os.path.exists(path) else {}
    return fmaps[path].get(line, os.path.basename(path))

PHASE = [("H pass (IMMA)", ("hpass8_mma", "hpass16_mma", "ott_mma_u8s8", "ott_mma_u8u8", "ott_pack_s16", "phase_hpass")),
         ("H pass (16-bit dp2a fallback)", ("hpass16", "hdot16", "hpass16il", "hdot16il", "hpass16_any", "hpass16il_any")),
         ("staging fix-up (16-bit)", ("phase_fixup16", "shift_mask16", "split_bytes16")),
         ("V pass taps (fast path)", ("vrow_bias8", "vrow_fma8", "ott_fma_rd", "ott_as_float", "ott_as_u32", "ott_i2f", "ott_hi16x2", "load_words", "sx_lo", "sx_hi", "ott_pack_u8")),
         ("V pass taps (exact / 16-bit rows)", ("vcolumns", "load_vals", "pack4", "pack2", "ott_clip")),
         ("V pass rows + stores", ("vpass", "phase_vpass", "vpass_store", "store_words", "zip8", "zip16", "ott_prmt")),
         ("copy items", ("copy_rows", "interleave_rows8", "item_copy")),
         ("staging (register path)", ("phase_stage", "load_vec_slow")),
         ("tables + item setup (producer)", ("phase_tables", "item_prepare")),
         ("kernel frame / sync / producer", ("ladder_kernel", "mbar_wait", "mbar_wait_backoff", "mbar_arrive", "mbar_init", "mbar_expect_tx", "mbar_add_tx",
                                             "bulk_copy_g2s", "tma_box_3d", "compute_sync", "fence_proxy_async", "fence_mbar_init", "smem_u32"))]
def phase_of(fn):
    for name, fns in PHASE:
        if fn in fns:
            return name
    return "other: " + fn

# ---- SASS offsets -> (file, line)
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(tmp) if f.startswith("kernels")][0]
dis = subprocess.check_output(["nvdisasm", "-g", "-c", os.path.join(tmp, cub)], stderr=subprocess.DEVNULL).decode()
inside, loc, offs = False, ("?", 0), {}
for line in dis.splitlines():
    if line.startswith("//---") and ".text." in line:
        inside = KERNEL in line
        continue
    if not inside:
        continue
    g = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if g:
        loc = (g.group(1), int(g.group(2)))
        continue
    g = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if g:
        offs[int(g.group(1), 16)] = (loc, g.group(2).strip())

# ---- ncu SASS page
src = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], stderr=subprocess.DEVNULL).decode()
rows = list(csv.reader(io.StringIO(src)))
# several kernels/launches may be in the report: take the first block for the kernel
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}
        blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
blk = max((b for b in blocks if NCU_KERNEL in b["name"]), key=lambda b: sum(int(x[5] or 0) for x in b["rows"][1:] if len(x) > 5 and x[5].isdigit()))
h, data = blk["rows"][0], blk["rows"][1:]
ia, ii, ismp = h.index("Address"), h.index("Instructions Executed"), h.index("# Samples")
stall_cols = [(x[6:], h.index(x)) for x in h if x.startswith("stall_") and "Not Issued" not in x]
base = int(data[0][ia], 16)
ph_inst, ph_smp, ph_stall, ph_ops = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter), collections.defaultdict(collections.Counter)
tot_i = tot_s = 0
for r in data:
    try:
        n, smp = int(r[ii]), int(r[ismp] or 0)
    except (ValueError, IndexError):
        continue
    off = int(r[ia], 16) - base
    (path, line), text = offs.get(off, (("?", 0), "?"))
    ph = phase_of(func_of(path, line))
    op = text.split()[1] if text.startswith("@") else text.split()[0]
    ph_inst[ph] += n; ph_smp[ph] += smp; tot_i += n; tot_s += smp
    ph_ops[ph][op.split(".")[0]] += n
    for name, c in stall_cols:
        ph_stall[ph][name] += int(r[c] or 0)
lines = []
lines.append("| phase | warp-instr / frame | share | stall-sample share | top opcodes | top stalls |")
lines.append("|---|---|---|---|---|---|")
for ph, n in ph_inst.most_common():
    ops = ", ".join("%s %.0f%%" % (o, 100.0 * c / n) for o, c in ph_ops[ph].most_common(5)) if n else ""
    st = ", ".join("%s %.0f%%" % (o, 100.0 * c / max(ph_smp[ph], 1)) for o, c in ph_stall[ph].most_common(3))
    lines.append("| %s | %.0f k | %.1f %% | %.1f %% | %s | %s |" % (ph, n / frames / 1e3, 100.0 * n / tot_i, 100.0 * ph_smp[ph] / max(tot_s, 1), ops, st))
lines.append("| total | %.0f k | | | %d SASS instructions in the kernel image | |" % (tot_i / frames / 1e3, len(offs)))
text = "\n".join(lines)
print(text)
if out_md:
    open(out_md, "w").write(text + "\n")
