#!/usr/bin/env python
"""Print the SASS of a kernel from an ncu report with per-instruction samples, restricted to a source line range.
Usage: tools/ncu_sass.py <report> <kernel regex> <lib.so> <file> <line_lo> <line_hi>"""
import csv, os, re, subprocess, sys, tempfile
rep, kre, lib, fname, lo, hi = sys.argv[1:7]
lo, hi = int(lo), int(hi)
csv.field_size_limit(10**9)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(out.splitlines()))
pick = os.environ.get("SMPC_NCU_PICK", "")
s0 = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and pick in r[1]][0]
kname = rows[s0][1]
hdr = rows[s0 + 1]
isamp, isrc, iex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
sass = []
for r in rows[s0 + 2:]:
    if r and r[0] == "Kernel Name":
        break
    if len(r) <= isamp or r[0] == "Address":
        continue
    st = sorted([(int(r[i] or 0), h[6:]) for i, h in stall_cols], reverse=True)[:2]
    sass.append((r[isrc].strip(), int(r[isamp] or 0), int(r[iex] or 0), st))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
cub = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cub], stdout=subprocess.PIPE, text=True).stdout.splitlines()
base = re.search(r"(k_\w+)", kname).group(1)
targs = re.compile(".*".join(("Li%sE" % a) if k == "int" else ("Lb%sE" % a) for k, a in re.findall(r"\((int|bool)\)(\d+)", kname)))
want, cur, lines = False, None, []
for ln in dis:
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m:
        want = base in m.group(1) and targs.search(m.group(1)) is not None
        continue
    if not want:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        lines.append(cur)
for idx, (loc, (ins, smp, ne, st)) in enumerate(zip(lines, sass)):
    if loc and loc[0] == fname and lo <= loc[1] <= hi:
        print("%5d L%-4d %7d smp %9d ex  %-70s %s" % (idx, loc[1], smp, ne, ins[:70], " ".join("%s:%d" % (n, v) for v, n in st if v)))
