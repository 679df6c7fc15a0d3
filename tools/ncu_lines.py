#!/usr/bin/env python
"""Correlate the stall samples of an ncu report with CUDA source lines (ncu's CLI only prints the
kernel's own file; inlined headers are lost). Usage:
    tools/ncu_lines.py <report.ncu-rep> <kernel regex> <lib.so> [top N]
Reads the SASS page of the report and the line table of the same cubin (nvdisasm -g), joins them by
instruction order, prints samples per (file:line) and per inlined-function region."""
import csv, os, re, subprocess, sys, tempfile, collections

def main():
    rep, kre, lib = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    csv.field_size_limit(10**9)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    # the page holds one block per matching launch; SMPC_NCU_PICK selects a substring of the demangled name
    pick = os.environ.get("SMPC_NCU_PICK", "")
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and pick in r[1]]
    if not starts:
        sys.exit("no kernel matches")
    s0 = starts[0]
    kname = rows[s0][1]
    hdr = rows[s0 + 1]
    isamp = hdr.index("# Samples")
    isrc = hdr.index("Source")
    iexec = hdr.index("Instructions Executed")
    iwf = hdr.index("L1 Wavefronts Shared") if "L1 Wavefronts Shared" in hdr else None
    iwfi = hdr.index("L1 Wavefronts Shared Ideal") if "L1 Wavefronts Shared Ideal" in hdr else None
    sass = []
    for r in rows[s0 + 2:]:
        if r and r[0] == "Kernel Name":
            break  # next launch
        if len(r) <= isamp or r[0] == "Address":
            continue
        sass.append((r[isrc].strip(), int(r[isamp] or 0), int(r[iexec] or 0), int(r[iwf] or 0) if iwf is not None else 0,
                     int(r[iwfi] or 0) if iwfi is not None else 0))
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
    # mangled name fragment from the demangled kernel name
    base = re.search(r"(k_\w+)", kname).group(1)
    # the library holds one cubin per translation unit: take the one whose symbol table has the kernel
    cubs = sorted([os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")], key=os.path.getsize, reverse=True)
    cub = next((c for c in cubs if base.encode() in open(c, "rb").read()), cubs[0])
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], stdout=subprocess.PIPE, text=True).stdout.splitlines()
    # template arguments in order, as they appear in the mangled name (Li36ELi24E..., Lb1E)
    targs = re.compile(".*".join(("Li%sE" % a) if k == "int" else ("Lb%sE" % a) for k, a in re.findall(r"\((int|bool)\)(\d+)", kname)))
    sect, cur, lines = None, None, []
    want = None
    for ln in dis:
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            nm = m.group(1)
            ok = re.search(r"\d+%s[IE]" % base, nm) is not None and targs.search(nm) is not None
            want = ok
            continue
        if not want:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            lines.append((cur, m.group(2).strip()))
    if len(lines) != len(sass):
        print("warning: instruction counts differ: disasm %d vs ncu %d" % (len(lines), len(sass)))
    per = collections.Counter()
    ex = collections.Counter()
    tot = 0
    wf = collections.Counter()
    wfi = collections.Counter()
    for (loc, _), (_, smp, ne, w, wi) in zip(lines, sass):
        per[loc] += smp
        ex[loc] += ne
        wf[loc] += w
        wfi[loc] += wi
        tot += smp
    print("kernel:", kname, " total samples:", tot)
    for loc, s in per.most_common(top):
        print("%6.2f%%  %8d samples  %10d inst  %s:%s  smem_wavefronts %d ideal %d" % (100.0 * s / max(tot, 1), s, ex[loc], loc[0] if loc else "?", loc[1] if loc else "?", wf[loc], wfi[loc]))

if __name__ == "__main__":
    main()
