#!/usr/bin/env python
"""SASS census of libsmpc_cuda.so: per kernel, instruction totals and the mnemonics that matter (profiles/r02_sass_census.txt).
usage: python tools/sass_census.py > profiles/r02_sass_census.txt"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "simple-mpc_b200", "libsmpc_cuda.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), stdout=subprocess.PIPE, text=True).stdout.split("\n")
cols = ["DFMA", "DMUL", "DADD", "MUFU", "MUFU.RSQ64H", "MUFU.RCP64H", "LDS", "STS", "LDG", "STG", "BAR", "UBLKCP", "SYNCS", "SHFL", "DMMA"]
print("# SASS census of simple-mpc_b200/libsmpc_cuda.so (cuobjdump -sass, sm_100a), per kernel: instruction totals and the mnemonics that matter")
print("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier try_wait/arrive, DFMA/DMUL/DADD = FP64 pipe, MUFU.RSQ64H = rsqrt() pivots,")
print("# LDS/STS = shared memory, BAR = __syncthreads, SHFL = warp shuffles. No UTC*MMA / LDTM: the path is fp64 and tcgen05 has no fp64 kind.")
print("%-110s %9s" % ("kernel", "total") + "".join(" %9s" % c for c in cols))
blocks = re.split(r"\s+Function : ", sass)[1:]
for name, blk in zip(names, blocks):
    ins = re.findall(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", blk, flags=re.M)
    if not ins:
        continue
    cnt = {c: 0 for c in cols}
    for i in ins:
        base = i.split(".")[0]
        if base in cnt:
            cnt[base] += 1
        if i.startswith("MUFU.RSQ64H"):
            cnt["MUFU.RSQ64H"] += 1
        if i.startswith("MUFU.RCP64H"):
            cnt["MUFU.RCP64H"] += 1
    short = re.sub(r"smpc::|\(anonymous namespace\)::|smpc_ctx::", "", name)
    print("%-110s %9d" % (short[:110], len(ins)) + "".join(" %9d" % cnt[c] for c in cols))
