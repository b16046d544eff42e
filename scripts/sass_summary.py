#!/usr/bin/env python
"""Per-kernel SASS summary of lichee_b200/libLICHeE_b200.so (`cuobjdump -sass`): instruction count and the mnemonics that
show what the kernels are made of.  python scripts/sass_summary.py > profiles/<round>_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "lichee_b200", "libLICHeE_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
cols = ["UBLKCP", "SYNCS", "VOTE", "REDUX", "MATCH", "SHFL", "LDS", "STS", "LDG", "STG", "DADD", "DMUL", "DFMA", "MUFU", "ATOM", "UTCHMMA", "HMMA", "IMMA"]
kern, order = collections.OrderedDict(), []
cur = None
arch = set(re.findall(r"arch = (sm_\w+)", out))
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = name.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        cur = kern.setdefault(name, collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur is not None:
        op = m.group(1)
        cur["instr"] += 1
        for c in cols:
            if op == c or op.startswith(c + ".") or (c == "ATOM" and op.startswith(("ATOM", "RED.", "ATOMS", "ATOMG"))):
                cur[c] += 1
print("# SASS summary of lichee_b200/libLICHeE_b200.so (`cuobjdump -sass`), per kernel: instruction count and the mnemonics that show what")
print("# the kernels are made of.  arch = %s.  UBLKCP = TMA bulk copy (cp.async.bulk), SYNCS = mbarrier, VOTE/REDUX/MATCH/SHFL = warp" % ", ".join(sorted(arch)))
print("# votes, reductions, match and shuffles; no tensor-core instruction (UTC*MMA / HMMA / IMMA) anywhere: nothing on this path is a")
print("# dense contraction.")
print("%-34s %6s " % ("kernel", "instr") + " ".join("%7s" % c for c in cols))
for name, c in kern.items():
    print("%-34s %6d " % (name[:34], c["instr"]) + " ".join("%7d" % c[x] for x in cols))
