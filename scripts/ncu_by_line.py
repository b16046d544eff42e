#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an ncu report.
ncu's CSV source page is SASS-only, so the SASS rows are aligned (by order) with `nvdisasm -g` of the
library's cubin, which carries the line table.  usage: ncu_by_line.py report.ncu-rep kernel_substring [top]"""
import csv, re, subprocess, sys, tempfile, os, glob
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))[2:]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "lichee_b200", "libLICHeE_b200.so")], cwd=tmp, capture_output=True)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the kernel's section
start = next(i for i, l in enumerate(sass) if l.startswith("\t.section\t.text.") and kname in l)
end = next((i for i in range(start + 1, len(sass)) if sass[i].startswith("\t.section\t.text.")), len(sass))
line = ("?", 0)
inst = []
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]*)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        inst.append((line, l.split("*/", 1)[1].strip()))
print(f"# {len(rows)} profiled SASS rows, {len(inst)} disassembled instructions", file=sys.stderr)
n = min(len(rows), len(inst))
agg = {}
for (ln, _), r in zip(inst[:n], rows[:n]):
    a = agg.setdefault(ln, [0, 0])
    a[0] += int(r[5]); a[1] += int(r[4])
ti = sum(a[0] for a in agg.values()); ts = sum(a[1] for a in agg.values())
srcs = {}
def text(f, ln):
    if f not in srcs:
        p = os.path.join(root, "lichee_b200", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return srcs[f][ln - 1].strip()[:105] if 0 < ln <= len(srcs[f]) else ""
print(f"total warp instructions {ti}, samples {ts}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f[:14]:14s}{ln:5d} {100*a[0]/ti:5.1f}% inst {100*a[1]/max(ts,1):5.1f}% smp | {text(f, ln)}")
