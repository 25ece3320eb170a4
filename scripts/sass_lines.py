"""Attribute an ncu SASS source page to CUDA source lines (the CLI's CUDA view carries no metrics).
usage: python scripts/sass_lines.py <report.ncu-rep> <kernel regex (mangled name in the cubin)> <lib.so> [top N] [ncu kernel-name regex]
Needs the library the report was captured from (built with -lineinfo)."""
import csv
import os
import re
import subprocess
import sys
import tempfile

rep, kern, lib = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
kern_ncu = sys.argv[5] if len(sys.argv) > 5 else kern
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# per function: list of (line, sass)
funcs, cur, line = {}, None, ("", 0)
for l in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m and cur:
        funcs[cur].append((line, m.group(2)))
name = [f for f in funcs if re.search(kern, f)]
assert name, list(funcs)[:20]
ins = funcs[name[0]]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern_ncu], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = [i for i, r in enumerate(rows) if "Instructions Executed" in r][0]
h = rows[hi]
ie, ss = h.index("Instructions Executed"), h.index("# Samples")
body = [r for r in rows[hi + 1:] if len(r) > ie]
print(f"# {name[0]}: {len(ins)} SASS instructions in the cubin, {len(body)} in the report")
n = min(len(ins), len(body))
agg = {}
tot_i = tot_s = 0.0
for (ln, _), r in zip(ins[:n], body[:n]):
    a = agg.setdefault(ln, [0.0, 0.0])
    a[0] += float(r[ie] or 0)
    a[1] += float(r[ss] or 0)
    tot_i += float(r[ie] or 0)
    tot_s += float(r[ss] or 0)
srcdir = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc")
srcs = {f: open(os.path.join(srcdir, f)).read().splitlines() for f in os.listdir(srcdir) if f.endswith((".cuh", ".cu", ".inl"))}
print(f"# total warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
print("# inst%  stall%  line  source")
for ln, (i, s) in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    src = srcs.get(ln[0], [])
    t = src[ln[1] - 1].strip()[:100] if 0 < ln[1] <= len(src) else ""
    print(f"{100 * i / tot_i:6.2f} {100 * s / max(tot_s, 1):6.2f} {ln[0][:16]:>16s}:{ln[1]:<5d} {t}")
