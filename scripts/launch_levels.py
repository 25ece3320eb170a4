#!/usr/bin/env python
"""Per-level kernel list of one build from an ncu launch-list csv (gpu__time_duration.sum).
usage: launch_levels.py launches.csv [build_index=3] [levels...]"""
import csv, sys
path = sys.argv[1]
build = int(sys.argv[2]) if len(sys.argv) > 2 else 3
want = [int(x) for x in sys.argv[3:]]
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
rows = [(x['Kernel Name'], float(x['Metric Value']), x['Grid Size']) for x in csv.DictReader(lines)
        if x.get('Metric Name') == 'gpu__time_duration.sum']
hem = [r for r in rows if 'hem::' in r[0] or 'cub' in r[0].lower()]
lv, cur = [], []
for r in hem:
    cur.append(r)
    if 'k_publish' in r[0]:
        lv.append(cur); cur = []
# builds: a level whose first kernels include k_boxhist starts a build
starts = [i for i, L in enumerate(lv) if any('k_boxhist' in r[0] for r in L)]
b0 = starts[build]; b1 = starts[build + 1] if build + 1 < len(starts) else len(lv)
tot = 0.0
for k, L in enumerate(lv[b0:b1]):
    s = sum(r[1] for r in L) / 1e3; tot += s
    print("LEVEL %d: %d launches, %.1f us" % (k, len(L), s))
    if not want or k in want:
        for r in L:
            print("   %-64s %8.1f us  %s" % (r[0][:64], r[1] / 1e3, r[2]))
print("build total %.1f us" % tot)
