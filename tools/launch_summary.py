#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`) of bench.py:
    python tools/launch_summary.py gpurun_out/launches.csv [paths_per_step] > profiles/rNN_launches.txt
Launches are grouped by kernel name and grid size (the whole-batch launches of the timed step and the per-chunk
launches of the host pipeline have different grids); the step's kernels are the ones of the whole-batch group."""
import csv
import re
import statistics
import sys
from collections import defaultdict

path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors='replace')) if len(r) > 14 and r[0].isdigit()]
byname = defaultdict(list)
for r in rows:
    name = re.sub(r'^void ', '', r[4])
    name = re.sub(r'\(.*$', '', name).replace('<unnamed>::', '')
    byname[name].append(float(r[14]) / 1e3)
print(f'# {path}: per-launch device time (us; under ncu: cold-cache, serialised)')
STEP = ['uam_k_bin_hist', 'uam_k_bin_scan', 'uam_k_bin_scatter', 'uam_k_score_groups', 'uam_k_reduce_paths', 'uam_k_best']
# a step kernel is launched on the whole batch (the timed step) and on the chunks of the host pipeline (e2e leg, a
# quarter of the batch each): the whole-batch launches are the ones above 45 % of that kernel's longest launch
whole, chunk, other = {}, {}, {}
for name, v in byname.items():
    if any(name.startswith(k) for k in STEP):
        m = max(v)
        whole[name] = [x for x in v if x > 0.45 * m]
        if len(whole[name]) == 1 and len(v) > 1 or len(v) == 1:      # a one-off launch (set-up call), not a step kernel
            other[name] = whole.pop(name)
            continue
        rest = [x for x in v if x <= 0.45 * m]
        if rest:
            chunk[name] = rest
    else:
        other[name] = v
order = lambda d: sorted(d, key=lambda n: [n.startswith(k) for k in STEP].index(True))
tot = sum(statistics.median(v) for v in whole.values())
print('# whole-batch launches (the timed step):')
for n in order(whole):
    v = whole[n]
    print(f'{n:34s} n = {len(v):3d}  median {statistics.median(v):9.1f} us  share of the step {statistics.median(v) / tot:.3f}')
print(f'sum of the step kernels {tot:9.1f} us')
print('# host-pipeline chunk launches (e2e leg):')
for n in order(chunk):
    print(f'{n:34s} n = {len(chunk[n]):3d}  median {statistics.median(chunk[n]):9.1f} us')
print('# one-off builds, waypoint mode:')
for n in sorted(other):
    print(f'{n:34s} n = {len(other[n]):3d}  median {statistics.median(other[n]):9.1f} us')
