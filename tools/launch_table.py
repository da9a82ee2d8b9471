#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel shares and
the per-launch table of the first step.  Usage: tools/launch_table.py launches.csv [steps]"""
import csv, re, sys
path = sys.argv[1]; steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lines = [l for l in open(path) if not l.startswith('==')]
rows = list(csv.DictReader(lines))
tot = {}; seq = []
for r in rows:
    name = r['Kernel Name']; v = float(r['Metric Value']); u = r['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    short = re.sub(r'[<(].*', '', name).replace('void ', '').replace('yb::', '')
    tot.setdefault(short, [0, 0]); tot[short][0] += v; tot[short][1] += 1
    m = re.search(r'<(.*?)>', name)
    seq.append((short, m.group(1) if m else '', v, r['Grid Size']))
T = sum(v[0] for v in tot.values())
for k, v in sorted(tot.items(), key=lambda x: -x[1][0]):
    print(f"{k:28s} {v[1]//steps:4d}/step {v[0]/steps:10.1f} us/step {100*v[0]/T:5.1f}%")
print(f"{'total':28s} {len(rows)//steps:4d}/step {T/steps:10.1f} us/step")
per = len(rows) // steps
if '-v' in sys.argv:
    for i, (s, t, v, g) in enumerate(seq[:per]):
        print(f"{i:3d} {s[:24]:24s} {t[:28]:28s} {g:16s} {v:8.1f}")
