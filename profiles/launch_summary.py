#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and the round series."""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
seq = [(r[4].split('(')[0].replace('void ', '').replace('b200fit::', ''), float(r[-1]) / 1e3) for r in rows]
tot = collections.defaultdict(lambda: [0, 0.0])
for n, t in seq:
    tot[n][0] += 1
    tot[n][1] += t
allt = sum(t for _, t in seq)
print("%d launches, %.1f us total (ncu: cold-cache, serialised -- compare SHARES)" % (len(seq), allt))
for n, (c, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-44s n=%4d total %9.1f us (%5.1f%%)  mean %8.2f us" % (n[:44], c, t, 100 * t / allt, t / c))
for tag in ('<1>', '<2>'):
    ev = [t for n, t in seq if n.startswith('fit_eval_kernel' + tag)]
    so = [t for n, t in seq if n.startswith('fit_solve_kernel' + tag)]
    if ev:
        pick = [0, 1, 2, 5, 10, 20, 40, 80, 160, len(ev) - 1]
        print("rounds%s=%d  eval us @round:" % (tag, len(ev)), {i: round(ev[i], 1) for i in pick if i < len(ev)})
        print("            solve us @round:", {i: round(so[i], 1) for i in pick if i < len(so)})
