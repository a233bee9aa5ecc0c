#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv --print-source sass,cuda` per source region.
usage: python profiles/src_summary.py mix.csv 'fit_kernel<(int)2>' [top_n]"""
import csv
import sys

path, kern = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
csv.field_size_limit(10 ** 9)
cur_file = cur_fn = hdr = None
agg = {}
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and r[0].isdigit() and kern in (cur_fn or ''):
        ix = {h: i for i, h in enumerate(hdr)}
        try:
            vals = [int(r[ix[k]] or 0) for k in ('Instructions Executed', '# Samples', 'stall_no_inst',
                                                 'Thread Instructions Executed')]
        except (ValueError, IndexError):
            continue
        a = agg.setdefault((cur_file, int(r[0])), [0, 0, r[1][:100], 0, 0])
        a[0] += vals[0]
        a[1] += vals[1]
        a[3] += vals[2]
        a[4] += vals[3]
tot_i = sum(a[0] for a in agg.values()) or 1
tot_s = sum(a[1] for a in agg.values()) or 1
print("kernel %s: total warp-inst %d, samples %d" % (kern, tot_i, tot_s))
byfile = {}
for (f, l), a in agg.items():
    b = byfile.setdefault(f, [0, 0, 0])
    b[0] += a[0]; b[1] += a[1]; b[2] += a[3]
for f, b in sorted(byfile.items(), key=lambda kv: -kv[1][1]):
    print("  %-40s inst %5.1f%% samples %5.1f%% (no_inst %5.1f%%)" % (f, 100 * b[0] / tot_i, 100 * b[1] / tot_s, 100 * b[2] / tot_s))
print("top lines by samples:")
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][int(__import__("os").environ.get("SORT_INST", "0") == "0")])[:topn]:
    print("  %-22s:%4d inst %5.1f%% samp %5.1f%% thr/inst %4.1f | %s" % (k[0], k[1], 100 * a[0] / tot_i, 100 * a[1] / tot_s, a[4] / max(a[0], 1), a[2].strip()))
