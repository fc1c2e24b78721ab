"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list."""
import collections
import csv
import sys

for f in sys.argv[1:]:
    rows = [r for r in csv.reader(l for l in open(f) if l.startswith('"'))]
    hdr = rows[0]
    ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
    d = collections.defaultdict(list)
    for r in rows[1:]:
        v = float(r[vi].replace(',', ''))
        v = v / 1000 if r[ui] in ('ns', 'nsecond') else (v * 1000 if r[ui] in ('ms', 'msecond') else v)
        d[r[ki][:70]].append(v)
    print(f)
    tot = sum(sum(v) for v in d.values())
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print("  %-70s n=%4d  avg %8.1f us  tot %9.1f us (%4.1f%%)" % (k, len(v), sum(v) / len(v), sum(v), 100 * sum(v) / tot))
    print("  total %.1f us" % tot)
