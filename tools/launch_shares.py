"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time and share per kernel name."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ui], 1e-6)
    name = r[ki].split('(')[0][:70]
    t = tot.setdefault(name, [0.0, 0])
    t[0] += v * scale; t[1] += 1
total = sum(v[0] for v in tot.values())
print('%-72s %10s %8s %7s' % ('kernel', 'total ms', 'launches', 'share'))
for k, (ms, n) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print('%-72s %10.3f %8d %6.1f%%' % (k, ms, n, 100 * ms / total))
print('%-72s %10.3f' % ('all captured launches', total))
