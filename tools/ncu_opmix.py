"""Executed-instruction mix and stall samples per opcode / per address range from `ncu --page source --csv` (SASS view).
Usage: python tools/ncu_opmix.py src.csv [per_unit_divisor]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
h = rows[hi]
iS, iE, iN = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
iW = h.index('L1 Wavefronts Shared')
ops = collections.Counter(); samp = collections.Counter(); wav = collections.Counter()
tot = 0
for r in rows[hi + 1:]:
    if len(r) <= iE or not r[iE]:
        continue
    src = r[iS].strip()
    m = re.match(r'(@!?U?P\w+\s+)?([A-Z0-9_]+)', src)
    if not m:
        continue
    op = m.group(2)
    e = float(r[iE]); tot += e
    ops[op] += e; samp[op] += float(r[iN] or 0)
    try: wav[op] += float(r[iW] or 0)
    except ValueError: pass
print('total warp instructions %.0f (%.1f per unit)' % (tot, tot / div))
ts = sum(samp.values())
for op, e in ops.most_common(40):
    print('%-10s %12.0f  %8.1f per unit  %5.1f %% of instr  %5.1f %% of samples  smem wavefronts/unit %.1f' % (op, e, e / div, 100 * e / tot, 100 * samp[op] / max(ts, 1), wav[op] / div))
