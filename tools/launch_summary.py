"""Summarise an ncu gpu__time_duration launch list: per-kernel totals for the last frame (between advance_kernel launches)."""
import csv, collections, re, sys
path = sys.argv[1]
rows = [r for r in csv.reader(open(path)) if len(r) > 5]
hdr = rows[0]; data = rows[1:]
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
gi = hdr.index('Grid Size')
def us(r):
    v = float(r[vi].replace(',', '')); u = r[ui]
    return v / 1000 if u in ('ns', 'nsecond') else v if u in ('us', 'usecond') else v * 1000
def nm(r):
    n = r[ki].replace('void ', '').replace('unnamed>::', '')
    return re.sub(r'\(.*', '', n)
names = [nm(r) for r in data]
adv = [i for i, n in enumerate(names) if n.startswith('advance_kernel')]
start, end = adv[-2] + 1, adv[-1] + 1
seq = []
for i in range(start, end):
    seq.append((names[i], data[i][gi], us(data[i])))
tot = sum(x[2] for x in seq)
print('frame launches', len(seq), 'sum us', round(tot, 1))
if len(sys.argv) > 2 and sys.argv[2] == 'seq':
    for n, g, t in seq: print(f"{n[:50]:50s} {g:16s} {t:8.1f}")
else:
    agg = collections.OrderedDict()
    for n, g, t in seq:
        a = agg.setdefault(n.split('<')[0], [0, 0.0]); a[0] += 1; a[1] += t
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:50]:50s} n={v[0]:3d} us={v[1]:8.1f} {100*v[1]/tot:5.1f}%")
