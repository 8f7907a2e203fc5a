"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: one step of the
bench (between two im2col_stem launches), grouped by (kernel, grid)."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
seq = []
for row in csv.DictReader(lines):
    if row.get('Metric Name') != 'gpu__time_duration.sum':
        continue
    v = float(row['Metric Value'].replace(',', ''))
    u = row['Metric Unit']
    v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)
    name = re.sub(r'\(.*', '', row['Kernel Name'])
    name = re.sub(r'<unnamed>::|void |at::native::|cub::', '', name)[:44]
    seq.append((name, row['Grid Size'], v))
starts = [i for i, s in enumerate(seq) if 'im2col' in s[0]]
# a full step = trunk + window graph: the widest gap between consecutive stem launches (the window prefill of a
# timed() call launches trunk graphs only)
gaps = [(starts[i + 1] - starts[i], i) for i in range(len(starts) - 1)]
if gaps:
    big = [g for g in gaps if g[0] > 0.75 * max(gaps)[0]]
    common = collections.Counter(g[0] for g in big).most_common(1)[0][0]     # a steady-state step, not a leg boundary
    i = [g[1] for g in big if g[0] == common][-1]
    a, b = starts[i], starts[i + 1]
else:
    a, b = 0, len(seq)
step = seq[a:b]
tot = sum(s[2] for s in step)
print('one step: %d launches, %.1f us summed kernel time' % (len(step), tot))
agg = collections.OrderedDict()
for k, g, v in step:
    e = agg.setdefault((k, g), [0, 0.0])
    e[0] += 1
    e[1] += v
print('%-46s %-16s %5s %9s %6s' % ('kernel', 'grid', 'n', 'us', '%'))
for (k, g), (n, t) in agg.items():
    print('%-46s %-16s %5d %9.1f %5.1f%%' % (k, g, n, t, 100 * t / tot))
byk = collections.defaultdict(float)
for k, g, v in step:
    byk[k] += v
print('--- by kernel')
for k, t in sorted(byk.items(), key=lambda x: -x[1])[:12]:
    print('%-46s %9.1f us %5.1f%%' % (k, t, 100 * t / tot))
