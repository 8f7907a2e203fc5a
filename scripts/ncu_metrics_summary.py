"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --csv` log of the igemm launches of
N bench steps: per (kernel, grid) launches, summed time, time-weighted tensor-pipe-active %, DRAM bytes.
Usage: python scripts/ncu_metrics_summary.py log.csv [steps] [out.json]
With out.json (profiles/igemm_ncu_step.json) it also writes the two figures bench.py's roofline block quotes
(traffic_bytes_per_launch, tensor_pipe_active_pct) together with their provenance - bench.py holds no constants."""
import collections
import csv
import json
import re
import sys

steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lines = [l for l in open(sys.argv[1]) if not l.startswith('==')]
per = collections.OrderedDict()
for row in csv.DictReader(lines):
    k = per.setdefault(row['ID'], {'name': row['Kernel Name'], 'grid': row['Grid Size']})
    v = float(row['Metric Value'].replace(',', ''))
    u, m = row['Metric Unit'], row['Metric Name']
    if m == 'gpu__time_duration.sum':
        v = v / 1e3 if u == 'ns' else (v * 1e3 if u == 'ms' else v)           # us
    if m.startswith('dram__bytes'):
        v = v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1.0, 'Gbyte': 1e3}.get(u, 1e-6)   # MB
    k[m] = v
agg = collections.OrderedDict()
T = 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'
for k in per.values():
    name = re.sub(r'\(.*', '', k['name'])
    name = re.sub(r'<unnamed>::|void ', '', name)
    e = agg.setdefault((name, k['grid']), [0, 0.0, 0.0, 0.0, 0.0])
    t = k.get('gpu__time_duration.sum', 0.0)
    e[0] += 1
    e[1] += t
    e[2] += t * k.get(T, 0.0)
    e[3] += k.get('dram__bytes_read.sum', 0.0)
    e[4] += k.get('dram__bytes_write.sum', 0.0)
n = sum(e[0] for e in agg.values())
tt = sum(e[1] for e in agg.values())
print('%d igemm launches (%.0f per step), %.2f ms per step, time-weighted tensor-pipe active %.1f %%, '
      'DRAM read %.1f MB + write %.1f MB per launch (average)'
      % (n, n / steps, tt / steps / 1e3, sum(e[2] for e in agg.values()) / tt,
         sum(e[3] for e in agg.values()) / n, sum(e[4] for e in agg.values()) / n))
print('%-34s %-14s %5s %9s %8s %10s %10s' % ('kernel', 'grid', 'n', 'ms/step', 'tensor%', 'dramR MB', 'dramW MB'))
for (name, grid), e in agg.items():
    print('%-34s %-14s %5d %9.2f %8.1f %10.1f %10.1f' % (name, grid, e[0], e[1] / steps / 1e3, e[2] / e[1], e[3] / e[0],
                                                         e[4] / e[0]))

if len(sys.argv) > 3:
    json.dump({'traffic_bytes_per_launch': (sum(e[3] for e in agg.values()) + sum(e[4] for e in agg.values())) / n * 1e6,
               'tensor_pipe_active_pct': sum(e[2] for e in agg.values()) / tt,
               'igemm_launches_per_step': n / steps, 'igemm_ms_per_step_under_ncu': tt / steps / 1e3,
               'source': 'profiles/igemm_ncu_step.json <- scripts/ncu_metrics_summary.py on an ncu range capture of '
                         'bench.py\'s timed region (HVR_NCU_RANGE=1, --steps %d); per-launch table: '
                         'profiles/r02_igemm_ncu_metrics_step_V7.txt' % int(steps)},
              open(sys.argv[3], 'w'), indent=1)
