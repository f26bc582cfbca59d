"""Print an ncu --csv --log-file launch list as one line per launch (development aid)."""
import csv, sys
from collections import OrderedDict
rows = list(csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"')))
d = OrderedDict()
for r in rows:
    d.setdefault(r['ID'], {'k': r['Kernel Name'][:24], 'grid': r['Grid Size']})[r['Metric Name']] = r['Metric Value']
short = {'gpu__time_duration.sum': 't_ns', 'dram__bytes_read.sum': 'rd', 'dram__bytes_write.sum': 'wr', 'smsp__inst_executed.sum': 'inst',
         'sm__warps_active.avg.pct_of_peak_sustained_active': 'occ%', 'smsp__thread_inst_executed_per_inst_executed.ratio': 'lanes'}
for k, v in d.items():
    print(k, v['k'], ' '.join(f"{short.get(m, m)}={v[m]}" for m in v if m not in ('k', 'grid')))
