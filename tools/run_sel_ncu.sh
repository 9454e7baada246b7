#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "select or round or flow or full_size or edges or tagg" ) > gpurun_out/pytest_quick.log 2>&1
tail -2 gpurun_out/pytest_quick.log
for cfg in ich55k cxr14_64c effb0_85k; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:tag_select|fill_loss" -s 8 -c 8 --csv --log-file gpurun_out/sel_launch_$cfg.csv python bench.py --config $cfg --steps 3 --warmup 1 --skip-e2e --skip-cpu-baseline --extra-configs none --no-graph > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/sel_launch_$cfg.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; k=h.index('Kernel Name'); v=h.index('Metric Value')
from collections import defaultdict
d=defaultdict(list)
for r in rows[hi+1:]:
    if len(r)>v: d[r[k].split('(')[0]].append(float(r[v].replace(',','')))
for n,x in d.items(): print('$cfg', n, 'n=%d'%len(x), 'mean %.2f us'%(sum(x)/len(x)/1000), 'min %.2f'%(min(x)/1000))
PY
done
( timeout 600 python bench.py --skip-e2e --skip-cpu-baseline --steps 300 ) > gpurun_out/bench_q.json 2>> gpurun_out/bench_quick.err
python tools/show_bench.py gpurun_out/bench_q.json | grep "ms_per_step\|select\|loss  "
