#!/bin/bash
mkdir -p gpurun_out
for cfg in "55000 1024 5 8" "47112 1024 14 8" "85000 1280 14 1"; do timeout 120 tools/bin/exp_sim_pdl_table $cfg 30; FMLP_SIM_PDL=0 timeout 120 tools/bin/exp_sim_pdl_table $cfg 30 | sed 's/pdl_table/no_pdl/'; done | python tools/show_exp.py /dev/stdin
( timeout 900 python -m pytest tests -m gpu -x -q -k "sim or round or flow or full_size or edges or tagg or pool" ) > gpurun_out/pytest_quick.log 2>&1
tail -2 gpurun_out/pytest_quick.log
( timeout 600 python bench.py --skip-e2e --skip-cpu-baseline --steps 300 ) > gpurun_out/bench_q.json 2>> gpurun_out/bench_quick.err
python tools/show_bench.py gpurun_out/bench_q.json | grep -v "loss_sweep"
python -c "
import json; d=json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1]); print(d['timing']['schedule']); [print(c['name'], c.get('schedule')) for c in d['configs']]"
