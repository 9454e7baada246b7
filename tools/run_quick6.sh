#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "loss or round or flow or local_update" ) > gpurun_out/pytest_quick.log 2>&1
tail -2 gpurun_out/pytest_quick.log
( timeout 600 python bench.py --skip-e2e --skip-cpu-baseline --steps 300 ) > gpurun_out/bench_q.json 2>> gpurun_out/bench_quick.err
python tools/show_bench.py gpurun_out/bench_q.json | grep "ms_per_step\|loss  \|select"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:fill_loss" -s 4 -c 4 --csv --log-file gpurun_out/fl.csv python bench.py --config cxr14_64c --steps 3 --warmup 1 --skip-e2e --skip-cpu-baseline --extra-configs none --no-graph > /dev/null 2>&1; grep fill_loss gpurun_out/fl.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
