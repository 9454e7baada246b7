#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "fedavg or aggregat or edges or dist or local_update" ) > gpurun_out/pytest_quick.log 2>&1
tail -2 gpurun_out/pytest_quick.log
( timeout 600 python bench.py --skip-cpu-baseline --extra-configs none --steps 300 ) > gpurun_out/bench_q.json 2>> gpurun_out/bench_quick.err
python tools/show_bench.py gpurun_out/bench_q.json | grep "ms_per_step\|^api"
