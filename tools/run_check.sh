#!/bin/bash
# short GPU check: selected parity tests ($PYTEST_K) + the bench line without the CPU / e2e legs
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "${PYTEST_K:-round or flow or loss}" ) > gpurun_out/pytest_quick.log 2>&1
tail -2 gpurun_out/pytest_quick.log
( timeout 600 python bench.py --skip-e2e --skip-cpu-baseline --steps 300 ) > gpurun_out/bench_q.json 2>> gpurun_out/bench_quick.err
python tools/show_bench.py gpurun_out/bench_q.json | grep -v loss_sweep
