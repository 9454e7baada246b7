#!/bin/bash
# quick GPU check: selected parity tests + the bench line without the CPU / e2e legs
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS:-} ) > gpurun_out/pytest_quick.log 2>&1
tail -4 gpurun_out/pytest_quick.log
( timeout 600 python bench.py --skip-e2e --skip-cpu-baseline ${BENCH_ARGS:-} ) > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -3 gpurun_out/bench_quick.err
python tools/show_bench.py gpurun_out/bench_quick.json
