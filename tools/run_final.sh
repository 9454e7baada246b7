#!/bin/bash
# last check of the round: full GPU suite, default bench line, main-stream priority variant
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r02.log 2>&1
tail -2 gpurun_out/pytest_gpu_r02.log
( timeout 300 python bench.py ) > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
python tools/show_bench.py gpurun_out/bench_r02.json | grep "ms_per_step\|^api\|^clocks"
for pr in -1; do
( FMLP_MAIN_PRIORITY=$pr timeout 300 python bench.py --skip-e2e --skip-cpu-baseline --steps 300 ) > gpurun_out/bench_mainprio.json 2>> gpurun_out/bench_quick.err
echo "--- main priority $pr: $(python tools/show_bench.py gpurun_out/bench_mainprio.json | grep ms_per_step | sed 's/value.*//' | tr '\n' ' ')"
done
