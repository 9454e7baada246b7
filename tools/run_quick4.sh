#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "select or round or flow or full_size or edges or tagg" ) > gpurun_out/pytest_quick.log 2>&1
tail -5 gpurun_out/pytest_quick.log
( timeout 600 python bench.py --skip-e2e --skip-cpu-baseline --steps 300 ) > gpurun_out/bench_q.json 2>> gpurun_out/bench_quick.err
python tools/show_bench.py gpurun_out/bench_q.json | grep "ms_per_step\|select"
for cl in 2 4; do
( FMLP_SELECT_CLUSTER=$cl timeout 600 python bench.py --skip-e2e --skip-cpu-baseline --steps 300 ) > gpurun_out/bench_q.json 2>> gpurun_out/bench_quick.err
echo "--- FMLP_SELECT_CLUSTER=$cl"; python tools/show_bench.py gpurun_out/bench_q.json | grep "ms_per_step\|select"
done
tail -3 gpurun_out/bench_quick.err
