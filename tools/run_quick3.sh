#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in concurrent sim_first proto_first; do
( FMLP_ROUND_SCHEDULE=$v timeout 600 python bench.py --skip-e2e --skip-cpu-baseline --steps 300 ) > gpurun_out/bench_q.json 2>> gpurun_out/bench_quick.err
echo "--- rep $rep $v: $(python tools/show_bench.py gpurun_out/bench_q.json | grep ms_per_step | sed 's/value.*//' | tr '\n' ' ')"
done; done
tail -3 gpurun_out/bench_quick.err
