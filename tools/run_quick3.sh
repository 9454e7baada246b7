#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "proto or round or flow or local_update" ) > gpurun_out/pytest_quick.log 2>&1
tail -2 gpurun_out/pytest_quick.log
for v in "FMLP_PROTO_TMA=1 FMLP_PROTO_CTAS=1" "FMLP_PROTO_TMA=1 FMLP_PROTO_CTAS=2" "FMLP_PROTO_TMA=0"; do
( env $v timeout 600 python bench.py --skip-e2e --skip-cpu-baseline --steps 200 ) > gpurun_out/bench_q.json 2>> gpurun_out/bench_quick.err
echo "--- $v"; python tools/show_bench.py gpurun_out/bench_q.json | grep "proto \|ms_per_step"
done
tail -3 gpurun_out/bench_quick.err
