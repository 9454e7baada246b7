#!/bin/bash
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q -k "proto or round or flow or local_update or full_size" ) > gpurun_out/pytest_quick.log 2>&1
tail -4 gpurun_out/pytest_quick.log
( timeout 600 python bench.py --skip-e2e --skip-cpu-baseline ) > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
tail -3 gpurun_out/bench_quick.err
python tools/show_bench.py gpurun_out/bench_quick.json
for kb in 160 227; do
( FMLP_PROTO_SMEM_KB=$kb timeout 600 python bench.py --skip-e2e --skip-cpu-baseline --extra-configs none --steps 200 ) > gpurun_out/bench_quick_kb$kb.json 2>> gpurun_out/bench_quick.err
echo "--- FMLP_PROTO_SMEM_KB=$kb"; python tools/show_bench.py gpurun_out/bench_quick_kb$kb.json
done
