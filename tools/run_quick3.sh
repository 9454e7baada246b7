#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in "FMLP_SIM_REQUEST_SMEM_KB=227 FMLP_PROTO_PAD_SMEM_KB=0" "FMLP_SIM_REQUEST_SMEM_KB=227 FMLP_PROTO_PAD_SMEM_KB=40" "FMLP_SIM_REQUEST_SMEM_KB=200 FMLP_PROTO_PAD_SMEM_KB=40" "FMLP_SIM_REQUEST_SMEM_KB=0 FMLP_PROTO_PAD_SMEM_KB=48" "FMLP_SIM_REQUEST_SMEM_KB=0 FMLP_PROTO_PAD_SMEM_KB=56"; do
( env $v timeout 600 python bench.py --skip-e2e --skip-cpu-baseline --steps 300 ) > gpurun_out/bench_q.json 2>> gpurun_out/bench_quick.err
echo "--- rep $rep $v: $(python tools/show_bench.py gpurun_out/bench_q.json | grep ms_per_step | sed 's/value.*//' | tr '\n' ' ')"
done; done
tail -3 gpurun_out/bench_quick.err
