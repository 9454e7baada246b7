#!/bin/bash
# one ncu --set full capture of the kernels matching $K in a short bench run ($BENCH_ARGS), summarised
mkdir -p gpurun_out
K=${K:-proto_accum}
B="python bench.py --steps 3 --warmup 1 --skip-e2e --skip-cpu-baseline --extra-configs none --no-graph ${BENCH_ARGS:-}"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$K" -s ${SKIP:-6} -c ${COUNT:-2} -f -o gpurun_out/prof_one $B > gpurun_out/ncu_one.log 2>&1
tail -2 gpurun_out/ncu_one.log | cut -c1-300
python tools/ncu_summary.py gpurun_out/prof_one.ncu-rep > gpurun_out/ncu_one_summary.txt 2>&1
cat gpurun_out/ncu_one_summary.txt
