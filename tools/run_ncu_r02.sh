#!/bin/bash
# ncu evidence for round 2 (one GPU): launch list of a bench run + full captures of every kernel of the round
mkdir -p gpurun_out
K='regex:tag_sim_kernel|sim_quad_table|tag_select|fill_loss|proto_accum|proto_finalize|fedavg_flat|agg_tails'
B="python bench.py --steps 3 --warmup 1 --skip-e2e --skip-cpu-baseline --extra-configs none --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 40 -c 40 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_launches_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "$K" -s 40 -c 10 -f -o gpurun_out/prof_r02_round $B > gpurun_out/ncu_round_r02.log 2>&1
tail -2 gpurun_out/ncu_round_r02.log
# the C = 14 workloads: similarity kernel at 14 class vectors, clustered selection at 85,000 rows
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tag_sim_kernel|fill_loss" -s 6 -c 2 -f -o gpurun_out/prof_r02_cxr14 $B --config cxr14_64c > gpurun_out/ncu_cxr14_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:tag_select|proto_finalize" -s 6 -c 2 -f -o gpurun_out/prof_r02_effb0 $B --config effb0_85k > gpurun_out/ncu_effb0_r02.log 2>&1
