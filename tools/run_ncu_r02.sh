#!/bin/bash
# ncu evidence for round 2 (one GPU): launch list of a bench run + full captures of every kernel of the round
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 1 --skip-e2e --skip-cpu-baseline --extra-configs none --no-graph"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fmlp -s 40 -c 40 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/ncu_launches_r02.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fmlp -s 40 -c 10 -f -o gpurun_out/prof_r02_round $B > gpurun_out/ncu_round_r02.log 2>&1
tail -2 gpurun_out/ncu_round_r02.log
B14="python bench.py --steps 3 --warmup 1 --skip-e2e --skip-cpu-baseline --extra-configs none --no-graph --config cxr14_64c"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tag_sim_kernel|tag_select_kernel|fill_loss" -s 6 -c 3 -f -o gpurun_out/prof_r02_c14 $B14 > gpurun_out/ncu_c14_r02.log 2>&1
tail -2 gpurun_out/ncu_c14_r02.log
