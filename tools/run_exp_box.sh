#!/bin/bash
mkdir -p gpurun_out
for cfg in "47112 1024 14 8" "85000 1280 14 1"; do
  for v in tools/bin/exp_sim_a tools/bin/exp_sim_b tools/bin/exp_sim_c; do timeout 120 $v $cfg 30 2>> gpurun_out/exp_box.err; done
done | tee gpurun_out/exp_sim_w16.jsonl | python tools/show_exp.py /dev/stdin
tail -3 gpurun_out/exp_box.err
