#!/bin/bash
mkdir -p gpurun_out
for cfg in "47112 1024 14 8" "85000 1280 14 1" "55000 1024 9 8" "55000 1000 14 3" "300 64 14 2" "40000 2048 16 4"; do
  for v in tools/bin/exp_sim_wide*; do timeout 120 $v $cfg 30 2>> gpurun_out/exp_box.err; done
done | tee gpurun_out/exp_sim_wide.jsonl | python tools/show_exp.py /dev/stdin
tail -3 gpurun_out/exp_box.err
