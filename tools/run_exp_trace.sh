#!/bin/bash
mkdir -p gpurun_out
out=gpurun_out/exp_sim_trace.jsonl
: > $out
for cfg in "55000 1024 5 8" "47112 1024 14 8" "85000 1280 14 1"; do
  for v in tools/bin/exp_sim_*; do
    timeout 120 $v $cfg 30 >> $out 2>> gpurun_out/exp_sim_trace.err
  done
done
cat $out
for kb in 227; do
  for cfg in "55000 1024 5 8" "47112 1024 14 8"; do
    FMLP_SIM_SMEM_KB=$kb timeout 120 tools/bin/exp_sim_b_tfirst $cfg 30 2>> gpurun_out/exp_sim_trace.err | sed "s/\"tfirst\"/\"tfirst_kb$kb\"/"
  done
done
