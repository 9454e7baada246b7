#!/bin/bash
# GPU-side sweep of the tag_sim variants built by tools/build_exp_sim.sh -> gpurun_out/exp_sim.jsonl
mkdir -p gpurun_out
out=gpurun_out/exp_sim.jsonl
: > $out
for v in tools/bin/exp_sim_*; do
  for cfg in "55000 1024 5 8" "47112 1024 14 8" "85000 1280 14 1"; do
    timeout 120 $v $cfg 30 >> $out 2>> gpurun_out/exp_sim.err || echo "{\"variant\": \"$v\", \"cfg\": \"$cfg\", \"failed\": $?}" >> $out
  done
done
for st in 2 3 4; do
  FMLP_SIM_STAGES=$st timeout 120 tools/bin/exp_sim_w4_rs2_rl2 55000 1024 5 8 30 >> $out 2>> gpurun_out/exp_sim.err
done
FMLP_SIM_STAGES=2 timeout 120 tools/bin/exp_sim_w8_rs2_rl2 55000 1024 5 8 30 >> $out 2>> gpurun_out/exp_sim.err
echo '{"note": "PDL off below"}' >> $out
FMLP_SIM_PDL=0 timeout 120 tools/bin/exp_sim_w8_rs2_rl2 55000 1024 5 8 30 >> $out 2>> gpurun_out/exp_sim.err
FMLP_SIM_PDL=0 timeout 120 tools/bin/exp_sim_w8_rs2_rl2 47112 1024 14 8 30 >> $out 2>> gpurun_out/exp_sim.err
cat $out
