#!/bin/bash
# Builds tools/bin/exp_sim_<variant> for a sweep of tag_sim tuning knobs (+ the round-1 kernel from git
# history as "old") — run here (nvcc cross-compiles), the binaries travel to the GPU box with gpurun.
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Iinclude -Ifedmlp_b200/csrc"
build() {  # name, extra defines, source
  nvcc $FLAGS $2 -DEXP_VARIANT="\"$1\"" tools/exp_sim.cu $3 -o tools/bin/exp_sim_$1 &
}
git show 871a0e9:fedmlp_b200/csrc/tag_sim.cu > /tmp/tag_sim_r01.cu
# (the round-1 kernel has the old ABI; its numbers are in profiles/r02_exp_sim_a.jsonl)
for W in 4 8; do for RS in 1 2; do for RL in 2 4; do
  build w${W}_rs${RS}_rl${RL} "-DFMLP_SIM_W=$W -DFMLP_SIM_RT_SMALL=$RS -DFMLP_SIM_RT_LARGE=$RL" fedmlp_b200/csrc/tag_sim.cu
done; done; done
wait
ls -la tools/bin
