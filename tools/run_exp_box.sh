#!/bin/bash
mkdir -p gpurun_out
for cfg in "55000 1024 5 8" "47112 1024 14 8" "85000 1280 14 1" "55000 1000 14 3" "300 64 14 2" "40000 2048 16 4" "7 8 3 1"; do
  for v in tools/bin/exp_sim_epi*; do timeout 120 $v $cfg 30 2>> gpurun_out/exp_box.err; done
done | tee gpurun_out/exp_sim_epi.jsonl | python tools/show_exp.py /dev/stdin
tail -3 gpurun_out/exp_box.err
( timeout 900 python -m pytest tests -m gpu -x -q -k "sim or round or flow or full_size or edges or tagg or pool or local_update" ) > gpurun_out/pytest_quick.log 2>&1
tail -2 gpurun_out/pytest_quick.log
