#!/bin/bash
# Round-2 evidence run (one GPU): GPU parity suite, default bench line, reference arm, ncu launch list + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r02.log 2>&1
tail -3 gpurun_out/pytest_gpu_r02.log
( time timeout 600 python bench.py ) > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
tail -c 600 gpurun_out/bench_r02.json
( time timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err
tail -c 400 gpurun_out/bench_r02_reference.json
bash tools/run_ncu_r02.sh
python tools/ncu_summary.py gpurun_out/prof_r02_round.ncu-rep > gpurun_out/ncu_summary_r02.txt 2>&1 || true
ls -la gpurun_out
