#!/bin/bash
# Round-2 evidence run (one GPU): smoke, GPU parity suite, default bench line, reference arm, ncu launch list + full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
( time timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" ) > gpurun_out/smoke_r02.log 2>&1
tail -4 gpurun_out/smoke_r02.log
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r02.log 2>&1
tail -3 gpurun_out/pytest_gpu_r02.log
( time timeout 600 python bench.py ) > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
python tools/show_bench.py gpurun_out/bench_r02.json | grep -v loss_sweep
( time timeout 400 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err
tail -c 300 gpurun_out/bench_r02_reference.json
bash tools/run_ncu_r02.sh
for n in round cxr14 effb0; do python tools/ncu_summary.py gpurun_out/prof_r02_$n.ncu-rep > gpurun_out/ncu_summary_r02_$n.txt 2>&1 || true; done
ls -la gpurun_out | head -40
