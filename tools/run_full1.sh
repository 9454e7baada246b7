#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu_r02.log 2>&1
tail -3 gpurun_out/pytest_gpu_r02.log
( time timeout 600 python bench.py ) > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err
tail -4 gpurun_out/bench_r02.err
python tools/show_bench.py gpurun_out/bench_r02.json
