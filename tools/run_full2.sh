#!/bin/bash
# 2-GPU validation: multi-rank GPU tests + the bench line at N=2 (with e2e) + reference arm under torchrun
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dist.py -m gpu -x -q ) > gpurun_out/pytest_gpu_multi_r02.log 2>&1
tail -3 gpurun_out/pytest_gpu_multi_r02.log
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10 ) > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err
echo "rc=$?"; tail -3 gpurun_out/bench_r02_2gpu.err
python tools/show_bench.py gpurun_out/bench_r02_2gpu.json
