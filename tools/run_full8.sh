#!/bin/bash
# 8-GPU check of the bench line (default collective), short
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 200 --warmup 10 --skip-e2e ) > gpurun_out/bench_r02_8gpu.json 2> gpurun_out/bench_r02_8gpu.err
echo "rc=$?"; tail -3 gpurun_out/bench_r02_8gpu.err
python tools/show_bench.py gpurun_out/bench_r02_8gpu.json | grep -v loss_sweep
python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_8gpu.json').read().strip().splitlines()[-1]); print(d['collective'][:200]); print([ (c['name'], c['parity']['parity_ok']) for c in d['configs']])"
