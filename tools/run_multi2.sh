#!/bin/bash
# 2-GPU validation of the round-2 aggregation exchange: tests, peer-signal microbenchmark, bench at N=2
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_2gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/pytest_multi2.log 2>&1; tail -8 gpurun_out/pytest_multi2.log
timeout 120 tools/bin/exp_peer_signal 28 4 50 > gpurun_out/exp_peer_signal.txt 2>&1; tail -12 gpurun_out/exp_peer_signal.txt
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 10"
timeout 400 $B > gpurun_out/bench_r02_2gpu_queue.json 2> gpurun_out/bench_r02_2gpu_queue.err; echo "queue rc=$?"
FMLP_ARQ_MULTICAST=0 timeout 300 $B --skip-e2e --extra-configs none > gpurun_out/bench_r02_2gpu_queue_p2p.json 2> gpurun_out/bench_r02_2gpu_queue_p2p.err; echo "queue-p2p rc=$?"
timeout 300 $B --skip-e2e --extra-configs none --collective fused_r01 > gpurun_out/bench_r02_2gpu_r01.json 2> gpurun_out/bench_r02_2gpu_r01.err; echo "r01 rc=$?"
timeout 300 $B --skip-e2e --extra-configs none --collective nccl > gpurun_out/bench_r02_2gpu_nccl.json 2> gpurun_out/bench_r02_2gpu_nccl.err; echo "nccl rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r02_2gpu_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['ms_per_step'],4), {k:v.get('ms') for k,v in d['kernels'].items()}, d.get('parity_ok'), (d.get('e2e') or {}).get('ms_per_step'), d['collective'][:60])
    except Exception as e: print(f, 'ERR', e)
PY
