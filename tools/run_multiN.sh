#!/bin/bash
# usage: tools/run_multiN.sh <N> [sweep?]  — aggregation sweep + bench.py at N GPUs with every collective variant
N=${1:-2}; SWEEP=${2:-1}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "$SWEEP" = "1" ]; then
  FMLP_SWEEP_SHORT=1 timeout 400 $T tools/arq_sweep.py > gpurun_out/arq_sweep_${N}gpu.jsonl 2> gpurun_out/arq_sweep_${N}gpu.err; echo "sweep rc=$?"
fi
B="$T bench.py --gpus $N --steps 200 --warmup 10"
timeout 500 $B > gpurun_out/bench_r02_${N}gpu_queue_split.json 2> gpurun_out/bench_r02_${N}gpu_queue_split.err; echo "queue_split rc=$?"
timeout 300 $B --skip-e2e --extra-configs none --collective queue > gpurun_out/bench_r02_${N}gpu_queue.json 2> gpurun_out/bench_r02_${N}gpu_queue.err; echo "queue rc=$?"
timeout 300 $B --skip-e2e --extra-configs none --collective fused_r01 > gpurun_out/bench_r02_${N}gpu_r01.json 2> gpurun_out/bench_r02_${N}gpu_r01.err; echo "r01 rc=$?"
timeout 300 $B --skip-e2e --extra-configs none --collective nccl > gpurun_out/bench_r02_${N}gpu_nccl.json 2> gpurun_out/bench_r02_${N}gpu_nccl.err; echo "nccl rc=$?"
FMLP_ARQ_MULTICAST=0 timeout 300 $B --skip-e2e --extra-configs none > gpurun_out/bench_r02_${N}gpu_queue_split_p2p.json 2> gpurun_out/bench_r02_${N}gpu_queue_split_p2p.err; echo "queue_split p2p rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r02_${N}gpu_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],4), {k:v.get('ms') for k,v in d['kernels'].items()}, 'parity', d.get('parity_ok'), 'e2e', (d.get('e2e') or {}).get('ms_per_step'))
        for c in d.get('configs') or []: print('    ', c.get('name'), c.get('ms_per_step'), (c.get('parity') or {}).get('parity_ok'), c.get('error'))
    except Exception as e: print(f, 'ERR', e)
PY
[ -f gpurun_out/arq_sweep_${N}gpu.jsonl ] && cat gpurun_out/arq_sweep_${N}gpu.jsonl
