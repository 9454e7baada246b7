#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
B="$T bench.py --gpus $N --steps 200 --warmup 10 --skip-e2e --extra-configs none"
run() { # tag, env... -- args
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 240 $B "$@" > gpurun_out/bench_var8_${N}gpu_$tag.json 2> gpurun_out/bench_var8_${N}gpu_$tag.err; echo "$tag rc=$?"
}
rm -f gpurun_out/bench_var8_${N}gpu_*.json
A4=$PWD/tools/bin/lib_ar_t480.so
Q5=$PWD/tools/bin/lib_arq_t544x2.so
run push_t480_1 FEDMLP_B200_LIB=$A4 FMLP_AR_CTAS_PER_SM=1 -- --collective push_split
run push_default X=1 -- --collective push_split
run r01_2stream X=1 -- --collective fused_r01 --streams 2
run queue_split_nvls X=1 -- --collective queue_split
run queue_split_p2p FMLP_ARQ_MULTICAST=0 -- --collective queue_split
run queue_split_nvls_t544_c148 FEDMLP_B200_LIB=$Q5 FMLP_ARQ_CTAS=148 -- --collective queue_split
run nccl X=1 -- --collective nccl --streams 2
FMLP_SWEEP_SHORT=1 timeout 300 $T tools/arq_sweep.py > gpurun_out/arq_sweep_${N}gpu.jsonl 2> gpurun_out/arq_sweep_${N}gpu.err; echo "sweep rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_var8_${N}gpu_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],4), {k:v.get('ms') for k,v in d['kernels'].items()}, 'parity', d.get('parity_ok'))
    except Exception as e: print(f, 'ERR', e)
PY
cat gpurun_out/arq_sweep_${N}gpu.jsonl
