#!/bin/bash
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
B="$T bench.py --gpus $N --steps 200 --warmup 10 --skip-e2e --extra-configs none"
run() { # tag, env... -- args
  tag=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 200 $B "$@" > gpurun_out/bench_var4_${N}gpu_$tag.json 2> gpurun_out/bench_var4_${N}gpu_$tag.err; echo "$tag rc=$?"
}
rm -f gpurun_out/bench_var4_${N}gpu_*.json
A4=$PWD/tools/bin/lib_ar_t480.so
A2=$PWD/tools/bin/lib_ar_t224.so
run push_default X=1 -- --collective push_split
run push_t480_1 FEDMLP_B200_LIB=$A4 FMLP_AR_CTAS_PER_SM=1 -- --collective push_split
run push_t480_2 FEDMLP_B200_LIB=$A4 FMLP_AR_CTAS_PER_SM=2 -- --collective push_split
run push_t224_2 FEDMLP_B200_LIB=$A2 FMLP_AR_CTAS_PER_SM=2 -- --collective push_split
run push_t224_1 FEDMLP_B200_LIB=$A2 FMLP_AR_CTAS_PER_SM=1 -- --collective push_split
run push_t480_1_s2 FEDMLP_B200_LIB=$A4 FMLP_AR_CTAS_PER_SM=1 -- --collective push_split --streams 2
run r01_2stream X=1 -- --collective fused_r01 --streams 2
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_var4_${N}gpu_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],4), {k:v.get('ms') for k,v in d['kernels'].items()}, 'parity', d.get('parity_ok'))
    except Exception as e: print(f, 'ERR', e)
PY
