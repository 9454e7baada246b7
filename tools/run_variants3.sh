#!/bin/bash
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
B="$T bench.py --gpus $N --steps 200 --warmup 10 --skip-e2e --extra-configs none"
run() { # tag, env...
  tag=$1; shift
  env "$@" timeout 200 $B > gpurun_out/bench_var3_${N}gpu_$tag.json 2> gpurun_out/bench_var3_${N}gpu_$tag.err; echo "$tag rc=$?"
}
rm -f gpurun_out/bench_var3_${N}gpu_*.json
L2=$PWD/tools/bin/lib_arq_t288x4.so
L5=$PWD/tools/bin/lib_arq_t544x2.so
run t544_c148_mc0 FEDMLP_B200_LIB=$L5 FMLP_ARQ_CTAS=148 FMLP_ARQ_MULTICAST=0
run t544_c148_mc0_noprio FEDMLP_B200_LIB=$L5 FMLP_ARQ_CTAS=148 FMLP_ARQ_MULTICAST=0 FMLP_AGG_PRIORITY=0
run t544_c296_mc0 FEDMLP_B200_LIB=$L5 FMLP_ARQ_MULTICAST=0
run t288_c296_mc0 FEDMLP_B200_LIB=$L2 FMLP_ARQ_CTAS=296 FMLP_ARQ_MULTICAST=0
run t288_c444_mc0 FEDMLP_B200_LIB=$L2 FMLP_ARQ_CTAS=444 FMLP_ARQ_MULTICAST=0
run t544_c148_mc1 FEDMLP_B200_LIB=$L5 FMLP_ARQ_CTAS=148 FMLP_ARQ_MULTICAST=1
run default_mc0 FMLP_ARQ_MULTICAST=0
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_var3_${N}gpu_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],4), {k:v.get('ms') for k,v in d['kernels'].items()}, 'parity', d.get('parity_ok'))
    except Exception as e: print(f, 'ERR', e)
PY
