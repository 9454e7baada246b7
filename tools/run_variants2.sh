#!/bin/bash
N=${1:-2}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
B="$T bench.py --gpus $N --steps 200 --warmup 10 --skip-e2e --extra-configs none"
for v in default t544x2 t288x4; do
  if [ $v = default ]; then unset FEDMLP_B200_LIB; else export FEDMLP_B200_LIB=$PWD/tools/bin/lib_arq_$v.so; fi
  for mc in 1 0; do
    FMLP_ARQ_MULTICAST=$mc timeout 200 $B > gpurun_out/bench_var_${N}gpu_${v}_mc$mc.json 2> gpurun_out/bench_var_${N}gpu_${v}_mc$mc.err; echo "$v mc=$mc rc=$?"
  done
done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_var_${N}gpu_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d['ms_per_step'],4), {k:v.get('ms') for k,v in d['kernels'].items()}, 'parity', d.get('parity_ok'))
    except Exception as e: print(f, 'ERR', e)
PY
