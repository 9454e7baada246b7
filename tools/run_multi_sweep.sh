# usage: bash tools/run_multi_sweep.sh [NGPU] ["chunk list"] [test?]
N=${1:-2}; CH=${2:-"1 2 4"}; T=${3:-1}
if [ "$T" = "1" ]; then timeout 300 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2; fi
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 10 --skip-e2e"
P='import sys,json; d=json.loads(sys.stdin.readlines()[-1]); print(round(d["ms_per_step"],4), round(d["value"]/1e6,1), {k:v["ms"] for k,v in d["kernels"].items()})'
for nc in $CH; do echo "== eager chunks=$nc"; FMLP_AR_CHUNKS=$nc timeout 150 $B 2>/dev/null | python -c "$P"; done
for nc in $CH; do echo "== graph chunks=$nc"; FMLP_AR_CHUNKS=$nc timeout 150 $B --graph-multi 1 2>/dev/null | python -c "$P"; done
