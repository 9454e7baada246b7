#!/bin/bash
mkdir -p gpurun_out
B="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 300 --warmup 10 --skip-e2e --extra-configs none"
run() { name=$1; shift; ( env "$@" timeout 300 $B $EXTRA ) > gpurun_out/bench_v2_$name.json 2> gpurun_out/bench_v2_$name.err; echo "--- $name rc=$? $(python tools/show_bench.py gpurun_out/bench_v2_$name.json | grep 'ms_per_step\|fedavg' | sed 's/value.*serial/serial/' | tr '\n' ' ')"; }
EXTRA="" run base FOO=1
EXTRA="" run pad40 FMLP_PROTO_PAD_SMEM_KB=40
EXTRA="" run pad24 FMLP_PROTO_PAD_SMEM_KB=24
EXTRA="--collective queue_split" run qsplit FOO=1
EXTRA="--collective queue_split" run qsplit_pad40 FMLP_PROTO_PAD_SMEM_KB=40
EXTRA="--streams 2" run s2 FOO=1
