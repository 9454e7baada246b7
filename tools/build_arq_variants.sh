#!/bin/bash
# libfedmlp_b200 variants with different CTA shapes of the work-queue aggregation kernel -> tools/bin/lib_arq_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin /tmp/arqv
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Ifedmlp_b200/csrc"
for v in "t544x2:-DFMLP_ARQ_THREADS=544 -DFMLP_ARQ_CTAS_PER_SM=2" "t288x4:-DFMLP_ARQ_THREADS=288 -DFMLP_ARQ_CTAS_PER_SM=4" "t288x6:-DFMLP_ARQ_THREADS=288 -DFMLP_ARQ_CTAS_PER_SM=6"; do
  name=${v%%:*}; defs=${v#*:}
  ( nvcc $FLAGS $defs -c fedmlp_b200/csrc/fedavg_allreduce_q.cu -o /tmp/arqv/arq_$name.o
    objs=$(ls fedmlp_b200/build/*.o | grep -v fedavg_allreduce_q.o)
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xlinker --exclude-libs=ALL -Xlinker -Bsymbolic -o tools/bin/lib_arq_$name.so $objs /tmp/arqv/arq_$name.o ) &
done
wait
ls -la tools/bin/*.so
# trace build (per-item timeline, tools/arq_trace.py)
nvcc $FLAGS -DFMLP_ARQ_TRACE -c fedmlp_b200/csrc/fedavg_allreduce_q.cu -o /tmp/arqv/arq_trace.o
objs=$(ls fedmlp_b200/build/*.o | grep -v fedavg_allreduce_q.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xlinker --exclude-libs=ALL -Xlinker -Bsymbolic -o tools/bin/lib_arq_trace.so $objs /tmp/arqv/arq_trace.o
