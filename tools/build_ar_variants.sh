#!/bin/bash
# libfedmlp_b200 variants with smaller CTAs of the round-1 push kernel (fedavg_allreduce.cu) -> tools/bin/lib_ar_<name>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/bin /tmp/arv
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Iinclude -Ifedmlp_b200/csrc"
for v in "t480:-DFMLP_AR_THREADS=480" "t224:-DFMLP_AR_THREADS=224"; do
  name=${v%%:*}; defs=${v#*:}
  ( nvcc $FLAGS $defs -c fedmlp_b200/csrc/fedavg_allreduce.cu -o /tmp/arv/ar_$name.o
    objs=$(ls fedmlp_b200/build/*.o | grep -v "fedavg_allreduce.o")
    nvcc -gencode arch=compute_100a,code=sm_100a -shared -Xlinker --exclude-libs=ALL -Xlinker -Bsymbolic -o tools/bin/lib_ar_$name.so $objs /tmp/arv/ar_$name.o ) &
done
wait
ls -la tools/bin/lib_ar_*.so
