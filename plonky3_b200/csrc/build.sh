#!/bin/bash
# Builds libp3gpu.so for sm_100a in-tree (plonky3_b200/libp3gpu.so).  Usage: build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
OUT=${P3GPU_OUT:-../libp3gpu.so}
OBJ=${P3GPU_OBJ:-../../build/obj}
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -ccbin /usr/bin/g++ $*"
mkdir -p $OBJ
pids=()
for f in ntt hash fri open peer air challenger query capi; do
  $NVCC $FLAGS -c $f.cu -o $OBJ/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o $OUT $OBJ/{ntt,hash,fri,open,peer,air,challenger,query,capi}.o
echo "built $(realpath $OUT)"
