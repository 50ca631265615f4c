#!/bin/bash
# Build a tuning variant of the library: scripts/build_variant.sh <name> <extra nvcc -D flags...>
# -> warpii_b200/variants/<name>.so (git-ignored; bench it with scripts/gpu_variants.sh)
set -eu
name=$1; shift
cd "$(dirname "$0")/../warpii_b200"
mkdir -p variants build/var_$name
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC $*"
for f in dgsem_stage_kernel dgsem_pencil_kernel dgsem_maxwell_kernel dgsem_general_kernel dgsem_aux_kernels warpii_gpu; do
  $NVCC $FLAGS -Xptxas -v -c csrc/$f.cu -o build/var_$name/$f.o 2> build/var_$name/$f.log
done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o variants/$name.so build/var_$name/*.o build/solver_capi.o -ldl
for k in Li2ELi4 Li3ELi4 Li3ELi5; do grep -A2 "pencil_stage_kernelI$k" build/var_$name/dgsem_pencil_kernel.log | grep -E "spill|Used" | tr '\n' ' '; echo " <- $name pencil $k"; done
grep -A2 "stage_kernel_generalILi2ELi4" build/var_$name/dgsem_general_kernel.log | grep -E "spill|Used" | tr '\n' ' '; echo " <- $name general <2,4>"
