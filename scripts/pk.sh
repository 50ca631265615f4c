#!/bin/bash
# compile the pencil kernel alone and print registers / spills per instantiation: scripts/pk.sh [extra -D flags]
mkdir -p build && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -ccbin /usr/bin/g++ -Xcompiler -fPIC -Xptxas -v "$@" -c csrc/dgsem_pencil_kernel.cu -o build/pk_scratch.o 2>&1 | grep -E "error|Compiling|Used|spill" | sed -e 's/ptxas info    : //' -e 's/Compiling entry function .*pencil_stage_kernelILi\([0-9]\)ELi\([0-9]\).*/<\1,\2>/' | paste - - - 
