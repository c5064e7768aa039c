#!/bin/bash
# build_variant.sh NAME "<extra nvcc flags>": experimental build of the CUDA library into
# gpurun_out is not shipped; variants go to ndt_b200/variants/libndt_b200_NAME.so
set -e
cd "$(dirname "$0")/../ndt_b200/csrc"
mkdir -p build ../variants
name=$1; shift
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -I../../include -I. "$@" -Xptxas -v -c kernels.cu -o build/kernels_$name.o 2> build/ptxas_$name.log
[ -f build/flatten.o ] || make build/flatten.o build/error.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libndt_b200_$name.so build/kernels_$name.o build/kdbuild.o build/flatten.o build/error.o -lm -ldl
grep -A2 "k_generationILi8ELb0" build/ptxas_$name.log | tr '\n' ' '; echo
