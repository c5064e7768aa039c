#!/bin/bash
# build_variant.sh NAME "<extra nvcc flags>": experimental build of the CUDA library (NP=8 only)
# into ndt_b200/variants/libndt_b200_NAME.so; select it with NDT_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/../ndt_b200/csrc"
mkdir -p build ../variants
name=$1; shift
FL="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 -Xcompiler -fPIC -I../../include -I."
NPV=${NPV:-8}
nvcc $FL "$@" -DNDT_NP=$NPV -Xptxas -v -c np_inst.cu -o build/np8_$name.o 2> build/ptxas_$name.log &
nvcc $FL "$@" -DNDT_ONLY_NP=$NPV -c kernels.cu -o build/kernels_$name.o &
wait
[ -f build/mgpu.o ] || make build/flatten.o build/error.o build/kdbuild.o build/mgpu.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../variants/libndt_b200_$name.so build/kernels_$name.o build/np8_$name.o build/kdbuild.o build/mgpu.o build/flatten.o build/error.o -lm -ldl -lpthread
grep -A2 "k_shadeILi${NPV}ELi1\|k_traceILi${NPV}ELi0" build/ptxas_$name.log | tr '\n' ' '; echo
