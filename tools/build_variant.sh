#!/bin/bash
# Experiment helper: tools/build_variant.sh NAME FILE.cu [-DDEFINE...]  ->  autograd.jl_b200/build_timing/libagb200_NAME.so
# (one source recompiled with extra defines, linked with the regular objects; run with AGB200_LIB=<that .so>).
set -e
cd "$(dirname "$0")/../autograd.jl_b200"
NAME=$1; SRC=$2; shift 2
mkdir -p build_timing
BASE=$(basename $SRC .cu)
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Wno-deprecated-gpu-targets "$@" -c csrc/$SRC -o build_timing/${BASE}_$NAME.o
nvcc -shared -Wno-deprecated-gpu-targets -o build_timing/libagb200_$NAME.so build_timing/${BASE}_$NAME.o $(ls build/*.o | grep -v "/$BASE.o") -lcudart -lcuda -ldl
echo built build_timing/libagb200_$NAME.so
