#!/bin/bash
# Build an experiment variant of libagb200.so: tools/build_variant.sh NAME "-DAGB_TC_NA=9 -DAGB_TC_NB=6 -DAGB_TC_NT=4"
# -> autograd.jl_b200/build_var/libagb200_NAME.so (select it with AGB200_LIB=...).  Only gemm_tc.cu is recompiled.
set -e
cd "$(dirname "$0")/.."
name=$1; defs=$2
mkdir -p autograd.jl_b200/build_var
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $defs \
  -c autograd.jl_b200/csrc/gemm_tc.cu -o autograd.jl_b200/build_var/gemm_tc_$name.o 2>&1 | grep -v deprecated || true
objs=$(ls autograd.jl_b200/build/*.o | grep -v gemm_tc.o)
nvcc -shared -o autograd.jl_b200/build_var/libagb200_$name.so $objs autograd.jl_b200/build_var/gemm_tc_$name.o -lcudart -lcuda -ldl 2>&1 | grep -v deprecated || true
rm -f autograd.jl_b200/build_var/gemm_tc_$name.o
ls -la autograd.jl_b200/build_var/libagb200_$name.so
