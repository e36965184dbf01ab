#!/bin/bash
# tuning builds of the specialised kernel: ./build_variant.sh <name> "<-D flags>" -> build_variants/libfluxcalc_<name>.so
# (used through FLUXCALC_LIB=... ; only spec_kernel.cu is recompiled, the other objects come from ./build)
set -e
cd "$(dirname "$0")"
mkdir -p build_variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -I../../../include $2 -c spec_kernel.cu -o build_variants/spec_kernel_$1.o
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o build_variants/libfluxcalc_$1.so build/kernels.o build_variants/spec_kernel_$1.o build/context.o build/level1.o build/nccl_dyn.o build/p2p_comm.o build/frontend.o -ldl
