#!/bin/bash
# Alternative build of the same ABI whose completion-counter polls end with the PTX-formal acquire
# (ld.acquire.gpu = LDG.STRONG.GPU + CCTL.IVALL once per rating; -DCARS_STRICT_ACQUIRE, csrc/sgd_kernels.cuh
# acquire_after_poll).  Select it with CARSKIT_B200_LIB=carskit_b200/libcarskit_b200_strict.so.
set -e
cd "$(dirname "$0")/.."
mkdir -p build
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
nvcc $F -DCARS_STRICT_ACQUIRE -c carskit_b200/csrc/engine.cu -o build/engine_strict.o &
nvcc $F -c carskit_b200/csrc/fm_engine.cu -o build/fm_engine.o &
wait
nvcc -shared -o carskit_b200/libcarskit_b200_strict.so build/engine_strict.o build/fm_engine.o $(ls build/multi_gpu.o 2>/dev/null) -ldl
