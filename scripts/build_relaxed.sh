#!/bin/bash
# Alternative build of the same ABI WITHOUT the ld.acquire.gpu that ends a successful completion-counter poll
# (-DCARS_RELAXED_POLL, csrc/sgd_kernels.cuh acquire_after_poll): only there to reproduce the A/B that showed the
# formal acquire costs nothing (profiles/r2).  Select it with CARSKIT_B200_LIB=carskit_b200/libcarskit_b200_relaxed.so.
set -e
cd "$(dirname "$0")/.."
mkdir -p build
F="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
nvcc $F -DCARS_RELAXED_POLL -c carskit_b200/csrc/engine.cu -o build/engine_relaxed.o &
nvcc $F -c carskit_b200/csrc/fm_engine.cu -o build/fm_engine.o &
wait
nvcc $F -c carskit_b200/csrc/ingest.cpp -o build/ingest.o
nvcc -shared -o carskit_b200/libcarskit_b200_relaxed.so build/engine_relaxed.o build/fm_engine.o build/ingest.o -ldl
