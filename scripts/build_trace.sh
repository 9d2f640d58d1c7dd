#!/bin/bash
# Developer build of the same ABI with per-rating stage timestamps in sgd_flagged_kernel (-DCARS_TRACE).
# Used only by scripts/trace_flagged.py via CARSKIT_B200_LIB; never loaded by tests, smoke() or bench.py.
set -e
cd "$(dirname "$0")/.."
mkdir -p build
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DCARS_TRACE -c carskit_b200/csrc/engine.cu -o build/engine_trace.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -c carskit_b200/csrc/fm_engine.cu -o build/fm_engine.o
nvcc -shared -o carskit_b200/libcarskit_b200_trace.so build/engine_trace.o build/fm_engine.o
