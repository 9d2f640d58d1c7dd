#!/bin/bash
mkdir -p gpurun_out
for cfg in "2 1" "4 1" "2 2" "4 2" "1 1"; do set -- $cfg
 a=$(CARS_FM_PPG_SHORT=$1 CARS_FM_PPG_LONG=$2 timeout 600 python bench.py --workload fm_k64_250Kx25Kx32c_25M --steps 3 --warmup 1 --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*')
 b=$(CARS_FM_PPG_SHORT=$1 CARS_FM_PPG_LONG=$2 timeout 900 python bench.py --workload fm_k64_5Mx500Kx32c_125M_per_gpu --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | grep -o '"ms_per_step": [0-9.]*')
 echo "ppg short=$1 long=$2 : 25M $a ; 125M $b"
done
