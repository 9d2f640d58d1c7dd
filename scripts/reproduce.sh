#!/bin/bash
# How the numbers under profiles/r2/ and in DESIGN.md were produced.  Run on a B200 box from the repo root (through
# `gpurun -- 'bash scripts/reproduce.sh <what>'` in the build environment).  Every step writes under gpurun_out/.
#   tests      the GPU parity suite (FAST, tagged rows, reference-bytecode vectors, JNI glue, rank, FM, ingest)
#   bench      the bench lines: EXACT (the default, with the parity digest at 100 M ratings), FAST uniform / Zipf(1.0)
#   evalk      predict / rank kernels: end-to-end line + launch list
#   ncu <mode> <workload> <kernel-regex> <tag>   one `ncu --set full` capture, summarised on the box (reports are ~25 MB)
#   multi N    the scaling bench as the driver launches it (N = 2, 4, 8) + the multi-GPU tests
#   fm         FM at the config-4 shard size (125 M rows, k = 64): bench line, launch list, ncu summaries of its kernels
#   litmus     the 128-byte line atomicity litmus test
set -u
mkdir -p gpurun_out
B=camf_ci_f64_1Mx100Kx32c_100M
case "${1:-tests}" in
  tests)
    timeout 2400 python -m pytest tests -m gpu -q | tee gpurun_out/pytest_gpu.log | tail -5
    python -c "import __graft_entry__ as g; g.smoke()" ;;
  bench)
    python bench.py --steps 10 --warmup 3 > gpurun_out/bench_exact_100M.json
    python bench.py --mode fast --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fast_uniform_100M.json
    python bench.py --mode fast --workload ${B}_zipf1.0 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fast_zipf_100M.json
    python bench.py --workload camf_ci_f64_100Kx10Kx32c_10M_zipf1.0 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_exact_zipf_10M.json
    python bench.py --workload camf_c_f10_frappe_shaped --mode fast --steps 20 --warmup 3 > gpurun_out/bench_config2_fast.json
    python bench.py --tuning "tagged=1;tagged_ctas=2" --steps 10 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/bench_tagged_100M.json ;;
  evalk)
    python scripts/bench_eval_kernels.py | tee gpurun_out/eval_kernels_e2e.jsonl
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/launches_eval.csv python scripts/bench_eval_kernels.py 10000000 4000 > /dev/null
    python scripts/launch_summary.py gpurun_out/launches_eval.csv ;;
  ncu)
    mode=$2; wl=$3; re=$4; tag=$5
    ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -o gpurun_out/prof_$tag \
        python bench.py --workload $wl --mode $mode --steps 1 --warmup 1 --no-cpu-baseline --no-parity > gpurun_out/ncu_$tag.log 2>&1
    python scripts/ncu_summary.py gpurun_out/prof_$tag.ncu-rep 25 | tee gpurun_out/ncu_summary_$tag.txt
    rm -f gpurun_out/prof_$tag.ncu-rep ;;
  multi)
    N=${2:-2}
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py \
        --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json
    timeout 1200 python -m pytest tests/test_multi_gpu.py -q | tail -3 ;;
  fm)
    F=fm_k64_5Mx500Kx32c_125M_per_gpu
    python bench.py --workload $F --steps 3 --warmup 1 > gpurun_out/bench_fm125M.json
    python bench.py --workload $F --steps 3 --warmup 1 --no-cpu-baseline --tuning "fm_runs=0" > gpurun_out/bench_fm125M_gather.json
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1200 --csv \
        --log-file gpurun_out/launches_fm.csv python bench.py --workload $F --steps 1 --warmup 0 --no-cpu-baseline > /dev/null
    python scripts/launch_summary.py gpurun_out/launches_fm.csv | tee gpurun_out/launches_fm_125M.txt
    for k in fm_run_reduce fm_piece_reduce fm_dense_reduce fm_row_update; do
      ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_$k \
          python bench.py --workload $F --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1
      python scripts/ncu_summary.py gpurun_out/prof_$k.ncu-rep 25 > gpurun_out/ncu_summary_${k}_125M.txt
      rm -f gpurun_out/prof_$k.ncu-rep
    done ;;
  litmus)
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/litmus/line_atomicity scripts/litmus/line_atomicity.cu
    ./scripts/litmus/line_atomicity | tee gpurun_out/litmus_line_atomicity.txt ;;
esac
