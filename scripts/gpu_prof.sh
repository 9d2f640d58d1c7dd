mkdir -p gpurun_out
# usage: gpu_prof.sh <tag> <kernel-regex> [env assignments...]
tag=$1; shift; re=$1; shift
env "$@" timeout 900 ncu --set full --clock-control none --import-source on -k regex:$re -s 1 -c 1 -o gpurun_out/prof_$tag python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_$tag.log
tail -2 gpurun_out/ncu_$tag.log
