mkdir -p gpurun_out
for t in 0 1 2; do
CARS_FLAGGED_TUNE=$t timeout 600 python bench.py --steps 5 --warmup 2 --no-cpu-baseline 2>&1 >/dev/null | grep "epochs in" | sed "s/^/tune=$t /"
done
