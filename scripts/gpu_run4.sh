mkdir -p gpurun_out
CARS_SCHEDULE=flagged CARS_WF_VARIANT=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgd_flagged -s 1 -c 1 -o gpurun_out/prof_fl_100M python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_fl.log
tail -3 gpurun_out/ncu_fl.log
CARS_SCHEDULE=wavefront CARS_WF_VARIANT=2 timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgd_wavefront -s 1 -c 1 -o gpurun_out/prof_wf_100M python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2> gpurun_out/ncu_wf.log
tail -3 gpurun_out/ncu_wf.log
