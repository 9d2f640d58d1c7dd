#!/bin/bash
mkdir -p gpurun_out
python -m tests.config1_depaulmovie .scratch/Movie_DePaulMovie/ratings.txt --gpu > gpurun_out/r49_config1.log 2>&1; tail -8 gpurun_out/r49_config1.log
python scripts/config2.py 10 > gpurun_out/r49_config2.log 2>&1; tail -1 gpurun_out/r49_config2.log
