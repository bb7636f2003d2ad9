#!/bin/bash
# DRAM traffic of the dominant kernel group (dgemm_kernel) over two timed EM iterations: ncu, two counters per launch.
# usage: bash tools/traffic_pass.sh <tag>
tag=${1:-rXX}
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile --predict-reps 1"
$cmd > /dev/null 2>&1 || { echo "plain run failed"; exit 1; }
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:dgemm_kernel -s 700 -c 340 \
    --csv --log-file gpurun_out/${tag}_dgemm_traffic.csv $cmd > gpurun_out/${tag}_traffic_run.log 2>&1
echo "ncu rc=$?"; wc -l gpurun_out/${tag}_dgemm_traffic.csv
