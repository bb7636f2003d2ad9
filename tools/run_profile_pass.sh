#!/bin/bash
# ncu launch list of the bench command (B200_PROFILING.md recipe): run the command without ncu first, then under ncu.
# usage: bash tools/run_profile_pass.sh <tag> [extra bench flags]
tag=${1:-rXX}; shift
cmd="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile --predict-reps 1 $*"
$cmd > gpurun_out/${tag}_pre_ncu.json 2> gpurun_out/${tag}_pre_ncu.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 1500 -c 1600 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu_run.log 2>&1
rc=$?
echo "ncu rc=$rc ($cmd)"; wc -l gpurun_out/${tag}_launches.csv
exit $rc
