#!/bin/bash
# ncu --set full capture of selected kernels of the C3 pipeline (tools/profile_pipeline.py), raw page + source page exported as CSV.
# usage: bash tools/ncu_kernels.sh <tag> <kernel-regex> [max launches]
tag=$1; re=$2; cnt=${3:-6}
cmd="python tools/profile_pipeline.py --steps 1"
ncu --set full --clock-control none --import-source on -k "regex:$re" -c $cnt -o gpurun_out/${tag} -f $cmd > gpurun_out/${tag}_run.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/${tag}.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2> /dev/null
ncu -i gpurun_out/${tag}.ncu-rep --page source --csv --print-source sass > gpurun_out/${tag}_source.csv 2> /dev/null
python tools/summarise_ncu_raw.py gpurun_out/${tag}_raw.csv "$tag" > gpurun_out/${tag}_summary.csv
rm -f gpurun_out/${tag}.ncu-rep
cat gpurun_out/${tag}_summary.csv
