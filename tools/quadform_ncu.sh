#!/bin/bash
# ncu --set full capture (with SASS-level stall sampling) of the warp-specialised quadratic form on S0 at C3.
# usage: bash tools/quadform_ncu.sh <tag>     (run tools/quadform_bench.py plain first: B200_PROFILING.md recipe)
tag=${1:-rXX}
export FRK_QUAD_NCU=1
ncu --set full --clock-control none --import-source on -k regex:quadform_ws_kernel -c 1 -o gpurun_out/${tag}_quadws -f \
    python tools/quadform_bench.py > gpurun_out/${tag}_quadws_run.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/${tag}_quadws.ncu-rep --page raw --csv > gpurun_out/${tag}_quadws_raw.csv 2> /dev/null
ncu -i gpurun_out/${tag}_quadws.ncu-rep --page source --csv > gpurun_out/${tag}_quadws_source.csv 2> /dev/null
python tools/summarise_ncu_raw.py gpurun_out/${tag}_quadws_raw.csv "ncu --set full --clock-control none, quadform_ws_kernel on S0 at C3 (tools/quadform_ncu.sh ${tag})" > gpurun_out/${tag}_quadws_summary.csv
rm -f gpurun_out/${tag}_quadws.ncu-rep
wc -c gpurun_out/${tag}_quadws_raw.csv gpurun_out/${tag}_quadws_source.csv
