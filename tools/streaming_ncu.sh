#!/bin/bash
# ncu --set full capture of the streaming kernels of the C3 pipeline (SRE() build, one EM iteration, predict):
# eval_basis (count/fill), grid lookup, binning, Gram on the ELL tiles, ELL quadratic form.  Run the command plain first
# (B200_PROFILING.md recipe), then under ncu; the raw page is exported as CSV so that only text travels back.
# usage: bash tools/streaming_ncu.sh <tag>
tag=${1:-rXX}
cmd="python tools/profile_pipeline.py --steps 1"
$cmd > gpurun_out/${tag}_pipeline.txt 2> gpurun_out/${tag}_pipeline.err || { echo "plain run failed"; tail -5 gpurun_out/${tag}_pipeline.err; exit 1; }
re='regex:eval_count_fast_kernel|eval_fill_warp_kernel|grid_lookup_kernel|bin_count_kernel|bin_scatter_kernel|bin_reduce_kernel|gram_ell|quadform_ell_kernel|quadform_ws_kernel|gather_fill_kernel|ell_fill_kernel|bucket_union_kernel|bucket_count_kernel'
ncu --set full --clock-control none --import-source on -k "$re" -c 22 -o gpurun_out/${tag}_streaming -f $cmd > gpurun_out/${tag}_streaming_run.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/${tag}_streaming.ncu-rep --page raw --csv > gpurun_out/${tag}_streaming_raw.csv 2> /dev/null
python tools/summarise_ncu_raw.py gpurun_out/${tag}_streaming_raw.csv "ncu --set full --clock-control none, C3 pipeline (tools/streaming_ncu.sh ${tag})" > gpurun_out/${tag}_streaming_ncu.csv
rm -f gpurun_out/${tag}_streaming.ncu-rep
wc -c gpurun_out/${tag}_streaming_raw.csv
