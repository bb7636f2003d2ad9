python bench.py > gpurun_out/r01f_bench.json 2> gpurun_out/r01f_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r01f_bench.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['predict']['value'], d['roofline']['frac'], d['cpu_baseline']['value'])"
for L in 1 3; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile --predict-reps 1 --lanes $L 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lanes',d['config']['nm_lanes'],d['value'],d['ms_per_step'],d['e2e']['value'],d.get('nelder_mead_last_mstep'))"; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --k-type unstructured > gpurun_out/r01f_bench_unstructured.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r01f_bench_unstructured.json')); print('unstructured', d['value'], d['ms_per_step'], d['e2e']['value'])"
bash tools/run_profile_pass.sh r01f
