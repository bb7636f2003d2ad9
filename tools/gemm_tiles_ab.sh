# A/B of the 64-row tile threshold (tiles per SM below which 64-row tiles are used) at the driver's setting
for v in 4 8 2 4 8; do
FRK_GEMM_SMALL_PER_SM=$v python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-profile --predict-reps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('small_per_sm $v: EM it/s',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2))"
done
