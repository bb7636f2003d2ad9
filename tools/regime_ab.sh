# A/B at the driver's bench setting once the big block's factorisations are memoised: GEMM tile heuristic x speculation width of small blocks
for cfg in "1 0" "4 0" "1 4" "4 4" "1 0" "4 0"; do
set -- $cfg
FRK_GEMM_SMALL_PER_SM=$1 FRK_SMALL_LANES=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-profile --predict-reps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('small_per_sm $1 small_lanes $2: EM it/s',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),'nm',d.get('nelder_mead_last_mstep'))"
done
