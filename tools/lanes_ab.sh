# A/B at the driver's setting after the memo: lanes of the large block (2 / 3) and of the small ones (4 / 6)
for cfg in "2 4" "3 4" "2 6" "3 6" "2 4" "3 4"; do
set -- $cfg
FRK_SMALL_LANES=$2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-profile --predict-reps 1 --lanes $1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('lanes $1 small_lanes $2: EM it/s',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),'nm',d.get('nelder_mead_last_mstep'))"
done
