# A/B of FRK_SMALL_LANES (speculation width for resolutions with <= 1024 functions) at the driver's bench setting (5 + 20 iterations)
for m in 0 4 8 0 4 8; do
FRK_SMALL_LANES=$m python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-profile --predict-reps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('small_lanes',$m,'EM it/s',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),'nm',d.get('nelder_mead_last_mstep'),'llk_last',d['llk_last'][-1])"
done
