# A/B of the saved-inverse reuse in the block-exponential M-step (FRK_NO_KINV_REUSE=1 switches it off), driver's setting and default
for rep in 1 2; do
for off in 1 0; do
if [ $off = 1 ]; then export FRK_NO_KINV_REUSE=1; else unset FRK_NO_KINV_REUSE; fi
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-profile --predict-reps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('reuse_off',$off,'5+20: EM it/s',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),'nm',d.get('nelder_mead_last_mstep'),'llk',d['llk_last'][-1], d['llk_trace_from_init'][-1])"
done; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile --predict-reps 1 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('3+10: EM it/s',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2))"
