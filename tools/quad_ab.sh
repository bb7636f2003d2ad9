for i in 1 2; do
for m in 24 8; do  # (FRK_QUAD_WS_MIN overrides the per-context default: 8, and 24 beside the Nelder-Mead lanes)
FRK_QUAD_WS_MIN=$m python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-profile --predict-reps 3 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ws_min',$m,'EM it/s',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),'predict',d['predict']['value'])"
done; done
