# N = 2 at the driver's bench setting (5 + 20 iterations): distributed Nelder-Mead candidates (default for N > 1) against replicated ones
for mode in "" "--replicate-nm"; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-profile --predict-reps 3 $mode 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=2 mode [$mode]','EM it/s',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),'predict',d['predict']['value'],'nm',d.get('nelder_mead_last_mstep'),'parity',d.get('parity'))"
done
