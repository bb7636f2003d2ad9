#!/bin/bash
# bench.py at N ranks in the multi-GPU modes (replicated candidates / distributed / distributed with two lanes); one line each
N=$1; shift
for mode in "" "--distribute-nm" "--distribute-nm --lanes 2"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-profile $mode 2>gpurun_out/scale_n${N}.err | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: continue
    print('N=$N mode=[$mode]', round(d['value'],1), 'it/s', round(d['ms_per_step'],3), 'ms  e2e', round(d['e2e']['value'],1), ' predict %.3g BAUs/s' % d['predict']['value'], ' parity', d.get('parity',{}).get('llk_rel'), d.get('parity',{}).get('ok'), d.get('nelder_mead_last_mstep'))"
done
