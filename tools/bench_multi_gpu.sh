#!/bin/bash
# usage: run_n.sh N [extra bench args]
N=$1; shift
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline "$@" 2> gpurun_out/err_n$N.log | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.readline())
print('N', d['n_gpus'], 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'e2e %.3e'%(d.get('e2e',{}).get('value') or 0), '|', d['config']['step_ms_first10'], d['config']['collective'][40:100])"
