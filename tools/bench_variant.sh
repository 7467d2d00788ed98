#!/bin/bash
# usage: run_var.sh label [bench args]
L=$1; shift
python bench.py "$@" > gpurun_out/bench_$L.json 2> gpurun_out/bench_$L.err || { echo "$L FAILED"; tail -5 gpurun_out/bench_$L.err; }
python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_$L.json').readlines()[-1])
print('$L', 'value %.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], 'kernel_ms %.3f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], 'cross/step %.3e'%d['config']['crossings_per_step'], 'cpu %.3e'%d.get('cpu_baseline',{}).get('value',0))"
