#!/bin/bash
mkdir -p gpurun_out
python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json')); p=d['parity_in_run']
print('value', d['value'], 'align_ms', d['roofline']['avg_launch_ms'], 'ok', p['ok'], p['max_rad'], p['max_m'], 'tol', p['tol_m'])
for m,v in p['per_oracle_per_arm'].items():
    for arm,w in v.items(): print(m, arm, w)
print('spread', p.get('oracle_f32_vs_f64'))"
VORS_NO_TILED=1 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2h_bench_generic.json 2> gpurun_out/r2h_bench_generic.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_generic.json')); p=d['parity_in_run']
print('GENERIC value', d['value'], 'align_ms', d['roofline']['avg_launch_ms'], 'ok', p['ok'], p['max_rad'], p['max_m'], 'tol', p['tol_m'])
for m,v in p['per_oracle_per_arm'].items():
    for arm,w in v.items(): print(m, arm, w)"
