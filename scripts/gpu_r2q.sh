#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; tail -6 gpurun_out/r2q_pytest.log
VORS_JOB_TIMES=1 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo rc=$?; grep "job times" gpurun_out/r2q_bench.err | tail -2
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench.json')); p=d['parity_in_run']; r=d['roofline']
print('value %.0f e2e %.0f align_ms %.3f frac %.3f' % (d['value'], d['e2e']['value'], r['avg_launch_ms'], r['frac']), 'parity', p['ok'], p['max_rad'], p['max_m'], p['share_within_1e-4'], 'cpu', d['cpu_baseline']['value'])"
