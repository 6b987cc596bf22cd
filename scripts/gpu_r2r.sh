#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_bench_path.py > gpurun_out/r2r_pytest.log 2>&1; tail -3 gpurun_out/r2r_pytest.log
VORS_JOB_TIMES=1 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo rc=$?
grep "job times" gpurun_out/r2r_bench.err | awk '{printf "%s/%s ", $9, $13} END{print ""}'
python -c "
import json; d=json.load(open('gpurun_out/r2r_bench.json')); p=d['parity_in_run']; r=d['roofline']
print('value %.0f e2e %.0f align_ms %.3f frac %.3f' % (d['value'], d['e2e']['value'], r['avg_launch_ms'], r['frac']), 'parity', p['ok'], p['max_rad'], p['max_m'], p['share_within_1e-4'])"
