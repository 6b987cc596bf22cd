#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2t_pytest.log 2>&1; tail -5 gpurun_out/r2t_pytest.log
LIB=visual-odometry-rs_b200/lib/libvors_b200.so
cp $LIB /tmp/stock.so
cp visual-odometry-rs_b200/lib_variants/timing.so $LIB
python bench.py --no-cpu-baseline --no-parity --steps 2 --warmup 3 > /tmp/t.json 2> gpurun_out/r2t_timing.txt; grep "^job 0" gpurun_out/r2t_timing.txt | tail -5
cp /tmp/stock.so $LIB
mv visual-odometry-rs_b200/lib_variants/timing.so /tmp/
python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err; echo rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r2t_bench.json')); p=d['parity_in_run']; r=d['roofline']
print('value %.0f e2e %.0f align_ms %.3f frac %.3f' % (d['value'], d['e2e']['value'], r['avg_launch_ms'], r['frac']), 'parity', p['ok'], p['max_rad'], p['max_m'], p['share_within_1e-4'])"
python scripts/bench_single.py > gpurun_out/r2t_single.json 2> gpurun_out/r2t_single.err; python -c "
import json
for r in json.load(open('gpurun_out/r2t_single.json')): print(r['shape'], r['levels'], r['mode'], 'gpu_ms %.3f cpu_ms %.3f speedup %.1f err %.1e %.1e' % (r['gpu_ms'], r['cpu_ms'], r['speedup'], r['max_pose_diff_rad'], r['max_pose_diff_m']))"
