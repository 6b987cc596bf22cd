#!/bin/bash
mkdir -p gpurun_out
LIB=visual-odometry-rs_b200/lib/libvors_b200.so
cp $LIB /tmp/stock.so
cp visual-odometry-rs_b200/lib_variants/timing.so $LIB
python bench.py --no-cpu-baseline --no-parity --steps 2 --warmup 3 > /tmp/t.json 2> gpurun_out/r2p_timing.txt; grep "^job 0" gpurun_out/r2p_timing.txt | tail -5
cp /tmp/stock.so $LIB
mv visual-odometry-rs_b200/lib_variants/timing.so /tmp/
VORS_JOB_TIMES=1 python bench.py --no-cpu-baseline --steps 20 --warmup 5 > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err; echo rc=$?; grep "job times" gpurun_out/r2p_bench.err | tail -3
python -c "
import json; d=json.load(open('gpurun_out/r2p_bench.json')); p=d['parity_in_run']; r=d['roofline']
print('value %.0f e2e %.0f align_ms %.3f frac %.3f' % (d['value'], d['e2e']['value'], r['avg_launch_ms'], r['frac']), 'parity', p['ok'], p['max_rad'], p['max_m'], 'tol', p['tol_m'])"
