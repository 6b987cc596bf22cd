#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; tail -12 gpurun_out/r2e_pytest.log
LIB=visual-odometry-rs_b200/lib/libvors_b200.so
cp $LIB /tmp/stock.so
cp visual-odometry-rs_b200/lib_variants/timing.so $LIB
python bench.py --no-cpu-baseline --no-parity --steps 2 --warmup 3 > /tmp/t.json 2> gpurun_out/r2e_timing.txt; grep "^job 0" gpurun_out/r2e_timing.txt | tail -5
cp /tmp/stock.so $LIB
mv visual-odometry-rs_b200/lib_variants/timing.so /tmp/
bash scripts/ab_bench.sh 2>&1 | tee gpurun_out/r2e_ab.txt
