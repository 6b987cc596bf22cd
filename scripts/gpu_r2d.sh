#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; tail -30 gpurun_out/r2d_pytest.log
LIB=visual-odometry-rs_b200/lib/libvors_b200.so
cp $LIB /tmp/stock.so
cp visual-odometry-rs_b200/lib_variants/timing.so $LIB
python bench.py --no-cpu-baseline --no-parity --steps 2 --warmup 3 > /tmp/t.json 2> gpurun_out/r2d_timing.txt; grep "^job" gpurun_out/r2d_timing.txt | tail -12
cp /tmp/stock.so $LIB
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_align -s 3 -c 1 -o gpurun_out/r2d_prof python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > gpurun_out/r2d_ncu.log 2>&1; tail -3 gpurun_out/r2d_ncu.log
