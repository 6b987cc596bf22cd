#!/bin/bash
mkdir -p gpurun_out
LIB=visual-odometry-rs_b200/lib/libvors_b200.so
cp $LIB /tmp/stock.so
for t in timing w8timing; do
  cp visual-odometry-rs_b200/lib_variants/$t.so $LIB
  python bench.py --no-cpu-baseline --no-parity --steps 2 --warmup 3 > /tmp/t.json 2> gpurun_out/r2f_$t.txt; echo "== $t"; grep "^job 0" gpurun_out/r2f_$t.txt | tail -5
  mv visual-odometry-rs_b200/lib_variants/$t.so /tmp/
done
cp /tmp/stock.so $LIB
bash scripts/ab_bench.sh 2>&1 | tee gpurun_out/r2f_ab.txt
python -m pytest tests/test_gpu_bench_path.py -m gpu -q -s > gpurun_out/r2f_pytest.log 2>&1; tail -6 gpurun_out/r2f_pytest.log
