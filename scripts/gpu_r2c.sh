#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r2c_pytest.log 2>&1; tail -25 gpurun_out/r2c_pytest.log
bash scripts/ab_bench.sh 2>&1 | tee gpurun_out/r2c_ab.txt
