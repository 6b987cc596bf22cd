#!/bin/bash
mkdir -p gpurun_out
scripts/ubench/texatlas > gpurun_out/r2b_texatlas.txt 2>&1; cat gpurun_out/r2b_texatlas.txt
python -m pytest tests/test_gpu_bench_path.py -m gpu -x -q -s > gpurun_out/r2b_pytest.log 2>&1; tail -8 gpurun_out/r2b_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2b_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2b_bench.json')); print(json.dumps(d['parity_in_run'], indent=1))"
