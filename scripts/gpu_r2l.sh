#!/bin/bash
mkdir -p gpurun_out
AB_STEPS=20 AB_WARMUP=5 AB_ARGS="--no-parity" bash scripts/ab_bench.sh 2>&1 | tee gpurun_out/r2l_ab.txt
