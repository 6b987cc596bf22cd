#!/bin/bash
# multi-GPU bench lines (run under `gpurun --gpus N`; every launch under its own `timeout`: a hung collective must not hold the box): default workload, config 4 (one alignment per GPU + gather), config 5 (1080p streams)
N=${1:-8}
mkdir -p gpurun_out
run() {  # name, extra args
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 $2 \
      > gpurun_out/r02_bench_$1_${N}gpu.json 2> gpurun_out/r02_bench_$1_${N}gpu.err; echo "$1 N=$N rc=$?"
  python -c "
import json,sys; d=json.load(open('gpurun_out/r02_bench_$1_${N}gpu.json')); r=d['roofline']
print('  value %.0f e2e %.0f ms/step %.3f align_ms %.3f frac %.3f parity %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], r['avg_launch_ms'], r['frac'], d['parity_in_run'].get('ok')))"
}
run c2 "--no-cpu-baseline"
run c4 "--config 4 --no-cpu-baseline"
run c5 "--config 5 --no-cpu-baseline"
