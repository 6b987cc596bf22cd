#!/bin/bash
# quick GPU check: parity tests (quiet) + bench summary
python -m pytest tests -m gpu -q -x 2>&1 | tail -4
python bench.py --no-cpu-baseline > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err
python -c "
import json; d=json.load(open('gpurun_out/q_bench.json')); r=d['roofline']; print('value %.0f e2e %.0f frac %.3f align_ms %.3f share %s posediff %s' % (d['value'], d['e2e']['value'], r['frac'], r['avg_launch_ms'], r['step_share'], d['config']['arms_max_abs_pose_diff']))"
tail -2 gpurun_out/q_bench.err
