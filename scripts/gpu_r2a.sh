#!/bin/bash
# round 2, GPU call A: parity tests (incl. the benchmarked-path test), default bench with parity_in_run, the other configs,
# and the texture-atlas microbenchmark.
mkdir -p gpurun_out
nvidia-smi -L
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
scripts/ubench/texatlas > gpurun_out/r2a_texatlas.txt 2>&1; cat gpurun_out/r2a_texatlas.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; tail -15 gpurun_out/r2a_pytest.log
python bench.py > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r2a_bench.err
for c in 1 3 5 4; do
  python bench.py --config $c --steps 20 --warmup 3 > gpurun_out/r2a_bench_c$c.json 2> gpurun_out/r2a_bench_c$c.err; echo "config $c rc=$?"; tail -2 gpurun_out/r2a_bench_c$c.err
done
python - <<'PY'
import json
for n in ["", "_c1", "_c3", "_c5", "_c4"]:
    try:
        d = json.load(open(f"gpurun_out/r2a_bench{n}.json"))
        print(n or "_c2", "value %.0f e2e %.0f frac %.3f align_ms %.3f parity %s cpu %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["avg_launch_ms"], d.get("parity_in_run"), d.get("cpu_baseline", {}).get("value")))
    except Exception as e:
        print(n, "ERR", e)
PY
