#!/bin/bash
# One short GPU call after a change: parity tests, the headline bench line, the two side-stream-bound configs and a launch list.
# Outputs under gpurun_out/ with the given tag (default r02b).
T=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${T}_bench_c2_1gpu.json 2> gpurun_out/${T}_bench_c2_1gpu.err; echo "c2 rc=$?"
for c in 3 5 1; do
  timeout 200 python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/${T}_bench_c${c}_1gpu.json 2> gpurun_out/${T}_bench_c${c}_1gpu.err; echo "c$c rc=$?"
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:vors --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2>&1
python - <<PY
import json
for n in ["c2", "c3", "c5", "c1"]:
    try:
        d = json.load(open(f"gpurun_out/${T}_bench_{n}_1gpu.json")); r = d["roofline"]; p = d["parity_in_run"]
        print(n, "value %.0f e2e %.0f ms/step %.3f align_ms %.3f frac %.3f share %s parity %s %.1e %.1e" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["step_share"], p.get("ok"), p.get("max_rad", -1), p.get("max_m", -1)))
    except Exception as e:
        print(n, "ERR", e)
PY
grep -c k_align gpurun_out/${T}_launches.csv
