#!/bin/bash
# Round-2 artifacts on one B200: GPU tests, bench lines of every BASELINE config (+ reference arm), launch list and ncu capture
# of the dominant kernel, single-stream latencies.  Outputs under gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_pytest.log 2>&1; tail -3 gpurun_out/r02_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c2_1gpu.json 2> gpurun_out/r02_bench_c2_1gpu.err; echo "c2 rc=$?"
for c in 1 3 4 5; do
  python bench.py --config $c --steps 20 --warmup 5 > gpurun_out/r02_bench_c${c}_1gpu.json 2> gpurun_out/r02_bench_c${c}_1gpu.err; echo "c$c rc=$?"
done
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r02_bench_reference.json 2> /dev/null
python scripts/bench_single.py > gpurun_out/r02_single.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name-base demangled -k regex:vors --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > /dev/null 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:k_align -s 9 -c 1 -o gpurun_out/r02_prof python bench.py --steps 2 --warmup 9 --no-cpu-baseline --no-parity > gpurun_out/r02_ncu.log 2>&1; tail -1 gpurun_out/r02_ncu.log
python - <<'PY'
import json
for n in ["c2", "c1", "c3", "c4", "c5"]:
    try:
        d = json.load(open(f"gpurun_out/r02_bench_{n}_1gpu.json")); r = d["roofline"]; p = d["parity_in_run"]
        print(n, "value %.0f e2e %.0f ms/step %.3f align_ms %.3f frac %.3f GB/s %.0f cpu %.1f parity %s %.1e %.1e" % (d["value"], d["e2e"]["value"], d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["achieved"], d.get("cpu_baseline", {}).get("value", -1), p.get("ok"), p.get("max_rad", -1), p.get("max_m", -1)))
    except Exception as e:
        print(n, "ERR", e)
PY
