#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_tests3.log
tail -15 gpurun_out/r2_tests3.log
for c in headline 4s storm sweep; do
  timeout 600 python bench.py --config $c --steps 2 --warmup 3 > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err || tail -5 gpurun_out/r2_bench_$c.err
done
timeout 900 python bench.py --config large8s --steps 2 --warmup 3 > gpurun_out/r2_bench_large8s.json 2> gpurun_out/r2_bench_large8s.err || tail -5 gpurun_out/r2_bench_large8s.err
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_traffic.csv python tools/profile_forward.py > gpurun_out/r2_traffic.log 2>&1
for c in headline 4s storm sweep large8s; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$c.json").read().strip().splitlines()[-1])
    print("$c", round(d["value"],3), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],3), "conv TF/s", round(d["roofline"]["achieved"],1), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    if "sweep" in d:
        for r in d["sweep"]: print("   ", r)
except Exception as e: print("$c failed", e)
PY
done
