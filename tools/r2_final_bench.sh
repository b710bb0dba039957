#!/bin/bash
# final measurements of round 2: tests, every bench configuration, reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2_final_tests.log; cat gpurun_out/r2_final_tests.log
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_final_headline.json 2> gpurun_out/r2_final_headline.err
for c in 4s storm sweep; do
  timeout 900 python bench.py --config $c --steps 2 --warmup 3 > gpurun_out/r2_final_$c.json 2> gpurun_out/r2_final_$c.err
done
timeout 1200 python bench.py --config large8s --steps 2 --warmup 3 > gpurun_out/r2_final_large8s.json 2> gpurun_out/r2_final_large8s.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_final_reference.json 2> gpurun_out/r2_final_reference.err
for c in headline 4s storm sweep large8s reference; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_final_$c.json").read().strip().splitlines()[-1])
    rf=d.get("roofline") or {}
    print("$c", round(d["value"],4), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],4), "conv TF/s", rf.get("achieved"), "frac", rf.get("frac"), (d.get("clocks") or {}).get("sm_mhz"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d.get("gpu_launches"))
    for r in d.get("sweep", []): print("   ", r)
except Exception as e: print("$c failed", e)
PY
done
