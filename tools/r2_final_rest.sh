#!/bin/bash
# the non-headline configurations with the final build (after the attention rewrite)
mkdir -p gpurun_out
for c in 4s storm sweep; do
  timeout 900 python bench.py --config $c --steps 2 --warmup 3 > gpurun_out/r2_final_$c.json 2> gpurun_out/r2_final_$c.err
done
timeout 1200 python bench.py --config large8s --steps 2 --warmup 3 > gpurun_out/r2_final_large8s.json 2> gpurun_out/r2_final_large8s.err
for c in 4s storm sweep large8s; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_final_$c.json").read().strip().splitlines()[-1])
    rf=d.get("roofline") or {}
    print("$c", round(d["value"],4), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],4), "conv TF/s", rf.get("achieved"), "frac", rf.get("frac"), (d.get("clocks") or {}).get("sm_mhz"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    for r in d.get("sweep", []): print("   ", r)
except Exception as e: print("$c failed", e)
PY
done
