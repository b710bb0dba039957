#!/bin/bash
mkdir -p gpurun_out
for m in 0 2 0 2; do
STORM_B200_PDL=$m timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_g$m.json 2> gpurun_out/r2_bench_g$m.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_g$m.json").read().strip().splitlines()[-1])
    print("pdl=$m", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("failed", e); print(open("gpurun_out/r2_bench_g$m.err").read()[-800:])
PY
done
