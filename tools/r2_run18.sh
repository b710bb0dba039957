#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_rp.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_rp.json 2> gpurun_out/r2_bench_rp.err
STORM_B200_TCR=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_r0.json 2> gpurun_out/r2_bench_r0.err
for f in r0 rp; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("$f failed", e); print(open("gpurun_out/r2_bench_$f.err").read()[-800:])
PY
done
