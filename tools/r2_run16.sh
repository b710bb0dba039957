#!/bin/bash
mkdir -p gpurun_out
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_r0.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_r0.json 2> gpurun_out/r2_bench_r0.err
STORM_B200_TCR=1 STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_r1.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_r1.json 2> gpurun_out/r2_bench_r1.err
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_r0b.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_r0b.json 2> gpurun_out/r2_bench_r0b.err
for f in r0 r1 r0b; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("$f failed", e); print(open("gpurun_out/r2_bench_$f.err").read()[-800:])
PY
done
