#!/bin/bash
# GPU call 1 of round 2: full -m gpu suite, then same-box A/B benches of the new knobs
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r2_tests1.log
B="python bench.py --steps 3 --warmup 2"
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_a.csv timeout 300 $B > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
STORM_B200_PF_DEPTH=1 STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_pf1.csv timeout 300 $B > gpurun_out/r2_bench_pf1.json 2>/dev/null
STORM_B200_PF_DEPTH=8 STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_pf8.csv timeout 300 $B > gpurun_out/r2_bench_pf8.json 2>/dev/null
STORM_B200_NO_LOWRES_SKIP=1 STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_nolow.csv timeout 300 $B > gpurun_out/r2_bench_nolow.json 2>/dev/null
tail -5 gpurun_out/r2_tests1.log
for f in a pf1 pf8 nolow; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("$f failed", e)
PY
done
