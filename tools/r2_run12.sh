#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -3
STORM_B200_TC_SA=3 STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_qsa4.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_qsa4.json 2> gpurun_out/r2_bench_qsa4.err
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_qsa3.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_qsa3.json 2> gpurun_out/r2_bench_qsa3.err
for f in qsa3 qsa4; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("$f failed", e)
PY
done
