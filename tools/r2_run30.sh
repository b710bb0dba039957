#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -2
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_i1.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_i1.json 2> gpurun_out/r2_bench_i1.err
STORM_B200_TC_GEMM2=0 STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_i0.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_i0.json 2> gpurun_out/r2_bench_i0.err
for f in i1 i0; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("failed", e); print(open("gpurun_out/r2_bench_$f.err").read()[-800:])
PY
done
