#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q -k "half_resolution" 2>&1 | tail -2
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_o0.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_o0.json 2> gpurun_out/r2_bench_o0.err
STORM_B200_OUT_PREC=1 STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_o1.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_o1.json 2> gpurun_out/r2_bench_o1.err
for f in o0 o1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("$f failed", e); print(open("gpurun_out/r2_bench_$f.err").read()[-800:])
PY
done
