#!/bin/bash
mkdir -p gpurun_out
STORM_B200_TCR=1 timeout 300 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -3
STORM_B200_TCR=1 timeout 300 python tools/conv_timeline.py 2>&1 | grep -E "b31|b27|b22|b13|b17"
STORM_B200_TCR=1 STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_r1.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_r1.json 2> gpurun_out/r2_bench_r1.err
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_rp.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_rp.json 2> gpurun_out/r2_bench_rp.err
for f in rp r1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("$f failed", e); print(open("gpurun_out/r2_bench_$f.err").read()[-800:])
PY
done
