#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -3
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_c1.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_c1.json 2> gpurun_out/r2_bench_c1.err
STORM_B200_TCR=0 STORM_B200_TCQ=2 timeout 300 python tools/conv_timeline.py > gpurun_out/r2_timeline_c1_Q.txt 2>&1
for f in c1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("$f failed", e); print(open("gpurun_out/r2_bench_$f.err").read()[-800:])
PY
done
cut -c1-150 gpurun_out/r2_timeline_c1_Q.txt
