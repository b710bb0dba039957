#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_tests2.log
timeout 300 python tools/conv_timeline.py > gpurun_out/r2_timeline2.txt 2>&1
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_b.csv timeout 300 python bench.py --steps 3 --warmup 2 > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -12 gpurun_out/r2_tests2.log; cat gpurun_out/r2_timeline2.txt
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_b.json").read().strip().splitlines()[-1])
print("b", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
PY
