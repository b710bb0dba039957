#!/bin/bash
mkdir -p gpurun_out
export STORM_B200_TCQ=1
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -15 > gpurun_out/r2_tests4_conv.log
cat gpurun_out/r2_tests4_conv.log | tail -8
if grep -q "passed" gpurun_out/r2_tests4_conv.log && ! grep -q "failed" gpurun_out/r2_tests4_conv.log; then
  timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_tests4.log; tail -6 gpurun_out/r2_tests4.log
  timeout 300 python tools/conv_timeline.py > gpurun_out/r2_timeline4_q.txt 2>&1; cat gpurun_out/r2_timeline4_q.txt
  STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_q1.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_q1.json 2> gpurun_out/r2_bench_q1.err
fi
export STORM_B200_TCQ=0
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_q0.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_q0.json 2> gpurun_out/r2_bench_q0.err
for f in q0 q1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("$f failed", e)
PY
done
