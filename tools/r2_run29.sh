#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q -k attention 2>&1 | tail -2
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_h1.json 2> gpurun_out/r2_bench_h1.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_h1.json").read().strip().splitlines()[-1])
    print("h1", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("failed", e); print(open("gpurun_out/r2_bench_h1.err").read()[-800:])
PY
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_traffic_h1.csv python tools/profile_forward.py > gpurun_out/r2_traffic_h1.log 2>&1
tail -1 gpurun_out/r2_traffic_h1.log
