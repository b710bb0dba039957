#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q -s -k "pair_kernel" 2>&1 | grep -E "rel-L2|passed|failed|Error" | tail -12
timeout 900 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -2
STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_e1.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_e1.json 2> gpurun_out/r2_bench_e1.err
STORM_B200_TCR_RH=16 STORM_B200_PROFILE_DUMP=gpurun_out/r2_dump_e0.csv timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_e0.json 2> gpurun_out/r2_bench_e0.err
for f in e1 e0; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("failed", e); print(open("gpurun_out/r2_bench_$f.err").read()[-800:])
PY
done
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
