#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -s -k attention 2>&1 | tail -12
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_j2.json 2> gpurun_out/r2_bench_j2.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_j2.json").read().strip().splitlines()[-1])
    print("j2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("failed", e); print(open("gpurun_out/r2_bench_j2.err").read()[-800:])
PY
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv -k regex:attn --log-file gpurun_out/r2_attn_j2.csv python tools/profile_forward.py > gpurun_out/r2_attn_j2.log 2>&1
grep -i "attn" gpurun_out/r2_attn_j2.csv | cut -d, -f5,15- | head -4
