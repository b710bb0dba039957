#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_f1.json 2> gpurun_out/r2_bench_f1.err
STORM_B200_PDL=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_f0.json 2> gpurun_out/r2_bench_f0.err
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_f1b.json 2> gpurun_out/r2_bench_f1b.err
for f in f1 f0 f1b; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"], d.get("gpu_launches"))
except Exception as e: print("failed", e); print(open("gpurun_out/r2_bench_$f.err").read()[-800:])
PY
done
