#!/bin/bash
mkdir -p gpurun_out
STORM_B200_GRAPH_DEBUG=1 timeout 600 python -m pytest tests/test_gpu_parity.py -q -s -k "graph_replay" 2>&1 | tail -12
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_g1.json 2> gpurun_out/r2_bench_g1.err
STORM_B200_GRAPH=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_g0.json 2> gpurun_out/r2_bench_g0.err
for f in g0 g1; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"], d["gpu_launches"])
except Exception as e: print("$f failed", e); print(open("gpurun_out/r2_bench_$f.err").read()[-1500:])
PY
done
