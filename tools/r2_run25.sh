#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q -k fir_resample 2>&1 | tail -5
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_d2.json 2> gpurun_out/r2_bench_d2.err
STORM_B200_RESAMPLE_TILED=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_bench_d2t.json 2> gpurun_out/r2_bench_d2t.err
for f in d2 d2t; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r2_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["clocks"]["sm_mhz"])
except Exception as e: print("failed", e); print(open("gpurun_out/r2_bench_$f.err").read()[-800:])
PY
done
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_traffic_d2.csv python tools/profile_forward.py > gpurun_out/r2_traffic_d2.log 2>&1
tail -1 gpurun_out/r2_traffic_d2.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -2
