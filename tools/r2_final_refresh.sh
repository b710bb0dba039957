#!/bin/bash
# refresh of the artefacts that depend on the attention rewrite (after tools/r2_final_all.sh)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 4 --warmup 3 > gpurun_out/r2_final_headline.json 2> gpurun_out/r2_final_headline.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_final_headline.json").read().strip().splitlines()[-1])
print("headline", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac"], d["clocks"], d["gpu_launches"], d["cpu_baseline"]["value"])
PY
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_traffic.csv python tools/profile_forward.py > gpurun_out/r2_traffic.log 2>&1
SEL="test_attention_core"
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_conv.py -q -x -k "$SEL and 0-" > gpurun_out/r2_racecheck_attn.log 2>&1
tail -2 gpurun_out/r2_racecheck_attn.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_conv.py -q -x -k "$SEL" > gpurun_out/r2_memcheck_attn.log 2>&1
tail -2 gpurun_out/r2_memcheck_attn.log
