#!/bin/bash
# profile collection for profiles/ (round 2)
mkdir -p gpurun_out
timeout 900 python bench.py --config sweep --steps 2 --warmup 3 > gpurun_out/r2_bench_sweep.json 2> gpurun_out/r2_bench_sweep.err || tail -5 gpurun_out/r2_bench_sweep.err
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_traffic.csv python tools/profile_forward.py > gpurun_out/r2_traffic.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tc -s 44 -c 8 -o gpurun_out/r2_ncu_full -f python tools/profile_forward.py > gpurun_out/r2_ncu_full.log 2>&1
ncu -i gpurun_out/r2_ncu_full.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_raw.csv 2>/dev/null
STORM_B200_TCQ=0 timeout 300 python tools/conv_timeline.py > gpurun_out/r2_timeline_P.txt 2>&1
STORM_B200_TCQ=1 timeout 300 python tools/conv_timeline.py > gpurun_out/r2_timeline_Q.txt 2>&1
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_conv.py -q -x -k "test_half_resolution_residual or (test_fused_conv and 0-case1) or (test_attention_core and 0-3-128-192)" > gpurun_out/r2_racecheck.log 2>&1
tail -15 gpurun_out/r2_racecheck.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_conv.py -q -x -k "test_half_resolution_residual or (test_fused_conv and 0-case) or test_attention_core" > gpurun_out/r2_memcheck.log 2>&1
tail -8 gpurun_out/r2_memcheck.log
python - <<PY
import json
for c in ["sweep"]:
    try:
        d=json.loads(open(f"gpurun_out/r2_bench_{c}.json").read().strip().splitlines()[-1])
        print(c, round(d["value"],3), round(d["ms_per_step"],1), "e2e", round(d["e2e"]["value"],3))
        for r in d.get("sweep", []): print("   ", r)
    except Exception as e: print(c, "failed", e)
PY
ls -la gpurun_out/r2_ncu_full* | head
