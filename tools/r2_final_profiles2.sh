#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tcr -s 27 -c 6 -o gpurun_out/r2_ncu_full_R -f python tools/profile_forward.py > gpurun_out/r2_ncu_full_R.log 2>&1
ncu -i gpurun_out/r2_ncu_full_R.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_R_raw.csv 2>/dev/null
STORM_B200_TCR=0 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_conv.py -q -x -k "test_half_resolution_residual or (test_fused_conv and 0-case) or (test_attention_core and 3-128-192-0)" > gpurun_out/r2_racecheck_PQ.log 2>&1
tail -3 gpurun_out/r2_racecheck_PQ.log
