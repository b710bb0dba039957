#!/bin/bash
# profile artefacts of round 2 (copied into profiles/ afterwards)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_conv.py -q -s -k "large_shapes or headline or rescale or dynamic_range or half_resolution or checkpoint or storm or discriminative or full_width or baseline_shape or trajectory or small_backbones" 2>&1 | grep -E "rel-L2|passed|failed|sweep prec|res_up" > gpurun_out/r2_parity_values.txt
tail -3 gpurun_out/r2_parity_values.txt
timeout 600 ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_traffic.csv python tools/profile_forward.py > gpurun_out/r2_traffic.log 2>&1
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tcr -s 14 -c 2 -o gpurun_out/r2_ncu_full_R -f python tools/profile_forward.py > gpurun_out/r2_ncu_full_R.log 2>&1
ncu -i gpurun_out/r2_ncu_full_R.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_R_raw.csv 2>/dev/null
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_tcq -s 8 -c 2 -o gpurun_out/r2_ncu_full_Q -f python tools/profile_forward.py > gpurun_out/r2_ncu_full_Q.log 2>&1
ncu -i gpurun_out/r2_ncu_full_Q.ncu-rep --page raw --csv > gpurun_out/r2_ncu_full_Q_raw.csv 2>/dev/null
STORM_B200_TCQ=0 STORM_B200_TCR=0 timeout 300 python tools/conv_timeline.py > gpurun_out/r2_conv_timeline_P.txt 2>&1
STORM_B200_TCQ=1 STORM_B200_TCR=0 timeout 300 python tools/conv_timeline.py > gpurun_out/r2_conv_timeline_Q.txt 2>&1
STORM_B200_TCR=1 timeout 300 python tools/conv_timeline.py > gpurun_out/r2_conv_timeline_R.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_conv.py -q -x -k "test_half_resolution_residual or (test_fused_conv and 0-case) or (test_attention_core and 3-128-192-0)" > gpurun_out/r2_racecheck.log 2>&1
tail -4 gpurun_out/r2_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_conv.py -q -x -k "test_half_resolution_residual or (test_fused_conv and 0-case) or test_attention_core or test_dynamic_range_sweep" > gpurun_out/r2_memcheck.log 2>&1
tail -4 gpurun_out/r2_memcheck.log
ls -la gpurun_out/r2_ncu_full_*.ncu-rep
