#!/bin/bash
mkdir -p gpurun_out
export STORM_B200_TCR=1
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -12
timeout 300 python tools/conv_timeline.py 2>&1 | grep -E "b31|b27|b22|b13|b17" 
