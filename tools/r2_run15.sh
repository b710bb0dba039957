#!/bin/bash
export STORM_B200_TCR=1
timeout 300 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -3
for cfg in "3 8" "4 6" "2 9" "5 5"; do set -- $cfg
  echo "== sa=$1 sw=$2"
  STORM_B200_TCR_SA=$1 STORM_B200_TCR_SW=$2 timeout 300 python tools/conv_timeline.py 2>&1 | grep -E "b31.conv0|b27.conv0|b27.conv1|b22.conv0"
done
