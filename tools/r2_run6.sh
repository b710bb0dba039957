#!/bin/bash
export STORM_B200_TCQ=1
for w in 1 2; do
  echo "== TCQ_W=$w"
  STORM_B200_TCQ_W=$w timeout 300 python tools/conv_timeline.py 2>&1 | grep -E "b31.conv0|b32.conv0|b13.conv0|b4.conv0|b32.conv1"
done
