#!/bin/bash
for cfg in "x x" "1 2" "1 3" "2 2"; do set -- $cfg; g=$1; c=$2; echo "--- TC_GROUPS=$g STAGES_CAP=$c"
  if [ "$g" = "x" ]; then MVPNET_B200_DEBUG=1 timeout 200 python tools/stage_bench.py 2>&1 | grep "FP4 tc\|FP3 tc\|mode=2" | sort | uniq | head -6
  else MVPNET_B200_TC_GROUPS=$g MVPNET_B200_TC_STAGES_CAP=$c MVPNET_B200_DEBUG=1 timeout 200 python tools/stage_bench.py 2>&1 | grep "FP4 tc\|FP3 tc\|mode=2" | sort | uniq | head -6; fi
done
