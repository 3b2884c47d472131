#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc2.py -m gpu -x -q 2>&1 | tail -3
MVPNET_B200_DEBUG=1 timeout 200 python tools/stage_bench.py 2>&1 | grep "tc2\|SA1\|SA2\|FA" | sort | uniq | head -20
echo "--- TA_ALIGN=32"
MVPNET_B200_TC2_TA_ALIGN=32 timeout 200 python tools/stage_bench.py 2>&1 | grep "tc2" | head
echo "--- GROUPS=3"
MVPNET_B200_TC2_GROUPS=3 timeout 200 python tools/stage_bench.py 2>&1 | grep "tc2" | head
