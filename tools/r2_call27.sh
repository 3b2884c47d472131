#!/bin/bash
timeout 200 python -m pytest tests/test_gpu_tc2.py -m gpu -x -q 2>&1 | tail -2
MVPNET_B200_TC2_TWO_CTAS=1 timeout 200 python -m pytest tests/test_gpu_tc2.py -m gpu -x -q 2>&1 | tail -2
MVPNET_B200_DEBUG=1 MVPNET_B200_TC2_TWO_CTAS=1 timeout 200 python tools/stage_bench.py 2>&1 | grep "tc2" | sort | uniq | head
timeout 200 python tools/stage_bench.py 2>&1 | grep "tc2" | head -4
