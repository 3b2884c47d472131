#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc2.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/stage_bench.py 2>&1 | grep "tc2\|FP" | head -20
