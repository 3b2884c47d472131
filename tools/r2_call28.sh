#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_golden.py tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python tools/stage_bench.py 2>&1 | grep "tc2\|SA1 rel" | head -6
