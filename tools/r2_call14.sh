#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fps" 2>&1 | tail -4
timeout 100 python tools/fps_prof.py 3 32768 8192 2>&1 | tail -1
timeout 100 python tools/fps_prof.py 1 200000 8192 2>&1 | tail -1
timeout 100 python tools/fps_prof.py 1 20000 2048 2>&1 | tail -1
