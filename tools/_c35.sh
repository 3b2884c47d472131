#!/bin/bash
for w in 32 16 8; do echo "--- FPS_WARPS=$w"; MVPNET_B200_FPS_WARPS=$w timeout 100 python tools/fps_prof.py 32 8192 2048 2>&1 | tail -1; MVPNET_B200_FPS_WARPS=$w timeout 100 python tools/fps_prof.py 1 8192 2048 2>&1 | tail -1; done
MVPNET_B200_FPS_WARPS=16 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vs_reference_kernels.py -m gpu -x -q -k fps 2>&1 | tail -2
MVPNET_B200_FPS_WARPS=8 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vs_reference_kernels.py -m gpu -x -q -k fps 2>&1 | tail -2
