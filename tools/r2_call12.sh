#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vs_reference_kernels.py tests/test_gpu_scene.py -m gpu -x -q -k "fps or scene" 2>&1 | tail -6
for args in "32 8192 2048" "32 2048 512" "1 200000 8192" "3 32768 8192"; do timeout 200 python tools/fps_prof.py $args 2>&1 | tail -1; done
MVPNET_B200_FPS=generic timeout 200 python tools/fps_prof.py 3 32768 8192 2>&1 | tail -1
