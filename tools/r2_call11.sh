#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vs_reference_kernels.py tests/test_gpu_golden.py -m gpu -x -q -k "fps or pn2 or mvpnet" 2>&1 | tail -6
for args in "32 8192 2048" "32 2048 512" "32 512 128" "32 128 32" "1 8192 2048"; do timeout 100 python tools/fps_prof.py $args 2>&1 | tail -1; done
MVPNET_B200_FPS=regs timeout 100 python tools/fps_prof.py 32 8192 2048 2>&1 | tail -1
