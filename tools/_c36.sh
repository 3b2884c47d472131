#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vs_reference_kernels.py tests/test_gpu_golden.py -m gpu -x -q -k "fps or pn2 or mvpnet" 2>&1 | tail -2
for n in "32 8192 2048" "32 4096 1024" "32 2048 512" "32 512 128"; do timeout 100 python tools/fps_prof.py $n 2>&1 | tail -1; done
echo "--- BUCKET_MIN=512"
for n in "32 2048 512" "32 512 128" "32 1024 256"; do MVPNET_B200_FPS_BUCKET_MIN=512 timeout 100 python tools/fps_prof.py $n 2>&1 | tail -1; done
MVPNET_B200_FPS_BUCKET_MIN=100 timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vs_reference_kernels.py -m gpu -x -q -k "fps" 2>&1 | tail -2
