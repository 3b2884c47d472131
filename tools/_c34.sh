#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_grid.py tests/test_gpu_unproject.py tests/test_gpu_golden.py tests/test_gpu_vs_reference_kernels.py tests/test_scene.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/ops_prof.py 2>&1 | grep "knn_pixels\|knn_distance"
