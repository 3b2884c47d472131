#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc2.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_c4_pytest.log
cat gpurun_out/r2_c4_pytest.log
timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vs_reference_kernels.py -m gpu -x -q -k fps 2>&1 | tail -6
timeout 300 python tools/stage_bench.py 2>&1 | tail -20 | tee gpurun_out/r2_c4_stage.log
