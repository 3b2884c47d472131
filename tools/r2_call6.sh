#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_train.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r2_c6_pytest.log
cat gpurun_out/r2_c6_pytest.log
timeout 300 python tools/stage_bench.py 2>&1 | tail -20 | tee gpurun_out/r2_c6_stage.log
MVPNET_B200_TC2_GROUPS=2 timeout 300 python tools/stage_bench.py 2>&1 | grep "tc2" | tee -a gpurun_out/r2_c6_stage.log
