#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fps_large" 2>&1 | tail -4
timeout 100 python tools/fps_prof.py 3 32768 8192 2>&1 | tail -1
timeout 100 python tools/fps_prof.py 1 200000 8192 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_c13_pytest.log
cat gpurun_out/r2_c13_pytest.log
timeout 200 python tools/ops_prof.py 2>&1 | tail -14 | tee gpurun_out/r2_c13_ops.log
