#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_tc2.py tests/test_gpu_train.py tests/test_gpu_golden.py -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -30 > gpurun_out/r2_c7_pytest.log
cat gpurun_out/r2_c7_pytest.log
MVPNET_B200_DEBUG=1 timeout 300 python tools/stage_bench.py 2>&1 | grep -v "^\[tc" | tail -20 | tee gpurun_out/r2_c7_stage.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tc2_kernel -c 2 -o /tmp/tc2b python tools/stage_bench.py > gpurun_out/r2_c7_ncu.log 2>&1
python tools/ncu_summary.py /tmp/tc2b.ncu-rep > gpurun_out/r2_c7_tc2_summary.md 2>&1
cat gpurun_out/r2_c7_tc2_summary.md
python tools/ncu_lines.py /tmp/tc2b.ncu-rep 0 45 > gpurun_out/r2_c7_tc2_lines_sa1.txt 2>&1
