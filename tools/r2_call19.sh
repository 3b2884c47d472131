#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tc2.py tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python tools/stage_bench.py 2>&1 | grep "tc2\|FP" | head -20
MVPNET_B200_DEBUG=1 timeout 200 python tools/fp4_prof.py 2>&1 | tail -3
timeout 400 ncu --set full --import-source on --clock-control none -k regex:tc_fused_mlp_kernel -c 2 -o /tmp/fp4 python tools/fp4_prof.py > gpurun_out/r2_c19_ncu.log 2>&1
python tools/ncu_summary.py /tmp/fp4.ncu-rep 2>&1 | tail -6
python tools/ncu_lines.py /tmp/fp4.ncu-rep 0 40 > gpurun_out/r2_c19_fp4_lines.txt 2>&1
head -70 gpurun_out/r2_c19_fp4_lines.txt | cut -c1-170
