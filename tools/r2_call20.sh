#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none -k regex:"tc_convg|maxpool|unfold" -c 13 -o /tmp/convg python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/r2_c20_ncu.log 2>&1
tail -2 gpurun_out/r2_c20_ncu.log
python tools/ncu_summary.py /tmp/convg.ncu-rep > gpurun_out/r2_c20_convg.md 2>&1
cat gpurun_out/r2_c20_convg.md
MVPNET_B200_DEBUG=1 timeout 300 python bench.py --steps 2 --warmup 3 --no-extras 2>&1 | grep "convg\|conv_general" | sort | uniq -c | head -20
