#!/bin/bash
mkdir -p gpurun_out
MVPNET_OPS_ONCE=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"kp_query" -c 1 -o /tmp/kpq python tools/ops_prof.py > /dev/null 2>&1
python tools/ncu_summary.py /tmp/kpq.ncu-rep | tail -4 | cut -c1-220
python tools/ncu_lines.py /tmp/kpq.ncu-rep 0 40 > gpurun_out/r2_kp_query_lines.txt 2>&1
head -75 gpurun_out/r2_kp_query_lines.txt | cut -c1-180
