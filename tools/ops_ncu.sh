#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/ops_prof.py 2>&1 | tail -14 | tee gpurun_out/r2_ops_times.log
MVPNET_OPS_ONCE=1 timeout 600 ncu --set full --clock-control none -k regex:"unproject_kernel|kp_query|fps_regs|pg_ball_query|group_points|pg_knn3|interpolate|dl_gather|dl_fill|dl_count|transpose_kernel|seg_" -c 40 \
   -o /tmp/ops_r2 python tools/ops_prof.py > gpurun_out/ncu_ops_r2.log 2>&1
ncu -i /tmp/ops_r2.ncu-rep --page raw --csv > gpurun_out/ops_r2_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/ops_r2_raw.csv > gpurun_out/ncu_ops_r2.md 2>&1
cat gpurun_out/ncu_ops_r2.md | cut -c1-230
MVPNET_OPS_ONCE=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:"pg_knn3" -c 1 -o /tmp/knn3 python tools/ops_prof.py > /dev/null 2>&1
python tools/ncu_lines.py /tmp/knn3.ncu-rep 0 30 > gpurun_out/r2_knn3_lines.txt 2>&1
head -50 gpurun_out/r2_knn3_lines.txt | cut -c1-170
