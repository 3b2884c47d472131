#!/bin/bash
# Round-2 evidence run (one B200): bench line + clocks, launch list of one step, ncu --set full of the tensor-core kernels of
# one step (raw page exported on the box), ncu --set full of the six extension ops / data-side kernels at model shapes.
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_${tag}.csv &
SMI=$!
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
kill $SMI
tail -2 gpurun_out/bench_${tag}.err
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
   --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_${tag}.log 2>&1
python tools/launch_table.py gpurun_out/launches_${tag}.csv > gpurun_out/launches_${tag}.md 2>&1
head -40 gpurun_out/launches_${tag}.md
timeout 900 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none -k regex:"tc_conv3x3|tc_fused|tc2_kernel|tc_convg" -c 60 \
   -o /tmp/full_${tag} python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_full_${tag}.log 2>&1
tail -2 gpurun_out/ncu_full_${tag}.log
ncu -i /tmp/full_${tag}.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/full_${tag}_raw.csv > gpurun_out/ncu_full_${tag}.md 2>&1
python tools/ncu_traffic.py gpurun_out/full_${tag}_raw.csv > gpurun_out/ncu_traffic_${tag}.json 2>&1
cat gpurun_out/ncu_traffic_${tag}.json
MVPNET_OPS_ONCE=1 timeout 600 ncu --set full --clock-control none -k regex:"unproject_kernel|kp_query|fps_regs|pg_ball_query|group_points|pg_knn3|interpolate|dl_gather|dl_fill|dl_count|transpose_kernel|seg_" -c 40 \
   -o /tmp/ops_${tag} python tools/ops_prof.py > gpurun_out/ncu_ops_${tag}.log 2>&1
ncu -i /tmp/ops_${tag}.ncu-rep --page raw --csv > gpurun_out/ops_${tag}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/ops_${tag}_raw.csv > gpurun_out/ncu_ops_${tag}.md 2>&1
cat gpurun_out/ncu_ops_${tag}.md
du -sh gpurun_out
