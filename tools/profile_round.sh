#!/bin/bash
# GPU visit that produces the committed evidence: bench line, one-step launch list, ncu --set full of this package's
# tensor-core kernels inside the timed step (summarised to csv on the box: the .ncu-rep is too large to bring back).
tag=${1:-p}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_${tag}.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
kill $SMI
tail -2 gpurun_out/bench_${tag}.err
timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
   --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_${tag}.log 2>&1
timeout 900 ncu --nvtx --nvtx-include "timed/" --set full --clock-control none -k regex:"tc_conv3x3|tc_fused" -c 45 \
   -o /tmp/full_${tag} python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_full_${tag}.log 2>&1
tail -2 gpurun_out/ncu_full_${tag}.log
ncu -i /tmp/full_${tag}.ncu-rep --page raw --csv > gpurun_out/full_${tag}_raw.csv 2>/dev/null
ls -la gpurun_out/ | head -20
du -sh gpurun_out
