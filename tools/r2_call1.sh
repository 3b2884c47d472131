#!/bin/bash
# round 2, GPU call 1: hardware probe, FPS tie tests, FPS timing + ncu source view, baseline bench
mkdir -p gpurun_out
timeout 120 ./tools/probe_tma_gather.bin > gpurun_out/r2_probe.log 2>&1; echo "probe rc=$?" >> gpurun_out/r2_probe.log
cat gpurun_out/r2_probe.log
timeout 600 python -m pytest tests/test_gpu_vs_reference_kernels.py tests/test_gpu_ops.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_c1_pytest.log
cat gpurun_out/r2_c1_pytest.log
timeout 120 python tools/fps_prof.py 32 8192 2048 2>&1 | tail -1
timeout 120 python tools/fps_prof.py 32 2048 512 2>&1 | tail -1
timeout 120 python tools/fps_prof.py 1 8192 2048 2>&1 | tail -1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:fps_regs -c 1 -o /tmp/fps python tools/fps_prof.py 32 8192 2048 > gpurun_out/r2_ncu_fps.log 2>&1
timeout 120 python tools/ncu_lines.py /tmp/fps.ncu-rep 0 30 > gpurun_out/r2_fps_lines.txt 2>&1
head -60 gpurun_out/r2_fps_lines.txt
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_base.json 2> gpurun_out/bench_r2_base.err
tail -c 3000 gpurun_out/bench_r2_base.json
