#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_ops.py tests/test_gpu_grid.py tests/test_gpu_unproject.py tests/test_gpu_golden.py tests/test_gpu_vs_reference_kernels.py -m gpu -x -q 2>&1 | tail -4
timeout 200 python tools/ops_prof.py 2>&1 | tail -12
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_c17_$i.json 2> gpurun_out/bench_r2_c17_$i.err; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_c17_$i.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["clocks"])
PY
done
