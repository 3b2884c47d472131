#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_c10_pytest.log
cat gpurun_out/r2_c10_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_c10.json 2> gpurun_out/bench_r2_c10.err
tail -5 gpurun_out/bench_r2_c10.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_c10.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"], d["gpu_launches"], d.get("gpu_launches_note"))
print(json.dumps(d.get("sub_lines"), indent=1)[:6000])
PY
