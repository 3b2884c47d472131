#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 200 python tools/stage_bench.py 2>&1 | grep "tc2\|FP" | head
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_c26.json 2> gpurun_out/bench_r2_c26.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_c26.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["fused_mlp_family"])
print(d["north_star_targets"], d["cpu_baseline"]["value"])
for k, v in sorted(d["stages"].items(), key=lambda kv: -kv[1]["ms"])[:16]:
    print("%-24s %.4f %s" % (k, v["ms"], v.get("tensor_frac_issued", "")))
PY
