#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r2_c8_pytest.log
cat gpurun_out/r2_c8_pytest.log
timeout 300 python tools/stage_bench.py 2>&1 | tail -14 | tee gpurun_out/r2_c8_stage.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tc2_kernel -c 4 -o /tmp/tc2c python tools/stage_bench.py > gpurun_out/r2_c8_ncu.log 2>&1
python tools/ncu_summary.py /tmp/tc2c.ncu-rep > gpurun_out/r2_c8_tc2_summary.md 2>&1
cat gpurun_out/r2_c8_tc2_summary.md
python tools/ncu_lines.py /tmp/tc2c.ncu-rep 0 40 > gpurun_out/r2_c8_tc2_lines_sa1.txt 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_c8.json 2> gpurun_out/bench_r2_c8.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_c8.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["fused_mlp_family"])
for k, v in sorted(d["stages"].items(), key=lambda kv: -kv[1]["ms"]):
    print("%-24s %.4f %s" % (k, v["ms"], v.get("tensor_frac_issued", "")))
PY
