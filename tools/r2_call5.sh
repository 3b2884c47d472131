#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_c5_pytest.log
cat gpurun_out/r2_c5_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_c5.json 2> gpurun_out/bench_r2_c5.err
tail -3 gpurun_out/bench_r2_c5.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_c5.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k, v in sorted(d["stages"].items(), key=lambda kv: -kv[1]["ms"]):
    print("%-24s %.4f %s" % (k, v["ms"], v.get("tensor_frac_issued", "")))
PY
timeout 600 ncu --set full --import-source on --clock-control none -k regex:tc2_kernel -c 6 -o /tmp/tc2 python tools/stage_bench.py > gpurun_out/r2_c5_ncu.log 2>&1
tail -3 gpurun_out/r2_c5_ncu.log
python tools/ncu_summary.py /tmp/tc2.ncu-rep > gpurun_out/r2_c5_tc2_summary.md 2>&1
cat gpurun_out/r2_c5_tc2_summary.md
for k in 0 1 2 3 4 5; do python tools/ncu_lines.py /tmp/tc2.ncu-rep $k 25 > gpurun_out/r2_c5_tc2_lines_$k.txt 2>&1; done
cp /tmp/tc2.ncu-rep gpurun_out/r2_tc2.ncu-rep
