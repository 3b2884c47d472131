#!/bin/bash
# one GPU visit: gpu test suite, bench (with stages), optionally a one-step launch list under ncu; every leg under its own timeout
tag=${1:-x}; tests=${2:-tests}; ncu_pass=${3:-0}
mkdir -p gpurun_out
timeout 900 python -m pytest $tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
MVPNET_B200_DEBUG=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
grep "tc_fused\|tc_conv" gpurun_out/bench_${tag}.err | sort | uniq -c | sort -k1,1nr | head -30
grep -v "tc_fused\|tc_conv" gpurun_out/bench_${tag}.err | tail -5
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("cpu_baseline", {}).get("value"))
print(json.dumps(d["north_star_targets"]))
for k, v in d["stages"].items():
    print("%-24s %.4f %s" % (k, v["ms"], v.get("tensor_frac_issued", "")))
PY
if [ "$ncu_pass" = "1" ]; then
  timeout 600 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
     --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 3 --no-extras > gpurun_out/ncu_${tag}.log 2>&1
  python tools/launch_table.py gpurun_out/launches_${tag}.csv | head -45
fi
