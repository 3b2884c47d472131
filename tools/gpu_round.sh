#!/bin/bash
# one GPU visit: full gpu test suite, bench (with stages), one-step launch list under ncu; every leg under its own timeout
tag=${1:-x}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
MVPNET_B200_DEBUG=1 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
grep tc_fused gpurun_out/bench_${tag}.err | sort | uniq -c | head -20
tail -3 gpurun_out/bench_${tag}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("cpu_baseline", {}).get("value"))
print(json.dumps(d["north_star_targets"]))
for k, v in d["stages"].items():
    print("%-24s %.4f %s" % (k, v["ms"], v.get("tensor_frac_issued", "")))
PY
