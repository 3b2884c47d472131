#!/bin/bash
# tc kernel parity + bench, each under its own timeout so a protocol bug cannot hang the box
tag=${1:-x}
timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_golden.py tests/test_gpu_scene.py -x -q 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log
tail -15 gpurun_out/${tag}_pytest.log
MVPNET_B200_DEBUG=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
grep tc_fused gpurun_out/bench_${tag}.err | sort | uniq -c | head -20
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"])
for k, v in d["stages"].items():
    print("%-24s %.4f %s" % (k, v["ms"], v.get("tensor_frac_issued", "")))
PY
