#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_golden.py -m gpu -x -q 2>&1 | tail -3
for i in 1 2; do timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_c24_$i.json 2> gpurun_out/bench_r2_c24_$i.err; tail -2 gpurun_out/bench_r2_c24_$i.err; python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_c24_$i.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["mode"][:60])
print({k: v["ms"] for k, v in d["stages"].items() if k in ("knn_pixels", "fps1", "ball_query1", "feature_propagation4", "set_abstraction1", "feature_aggregation", "decode_inputs")})
PY
done
