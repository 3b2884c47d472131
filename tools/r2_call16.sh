#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "fps" 2>&1 | tail -3
timeout 100 python tools/fps_prof.py 3 32768 8192 2>&1 | tail -1
timeout 100 python tools/fps_prof.py 1 200000 8192 2>&1 | tail -1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2_c16.json 2> gpurun_out/bench_r2_c16.err
tail -3 gpurun_out/bench_r2_c16.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_c16.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["fused_mlp_family"]["ms_per_step"], d["north_star_targets"])
for k, v in sorted(d["stages"].items(), key=lambda kv: -kv[1]["ms"])[:14]:
    print("%-24s %.4f %s" % (k, v["ms"], v.get("tensor_frac_issued", "")))
s = d["sub_lines"]
print(s["b1_latency"], s["config5_whole_scene_pn2ssg"]["200k_points_b1"]["forward_ms"], s["config5_whole_scene_pn2ssg"]["reference_test_shape_b3_x_32768"]["forward_ms"], s["scene_pipeline"]["scenes_per_s"])
PY
