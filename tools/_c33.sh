#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/clocks_r2b.csv &
SMI=$!
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err
kill $SMI
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2b.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["gpu_launches"], d["roofline"]["frac"], d["roofline"]["traffic"], d["fused_mlp_family"]["ms_per_step"], d["fused_mlp_family"]["tensor_frac_issued"])
print(d["north_star_targets"], d["cpu_baseline"]["value"], d["clocks"])
s = d["sub_lines"]; print(s["b1_latency"], s["config2_pn2ssg"]["forward_eval_fused_b1_ms"], s["config2_pn2ssg"]["forward_eval_fused_b32_ms"], s["config2_pn2ssg"]["train_step_fwd_bwd_b1"]["ms"], s["config2_pn2ssg"]["train_step_fwd_bwd_b32"]["ms"], s["config5_whole_scene_pn2ssg"]["200k_points_b1"]["forward_ms"], s["config5_whole_scene_pn2ssg"]["reference_test_shape_b3_x_32768"]["forward_ms"], s["scene_pipeline"])
for k, v in sorted(d["stages"].items(), key=lambda kv: -kv[1]["ms"]):
    print("%-24s %.4f %s %s" % (k, v["ms"], v.get("tensor_frac_issued", ""), v.get("hbm_frac", "")))
PY
