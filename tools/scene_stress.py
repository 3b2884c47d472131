"""BASELINE config 5: whole-scene PN2SSG (pn2ssg_scene.yaml shape) on ~200k synthetic points, 1 GPU:
time per forward and peak memory, fused path.  Run on the GPU box: python tools/scene_stress.py [N]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvpnet_b200 import synthetic, engine
from mvpnet_b200.modules import PN2SSG
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
rng = np.random.RandomState(0)
pts = rng.uniform([0, 0, 0], [6.0, 8.0, 2.7], (n, 3)); sel = rng.rand(n)
pts[sel < 0.4, 2] = 0.0; pts[(sel >= 0.4) & (sel < 0.6), 0] = 0.0; pts[(sel >= 0.6) & (sel < 0.8), 1] = 8.0
pts = (pts + rng.randn(n, 3) * 0.005).astype(np.float32)
net = synthetic.fill_parameters(PN2SSG(0, 20, num_centroids=(8192, 2048, 512, 128)), seed=8).eval().cuda()
batch = {'points': torch.from_numpy(np.ascontiguousarray(pts.T))[None].cuda()}
torch.cuda.reset_peak_memory_stats()
with torch.no_grad():
    net.fast_forward(batch); torch.cuda.synchronize()
    with engine.profile() as prof:
        t0 = time.perf_counter(); out = net.fast_forward(batch)['seg_logit']; torch.cuda.synchronize(); dt = time.perf_counter() - t0
    st = {k: round(float(np.median(v)), 3) for k, v in prof.summary().items()}
print({'points': n, 'forward_ms': round(dt * 1e3, 2), 'peak_mem_MB': round(torch.cuda.max_memory_allocated() / 2**20, 1), 'stages_ms': st})
