"""Phase clocks of the tc2 kernels (MVPNET_B200_TC2_PROF=1): where a tile group's time goes."""
import os
os.environ['MVPNET_B200_TC2_PROF'] = '1'
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import mvpnet_b200
from mvpnet_b200 import engine, synthetic
from mvpnet_b200.modules import SharedMLP
ext = mvpnet_b200.load_ext()
dev, B = 'cuda', 32
with torch.no_grad():
    pts = torch.from_numpy(np.stack([synthetic.room_points(8192, s)[0] for s in range(B)])).to(dev)
    idx = ext.fps_cuda.farthest_point_sample(pts, 2048)
    new = torch.gather(pts, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    nbr = ext.ball_query_cuda.ball_query(new, pts, 0.1, 32)
    mlp = synthetic.fill_parameters(SharedMLP(67, (32, 32, 64), ndim=2), seed=1).eval().to(dev)
    fs = engine.split_rows(torch.randn(B, 8192, 64, device=dev))
    tc = engine.TcChain(engine._mlp_layers(mlp), 67, dev)
    for _ in range(2):
        ext.fused_cuda.tc2_set_abstraction(fs, pts, new, nbr, *tc.args(), True, True)
    ext.fused_cuda.tc2_prof_dump('warmup (discard)')
    ext.fused_cuda.tc2_set_abstraction(fs, pts, new, nbr, *tc.args(), True, True)
    ext.fused_cuda.tc2_prof_dump('SA1, one launch, %d tiles per CTA' % (B * 2048 // 4 // 148))
    nv, h, w, hp, wp = 5, 120, 160, 128, 160
    mlp = synthetic.fill_parameters(SharedMLP(68, (64, 64, 64), ndim=2), seed=3).eval().to(dev)
    rs = engine.split_rows(torch.randn(B * nv, hp, wp, 64, device=dev))
    pix = torch.rand(B, nv * h * w, 3, device=dev)
    base = torch.randint(0, nv * h * w - 200, (B, 8192, 1), device=dev)
    knn = torch.cat([base, base + 1, base + 160], 2)
    tcf = engine.TcChain(engine._mlp_layers(mlp), 68, dev)
    ext.fused_cuda.tc2_feature_aggregation(rs, nv, h, w, pix, pts, knn, True, *tcf.args(), False, True)
    ext.fused_cuda.tc2_prof_dump('FA warmup (discard)')
    ext.fused_cuda.tc2_feature_aggregation(rs, nv, h, w, pix, pts, knn, True, *tcf.args(), False, True)
    ext.fused_cuda.tc2_prof_dump('FA, one launch, %d units per CTA' % (B * 8192 // 128 * 3 // 148))
