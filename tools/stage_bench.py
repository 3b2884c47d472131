"""Stage timing of the fused 3D kernels at the bench shapes (32 chunks): FA, SA1, SA2 on both kernel generations."""
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, '.')
import mvpnet_b200
from mvpnet_b200 import engine, synthetic
from mvpnet_b200.modules import SharedMLP

ext = mvpnet_b200.load_ext()
dev = 'cuda'
B = 32


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


with torch.no_grad():
    pts = torch.from_numpy(np.stack([synthetic.room_points(8192, s)[0] for s in range(B)])).to(dev)
    idx = ext.fps_cuda.farthest_point_sample(pts, 2048)
    new = torch.gather(pts, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    nbr = ext.ball_query_cuda.ball_query(new, pts, 0.1, 32)
    for name, widths, (xyz, nxyz, nb) in (('SA1', (32, 32, 64), (pts, new, nbr)),):
        mlp = synthetic.fill_parameters(SharedMLP(67, widths, ndim=2), seed=1).eval().to(dev)
        feat = torch.randn(B, xyz.size(1), 64, device=dev)
        fs = engine.split_rows(feat)
        tc = engine.TcChain(engine._mlp_layers(mlp), 67, dev)
        a = ext.fused_cuda.tc_set_abstraction(feat, xyz, nxyz, nb, *tc.args())
        b, _ = ext.fused_cuda.tc2_set_abstraction(fs, xyz, nxyz, nb, *tc.args(), True, True)
        print(name, 'rel diff tc2 vs tc: %.2e' % float((a - b).abs().max() / a.abs().max()))
        print(name, 'tc  %.4f ms' % timeit(lambda: ext.fused_cuda.tc_set_abstraction(feat, xyz, nxyz, nb, *tc.args())))
        print(name, 'tc2 %.4f ms (f32 + split out)' % timeit(lambda: ext.fused_cuda.tc2_set_abstraction(fs, xyz, nxyz, nb, *tc.args(), True, True)))
    # SA2
    idx2 = ext.fps_cuda.farthest_point_sample(new, 512)
    new2 = torch.gather(new, 1, idx2.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    nbr2 = ext.ball_query_cuda.ball_query(new2, new, 0.2, 32)
    mlp = synthetic.fill_parameters(SharedMLP(67, (64, 64, 128), ndim=2), seed=2).eval().to(dev)
    feat = torch.randn(B, 2048, 64, device=dev)
    fs = engine.split_rows(feat)
    tc = engine.TcChain(engine._mlp_layers(mlp), 67, dev)
    print('SA2 tc  %.4f ms' % timeit(lambda: ext.fused_cuda.tc_set_abstraction(feat, new, new2, nbr2, *tc.args())))
    print('SA2 tc2 %.4f ms' % timeit(lambda: ext.fused_cuda.tc2_set_abstraction(fs, new, new2, nbr2, *tc.args(), True, False)))
    # FA
    nv, h, w, hp, wp = 5, 120, 160, 128, 160
    mlp = synthetic.fill_parameters(SharedMLP(68, (64, 64, 64), ndim=2), seed=3).eval().to(dev)
    rows = torch.randn(B * nv, hp, wp, 64, device=dev)
    pix = torch.rand(B, nv * h * w, 3, device=dev)
    knn = torch.randint(0, nv * h * w, (B, 8192, 3), device=dev)
    # locality like the real 3-NN: the three pixels of a point are neighbours in one view
    base = torch.randint(0, nv * h * w - 200, (B, 8192, 1), device=dev)
    knn = torch.cat([base, base + 1, base + 160], 2)
    tc = engine.TcChain(engine._mlp_layers(mlp), 68, dev)
    f2d = rows[:, :h].permute(0, 3, 1, 2).reshape(B, nv, 64, h, w)
    rs = engine.split_rows(rows)
    a = ext.fused_cuda.tc_feature_aggregation(f2d, pix, pts, knn, True, *tc.args())
    b, _ = ext.fused_cuda.tc2_feature_aggregation(rs, nv, h, w, pix, pts, knn, True, *tc.args(), True, True)
    print('FA rel diff tc2 vs tc: %.2e' % float((a - b).abs().max() / a.abs().max()))
    print('FA tc  %.4f ms' % timeit(lambda: ext.fused_cuda.tc_feature_aggregation(f2d, pix, pts, knn, True, *tc.args())))
    print('FA tc2 %.4f ms' % timeit(lambda: ext.fused_cuda.tc2_feature_aggregation(rs, nv, h, w, pix, pts, knn, True, *tc.args(), False, True)))
    print('fps1 %.4f ms' % timeit(lambda: ext.fps_cuda.farthest_point_sample(pts, 2048), 5))
    print('fps2 %.4f ms' % timeit(lambda: ext.fps_cuda.farthest_point_sample(new, 512), 5))
    # FP4 (+ seg head) and FP3 on the round-1 kernel (weight ring shared by the tile groups)
    ki, kd = ext.knn_distance_cuda.knn_distance(pts, new, 3)
    mlp = synthetic.fill_parameters(SharedMLP(128, (128, 128, 128), ndim=1), seed=4).eval().to(dev)
    seg = synthetic.fill_parameters(SharedMLP(128, (128,), ndim=1), seed=5).eval().to(dev)
    head = synthetic.fill_parameters(torch.nn.Conv1d(128, 20, 1), seed=6).to(dev)
    layers = engine._mlp_layers(mlp) + engine._mlp_layers(seg) + [(head, None, False)]
    tcf = engine.TcChain(layers, 128, dev)
    sparse = torch.randn(B, 2048, 128, device=dev)
    print('FP4 tc  %.4f ms' % timeit(lambda: ext.fused_cuda.tc_feature_propagation(sparse, ki, kd, None, 1e-10, *tcf.args())))
    ki3, kd3 = ext.knn_distance_cuda.knn_distance(new, new2, 3)
    mlp3 = synthetic.fill_parameters(SharedMLP(320, (256, 128), ndim=1), seed=7).eval().to(dev)
    tc3 = engine.TcChain(engine._mlp_layers(mlp3), 320, dev)
    sp3 = torch.randn(B, 512, 256, device=dev)
    sk3 = torch.randn(B, 2048, 64, device=dev)
    print('FP3 tc  %.4f ms' % timeit(lambda: ext.fused_cuda.tc_feature_propagation(sp3, ki3, kd3, sk3, 1e-10, *tc3.args())))
