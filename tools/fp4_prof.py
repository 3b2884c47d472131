"""FP4 (+ segmentation head) on the round-1 fused kernel at the bench shape, alone: target of ncu --set full --import-source on."""
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
    ki, kd = ext.knn_distance_cuda.knn_distance(pts, new, 3)
    mlp = synthetic.fill_parameters(SharedMLP(128, (128, 128, 128), ndim=1), seed=4).eval().to(dev)
    seg = synthetic.fill_parameters(SharedMLP(128, (128,), ndim=1), seed=5).eval().to(dev)
    head = synthetic.fill_parameters(torch.nn.Conv1d(128, 20, 1), seed=6).to(dev)
    layers = engine._mlp_layers(mlp) + engine._mlp_layers(seg) + [(head, None, False)]
    tcf = engine.TcChain(layers, 128, dev)
    sparse = torch.randn(B, 2048, 128, device=dev)
    for _ in range(3):
        out = ext.fused_cuda.tc_feature_propagation(sparse, ki, kd, None, 1e-10, *tcf.args())
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        ext.fused_cuda.tc_feature_propagation(sparse, ki, kd, None, 1e-10, *tcf.args())
    e.record()
    torch.cuda.synchronize()
    print('FP4 tc %.4f ms' % (s.elapsed_time(e) / 10))
