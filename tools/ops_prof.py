"""The six extension ops + unprojection + pixel k-NN at the model's shapes (32 chunks), one call each: the target of
`ncu --set full` for per-kernel achieved DRAM GB/s (VERDICT r1 missing #6).  Also prints CUDA-event times."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import mvpnet_b200
from mvpnet_b200 import synthetic
from mvpnet_b200.data import invert_intrinsics

ext = mvpnet_b200.load_ext()
B = 32
dev = 'cuda'
chunks = [synthetic.make_chunk(seed=s) for s in range(4)]
rep = lambda a: torch.from_numpy(np.stack([a(c) for c in chunks])).repeat(B // 4, *([1] * np.stack([a(c) for c in chunks]).ndim)[1:]).to(dev)
pts = rep(lambda c: c['points'])
depth = rep(lambda c: c['depth'])
pose = rep(lambda c: c['pose'])
cam_inv = rep(lambda c: np.broadcast_to(invert_intrinsics(c['cam_matrix']), (5, 3, 3)).copy())
box = rep(lambda c: c['chunk_box'])


ONCE = os.environ.get('MVPNET_OPS_ONCE') == '1'      # one launch per op (under ncu)


def t(name, fn, n=5):
    if ONCE:
        return fn()
    fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        out = fn()
    e.record()
    torch.cuda.synchronize()
    print('%-34s %.4f ms' % (name, s.elapsed_time(e) / n))
    return out


with torch.no_grad():
    xyz32, mask, xyz64 = t('unproject 32x5x120x160', lambda: ext.unproject_cuda.unproject(depth, cam_inv, pose, box, True))
    t('knn_pixels 8192 x 96000, k=3', lambda: ext.unproject_cuda.knn_pixels(pts.double(), xyz64, mask.reshape(B, -1), 3))
    idx = t('fps 8192 -> 2048', lambda: ext.fps_cuda.farthest_point_sample(pts, 2048))
    cent = torch.gather(pts, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    nbr = t('ball_query 2048 x 8192 r=0.1 K=32', lambda: ext.ball_query_cuda.ball_query(cent, pts, 0.1, 32))
    feat = torch.randn(B, 64, 8192, device=dev)
    g = t('group_points fwd 64ch 2048x32', lambda: ext.group_points_cuda.group_points_forward(feat, nbr))
    t('group_points bwd (atomic)', lambda: ext.group_points_cuda.group_points_backward(g, nbr, 8192))
    t('group_points bwd (deterministic)', lambda: ext.group_points_cuda.group_points_backward_det(g, nbr, 8192))
    ki, kd = t('knn_distance 8192 x 2048', lambda: ext.knn_distance_cuda.knn_distance(pts, cent, 3))
    w = torch.rand(B, 8192, 3, device=dev)
    f2 = torch.randn(B, 128, 2048, device=dev)
    o = t('interpolate fwd 128ch 2048->8192', lambda: ext.interpolate_cuda.interpolate_forward(f2, ki, w))
    t('interpolate bwd (atomic)', lambda: ext.interpolate_cuda.interpolate_backward(o, ki, w, 2048))
    t('interpolate bwd (deterministic)', lambda: ext.interpolate_cuda.interpolate_backward_det(o, ki, w, 2048))
from mvpnet_b200 import train  # noqa: E402
lg = torch.randn(B, 20, 8192, device=dev, requires_grad=True)
lb = torch.randint(0, 20, (B, 8192), device=dev)
t('seg_loss fwd (loss + confusion)', lambda: train.seg_loss_and_confusion(lg, lb))
t('seg_loss fwd + bwd', lambda: train.seg_loss_and_confusion(lg, lb)[0].backward())
