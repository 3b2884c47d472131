"""GPU: BASELINE config 5 shapes — the whole-scene PN2SSG of configs/scannet/3d_baselines/pn2ssg_scene.yaml
(num_centroids 8192/2048/512/128, xyz-only input) on 32 768-point clouds (the reference's test shape,
test_3d_scene.py:48-49): the generic (L2-streaming) FPS, multi-tile ball query / 3-NN against the oracle, and
fused-vs-composed logits."""
import numpy as np
import pytest
import torch

import oracle
from mvpnet_b200 import synthetic

pytestmark = pytest.mark.gpu


def scene_points(n, seed):
    rng = np.random.RandomState(seed)
    pts = rng.uniform([0, 0, 0], [6.0, 8.0, 2.7], (n, 3))
    sel = rng.rand(n)
    pts[sel < 0.4, 2] = 0.0
    pts[(sel >= 0.4) & (sel < 0.6), 0] = 0.0
    pts[(sel >= 0.6) & (sel < 0.8), 1] = 8.0
    return (pts + rng.randn(n, 3) * 0.005).astype(np.float32)


def test_scene_ops_against_oracle():
    import mvpnet_b200
    ext = mvpnet_b200.load_ext()
    pts = scene_points(32768, 0)[None]
    t = torch.from_numpy(pts).cuda()
    idx = ext.fps_cuda.farthest_point_sample(t, 8192)          # N > 8192: generic kernel
    want = oracle.farthest_point_sample(pts, 8192)
    assert np.array_equal(idx.cpu().numpy(), want)
    cent = torch.gather(t, 1, idx.unsqueeze(-1).expand(-1, -1, 3)).contiguous()
    sub = cent[:, :1024].contiguous()                            # 1024 queries x 32768 keys: multi-tile search
    bq = ext.ball_query_cuda.ball_query(sub, t, 0.1, 32)
    assert np.array_equal(bq.cpu().numpy(), oracle.ball_query(sub.cpu().numpy(), pts, 0.1, 32))
    ki, kd = ext.knn_distance_cuda.knn_distance(t[:, :4096].contiguous(), cent, 3)
    oi, od = oracle.knn_distance(pts[:, :4096], cent.cpu().numpy(), 3)
    assert np.array_equal(ki.cpu().numpy(), oi) and np.array_equal(kd.cpu().numpy(), od)


def test_scene_network_fused_vs_composed():
    from mvpnet_b200.modules import PN2SSG
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = synthetic.fill_parameters(PN2SSG(0, 20, num_centroids=(8192, 2048, 512, 128)), seed=8).eval().cuda()
    pts = np.stack([scene_points(32768, s) for s in (1, 2)])
    batch = {'points': torch.from_numpy(np.ascontiguousarray(pts.transpose(0, 2, 1))).cuda()}
    with torch.no_grad():
        a = net.fast_forward(batch)['seg_logit']
        b = net(batch)['seg_logit']
    assert tuple(a.shape) == (2, 20, 32768)
    assert float((a - b).abs().max() / b.abs().max()) < 1e-4
