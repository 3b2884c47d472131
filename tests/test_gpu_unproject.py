"""GPU parity for the data side of FeatureAggregation: depth unprojection and the 2D->3D k-NN
against the CPU oracle (bit-exact: fp64 arithmetic with a fixed operation order, integer ids)."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def make_views(b, nv, h, w, seed):
    rng = np.random.RandomState(seed)
    depth = rng.uniform(0.5, 4.0, (b, nv, h, w)).astype(np.float32)
    depth[rng.rand(b, nv, h, w) < 0.1] = 0.0
    cam = np.array([[577.87, 0, 319.5], [0, 577.87, 239.5], [0, 0, 1]], np.float32)
    cam[0] /= 4
    cam[1] /= 4
    cam_inv = np.broadcast_to(np.linalg.inv(cam), (b, nv, 3, 3)).copy()
    pose = np.zeros((b, nv, 4, 4), np.float32)
    for i in range(b):
        for f in range(nv):
            q, _ = np.linalg.qr(rng.randn(3, 3))
            if np.linalg.det(q) < 0:
                q[:, 0] = -q[:, 0]
            pose[i, f, :3, :3] = q
            pose[i, f, :3, 3] = rng.uniform(-1, 1, 3)
            pose[i, f, 3, 3] = 1
    return depth, cam_inv, pose


@pytest.mark.parametrize('b,nv,h,w,box', [(1, 1, 120, 160, False), (2, 5, 120, 160, True), (1, 3, 17, 23, True)])
def test_unproject(b, nv, h, w, box):
    import mvpnet_b200
    ext = mvpnet_b200.load_ext()
    depth, cam_inv, pose = make_views(b, nv, h, w, 0)
    boxes = np.array([[-0.5, -0.5, 1.0, 1.0]] * b, np.float64) if box else None
    t = lambda a: torch.from_numpy(a).cuda()
    xyz32, mask, xyz64 = ext.unproject_cuda.unproject(t(depth), t(cam_inv), t(pose), t(boxes) if box else None, True)
    for i in range(b):
        w64, w32, wm = oracle.unproject(depth[i], cam_inv[i], pose[i], boxes[i] if box else None)
        assert np.array_equal(xyz64[i].cpu().numpy().reshape(nv, h, w, 3), w64)
        assert np.array_equal(xyz32[i].cpu().numpy(), w32)
        assert np.array_equal(mask[i].cpu().numpy().astype(bool), wm)


@pytest.mark.parametrize('nq,nv,k', [(2048, 1, 3), (1024, 5, 3), (257, 2, 1), (300, 2, 8)])
def test_knn_pixels(nq, nv, k):
    import mvpnet_b200
    ext = mvpnet_b200.load_ext()
    b, h, w = 2, 120, 160
    depth, cam_inv, pose = make_views(b, nv, h, w, 1)
    rng = np.random.RandomState(2)
    qs, idxs, d2s = [], [], []
    pix, msk = [], []
    for i in range(b):
        x64, _, m = oracle.unproject(depth[i], cam_inv[i], pose[i])
        flat = x64.reshape(-1, 3)
        valid = np.nonzero(m.reshape(-1))[0]
        q = flat[rng.choice(valid, nq)] + rng.randn(nq, 3) * 0.02
        q = q.astype(np.float32).astype(np.float64)   # chunk points are float32 in the reference
        ii, dd = oracle.knn_pixels(q, flat, m, k)
        qs.append(q); idxs.append(ii); d2s.append(dd); pix.append(flat); msk.append(m.reshape(-1))
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    for exhaustive in (False, True):       # uniform-grid search and exhaustive search: identical results
        gi, gd = ext.unproject_cuda.knn_pixels(t(np.stack(qs)), t(np.stack(pix)), t(np.stack(msk).astype(np.uint8)), k, exhaustive)
        assert np.array_equal(gi.cpu().numpy(), np.stack(idxs))
        assert np.array_equal(gd.cpu().numpy(), np.stack(d2s))


def test_knn_pixels_grid_edge_cases():
    """Far / outside-the-box queries, coplanar pixels, exact ties, a single occupied cell."""
    import mvpnet_b200
    ext = mvpnet_b200.load_ext()
    rng = np.random.RandomState(5)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    cases = []
    # (a) pixels on a plane (zero z extent), lattice => many exact distance ties; queries near and very far
    gx, gy = np.meshgrid(np.arange(60) * 0.05, np.arange(40) * 0.05)
    plane = np.stack([gx.ravel(), gy.ravel(), np.zeros(gx.size)], 1)
    q = np.concatenate([plane[rng.choice(len(plane), 200)] + [0.025, 0.025, 0.0], rng.uniform(-30, 30, (56, 3))])
    cases.append((plane, np.ones(len(plane), np.uint8), q))
    # (b) all valid pixels identical (one cell), plus masked-out decoys closer to the queries
    same = np.tile([[1.0, 2.0, 3.0]], (500, 1))
    decoy = rng.uniform(-1, 1, (500, 3))
    cases.append((np.concatenate([decoy, same]), np.concatenate([np.zeros(500, np.uint8), np.ones(500, np.uint8)]), rng.uniform(-1, 1, (64, 3))))
    # (c) two distant clusters: queries between them must walk many empty shells
    cl = np.concatenate([rng.randn(3000, 3) * 0.05, rng.randn(3000, 3) * 0.05 + [8.0, 0.0, 0.0]])
    cases.append((cl, np.ones(6000, np.uint8), np.stack([np.linspace(-1, 9, 128), np.zeros(128), np.zeros(128)], 1)))
    for pix, m, q in cases:
        for k in (1, 3, 5):
            want_i, want_d = oracle.knn_pixels(q, pix, m, k)
            gi, gd = ext.unproject_cuda.knn_pixels(t(q)[None], t(pix)[None], t(m)[None], k, False)
            assert np.array_equal(gi[0].cpu().numpy(), want_i)
            assert np.array_equal(gd[0].cpu().numpy(), want_d)


def test_knn_pixels_too_few_valid():
    import mvpnet_b200
    ext = mvpnet_b200.load_ext()
    pix = torch.zeros(1, 10, 3, dtype=torch.float64).cuda()
    mask = torch.zeros(1, 10, dtype=torch.uint8).cuda()
    mask[0, 4] = 1
    for exhaustive in (False, True):
        gi, _ = ext.unproject_cuda.knn_pixels(torch.ones(1, 2, 3, dtype=torch.float64).cuda(), pix, mask, 3, exhaustive)
        assert gi.cpu().numpy().tolist() == [[[4, -1, -1], [4, -1, -1]]]


def test_decode_stored_inputs_bit_exact():
    """uint8 HWC colour / uint16 mm depth -> network inputs on the device == the host conversions of
    scannet_2d3d.py:229-251 (float32, same operation order)."""
    import mvpnet_b200
    from mvpnet_b200 import engine
    rng = np.random.RandomState(0)
    rgb = rng.randint(0, 256, (2, 3, 24, 40, 3)).astype(np.uint8)
    mm = rng.randint(0, 65536, (2, 3, 24, 40)).astype(np.uint16)
    out = engine.decode_stored_inputs({'images_u8': torch.from_numpy(rgb).cuda(), 'depth_mm': torch.from_numpy(mm.view(np.int16)).cuda()})
    want = (rgb.astype(np.float32) / np.float32(255.0) - np.asarray(engine.IMAGE_MEAN, np.float32)) / np.asarray(engine.IMAGE_STD, np.float32)
    assert out['images'].shape == (2, 3, 3, 24, 40)
    assert np.array_equal(out['images'].cpu().numpy(), want.transpose(0, 1, 4, 2, 3))
    assert np.array_equal(out['depth'].cpu().numpy(), mm.astype(np.float32) / np.float32(1000.0))
