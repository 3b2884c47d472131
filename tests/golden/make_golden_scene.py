"""Golden fixture for the scene-level callers (mvpnet_b200/scene.py) from the REFERENCE's own numpy code:
mvpnet/utils/chunk_util.py (imported unmodified from /root/reference) and the vote accumulation of
test_mvpnet_3d.py:136-175 (restated line by line here; the script itself needs open3d / yacs to import).

    python tests/golden/make_golden_scene.py   ->  tests/golden/scene_chunks.npz
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from mvpnet_b200 import synthetic  # noqa: E402

spec = importlib.util.spec_from_file_location('chunk_util', '/root/reference/mvpnet/utils/chunk_util.py')
chunk_util = importlib.util.module_from_spec(spec)
spec.loader.exec_module(chunk_util)


def scene_points(seed=11, n=60000):
    """A 6 x 8 x 2.7 m room tiled from the synthetic surface sampler."""
    rng = np.random.RandomState(seed)
    tiles = []
    for i in range(3):
        for j in range(4):
            p, _ = synthetic.room_points(n // 12, seed=seed * 100 + i * 4 + j)
            tiles.append(p + np.array([i * 1.9, j * 1.9, 0.0], np.float32))
    # an isolated sparse cluster: below `thresh` in every chunk, so these points end without a prediction
    far = (rng.rand(300, 3) * 0.4 + np.array([9.0, 11.0, 0.5])).astype(np.float32)
    pts = np.concatenate(tiles + [far]).astype(np.float32)
    return pts[rng.permutation(len(pts))]


def reference_select_frames():
    """mvpnet/data/scannet_2d3d.py:20-30, taken from the source file by AST (the module itself imports open3d)."""
    import ast
    src = open('/root/reference/mvpnet/data/scannet_2d3d.py').read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'select_frames'][0]
    ns = {'np': np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'scannet_2d3d.py', 'exec'), ns)
    return ns['select_frames']


def overlap_matrix(seed=3, num_points=5000, num_frames=40):
    """Synthetic visibility: every frame sees the points inside a random cone-ish window; several exact ties."""
    rng = np.random.RandomState(seed)
    centre = rng.rand(num_frames, 1) * num_points
    width = rng.randint(200, 1500, size=(num_frames, 1))
    idx = np.arange(num_points)[None, :]
    m = (np.abs(idx - centre) < width) & (rng.rand(num_frames, num_points) < 0.8)
    m[7] = m[3]                                   # duplicated frames: equal coverage, the lower index must win
    m[21] = m[20]
    return np.ascontiguousarray(m.T)


def boundary_points(seed=21, n=20000):
    """Extent EXACTLY on a window boundary: x spans 2.5 m, y 3.0 m with chunk 1.5 / stride 0.5, so that
    (limit - chunk) / stride is an integer in exact arithmetic and the float32 / float64 subtraction of the reference
    (chunk_util.py:24-28) decides the window count; many points sit exactly on window edges (multiples of 0.5)."""
    rng = np.random.RandomState(seed)
    p = rng.rand(n, 3).astype(np.float32) * np.array([2.5, 3.0, 2.0], np.float32) + np.array([0.3, -1.7, 0.0], np.float32)
    p[:2000, :2] = (np.round(rng.rand(2000, 2) * np.array([5, 6])) * 0.5 + np.array([0.3, -1.7])).astype(np.float32)
    p[0, :2] = (0.3, -1.7)
    p[1, :2] = (np.float32(0.3) + np.float32(2.5), np.float32(-1.7) + np.float32(3.0))
    return p.astype(np.float32)


def mvpnet2d_golden():
    """Reference MVPNet2D (mvpnet/models/mvpnet_2d.py:7-34, imported unmodified) with the reference UNetResNet34 on the
    seed-0 synthetic chunk and the k-NN indices of rgbd_chunk.npz; group_points served by the oracle."""
    import types
    import torch
    import oracle
    from mvpnet_b200 import compat
    compat.install(modules=oracle.ext_modules(), reference_root='/root/reference')
    for missing in ('open3d', 'natsort'):
        sys.modules.setdefault(missing, types.ModuleType(missing))
    from mvpnet.models.mvpnet_2d import MVPNet2D
    from mvpnet.models.unet_resnet34 import UNetResNet34
    torch.set_grad_enabled(False)
    chunk = synthetic.make_chunk(seed=0)
    knn = np.load(os.path.join(HERE, 'rgbd_chunk.npz'))['knn_indices'].astype(np.int64)
    model = MVPNet2D(UNetResNet34(20, p=0.5, pretrained=False))
    synthetic.fill_parameters(model, seed=8).eval()
    out = model({'images': torch.from_numpy(chunk['images'])[None], 'knn_indices': torch.from_numpy(knn)[None]})['seg_logit']
    np.savez_compressed(os.path.join(HERE, 'mvpnet2d.npz'), logit_sample=out[0, :, ::4].numpy(), logit_absmax=np.float64(out.abs().max()),
                        logit_checksum=np.float64(out.double().sum()))
    print('mvpnet2d', tuple(out.shape), float(out.abs().max()))


def nearest_golden():
    """test_3d_scene.py:155-163: 1-NN label propagation with scikit-learn's ball tree, two votes of 8192 samples."""
    from sklearn.neighbors import NearestNeighbors
    pts = scene_points()[:40000]
    rng = np.random.RandomState(9)
    ind = np.stack([rng.choice(len(pts), size=8192, replace=False) for _ in range(2)])
    nn = []
    for v in range(2):
        nbrs = NearestNeighbors(n_neighbors=1, algorithm='ball_tree').fit(pts[ind[v]])
        nn.append(nbrs.kneighbors(pts[:, 0:3])[1][:, 0])
    np.savez_compressed(os.path.join(HERE, 'nearest_1nn.npz'), vote_indices=ind.astype(np.int32), nn_indices=np.stack(nn).astype(np.int32))
    print('nearest', np.stack(nn).shape)


def main():
    bp = boundary_points()
    bidx = chunk_util.scene2chunks_legacy(bp, chunk_size=(1.5, 1.5), stride=0.5, thresh=200, margin=(0.2, 0.2))
    np.savez_compressed(os.path.join(HERE, 'scene_boundary.npz'), points_checksum=np.float64(bp.astype(np.float64).sum()),
                        numpy_version=np.__version__, chunk_sizes=np.array([len(i) for i in bidx]),
                        chunk_index_checksums=np.array([int(i.astype(np.int64).sum()) for i in bidx]))
    print('boundary chunks', len(bidx), 'numpy', np.__version__)
    mvpnet2d_golden()
    nearest_golden()
    pts = scene_points()
    idx, bbox = chunk_util.scene2chunks_legacy(pts, chunk_size=(1.5, 1.5), stride=0.5, thresh=1000, margin=(0.2, 0.2), return_bbox=True)
    # vote accumulation exactly as test_mvpnet_3d.py:136-175, with seeded stand-in logits per chunk
    num_classes = 20
    rng = np.random.RandomState(5)
    logit_sum = np.zeros([len(pts), num_classes], dtype=np.float32)
    count = np.zeros(len(pts), dtype=np.uint8)
    seeds = []
    for c, ind in enumerate(idx):
        seeds.append(1000 + c)
        seg_logit = np.random.RandomState(1000 + c).randn(num_classes, len(ind) + 7).astype(np.float32)   # 7 padded columns
        seg_logit = seg_logit.T[:len(ind)]
        logit_sum[ind] += seg_logit
        count[ind] += 1
    mean = logit_sum / np.maximum(count[:, np.newaxis], 1)
    label = np.argmax(mean, axis=1)
    label[count == 0] = num_classes
    np.savez_compressed(os.path.join(HERE, 'scene_chunks.npz'), points_checksum=np.float64(pts.astype(np.float64).sum()),
                        num_chunks=len(idx), chunk_sizes=np.array([len(i) for i in idx]),
                        chunk_index_checksums=np.array([int(i.astype(np.int64).sum()) for i in idx]),
                        first_chunk=idx[0], bboxes=np.stack(bbox), count=count, label=label.astype(np.int64),
                        mean_sample=mean[::97], mean_checksum=np.float64(mean.astype(np.float64).sum()))
    sel = {str(k): np.array(reference_select_frames()(overlap_matrix(), k)) for k in (1, 3, 5, 12)}
    np.savez_compressed(os.path.join(HERE, 'select_frames.npz'), **sel)
    print('chunks', len(idx), 'points without prediction', int((count == 0).sum()), 'frames', sel['5'])


if __name__ == '__main__':
    main()
