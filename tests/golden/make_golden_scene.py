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


def main():
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
