"""Seeded synthetic inputs of the shapes BASELINE.json names (there is no ScanNet offline).

A chunk is a 1.9 x 1.9 x 2.5 m room (1.5 m chunk + 2 x 0.2 m margin, mvpnet/config/mvpnet_3d.py:20-22)
with a few axis-aligned boxes standing on the floor.  Scene points are sampled on its surfaces; RGB-D
views are ray-cast from random interior camera poses with ScanNet depth intrinsics scaled to
160 x 120 (scannet_2d3d.py:206-210), depth quantised to millimetres like the uint16 ScanNet PNGs and
~10 % of the pixels dropped to 0 (invalid).  Everything is numpy, deterministic in `seed`.
"""
import numpy as np

ROOM = np.array([1.9, 1.9, 2.5])
SCANNET_DEPTH_INTRINSICS = np.array([[577.870605, 0.0, 319.5, 0.0], [0.0, 577.870605, 239.5, 0.0],
                                     [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]], np.float32)


def _boxes(rng, n=3):
    out = []
    for _ in range(n):
        size = rng.uniform([0.3, 0.3, 0.3], [0.7, 0.7, 1.2])
        lo = np.array([rng.uniform(0.05, ROOM[0] - size[0] - 0.05), rng.uniform(0.05, ROOM[1] - size[1] - 0.05), 0.0])
        out.append((lo, lo + size))
    return out


def _planes(boxes):
    """Axis-aligned rectangles (axis, offset, lo2, hi2) of the room shell and the boxes."""
    rects = []
    lo, hi = np.zeros(3), ROOM
    for ax in range(3):
        o = [a for a in range(3) if a != ax]
        for off in (lo[ax], hi[ax]):
            rects.append((ax, off, lo[o], hi[o]))
    for blo, bhi in boxes:
        for ax in range(3):
            o = [a for a in range(3) if a != ax]
            for off in (blo[ax], bhi[ax]):
                if ax == 2 and off == 0.0:
                    continue
                rects.append((ax, off, blo[o], bhi[o]))
    return rects


def room_points(n, seed=0, jitter=0.005):
    """(n, 3) float32 points on floor (35 %), two walls (40 %), boxes (25 %) + normal jitter."""
    rng = np.random.RandomState(seed)
    boxes = _boxes(rng)
    pts = np.empty((n, 3))
    sel = rng.rand(n)
    for i in range(n):
        s = sel[i]
        if s < 0.35:
            p = [rng.uniform(0, ROOM[0]), rng.uniform(0, ROOM[1]), 0.0]
        elif s < 0.55:
            p = [0.0, rng.uniform(0, ROOM[1]), rng.uniform(0, ROOM[2])]
        elif s < 0.75:
            p = [rng.uniform(0, ROOM[0]), ROOM[1], rng.uniform(0, ROOM[2])]
        else:
            blo, bhi = boxes[rng.randint(len(boxes))]
            p = rng.uniform(blo, bhi)
            face = rng.randint(5)
            if face == 4:
                p[2] = bhi[2]
            else:
                p[face % 2] = (blo, bhi)[face // 2][face % 2]
        pts[i] = p
    pts += rng.randn(n, 3) * jitter
    return pts.astype(np.float32), boxes


def _look_at(eye, target):
    fwd = target - eye
    fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, np.array([0.0, 0.0, 1.0]))
    if np.linalg.norm(right) < 1e-6:
        right = np.array([1.0, 0.0, 0.0])
    right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    pose = np.eye(4)
    pose[:3, 0], pose[:3, 1], pose[:3, 2], pose[:3, 3] = right, down, fwd, eye   # camera x right, y down, z forward
    return pose.astype(np.float32)


def render_depth(pose, cam, boxes, h=120, w=160):
    """Ray-cast the room: z-depth (metres, float32, mm-quantised) of the nearest surface per pixel."""
    v, u = np.indices((h, w))
    inv = np.linalg.inv(cam[:3, :3].astype(np.float64))
    rays = (inv @ np.stack([u.ravel(), v.ravel(), np.ones(h * w)])).T          # camera frame, z == 1
    R, t = pose[:3, :3].astype(np.float64), pose[:3, 3].astype(np.float64)
    dirs = rays @ R.T
    best = np.full(h * w, np.inf)
    for ax, off, lo2, hi2 in _planes(boxes):
        o = [a for a in range(3) if a != ax]
        with np.errstate(divide='ignore', invalid='ignore'):
            s = (off - t[ax]) / dirs[:, ax]
        hit = t[None, o] + s[:, None] * dirs[:, o]
        ok = (s > 1e-4) & np.all(hit >= lo2 - 1e-9, axis=1) & np.all(hit <= hi2 + 1e-9, axis=1)
        best = np.where(ok & (s < best), s, best)
    depth = np.where(np.isfinite(best), best, 0.0)        # ray parameter == z-depth because rays have z == 1
    depth_mm = np.clip(np.round(depth * 1000.0), 0, 65535).astype(np.uint16)
    return depth_mm.reshape(h, w)


def _visible(points, pose, cam, depth_mm, tol=0.05):
    """Which points project into the view and agree with its depth map within `tol` metres."""
    h, w = depth_mm.shape
    R, t = pose[:3, :3].astype(np.float64), pose[:3, 3].astype(np.float64)
    pc = (points.astype(np.float64) - t) @ R
    z = pc[:, 2]
    with np.errstate(divide='ignore', invalid='ignore'):
        u = np.round(cam[0, 0] * pc[:, 0] / z + cam[0, 2]).astype(np.int64)
        v = np.round(cam[1, 1] * pc[:, 1] / z + cam[1, 2]).astype(np.int64)
    ok = (z > 0.05) & (u >= 0) & (u < w) & (v >= 0) & (v < h)
    dz = np.zeros(len(points))
    dz[ok] = depth_mm[v[ok], u[ok]] / 1000.0
    return ok & (np.abs(dz - z) < tol)


def make_chunk(seed=0, num_points=8192, num_views=5, h=120, w=160, drop=0.1, candidates=12):
    """One synthetic chunk: dict of numpy arrays
         points (np,3) f32, depth (nv,h,w) f32 metres, depth_mm (nv,h,w) u16, cam_matrix (4,4) f32 (already
         scaled to h x w), pose (nv,4,4) f32, images (nv,3,h,w) f32, chunk_box (4,) f64."""
    rng = np.random.RandomState(seed + 7919)
    points, boxes = room_points(num_points, seed)
    cam = SCANNET_DEPTH_INTRINSICS.copy()
    cam[0] /= 640.0 / w
    cam[1] /= 480.0 / h
    # candidate views, then the reference's greedy cover (select_frames, scannet_2d3d.py:20-30):
    # repeatedly take the view seeing the most not-yet-covered points
    cand = []
    for _ in range(max(num_views, candidates)):
        eye = rng.uniform([0.2, 0.2, 0.8], [ROOM[0] - 0.2, ROOM[1] - 0.2, 2.0])
        target = np.array([rng.uniform(0.0, ROOM[0]), rng.uniform(0.0, ROOM[1]), rng.uniform(0.0, 2.2)])
        if np.linalg.norm(target - eye) < 0.5:
            target = np.array([ROOM[0] - eye[0], ROOM[1] - eye[1], 0.3])
        pose = _look_at(eye, target)
        d = render_depth(pose, cam, boxes, h, w)
        cand.append((pose, d, _visible(points, pose, cam, d)))
    overlap = np.stack([c[2] for c in cand], axis=1)
    poses, depths = [], []
    for _ in range(num_views):
        f = int(overlap.sum(0).argmax())
        overlap[overlap[:, f]] = False
        pose, d, _ = cand[f]
        d = d.copy()
        d[rng.rand(h, w) < drop] = 0
        poses.append(pose)
        depths.append(d)
    depth_mm = np.stack(depths)
    images = rng.randn(num_views, 3, h, w).astype(np.float32)
    # scannet_2d3d.py:393: the box handed to get_rgbd_data already includes the 0.2 m chunk margin
    box = np.array([0.0, 0.0, ROOM[0], ROOM[1]], np.float64)
    return {'points': points, 'depth_mm': depth_mm, 'depth': depth_mm.astype(np.float32) / 1000.0,
            'cam_matrix': cam, 'pose': np.stack(poses), 'images': images, 'chunk_box': box}


def train_batch(batch, num_points=8192, channels=64, num_classes=20):
    """Seeded inputs of one training step (BASELINE config 2 / SURVEY §8 f4): room chunks (b, n, 3), features
    (b, c, n) ~ N(0, 1), labels (b, n) in [0, num_classes) from a coarse spatial grid with 10 % ignore_index (-100),
    class weights (num_classes,).  Returns (points ndarray, feature Tensor, label LongTensor, weight Tensor)."""
    import torch
    pts = np.stack([room_points(num_points, seed=100 + s)[0] for s in range(batch)])
    g = torch.Generator().manual_seed(1000 + batch)
    feat = torch.randn(batch, channels, num_points, generator=g)
    cell = np.floor(pts / np.array([0.5, 0.5, 0.6], np.float32)).astype(np.int64)
    label = (cell[..., 0] + 4 * cell[..., 1] + 7 * cell[..., 2]) % num_classes
    rng = np.random.RandomState(2000 + batch)
    label[rng.rand(batch, num_points) < 0.1] = -100
    weight = torch.linspace(0.5, 1.5, num_classes)
    return pts, feat, torch.from_numpy(label), weight


def fill_parameters(module, seed=0):
    """Deterministic, init-order-independent parameters and NON-TRIVIAL BatchNorm statistics, keyed by
    state_dict name, so two implementations of the same architecture get identical weights."""
    import torch
    sd = module.state_dict()
    for i, name in enumerate(sorted(sd.keys())):
        t = sd[name]
        if not t.is_floating_point():
            continue
        g = torch.Generator().manual_seed(seed * 100003 + i)
        if name.endswith('running_var'):
            v = torch.rand(t.shape, generator=g) * 0.5 + 0.75
        elif name.endswith('running_mean'):
            v = torch.randn(t.shape, generator=g) * 0.1
        elif name.endswith('bias'):
            v = torch.randn(t.shape, generator=g) * 0.1
        elif t.dim() == 1:                     # BN weight
            v = torch.rand(t.shape, generator=g) * 0.5 + 0.5
        else:                                  # conv / deconv / linear weight: keeps activations O(1) through ~40 layers
            fan_in = t[0].numel()
            v = torch.randn(t.shape, generator=g) * (1.5 / fan_in) ** 0.5
        t.copy_(v.to(t.dtype))
    return module
