"""Data side of FeatureAggregation on the GPU: depth unprojection + 2D->3D k-NN.

Replaces, per chunk, what the reference computes in DataLoader workers with numpy + scikit-learn
(mvpnet/data/scannet_2d3d.py:33-39 depth2xyz, :255-262 pose/validity, :273-281 chunk-box mask,
:298-313 ball-tree k-NN + remap to flat pixel ids) and emits the same dictionary entries
(`image_xyz`, `image_mask`, `knn_indices`) the model consumes (mvpnet_3d.py:98-109).
"""
import numpy as np
import torch

from . import load_ext


def invert_intrinsics(cam_matrix):
    """float32 inverse of the 3x3 intrinsics, computed exactly as the reference does
    (np.linalg.inv on the float32 matrix, scannet_2d3d.py:38) — on the host: it is 9 numbers of
    metadata and LAPACK's rounding is part of the contract."""
    cam = cam_matrix.detach().cpu().numpy() if torch.is_tensor(cam_matrix) else np.asarray(cam_matrix)
    cam = cam.astype(np.float32, copy=False)[..., :3, :3]
    return np.linalg.inv(cam).astype(np.float32)


def unproject_and_knn(depth, cam_matrix, pose, points, k=3, chunk_box=None, cam_inv=None):
    """depth (b,nv,h,w) f32 metres [CUDA]; cam_matrix (b,3|4,3|4) or (b,nv,..) intrinsics already scaled to
    h x w (host or device); pose (b,nv,4,4) f32 [CUDA]; points (b,np,3) f32 [CUDA]; chunk_box (b,4) f64
    {x0,y0,x1,y1} or None.  Returns image_xyz (b,nv,h,w,3) f32, image_mask (b,nv,h,w) bool,
    knn_indices (b,np,k) int64 flat pixel ids in [0, nv*h*w)."""
    ext = load_ext()
    b, nv, h, w = depth.shape
    if cam_inv is None:
        cam_inv = invert_intrinsics(cam_matrix)
    if not torch.is_tensor(cam_inv):
        cam_inv = torch.from_numpy(np.ascontiguousarray(cam_inv))
    if cam_inv.dim() == 3:                                   # one intrinsics per chunk -> per view
        cam_inv = cam_inv[:, None].expand(b, nv, 3, 3)
    cam_inv = cam_inv.to(depth.device, torch.float32).contiguous()
    box = None if chunk_box is None else chunk_box.to(depth.device, torch.float64).contiguous()
    from .engine import _stage
    with _stage('unproject'):
        xyz32, mask, xyz64 = ext.unproject_cuda.unproject(depth.contiguous(), cam_inv, pose.contiguous(), box, True)
    query = points.to(torch.float64).contiguous()            # sklearn promotes the float32 chunk points
    with _stage('knn_pixels'):
        index, _ = ext.unproject_cuda.knn_pixels(query, xyz64, mask.reshape(b, -1), k)
    return {'image_xyz': xyz32, 'image_mask': mask.bool(), 'knn_indices': index}
