"""Generate the golden fixtures under tests/golden/ by running the REFERENCE's own Python
(/root/reference, imported unmodified) on CPU, with its six CUDA extension modules served by the
C oracle (oracle/ — the reference has no CPU implementation of them).  Run here (the container
with /root/reference); the .npz files are committed and travel to the GPU box, the reference does
not.

    python tests/golden/make_golden.py

Fixtures (inputs are re-generated from seeds by mvpnet_b200.synthetic in the tests; checksums of
the inputs are stored to catch generator drift):
  rgbd_chunk.npz    reference ScanNet2D3DChunks.get_rgbd_data (depth2xyz, pose, masks, sklearn ball-tree
                    3-NN) on a synthetic chunk written to disk as PNG/txt  -> image_xyz, image_mask, knn_indices
  fa_c1.npz         BASELINE config 1: FeatureAggregation on 1 view, 2048 points
  pn2_small.npz     reference PN2SSG (narrow config, 2048 pts): per-level FPS indices + logits
  pn2_full.npz      reference PN2SSG(64, 20) default config on one 8192-pt chunk: logits
  mvpnet_c3.npz     BASELINE config 3: reference MVPNet3D(UNetResNet34, PN2SSG) on one chunk: logits
"""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from mvpnet_b200 import compat, synthetic  # noqa: E402

PN2_SMALL = dict(sa_channels=((16, 16, 32), (32, 32, 64), (64, 64, 128), (128, 128, 256)),
                 num_centroids=(512, 128, 32, 8), radius=(0.2, 0.4, 0.8, 1.6), max_neighbors=(32, 32, 32, 32),
                 fp_channels=((128, 128), (128, 128), (128, 64), (64, 64, 64)), seg_channels=(64,))


def checksum(a):
    return float(np.asarray(a, np.float64).sum())


def import_reference():
    compat.install(modules=oracle.ext_modules(), reference_root=REF)
    for missing in ('open3d', 'natsort'):                           # imported at module level, never used on this path
        sys.modules.setdefault(missing, types.ModuleType(missing))
    torch.set_grad_enabled(False)


def ref_rgbd(chunk, k=3):
    """Call the reference's get_rgbd_data on files written from a synthetic chunk."""
    from PIL import Image
    from mvpnet.data.scannet_2d3d import ScanNet2D3DChunks
    nv, h, w = chunk['depth_mm'].shape
    with tempfile.TemporaryDirectory() as d:
        scan = os.path.join(d, 'scene0000_00')
        for sub in ('color', 'depth', 'pose'):
            os.makedirs(os.path.join(scan, sub))
        for f in range(nv):
            rgb = np.clip(chunk['images'][f].transpose(1, 2, 0) * 40 + 128, 0, 255).astype(np.uint8)
            Image.fromarray(rgb).save(os.path.join(scan, 'color', '%d.png' % (f * 20)))
            Image.fromarray(chunk['depth_mm'][f]).save(os.path.join(scan, 'depth', '%d.png' % (f * 20)))
            np.savetxt(os.path.join(scan, 'pose', '%d.txt' % (f * 20)), chunk['pose'][f])
        ds = object.__new__(ScanNet2D3DChunks)
        ds.image_dir, ds.num_rgbd_frames, ds.k = d, nv, k
        ds.resize, ds.resize_scale = (w, h), (640.0 / w, 480.0 / h)
        ds.color_jitter, ds.image_normalizer, ds.flip = None, None, 0.0
        npts = chunk['points'].shape[0]
        cam640 = synthetic.SCANNET_DEPTH_INTRINSICS.copy()
        # one base point per frame, each seen by exactly one frame -> greedy selection keeps frame order
        data_dict = {'scan_id': 'scene0000_00', 'base_point_ind': np.arange(nv), 'frame_ids': [f * 20 for f in range(nv)],
                     'pointwise_rgbd_overlap': np.eye(nv, dtype=bool), 'cam_matrix': cam640}
        # numpy-1.x semantics for the margin arithmetic (the reference's era): float64 box
        out = ds.get_rgbd_data(data_dict, chunk['points'], chunk['chunk_box'], np.ones(npts, bool))
    return out


def main():
    import_reference()
    from mvpnet.models.mvpnet_3d import FeatureAggregation, MVPNet3D
    from mvpnet.models.pn2.pn2ssg import PN2SSG
    from mvpnet.models.unet_resnet34 import UNetResNet34
    from mvpnet.ops.group_points import group_points
    torch.set_num_threads(os.cpu_count())

    # ---- rgbd_chunk: the data side (a1 + a2) ----------------------------------------------------
    chunk = synthetic.make_chunk(seed=0)
    out = ref_rgbd(chunk)
    np.savez_compressed(os.path.join(HERE, 'rgbd_chunk.npz'),
                        image_xyz=out['image_xyz'], image_mask=np.packbits(out['image_mask']),
                        knn_indices=out['knn_indices'].astype(np.int32),
                        in_checksum=checksum(chunk['depth']) + checksum(chunk['pose']) + checksum(chunk['points']))
    print('rgbd_chunk', out['image_xyz'].shape, out['image_mask'].mean(), out['knn_indices'][:2])

    # ---- fa_c1: BASELINE config 1 ---------------------------------------------------------------
    c1 = synthetic.make_chunk(seed=1, num_points=2048, num_views=1)
    o1 = ref_rgbd(c1)
    g = torch.Generator().manual_seed(11)
    feat2d = torch.randn(1, 64, 1, 120, 160, generator=g)
    fa = synthetic.fill_parameters(FeatureAggregation(64), seed=3).eval()
    knn = torch.from_numpy(o1['knn_indices'])[None]
    f = group_points(feat2d.reshape(1, 64, -1), knn)
    xyz = group_points(torch.from_numpy(o1['image_xyz'])[None].permute(0, 4, 1, 2, 3).reshape(1, 3, -1), knn)
    pts = torch.from_numpy(c1['points'].T.copy())[None]
    y = fa(xyz, pts, f)
    np.savez_compressed(os.path.join(HERE, 'fa_c1.npz'), out=y.numpy(), knn_indices=o1['knn_indices'].astype(np.int32),
                        in_checksum=checksum(feat2d.numpy()) + checksum(c1['points']))
    print('fa_c1', y.shape, float(y.abs().max()))

    # ---- pn2_small -------------------------------------------------------------------------------
    pts_s, _ = synthetic.room_points(2048, seed=2)
    g = torch.Generator().manual_seed(12)
    feat_s = torch.randn(1, 16, 2048, generator=g)
    net = synthetic.fill_parameters(PN2SSG(16, 20, **PN2_SMALL), seed=4).eval()
    xyz = torch.from_numpy(pts_s.T.copy())[None]
    from mvpnet.ops.fps import farthest_point_sample
    from mvpnet.ops.ball_query import ball_query
    fps_idx, bq_idx, cur = [], [], xyz
    for m, r in zip(PN2_SMALL['num_centroids'], PN2_SMALL['radius']):
        idx = farthest_point_sample(cur, m)
        new = torch.gather(cur, 2, idx[:, None, :].expand(1, 3, m))
        fps_idx.append(idx.numpy()[0].astype(np.int32))
        bq_idx.append(ball_query(new, cur, r, 32).numpy()[0].astype(np.int32))
        cur = new
    logit = net({'points': xyz, 'feature': feat_s})['seg_logit']
    np.savez_compressed(os.path.join(HERE, 'pn2_small.npz'), logit=logit.numpy(),
                        **{'fps%d' % i: v for i, v in enumerate(fps_idx)}, **{'bq%d' % i: v for i, v in enumerate(bq_idx)},
                        in_checksum=checksum(pts_s) + checksum(feat_s.numpy()))
    print('pn2_small', logit.shape, float(logit.abs().max()))

    # ---- pn2_small_train: BASELINE config 2 backward — train-mode BatchNorm, loss = sum(logit * w), gradients
    torch.set_grad_enabled(True)
    net_t = synthetic.fill_parameters(PN2SSG(16, 20, dropout_prob=0.0, **PN2_SMALL), seed=4).train()
    feat_t = feat_s.clone().requires_grad_(True)
    logit_t = net_t({'points': xyz, 'feature': feat_t})['seg_logit']
    wgt = torch.randn(logit_t.shape, generator=torch.Generator().manual_seed(14))
    (logit_t * wgt).sum().backward()
    grads = {n: p_.grad.numpy() for n, p_ in net_t.named_parameters()
             if n in ('sa_modules.0.mlp.0.conv.weight', 'sa_modules.3.mlp.2.conv.weight', 'fp_modules.0.mlp.0.conv.weight',
                      'fp_modules.3.mlp.2.bn.weight', 'seg_logit.weight')}
    np.savez_compressed(os.path.join(HERE, 'pn2_small_train.npz'), logit=logit_t.detach().numpy(), feat_grad=feat_t.grad.numpy(),
                        **{'g_' + k.replace('.', '_'): v for k, v in grads.items()})
    print('pn2_small_train', float(logit_t.abs().max()), float(feat_t.grad.abs().max()))
    torch.set_grad_enabled(False)

    # ---- pn2_full: default PN2SSG(64, 20) on an 8192-pt chunk (BASELINE config 2, forward) ---------
    pts_f, _ = synthetic.room_points(8192, seed=0)
    g = torch.Generator().manual_seed(13)
    feat_f = torch.randn(1, 64, 8192, generator=g)
    net = synthetic.fill_parameters(PN2SSG(64, 20), seed=5).eval()
    logit = net({'points': torch.from_numpy(pts_f.T.copy())[None], 'feature': feat_f})['seg_logit']
    np.savez_compressed(os.path.join(HERE, 'pn2_full.npz'), logit=logit.numpy(),
                        in_checksum=checksum(pts_f) + checksum(feat_f.numpy()))
    print('pn2_full', logit.shape, float(logit.abs().max()))

    # ---- mvpnet_c3: BASELINE config 3 -------------------------------------------------------------
    net2d = UNetResNet34(20, p=0.5, pretrained=False)
    model = MVPNet3D(net2d, None, PN2SSG(64, 20), in_channels=64, mlp_channels=(64, 64, 64), reduction='sum', use_relation=True)
    synthetic.fill_parameters(model, seed=6).eval()
    batch = {'images': torch.from_numpy(chunk['images'])[None], 'image_xyz': torch.from_numpy(out['image_xyz'])[None],
             'knn_indices': torch.from_numpy(out['knn_indices'])[None], 'points': torch.from_numpy(chunk['points'].T.copy())[None]}
    logit = model(batch)['seg_logit']
    feat2d = model.net_2d({'image': batch['images'][0]})['feature']
    np.savez_compressed(os.path.join(HERE, 'mvpnet_c3.npz'), logit=logit.numpy(),
                        feat2d_checksum=checksum(feat2d.numpy()), feat2d_absmax=float(feat2d.abs().max()),
                        feat2d_sample=feat2d[:, :, ::16, ::16].numpy())
    print('mvpnet_c3', logit.shape, float(logit.abs().max()))


if __name__ == '__main__':
    main()
