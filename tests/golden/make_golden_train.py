"""Golden fixtures for BASELINE config 2 at FULL size and for the training step
(SURVEY §8 f4): the REFERENCE's own Python (`/root/reference`, unmodified: PN2SSG, SegLoss, SegAccuracy, SegIoU) in
TRAIN mode on CPU — batch-statistics BatchNorm, forward + backward — on 8192-point synthetic room chunks with the
default channel widths, batch 1 and batch 32 (the training batch, configs/scannet/mvpnet_3d_unet_resnet34_pn2ssg.yaml:38).
Its six CUDA extension modules are served by the C oracle.  The step follows train_mvpnet_3d.py:158-180:
preds = model(batch); loss = SegLoss(weight)(preds, batch); metrics; loss.backward().

    python tests/golden/make_golden_train.py          (needs /root/reference; ~10 GB of RAM for batch 32)

Stored (kept small: the logits of batch 32 alone would be 21 MB):
  loss, acc_tp / acc_n, conf_mat                 SegLoss value, SegAccuracy counts, SegIoU confusion matrix
  logit_sample [B,20,128], feat_grad_sample      every 64th point of the logits / of d loss / d feature
  gn_<param>                                     L2 norm of every parameter gradient
  g_<param>                                      full gradient of a few parameter tensors across the depth of the net
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference'
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from mvpnet_b200 import compat, synthetic  # noqa: E402

NUM_CLASSES = 20
FULL_GRADS = ('sa_modules.0.mlp.0.conv.weight', 'sa_modules.1.mlp.1.bn.weight', 'sa_modules.3.mlp.2.conv.weight',
              'fp_modules.0.mlp.0.bn.bias', 'fp_modules.3.mlp.2.conv.weight', 'mlp_seg.0.conv.weight', 'seg_logit.weight',
              'seg_logit.bias')


train_inputs = synthetic.train_batch   # seeded inputs of the training step, shared with tests/test_gpu_train.py


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / np.abs(b).max())


def run_step(batch, dtype, PN2SSG, SegLoss, SegAccuracy, SegIoU):
    pts, feat, label, weight = train_inputs(batch)
    net = synthetic.fill_parameters(PN2SSG(64, NUM_CLASSES, dropout_prob=0.0), seed=5).train().to(dtype)
    feat = feat.clone().to(dtype).requires_grad_(True)
    data = {'points': torch.from_numpy(pts.transpose(0, 2, 1).copy()).to(dtype), 'feature': feat, 'seg_label': label}
    preds = net(data)
    loss = sum(SegLoss(weight=weight.to(dtype))(preds, data).values())
    acc, iou = SegAccuracy(), SegIoU(NUM_CLASSES)
    with torch.no_grad():
        acc.update_dict(preds, data)
        iou.update_dict(preds, data)
    loss.backward()
    return {'pts': pts, 'feat': feat, 'net': net, 'preds': preds, 'loss': loss, 'acc': acc, 'iou': iou}


def main():
    """Every quantity is stored from a FLOAT64 run of the reference (the truth) together with the deviation of the
    reference's own FLOAT32 run from it (`noise_*`): train-mode gradients at batch 1 are ill-conditioned (batch statistics
    over as few as 128 samples, cancellation in the BatchNorm backward) — the reference's fp32 CPU result is itself up to
    1e-2 of max away from the exact gradient, so the GPU test asks for "as accurate as the reference's fp32", not for
    agreement with one particular fp32 rounding."""
    compat.install(modules=oracle.ext_modules(), reference_root=REF)
    for missing in ('open3d', 'natsort'):
        sys.modules.setdefault(missing, types.ModuleType(missing))
    from mvpnet.models.pn2.pn2ssg import PN2SSG
    from mvpnet.models.loss import SegLoss
    from mvpnet.models.metric import SegAccuracy, SegIoU
    torch.set_num_threads(os.cpu_count())
    for batch in (1, 32):
        r32 = run_step(batch, torch.float32, PN2SSG, SegLoss, SegAccuracy, SegIoU)
        r64 = run_step(batch, torch.float64, PN2SSG, SegLoss, SegAccuracy, SegIoU)
        lg32, lg64 = r32['preds']['seg_logit'].detach()[:, :, ::64].numpy(), r64['preds']['seg_logit'].detach()[:, :, ::64].numpy()
        fg32, fg64 = r32['feat'].grad[:, :, ::64].numpy(), r64['feat'].grad[:, :, ::64].numpy()
        out = {'loss': np.float64(r64['loss'].item()), 'loss_f32': np.float64(r32['loss'].item()),
               'acc_tp': np.int64(r32['acc'].sum), 'acc_n': np.int64(r32['acc'].count),
               'conf_mat': r32['iou'].mat.numpy().astype(np.int64), 'miou': np.float64(r32['iou'].global_avg),
               'logit_sample': lg64.astype(np.float32), 'noise_logit': np.float64(rel(lg32, lg64)),
               'feat_grad_sample': fg64.astype(np.float32), 'noise_feat_grad': np.float64(rel(fg32, fg64)),
               'in_checksum': np.float64(float(r32['pts'].astype(np.float64).sum()) + float(r32['feat'].detach().double().sum()))}
        p64 = dict(r64['net'].named_parameters())
        for name, p in r32['net'].named_parameters():
            key = name.replace('.', '_')
            g64 = p64[name].grad
            out['gn_' + key] = np.float64(g64.norm().item())
            out['noise_' + key] = np.float64(rel(p.grad.numpy(), g64.numpy()))
            if name in FULL_GRADS:
                out['g_' + key] = g64.numpy().astype(np.float32)
        # running statistics after the step pin the train-mode BatchNorm update (momentum 0.1)
        sd = r64['net'].state_dict()
        for name in ('sa_modules.0.mlp.0.bn.running_mean', 'sa_modules.0.mlp.0.bn.running_var', 'fp_modules.3.mlp.2.bn.running_var'):
            out['rs_' + name.replace('.', '_')] = sd[name].numpy().astype(np.float32)
        path = os.path.join(HERE, 'pn2_train_b%d.npz' % batch)
        np.savez_compressed(path, **out)
        worst = max((float(v), k) for k, v in out.items() if k.startswith('noise_'))
        print('pn2_train_b%d' % batch, 'loss %.6f (f32 %.6f)' % (out['loss'], out['loss_f32']), 'acc %d/%d' % (out['acc_tp'], out['acc_n']),
              'miou %.4f' % out['miou'], 'largest fp32-vs-fp64 deviation of the reference itself: %.2e (%s)' % worst,
              '%.0f KB' % (os.path.getsize(path) / 1024))


if __name__ == '__main__':
    main()
