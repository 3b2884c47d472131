"""Training-step pieces on the device: the loss, the metrics and the step of
mvpnet/train_mvpnet_3d.py:158-180, with the reference's class names and call signatures.

  SegLoss        mvpnet/models/loss.py:5-21       weighted cross entropy, ignore_index, mean over the non-ignored points
  SegAccuracy    mvpnet/models/metric.py:5-23     fraction of non-ignored points whose argmax equals the label
  SegIoU         mvpnet/models/metric.py:26-73    confusion matrix -> per-class IoU -> mean

One kernel pass over the logits (csrc/train_ops.cu) produces the loss terms, the log-sum-exp kept for the backward and
the confusion matrix both metrics derive from; reductions have a fixed order (the loss is the same bits on every run)
and nothing synchronises with the host until a number is read (`.item()` / `global_avg`), whereas the reference's
SegAccuracy calls `.item()` every step.  `train_step` is the body of the reference loop."""
import torch
from torch import nn

from . import load_ext


class _SegLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logit, label, weight, ignore_index):
        ext = load_ext()
        logit = logit.contiguous()
        out, lse, conf = ext.train_cuda.seg_loss_forward(logit, label.contiguous(), weight, ignore_index)
        ctx.save_for_backward(logit, label, lse, out)
        ctx.weight, ctx.ignore_index = weight, ignore_index
        ctx.mark_non_differentiable(conf)
        return out[0], conf

    @staticmethod
    def backward(ctx, grad_loss, _grad_conf):
        logit, label, lse, out = ctx.saved_tensors
        g = load_ext().train_cuda.seg_loss_backward(logit, label.contiguous(), ctx.weight, lse, out,
                                                   grad_loss.reshape(1).float().contiguous(), ctx.ignore_index)
        return g, None, None, None


def seg_loss_and_confusion(logit, label, weight=None, ignore_index=-100):
    """(loss scalar tensor with grad, confusion matrix int64 (C, C) of this batch) from one pass over the logits."""
    if weight is not None:
        weight = weight.to(device=logit.device, dtype=torch.float32).contiguous()
    return _SegLossFn.apply(logit, label, weight, ignore_index)


class SegLoss(nn.Module):
    """Segmentation loss — same constructor / forward as mvpnet.models.loss.SegLoss; `last_confusion` holds the
    confusion matrix of the last batch (shared with the metrics so the logits are read once)."""

    def __init__(self, weight=None, ignore_index=-100):
        super().__init__()
        self.weight = weight
        self.ignore_index = ignore_index
        self.last_confusion = None

    def forward(self, preds, labels):
        loss, conf = seg_loss_and_confusion(preds['seg_logit'], labels['seg_label'], self.weight, self.ignore_index)
        self.last_confusion = conf
        return {'seg_loss': loss}


def _confusion(preds, labels, ignore_index, given):
    if given is not None:
        return given
    logit = preds['seg_logit'].detach().contiguous()
    conf = torch.zeros(logit.size(1), logit.size(1), dtype=torch.int64, device=logit.device)
    load_ext().train_cuda.seg_confusion(logit, labels['seg_label'].contiguous(), ignore_index, conf)
    return conf


class SegAccuracy(object):
    """Segmentation accuracy (AverageMeter semantics of the reference: running sum / count); device-side counters."""
    name = 'seg_acc'

    def __init__(self, ignore_index=-100):
        self.ignore_index = ignore_index
        self.reset()

    def reset(self):
        self.sum_t, self.count_t = None, None

    def update_dict(self, preds, labels, confusion=None):
        conf = _confusion(preds, labels, self.ignore_index, confusion)
        tp, n = conf.diagonal().sum(), conf.sum()
        self.sum_t = tp if self.sum_t is None else self.sum_t + tp
        self.count_t = n if self.count_t is None else self.count_t + n

    @property
    def sum(self):
        return 0 if self.sum_t is None else int(self.sum_t.item())

    @property
    def count(self):
        return 0 if self.count_t is None else int(self.count_t.item())

    @property
    def global_avg(self):
        return self.sum / self.count if self.count else float('nan')

    def __str__(self):
        return '{:.4f}'.format(self.global_avg)


class SegIoU(object):
    """Segmentation IoU from the running confusion matrix (rows = label, columns = prediction)."""
    name = 'seg_iou'

    def __init__(self, num_classes, ignore_index=-100):
        self.num_classes = num_classes
        self.ignore_index = ignore_index
        self.mat = None

    def update_dict(self, preds, labels, confusion=None):
        conf = _confusion(preds, labels, self.ignore_index, confusion)
        self.mat = conf.clone() if self.mat is None else self.mat + conf

    def reset(self):
        self.mat = None

    @property
    def iou(self):
        h = self.mat.float()
        return torch.diag(h) / (h.sum(1) + h.sum(0) - torch.diag(h))

    @property
    def global_avg(self):
        return self.iou.mean().item()

    def __str__(self):
        return '{iou:.4f}'.format(iou=self.iou.mean().item())

    @property
    def summary_str(self):
        return str(self)


def train_step(model, loss_fn, data_batch, optimizer=None, metrics=(), max_grad_norm=0.0):
    """One iteration of mvpnet/train_mvpnet_3d.py:158-180: forward, loss, metrics (no_grad), backward, optional
    gradient clipping and optimizer step.  Returns (preds, loss_dict); nothing here synchronises with the host."""
    preds = model(data_batch)
    if optimizer is not None:
        optimizer.zero_grad()
    loss_dict = loss_fn(preds, data_batch)
    total_loss = sum(loss_dict.values())
    with torch.no_grad():
        conf = getattr(loss_fn, 'last_confusion', None)
        for metric in metrics:
            metric.update_dict(preds, data_batch, conf) if conf is not None else metric.update_dict(preds, data_batch)
    total_loss.backward()
    if max_grad_norm > 0:
        nn.utils.clip_grad_norm_(model.parameters(), max_norm=max_grad_norm)
    if optimizer is not None:
        optimizer.step()
    return preds, loss_dict
