"""The callers on either side of the per-chunk forward, device-resident:

  * `scene2chunks_legacy` — sliding xy chunks over a whole scene (mvpnet/utils/chunk_util.py:4-53), same corner
    arithmetic (float64), same inclusive comparisons, same chunk order; index lists stay on the device of `points`.
  * `VoteAccumulator` — what test_mvpnet_3d.py:136-175 does in numpy around `model(data_batch)`: per-point logit sums
    and prediction counts over the chunks, mean, arg-max, points without prediction labelled `num_classes`.  The adds
    happen chunk by chunk in call order (indices inside a chunk are unique), so the fp32 sums are the reference's.

  * `select_frames` — greedy choice of the RGB-D frames that cover the most still-uncovered base points
    (mvpnet/data/scannet_2d3d.py:20-30), ties resolved like numpy's argmax (lowest frame index).

Pure torch (index arithmetic and bandwidth-bound adds): runs on CPU tensors too, which is how the CPU tests pin it
against the reference's own numpy code (tests/golden/make_golden_scene.py).
"""
import math

import torch


def scene2chunks_legacy(points, chunk_size, stride, thresh=1000, margin=(0.2, 0.2), return_bbox=False):
    """points (num_points, 3) tensor -> list of int64 index tensors [, list of (6,) float64 bbox tensors]."""
    assert points.dim() == 2 and points.size(1) == 3
    pts = points.double()                      # the reference compares float32 coordinates with float64 corners
    chunk = torch.as_tensor(chunk_size, dtype=torch.float64, device=points.device)
    marg = torch.as_tensor(margin, dtype=torch.float64, device=points.device)
    coord_max, coord_min = pts.max(dim=0)[0], pts.min(dim=0)[0]
    limit = (coord_max - coord_min)[:2].cpu()
    num_chunks = [int(math.ceil((float(limit[a]) - float(chunk[a])) / stride)) + 1 for a in range(2)]
    cmin = coord_min.cpu()
    xy = pts[:, :2]
    chunk_indices, chunk_bboxes = [], []
    for i in range(num_chunks[0]):
        for j in range(num_chunks[1]):
            corner = torch.tensor([float(cmin[0]) + i * stride, float(cmin[1]) + j * stride], dtype=torch.float64, device=points.device)
            inside = ((xy >= corner) & (xy <= corner + chunk)).all(dim=1)
            if int(inside.sum()) < thresh:     # discard unqualified chunks
                continue
            mask = ((xy >= corner - marg) & (xy <= corner + chunk + marg)).all(dim=1)
            idx = torch.nonzero(mask, as_tuple=False).squeeze(1)
            chunk_indices.append(idx)
            if return_bbox:
                z = pts[idx, 2]
                chunk_bboxes.append(torch.cat([corner - marg, z.min().reshape(1), corner + chunk + marg, z.max().reshape(1)]))
    return (chunk_indices, chunk_bboxes) if return_bbox else chunk_indices


class VoteAccumulator:
    """Whole-scene logits from overlapping chunk predictions (test_mvpnet_3d.py:136-175)."""

    def __init__(self, num_points, num_classes, device):
        self.num_classes = num_classes
        self.logit_sum = torch.zeros(num_points, num_classes, dtype=torch.float32, device=device)
        self.count = torch.zeros(num_points, dtype=torch.int32, device=device)

    def add(self, chunk_ind, seg_logit):
        """chunk_ind (n,) int64 scene indices of the chunk's points; seg_logit (num_classes, >= n): the chunk's
        prediction, possibly padded with re-sampled points at the end (test_mvpnet_3d.py:147-163) — only the first n
        columns count."""
        n = chunk_ind.numel()
        self.logit_sum.index_add_(0, chunk_ind, seg_logit[:, :n].t().to(torch.float32))
        self.count.index_add_(0, chunk_ind, torch.ones(n, dtype=torch.int32, device=self.count.device))

    def add_batch(self, chunk_inds, seg_logits):
        """Several chunks of one forward: list of index tensors + (b, num_classes, np) logits, added in chunk order."""
        for b, ind in enumerate(chunk_inds):
            self.add(ind, seg_logits[b])

    def finalize(self):
        """-> (mean logits (num_points, num_classes), labels (num_points,) int64; `num_classes` where no prediction)."""
        mean = self.logit_sum / self.count.clamp(min=1).unsqueeze(1).to(torch.float32)
        label = mean.argmax(dim=1)
        label[self.count == 0] = self.num_classes
        return mean, label


def select_frames(rgbd_overlap, num_rgbd_frames):
    """rgbd_overlap (num_basepoints, num_frames) bool tensor: point p is seen by frame f.  Returns the list of
    `num_rgbd_frames` frame indices chosen greedily by remaining coverage (scannet_2d3d.py:20-30)."""
    overlap = rgbd_overlap.clone()             # the reference copies too: the input is not modified
    selected = []
    for _ in range(num_rgbd_frames):
        counts = overlap.sum(dim=0)
        # first index of the maximum, as numpy.argmax (torch.argmax does not promise which of several maxima)
        frame_idx = int(torch.nonzero(counts == counts.max(), as_tuple=False)[0])
        selected.append(frame_idx)
        overlap[overlap[:, frame_idx].clone()] = False  # every point covered by this frame stops counting (mask copied: it aliases the target)
    return selected
