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


def scene2chunks_legacy(points, chunk_size, stride, thresh=1000, margin=(0.2, 0.2), return_bbox=False, promotion='nep50'):
    """points (num_points, 3) tensor -> list of int64 index tensors [, list of (6,) float64 bbox tensors].

    All candidate windows are evaluated in ONE pass on the device of `points` (a windows x points membership matrix);
    the host is consulted a constant number of times per scene (bounding box, window counts, total index count) instead
    of once per candidate window.

    Arithmetic follows the reference line by line (chunk_util.py:22-47) for float32 `points`:
      limit = max - min in float32 (:24-26); num_chunks from float64 (limit - chunk_size) / stride (:28);
      corner = coord_min[a] + i * stride (:32): np.float32 + python float.  promotion='nep50' (NumPy >= 2, the
      environment the golden fixtures were generated in): the python float is cast to float32 and the sum is float32;
      promotion='legacy' (NumPy 1.x value-based casting, the reference's era): the sum is float64.
      `xy >= corner` compares float32 with the corner's dtype, `xy <= corner + chunk_size` and both margin tests
      compare in float64 (chunk_size / margin are float64 arrays) (:40, :44)."""
    import numpy as np
    assert points.dim() == 2 and points.size(1) == 3
    dev = points.device
    pts = points if points.dtype == torch.float32 else points.float()
    lohi = torch.stack([pts.min(dim=0)[0], pts.max(dim=0)[0]]).cpu().numpy()                 # host visit 1: 6 floats
    cmin, cmax = lohi[0], lohi[1]
    chunk = np.asarray(chunk_size, dtype=np.float64)
    marg = np.asarray(margin, dtype=np.float64)
    limit = (cmax - cmin).astype(np.float32)
    num_chunks = np.ceil((limit[:2].astype(np.float64) - chunk) / stride).astype(int) + 1
    corners = []
    for i in range(num_chunks[0]):
        for j in range(num_chunks[1]):
            if promotion == 'nep50':
                corners.append((np.float32(cmin[0]) + np.float32(i * stride), np.float32(cmin[1]) + np.float32(j * stride)))
            else:
                corners.append((np.float64(cmin[0]) + i * stride, np.float64(cmin[1]) + j * stride))
    if not corners:
        return ([], []) if return_bbox else []
    corner64 = torch.from_numpy(np.asarray(corners, dtype=np.float64)).to(dev)                # (W, 2), exact either way
    xy64 = pts[:, :2].double()[None]                                                          # (1, N, 2)
    c = corner64[:, None, :]                                                                  # (W, 1, 2)
    chunk_t, marg_t = torch.from_numpy(chunk).to(dev), torch.from_numpy(marg).to(dev)
    # float32 >= float32 equals the comparison of the exactly-converted float64 values
    inside = ((xy64 >= c) & (xy64 <= c + chunk_t)).all(dim=2)                                 # (W, N)
    with_margin = ((xy64 >= c - marg_t) & (xy64 <= c + chunk_t + marg_t)).all(dim=2)
    counts = torch.stack([inside.sum(dim=1), with_margin.sum(dim=1)]).cpu().numpy()           # host visit 2: 2 W integers
    keep = np.nonzero(counts[0] >= thresh)[0]                                                 # discard unqualified chunks
    if keep.size == 0:
        return ([], []) if return_bbox else []
    keep_t = torch.from_numpy(keep).to(dev)
    sel = with_margin.index_select(0, keep_t)
    flat = torch.nonzero(sel, as_tuple=False)[:, 1]                                           # host visit 3 (size of the result)
    chunk_indices = list(torch.split(flat, [int(n) for n in counts[1][keep]]))
    if not return_bbox:
        return chunk_indices
    z = pts[:, 2].double()[None].expand(sel.size(0), -1)
    inf = torch.full_like(z, float('inf'))
    zmin, zmax = torch.where(sel, z, inf).min(dim=1)[0], torch.where(sel, z, -inf).max(dim=1)[0]
    ck = corner64.index_select(0, keep_t)
    boxes = torch.cat([ck - marg_t, zmin[:, None], ck + chunk_t + marg_t, zmax[:, None]], dim=1)
    return chunk_indices, list(boxes)


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


def select_frames_device(rgbd_overlap, num_rgbd_frames):
    """select_frames without a host round trip per pick: returns an int64 tensor (num_rgbd_frames,) on the device of
    `rgbd_overlap`; same greedy rule and the same lowest-index tie break (scannet_2d3d.py:20-30)."""
    overlap = rgbd_overlap.clone()
    nf = overlap.size(1)
    ar = torch.arange(nf, device=overlap.device)
    picks = []
    for _ in range(num_rgbd_frames):
        counts = overlap.sum(dim=0)
        first = torch.where(counts == counts.max(), ar, torch.full_like(ar, nf)).min()        # first maximum, stays on the device
        picks.append(first)
        covered = overlap.index_select(1, first.reshape(1)).squeeze(1)
        overlap = overlap & ~covered[:, None]
    return torch.stack(picks)


def propagate_nearest(points, vote_points, vote_logits):
    """Whole-scene prediction from subsampled votes (test_3d_scene.py:152-165): every scene point takes the logits of its
    nearest sampled point in each vote (sklearn NearestNeighbors(1, 'ball_tree') there; this package's exact grid k-NN
    kernel with k = 1 here, float64 like sklearn, ties to the lowest index), averaged over the votes.
    points (n, 3) f32; vote_points (v, m, 3) f32; vote_logits (v, c, m) f32 -> (mean logits (n, c), labels (n,))."""
    from . import load_ext
    ext = load_ext()
    v, m = vote_points.size(0), vote_points.size(1)
    query = points.double()[None].expand(v, -1, -1).contiguous()
    mask = torch.ones(v, m, dtype=torch.uint8, device=points.device)
    index, _ = ext.unproject_cuda.knn_pixels(query, vote_points.double().contiguous(), mask, 1)      # (v, n, 1)
    gathered = torch.gather(vote_logits, 2, index[:, :, 0].unsqueeze(1).expand(-1, vote_logits.size(1), -1))   # (v, c, n)
    mean = (gathered.sum(dim=0) / v).t().contiguous()
    return mean, mean.argmax(dim=1)


def whole_scene_forward(net, points, nb_pts, vote_indices):
    """test_3d_scene.py:120-165 on the device: `vote_indices` (v, nb_pts) int64 are the sampled point ids of every vote
    (the script draws them with np.random.choice); one batched forward of the votes, then 1-NN propagation + mean."""
    vp = points.index_select(0, vote_indices.reshape(-1)).reshape(vote_indices.size(0), nb_pts, 3)
    batch = {'points': vp.transpose(1, 2).contiguous()}
    with torch.no_grad():
        logits = net.fast_forward(batch)['seg_logit'] if hasattr(net, 'fast_forward') and not net.training else net(batch)['seg_logit']
    return propagate_nearest(points, vp, logits)


def chunked_scene_forward(forward, chunk_indices, batch_of, num_points, num_classes, device):
    """test_mvpnet_3d.py:136-175 around any per-chunk forward: `batch_of(ind)` builds the model input of one chunk from
    its scene indices, `forward(batch)` returns seg_logit (1, num_classes, >= len(ind)).  Votes are accumulated on the
    device in chunk order; nothing synchronises with the host inside the loop.  -> (mean logits, labels)."""
    votes = VoteAccumulator(num_points, num_classes, device)
    with torch.no_grad():
        for ind in chunk_indices:
            votes.add(ind, forward(batch_of(ind))[0])
    return votes.finalize()
