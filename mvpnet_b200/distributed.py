"""Chunk sharding across ranks (one process per GPU) and the single collective of the path.

Chunks are independent units of work (the reference loops over them one at a time,
mvpnet/test_mvpnet_3d.py:142-164): rank r of W takes the contiguous block
[r*ceil(C/W), min((r+1)*ceil(C/W), C)), weights are replicated, nothing is exchanged inside the
forward, and the per-chunk logits are collected with ONE all-gather (NCCL over NVLink on GPUs; the
same code runs on gloo for the CPU tests)."""
import torch
import torch.distributed as dist


def chunk_shard(num_chunks, rank, world_size):
    """Contiguous shard [lo, hi) of `num_chunks` for `rank`; the last ranks may get fewer (or zero)."""
    per = (num_chunks + world_size - 1) // world_size
    lo = min(rank * per, num_chunks)
    return lo, min(lo + per, num_chunks)


def all_gather_chunks(local, num_chunks, out=None):
    """local: (n_local, ...) results of this rank's shard (n_local <= ceil(C/W)).  Returns the
    (num_chunks, ...) tensor in global chunk order on every rank.  One all_gather_into_tensor; ragged
    shards are zero-padded to the common shard size and the padding is dropped afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    per = (num_chunks + world - 1) // world
    if local.size(0) != per:
        pad = local.new_zeros((per - local.size(0),) + tuple(local.shape[1:]))
        local = torch.cat([local, pad], dim=0)
    if out is None:
        out = local.new_empty((world * per,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, local.contiguous())
    return out[:num_chunks]
