"""CPU, world_size 2, gloo on 127.0.0.1: the N>1 host logic (chunk sharding + the one all-gather)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvpnet_b200.distributed import all_gather_chunks, chunk_shard


def test_chunk_shard_covers_everything_once():
    for c in (1, 7, 32, 256, 257):
        for w in (1, 2, 3, 8):
            spans = [chunk_shard(c, r, w) for r in range(w)]
            covered = [i for lo, hi in spans for i in range(lo, hi)]
            assert covered == list(range(c))
            assert max(hi - lo for lo, hi in spans) == (c + w - 1) // w
    assert chunk_shard(256, 3, 8) == (96, 128)


def _fake_logits(chunk_id):
    g = torch.Generator().manual_seed(chunk_id)
    return torch.randn(20, 64, generator=g)


def _worker(rank, world, port, num_chunks, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = chunk_shard(num_chunks, rank, world)
        local = torch.stack([_fake_logits(i) for i in range(lo, hi)]) if hi > lo else torch.empty(0, 20, 64)
        full = all_gather_chunks(local, num_chunks)
        want = torch.stack([_fake_logits(i) for i in range(num_chunks)])
        ret[rank] = bool(torch.equal(full, want))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('num_chunks', [8, 5])
def test_all_gather_chunks_world2(num_chunks):
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(2, port, num_chunks, ret), nprocs=2, join=True)
    assert ret.get(0) is True and ret.get(1) is True
