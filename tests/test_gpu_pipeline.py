"""engine.PipelinedForward: uploads, compute and downloads of consecutive submissions overlap on separate streams, and
every submission returns ITS OWN result (buffer-reuse hazards: input set overwritten before it was read, static output
overwritten before it was copied out, pinned output reused before the download finished)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_pipelined_forward_returns_each_submissions_result():
    from mvpnet_b200 import engine
    dev = torch.device('cuda', 0)
    static_out = torch.empty(1 << 22, device=dev)            # stands in for a CUDA-graph's static output buffer

    def forward(d):
        y = d['x']
        for _ in range(20):                                   # long enough that copies and compute really overlap
            y = torch.sin(y) + d['b']
        static_out.copy_(y)
        return static_out

    host = [{'x': torch.full((1 << 22,), float(i)).pin_memory(), 'b': torch.full((1 << 22,), 0.5 * i).pin_memory()} for i in range(7)]
    pipe = engine.PipelinedForward(forward, host[0], dev, depth=2)
    got = []
    for h in host:
        out, ev = pipe.submit(h)
        got.append((out, ev))
        if len(got) >= 2:                                     # consume with one submission of lag, like a serving loop
            o, e = got[-2]
            e.synchronize()
            i = len(got) - 2
            want = torch.full((4,), float(i))
            for _ in range(20):
                want = torch.sin(want) + 0.5 * i
            assert torch.allclose(o[:4], want, atol=1e-4) and torch.allclose(o[-4:], want, atol=1e-4), i
    torch.cuda.synchronize()
