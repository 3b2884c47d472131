"""Timing of the 2D network (out of kernel scope, cuDNN) under memory-format variants, strict fp32."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
x = torch.randn(160, 3, 120, 160, device='cuda')
def run(tag, model, inp, n=5):
    with torch.no_grad():
        for _ in range(3): model.net_2d.features(inp)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n): y = model.net_2d.features(inp)
        e.record(); torch.cuda.synchronize()
    print(tag, 'ms/step %.2f' % (s.elapsed_time(e) / n), 'out strides', y.stride())
m = bench.build_model('cuda')
run('nchw', m, x)
m.net_2d.to(memory_format=torch.channels_last)
run('channels_last weights, nchw input', m, x)
run('channels_last weights + input', m, x.contiguous(memory_format=torch.channels_last))
torch.backends.cudnn.allow_tf32 = True
run('[tf32, not parity-valid] channels_last', m, x.contiguous(memory_format=torch.channels_last))
