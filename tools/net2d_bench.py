"""Device time of the whole 2D network (net2d.FastUNetResNet34.features_rows) on 160 views of 120 x 160, eager launches."""
import sys, os, warnings
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvpnet_b200 import net2d, synthetic
from mvpnet_b200.unet import UNetResNet34

with warnings.catch_warnings():
    warnings.simplefilter('ignore')
    net = UNetResNet34(20, p=0.5, pretrained=False)
synthetic.fill_parameters(net, seed=4)
net = net.cuda().eval()
fast = net2d.FastUNetResNet34(net)
x = torch.randn(int(os.environ.get('VIEWS', '160')), 3, 120, 160, device='cuda')
for _ in range(3):
    y = fast.features_rows(x)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
it = 10
s.record()
for _ in range(it):
    y = fast.features_rows(x)
e.record()
torch.cuda.synchronize()
print('net2d eager: %.3f ms per pass, checksum %.6e' % (s.elapsed_time(e) / it, float(y.float().sum())))
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    y = fast.features_rows(x)
g.replay(); torch.cuda.synchronize()
s.record()
for _ in range(it):
    g.replay()
e.record()
torch.cuda.synchronize()
print('net2d graph: %.3f ms per pass, checksum %.6e' % (s.elapsed_time(e) / it, float(y.float().sum())))
