"""A few tensor-core convolution launches at the UNet's shapes, for `ncu --set full -k regex:tc_conv`."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvpnet_b200 import net2d

N = int(os.environ.get('VIEWS', '160'))
which = os.environ.get('SHAPES', 'layer1,layer3').split(',')
shapes = {'layer1': (64, 80, 64, 0, 64), 'decoder0': (128, 160, 64, 64, 64), 'layer2': (32, 40, 128, 0, 128),
          'layer3': (16, 20, 256, 0, 256), 'layer4': (8, 10, 512, 0, 512)}
for name in which:
    h, w, c1, c2, co = shapes[name]
    x1 = torch.randn(N, h, w, c1, device='cuda')
    x2 = torch.randn(N, h, w, c2, device='cuda') if c2 else None
    r = torch.randn(N, h, w, co, device='cuda')
    packed, bias = net2d.pack_conv3x3(torch.randn(co, c1 + c2, 3, 3, device='cuda') * 0.05, torch.zeros(co, device='cuda'))
    p1 = net2d.Planar.from_nhwc(x1); p2 = None if x2 is None else net2d.Planar.from_nhwc(x2); pr = net2d.Planar.from_nhwc(r)
    for _ in range(2):
        net2d.conv3x3(p1, packed, bias, x2=p2, residual=pr, relu=True)
    torch.cuda.synchronize()
