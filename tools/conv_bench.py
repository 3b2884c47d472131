"""Device time of the tensor-core 3x3 convolution at the UNet's shapes (160 views) next to cuDNN fp32."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from mvpnet_b200 import net2d

torch.backends.cudnn.allow_tf32 = False
torch.backends.cudnn.benchmark = True
N = int(os.environ.get('VIEWS', '160'))
shapes = [('layer1', 64, 80, 64, 0, 64), ('decoder0', 128, 160, 64, 64, 64), ('decoder1', 64, 80, 64, 64, 64), ('layer2', 32, 40, 128, 0, 128),
          ('decoder2', 32, 40, 128, 128, 128), ('layer3', 16, 20, 256, 0, 256), ('decoder3', 16, 20, 256, 256, 256), ('layer4', 8, 10, 512, 0, 512)]


def t(fn, it=5):
    fn(); fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it


for name, h, w, c1, c2, co in shapes:
    x1 = torch.randn(N, h, w, c1, device='cuda')
    x2 = torch.randn(N, h, w, c2, device='cuda') if c2 else None
    wt = torch.randn(co, c1 + c2, 3, 3, device='cuda') * 0.05
    b = torch.zeros(co, device='cuda')
    packed, bias = net2d.pack_conv3x3(wt, b)
    p1 = net2d.Planar.from_nhwc(x1); p2 = None if x2 is None else net2d.Planar.from_nhwc(x2)
    mine = t(lambda: net2d.conv3x3(p1, packed, bias, x2=p2, relu=True))
    xin = (x1 if x2 is None else torch.cat([x1, x2], 3)).permute(0, 3, 1, 2).contiguous()
    ref = 0.0 if os.environ.get('NO_CUDNN') else t(lambda: F.relu_(F.conv2d(xin, wt, b, padding=1)))
    gf = 2 * N * h * w * (c1 + c2) * co * 9 / 1e9
    print('%-9s %3dx%-3d %3d+%-3d->%3d  tc %.3f ms (%.0f TF/s fp32-equiv, %.0f issued)   cudnn-nchw %.3f ms  x%.1f' %
          (name, h, w, c1, c2, co, mine, gf / mine, 3 * gf / mine, ref, ref / mine), flush=True)
