"""Device time of the general tensor-core convolution (csrc/tc_convg.cu) at the UNet's shapes (160 views of 128 x 160)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mvpnet_b200 import net2d

N = int(os.environ.get('VIEWS', '160'))


def t(fn, it=10):
    fn(); fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it


def planar(n, h, w, c):
    return net2d.Planar.from_nhwc(torch.randn(n, h, w, c, device='cuda'))


rows = []
# stem: 7x7 on the row-unfolded image
img = torch.randn(N, 3, 128, 160, device='cuda')
x = net2d.Planar(net2d.load_ext().fused_cuda.unfold_stem(img), N, 128, 160, 32)
stem = net2d.pack_stem7x7(torch.randn(64, 3, 7, 7, device='cuda') * 0.05, torch.zeros(64, device='cuda'))
rows.append(('stem 7x7 3->64 128x160', t(lambda: net2d.conv_general(x, *stem, stride=1, relu=True)), N * 128 * 160 * (32 + 64) * 4))
for name, h, w, ci, co in (('layer2.0', 64, 80, 64, 128), ('layer3.0', 32, 40, 128, 256), ('layer4.0', 16, 20, 256, 512)):
    xx = planar(N, h, w, ci)
    c3 = net2d.pack_conv_taps(torch.randn(co, ci, 3, 3, device='cuda') * 0.05, torch.zeros(co, device='cuda'))
    c1 = net2d.pack_conv_taps(torch.randn(co, ci, 1, 1, device='cuda') * 0.05, torch.zeros(co, device='cuda'))
    io = N * (h * w * ci + h * w // 4 * co) * 4
    rows.append(('%s 3x3/s2 %d->%d %dx%d' % (name, ci, co, h, w), t(lambda: net2d.conv_general(xx, *c3, stride=2, relu=True)), io))
    rows.append(('%s 1x1/s2 %d->%d %dx%d' % (name, ci, co, h, w), t(lambda: net2d.conv_general(xx, *c1, stride=2, relu=False)), io))
for name, h, w, ci, co in (('deconv4', 8, 10, 512, 256), ('deconv3', 16, 20, 256, 128), ('deconv2', 32, 40, 128, 64), ('deconv1', 64, 80, 64, 64)):
    xx = planar(N, h, w, ci)
    d = net2d.pack_deconv2x2(torch.randn(ci, co, 2, 2, device='cuda') * 0.05, torch.zeros(co, device='cuda'))
    rows.append(('%s 2x2T %d->%d %dx%d' % (name, ci, co, h, w), t(lambda: net2d.deconv2x2(xx, *d, relu=True)), N * h * w * (ci + 4 * co) * 4))
tot = 0.0
for name, ms, io in rows:
    tot += ms
    print('%-34s %.4f ms   %7.1f MB in+out  %6.0f GB/s' % (name, ms, io / 1e6, io / ms / 1e6), flush=True)
print('total %.4f ms' % tot)
