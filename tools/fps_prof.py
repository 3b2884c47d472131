"""FPS at the model's level-1 shape (32 clouds x 8192 points -> 2048), for timing and for ncu."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
import mvpnet_b200
from mvpnet_b200 import synthetic

ext = mvpnet_b200.load_ext()
b = int(sys.argv[1]) if len(sys.argv) > 1 else 32
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
m = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
if n <= 8192:
    pts = torch.from_numpy(np.stack([synthetic.room_points(n, s)[0] for s in range(b)])).cuda()
else:
    rng = np.random.RandomState(0)
    pts = torch.from_numpy((rng.rand(b, n, 3) * np.array([6, 8, 2.7])).astype(np.float32)).cuda()
for _ in range(3):
    ext.fps_cuda.farthest_point_sample(pts, m)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    ext.fps_cuda.farthest_point_sample(pts, m)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
print('fps b=%d n=%d m=%d: %.3f ms  (%.3f us / iteration)' % (b, n, m, ms, ms * 1e3 / (m - 1)))
