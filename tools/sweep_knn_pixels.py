"""Timing sweep of the 2D->3D k-NN (cell size) on the bench workload.  Run on the GPU box:
   for s in 1 2 3 4 6; do MVPNET_B200_KP_CELL_SCALE=$s python tools/sweep_knn_pixels.py; done"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from mvpnet_b200.data import unproject_and_knn
host, _ = bench.make_host_batch(list(range(int(os.environ.get('CHUNKS', '32')))), pin=False)
dev = {k: v.cuda() for k, v in host.items()}
def run():
    return unproject_and_knn(dev['depth'], None, dev['pose'], dev['points'], k=3, chunk_box=dev['chunk_box'], cam_inv=dev['cam_inv'])
ref = run()['knn_indices']
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    out = run()
e.record(); torch.cuda.synchronize()
print('cell_scale', os.environ.get('MVPNET_B200_KP_CELL_SCALE', 'default'), 'ms/step %.3f' % (s.elapsed_time(e) / 5), 'checksum', int(out['knn_indices'].sum()))
